// phmm_sw.cu -- Smith-Waterman haplotype -> reference alignment with back-track on sm_100a.
//
// Replaces, for a BATCH of (reference window, haplotype) pairs, hc::IntelSWAligner::align
// (smithwaterman/intel_smithwaterman.hpp:29-44) = runSWOnePairBT_avx2 (native/PairWiseSW.h:41-447) with the
// SOFTCLIP overhang strategy, plus the aligner's all-match shortcut.  Integer work, bit-exact: same scores,
// same back-track bits, same end-cell tie rules, hence the same CIGAR and offset.
//
// Not a translation.  The reference sweeps anti-diagonals with 8-lane AVX2 vectors over rolling H/E/F arrays
// and streams 16-bit back-track words; here
//   * one WARP per alignment; lane l owns C consecutive haplotype columns (C = 4..32 by haplotype length)
//     with H(row-1), F and the haplotype bases of its columns in registers;
//   * rows stream through the lanes as a wavefront (lane l works on row t - l at step t): the left
//     neighbour's H and E of the row arrive by shuffle, the diagonal is the value received one step earlier;
//   * back-track codes are FOUR BITS per cell (direction 0..2 | extension flags 4, 8: PairWiseSW.h:60-70), two
//     cells per byte, C / 2 contiguous bytes per lane and row, in a global scratch matrix whose pitch is the
//     alignment's own ceil(ncol / 2) rounded up to 16 bytes (a 415 x 415 pair: 86 KB; one byte per cell at a pitch
//     of 32 C took 212 KB).  The host caps a launch by a scratch budget and walks larger batches in pieces;
//   * the last row and last column of H go to shared memory; lane 0 replays the reference's anti-diagonal
//     order over them to pick the end cell (PairWiseSW.h:329-357), walks the back-track matrix
//     (getCIGAR, :367-520), merges equal neighbours and writes (op, length) pairs in CIGAR order.
// The cell update is PairWiseSW.h:123-159 (MAIN_CODE) operation for operation.
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <algorithm>
#include <cstdlib>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <vector>
#include <cuda_runtime.h>

#include "../../include/phmm.h"

namespace {

constexpr int kSwMaxLen = PHMM_SW_MAX_LEN + 1;           // MAX_SEQ_LEN of the reference (smithwaterman_common.h)
constexpr int32_t kLowInit = INT32_MIN / 2;              // LOW_INIT_VALUE
constexpr int32_t kMinCutoff = -100000000;               // MATRIX_MIN_CUTOFF
enum : int { kMatch = 0, kInsert = 1, kDelete = 2, kInsertExt = 4, kDeleteExt = 8, kSoftclip = 9 };

struct SwJob {
    int32_t ref_off, nrow, alt_off, ncol;
    int64_t bt_off;                                      // this alignment's back-track matrix in the scratch
    int32_t index;                                       // position in the caller's batch
    int32_t pitch;                                       // bytes per back-track row: ceil(ncol / 2) rounded up to 16
};

struct SwArgs {
    const uint8_t* ref_bases;
    const uint8_t* alt_bases;
    const SwJob* jobs;
    int32_t n_jobs;
    int32_t w_match, w_mismatch, w_open, w_extend;
    uint8_t* bt;
    int64_t cap_elems;                                   // capacity of the compact element arrays
    unsigned long long* cursor;                          // elements handed out so far
    int32_t* offset; int32_t* n_elems; int64_t* elem_start; uint8_t* ops; int32_t* lens;
};

constexpr int kSwMaxElems = 2 * (PHMM_SW_MAX_LEN + 1);   // a CIGAR of two such sequences cannot have more elements

template <int C>
__global__ void __launch_bounds__(32) sw_kernel(const SwArgs a)
{
    __shared__ uint8_t s_seq1[kSwMaxLen];
    __shared__ int32_t s_lastrow[kSwMaxLen + 1];
    __shared__ int32_t s_lastcol[kSwMaxLen + 1];
    __shared__ uint8_t s_ops[kSwMaxElems];
    __shared__ int32_t s_lens[kSwMaxElems];
    const int lane = threadIdx.x;
    if ((int)blockIdx.x >= a.n_jobs) return;
    const SwJob job = a.jobs[blockIdx.x];
    const int nrow = job.nrow, ncol = job.ncol;
    const uint8_t* seq1 = a.ref_bases + job.ref_off;
    const uint8_t* seq2 = a.alt_bases + job.alt_off;
    uint8_t* bt = a.bt + job.bt_off;
    const int PITCH = job.pitch;

    for (int i = lane; i < nrow; i += 32) s_seq1[i] = seq1[i];
    const int j0 = lane * C;                             // columns j0+1 .. j0+C (1-based)
    int32_t Hp[C], F[C];
    uint8_t s2[C];
#pragma unroll
    for (int c = 0; c < C; c++) {
        Hp[c] = 0; F[c] = kLowInit;                      // row 0: H = 0, F = LOW (PairWiseSW.h:204-224,318-327)
        s2[c] = (j0 + c < ncol) ? seq2[j0 + c] : (uint8_t)0;
    }
    __syncwarp();

    int32_t outH = 0, outE = kLowInit;                   // H, E of this lane's last column on its current row
    int32_t dgL = 0;                                     // H(row-1, j0)
    const int own = (ncol - 1) / C;                      // lane that owns column ncol
    const int steps = nrow + 31;
    for (int t = 0; t < steps; ++t) {
        const int i = t - lane + 1;                      // 1-based row of this lane at this step
        int32_t inH = __shfl_up_sync(0xffffffffu, outH, 1);
        int32_t inE = __shfl_up_sync(0xffffffffu, outE, 1);
        if (lane == 0) { inH = 0; inE = kLowInit; }      // column 0: H = 0, E = LOW
        if (i >= 1 && i <= nrow) {
            const uint8_t b1 = s_seq1[i - 1];
            int32_t hl = inH, e = inE, hd = dgL;
            __align__(16) uint8_t row[C / 2];
#pragma unroll
            for (int c = 0; c < C; c++) {
                // MAIN_CODE (PairWiseSW.h:123-159)
                const int32_t ext_h = e + a.w_extend, open_h = hl + a.w_open;
                const int32_t e11 = max(open_h, ext_h);
                int ext = (open_h > ext_h) ? 0 : kInsertExt;
                const int32_t ext_v = F[c] + a.w_extend, open_v = Hp[c] + a.w_open;
                const int32_t f11 = max(ext_v, open_v);
                if (!(open_v > ext_v)) ext |= kDeleteExt;
                const int32_t m11 = hd + (b1 == s2[c] ? a.w_match : a.w_mismatch);
                int32_t h11 = max(kMinCutoff, m11);
                int b = kMatch;
                if (e11 > h11) { b = kInsert; h11 = e11; }
                if (f11 > h11) { b = kDelete; h11 = f11; }
                hd = Hp[c];
                Hp[c] = h11; F[c] = f11; e = e11; hl = h11;
                if (c & 1) row[c >> 1] |= (uint8_t)((b | ext) << 4);      // odd column: high nibble
                else row[c >> 1] = (uint8_t)(b | ext);
            }
            // this lane's C / 2 bytes of the row; lanes wholly beyond the haplotype have nothing to keep
            if (j0 < ncol) {
                uint8_t* dst = bt + (size_t)(i - 1) * PITCH + (j0 >> 1);
                if (C == 32) *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(row);
                else if (C == 16) *reinterpret_cast<uint2*>(dst) = *reinterpret_cast<const uint2*>(row);
                else if (C == 8) *reinterpret_cast<uint32_t*>(dst) = *reinterpret_cast<const uint32_t*>(row);
                else *reinterpret_cast<uint16_t*>(dst) = *reinterpret_cast<const uint16_t*>(row);
            }
            dgL = inH;
            outH = hl; outE = e;
            if (lane == own) s_lastcol[i] = Hp[ncol - 1 - j0];
        }
    }
#pragma unroll
    for (int c = 0; c < C; c++) if (j0 + c < ncol) s_lastrow[j0 + c + 1] = Hp[c];
    __threadfence_block();
    __syncwarp();
    if (lane != 0) return;

    // best end cell: on every anti-diagonal first the last-row cell, then the last-column cell (:329-357)
    int32_t maxScore = INT32_MIN; int max_i = 0, max_j = 0;
    for (int ad = 1; ad <= nrow + ncol; ad++) {
        if (ad >= nrow + 1) {
            const int j = ad - nrow; const int32_t score = s_lastrow[j];
            if (maxScore < score || (maxScore == score && abs(nrow - j) < abs(max_i - max_j))) { maxScore = score; max_i = nrow; max_j = j; }
        }
        if (ad >= ncol + 1) {
            const int i = ad - ncol; const int32_t score = s_lastcol[i];
            if (maxScore < score || (maxScore == score && (max_j == ncol || abs(i - ncol) <= abs(max_i - max_j)))) { maxScore = score; max_i = i; max_j = ncol; }
        }
    }
    // getCIGAR (:367-520), SOFTCLIP; elements are collected back to front in shared memory and merged on the
    // way, then copied in CIGAR order into a slice of the compact output reserved with one atomic
    int n = 0;
    auto push = [&](int op, int len) {
        if (n > 0 && s_ops[n - 1] == (uint8_t)op) { s_lens[n - 1] += len; return; }
        s_ops[n] = (uint8_t)op; s_lens[n] = len; n++;
    };
    int i = max_i, j = max_j;
    if (j < ncol) push(kSoftclip, ncol - j);
    int state = 0;
    while (i > 0 && j > 0) {
        const int btr = (bt[(size_t)(i - 1) * PITCH + ((j - 1) >> 1)] >> (((j - 1) & 1) * 4)) & 15;
        if (state == kInsertExt) { j--; push(kInsert, 1); state = btr & kInsertExt; }
        else if (state == kDeleteExt) { i--; push(kDelete, 1); state = btr & kDeleteExt; }
        else switch (btr & 3) {
            case kMatch:  i--; j--; push(kMatch, 1); state = 0; break;
            case kInsert: j--; push(kInsert, 1); state = btr & kInsertExt; break;
            default:      i--; push(kDelete, 1); state = btr & kDeleteExt; break;
        }
    }
    if (j > 0) push(kSoftclip, j);
    a.offset[job.index] = i;
    a.n_elems[job.index] = n;
    const long long start = (long long)atomicAdd(a.cursor, (unsigned long long)n);
    a.elem_start[job.index] = start;
    if (start + n > a.cap_elems) return;                 // the host sees the cursor beyond the capacity
    for (int k = 0; k < n; k++) {
        const int op = s_ops[n - 1 - k];
        a.ops[start + k] = op == kMatch ? 'M' : op == kInsert ? 'I' : op == kDelete ? 'D' : 'S';
        a.lens[start + k] = s_lens[n - 1 - k];
    }
}

struct DevBuf {
    void* p = nullptr; size_t cap = 0;
    cudaError_t reserve(size_t n) {
        if (n <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        const size_t want = n + n / 4 + 4096;
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};
struct HostBuf {
    void* p = nullptr; size_t cap = 0;
    cudaError_t reserve(size_t n) {
        if (n <= cap) return cudaSuccess;
        if (p) cudaFreeHost(p);
        p = nullptr; cap = 0;
        const size_t want = n + n / 4 + 4096;
        cudaError_t e = cudaMallocHost(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
};

// Everything the aligner owns on ONE device: a stream, two events, pinned staging and device buffers, all
// grow-only and reused from call to call.  Keyed by the CUDA ordinal (a call for another device gets that device's
// context, never this one's pointers), shared by all threads, calls on one device serialised by its mutex;
// phmm_sw_release() frees the lot.
struct SwCtx {
    std::mutex mu;
    cudaStream_t stream = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    HostBuf h_in, h_out;
    DevBuf d_in, d_out, d_bt;
    void release() {
        h_in.release(); h_out.release(); d_in.release(); d_out.release(); d_bt.release();
        if (e0) cudaEventDestroy(e0);
        if (e1) cudaEventDestroy(e1);
        if (stream) cudaStreamDestroy(stream);
        e0 = e1 = nullptr; stream = nullptr;
    }
};
std::mutex g_sw_mu;
std::map<int, std::unique_ptr<SwCtx>> g_sw_ctx;

constexpr size_t kAlign = 256;
inline size_t align_up(size_t x) { return (x + kAlign - 1) / kAlign * kAlign; }
// back-track scratch of one launch; larger batches are walked in pieces (PHMM_SW_SCRATCH_MB overrides)
size_t scratch_budget()
{
    static const size_t b = [] { const char* v = getenv("PHMM_SW_SCRATCH_MB"); return (size_t)(v ? std::max(1, atoi(v)) : 1024) << 20; }();
    return b;
}

#define SW_TRY(expr) do { cudaError_t e__ = (expr); if (e__ != cudaSuccess) { \
    std::fprintf(stderr, "phmm_sw_align: %s: %s\n", #expr, cudaGetErrorString(e__)); return PHMM_ERR_CUDA; } } while (0)

// intel_smithwaterman.hpp:47-58
bool all_match(const uint8_t* ref, int nref, const uint8_t* alt, int nalt)
{
    if (nref != nalt) return false;
    int mismatch = 0;
    for (int i = 0; mismatch <= 2 && i < nref; i++) if (alt[i] != ref[i]) mismatch++;
    return mismatch <= 2;
}

struct DeviceGuard {
    int saved = -1;
    DeviceGuard() { if (cudaGetDevice(&saved) != cudaSuccess) saved = -1; }
    ~DeviceGuard() { if (saved >= 0) cudaSetDevice(saved); }
};

}  // namespace

extern "C" void phmm_sw_release(void)
{
    std::lock_guard<std::mutex> lk(g_sw_mu);
    DeviceGuard guard;
    for (auto& kv : g_sw_ctx) {
        std::lock_guard<std::mutex> lk2(kv.second->mu);
        if (cudaSetDevice(kv.first) == cudaSuccess) kv.second->release();
    }
    g_sw_ctx.clear();
}

extern "C" int phmm_sw_align(int32_t device, const phmm_sw_batch* b, phmm_sw_result* r)
{
    if (!b || !r || b->n < 0) return PHMM_ERR_INVALID_ARG;
    if (b->n == 0) return PHMM_OK;
    if (!b->ref_off || !b->alt_off || !b->ref_bases || !b->alt_bases || !r->offset || !r->elem_beg || !r->ops || !r->lens ||
        r->cap_elems < b->n)
        return PHMM_ERR_INVALID_ARG;
    int visible = 0;
    if (cudaGetDeviceCount(&visible) != cudaSuccess || visible == 0 || device < 0 || device >= visible) return PHMM_ERR_NO_DEVICE;
    DeviceGuard guard;                                   // the caller's current device is left as it was
    SW_TRY(cudaSetDevice(device));
    SwCtx* ctx;
    {
        std::lock_guard<std::mutex> lk(g_sw_mu);
        auto& slot = g_sw_ctx[device];
        if (!slot) slot.reset(new SwCtx());
        ctx = slot.get();
    }
    std::lock_guard<std::mutex> lk(ctx->mu);
    // (each object on its own: a call that failed half-way through this block is completed by the next one)
    if (!ctx->stream) SW_TRY(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    if (!ctx->e0) SW_TRY(cudaEventCreate(&ctx->e0));
    if (!ctx->e1) SW_TRY(cudaEventCreate(&ctx->e1));
    cudaStream_t st = ctx->stream;

    // the aligner's all-match shortcut on the host; everything else grouped by columns-per-lane class
    const int n = b->n;
    std::vector<SwJob> jobs[4];                          // C = 4, 8, 16, 32
    std::vector<int32_t> n_elems(n, 0);
    std::vector<int32_t> match_len(n, 0);                // > 0: shortcut, the CIGAR is "<len>M"
    for (int k = 0; k < n; k++) {
        const int nref = b->ref_off[k + 1] - b->ref_off[k], nalt = b->alt_off[k + 1] - b->alt_off[k];
        if (nref < 1 || nalt < 1) return PHMM_ERR_INVALID_ARG;                     // the reference throws (:33-34)
        if (nref > PHMM_SW_MAX_LEN || nalt > PHMM_SW_MAX_LEN) return PHMM_ERR_UNSUPPORTED;   // its arrays end there
        const uint8_t* ref = b->ref_bases + b->ref_off[k];
        const uint8_t* alt = b->alt_bases + b->alt_off[k];
        if (all_match(ref, nref, alt, nalt)) { r->offset[k] = 0; n_elems[k] = 1; match_len[k] = nref; continue; }
        const int cls = nalt <= 128 ? 0 : nalt <= 256 ? 1 : nalt <= 512 ? 2 : 3;
        const int pitch = (((nalt + 1) / 2) + 15) / 16 * 16;
        jobs[cls].push_back(SwJob{b->ref_off[k], nref, b->alt_off[k], nalt, 0, k, pitch});
    }
    std::vector<SwJob> all;
    for (auto& v : jobs) all.insert(all.end(), v.begin(), v.end());
    const size_t n_jobs = all.size();
    r->kernel_ms = 0.f;
    const uint8_t* h_ops = nullptr;
    const int32_t* h_lens = nullptr;
    const int64_t* h_start = nullptr;
    if (n_jobs) {
        // launches: consecutive jobs of one class whose back-track matrices fit the scratch budget
        struct Piece { size_t first, count; int cls; size_t bt_bytes; };
        std::vector<Piece> pieces;
        size_t max_bt = 0;
        {
            size_t at = 0;
            for (int cls = 0; cls < 4; cls++) {
                const size_t end = at + jobs[cls].size();
                while (at < end) {
                    Piece pc{at, 0, cls, 0};
                    while (at < end) {
                        const size_t need = (size_t)all[at].nrow * all[at].pitch;
                        if (pc.count && pc.bt_bytes + need > scratch_budget()) break;
                        all[at].bt_off = (int64_t)pc.bt_bytes;
                        pc.bt_bytes += need; pc.count++; at++;
                    }
                    max_bt = std::max(max_bt, pc.bt_bytes);
                    pieces.push_back(pc);
                }
            }
        }
        // ONE pinned input block [ref bases | haplotype bases | jobs] and one upload
        const size_t ref_bytes = (size_t)b->ref_off[n], alt_bytes = (size_t)b->alt_off[n];
        const int64_t cap = r->cap_elems;
        const size_t i_ref = 0, i_alt = align_up(ref_bytes), i_jobs = align_up(i_alt + alt_bytes), in_bytes = i_jobs + sizeof(SwJob) * n_jobs;
        // outputs: [cursor | offset | n_elems | elem_start] come back first, then the used part of [ops | lens]
        const size_t o_cursor = 0, o_off = 256, o_ne = align_up(o_off + sizeof(int32_t) * n), o_start = align_up(o_ne + sizeof(int32_t) * n),
                     o_ops = align_up(o_start + sizeof(int64_t) * n), o_lens = align_up(o_ops + (size_t)cap), out_bytes = o_lens + sizeof(int32_t) * (size_t)cap;
        SW_TRY(ctx->h_in.reserve(in_bytes)); SW_TRY(ctx->d_in.reserve(in_bytes));
        SW_TRY(ctx->h_out.reserve(out_bytes)); SW_TRY(ctx->d_out.reserve(out_bytes));
        SW_TRY(ctx->d_bt.reserve(max_bt));
        uint8_t* hi = (uint8_t*)ctx->h_in.p;
        std::memcpy(hi + i_ref, b->ref_bases, ref_bytes);
        std::memcpy(hi + i_alt, b->alt_bases, alt_bytes);
        std::memcpy(hi + i_jobs, all.data(), sizeof(SwJob) * n_jobs);
        SW_TRY(cudaMemcpyAsync(ctx->d_in.p, hi, in_bytes, cudaMemcpyHostToDevice, st));
        uint8_t* dout = (uint8_t*)ctx->d_out.p;
        SW_TRY(cudaMemsetAsync(dout + o_cursor, 0, 8, st));

        SwArgs a{};
        a.ref_bases = (const uint8_t*)ctx->d_in.p + i_ref; a.alt_bases = (const uint8_t*)ctx->d_in.p + i_alt;
        a.w_match = b->w_match; a.w_mismatch = b->w_mismatch; a.w_open = b->w_open; a.w_extend = b->w_extend;
        a.bt = (uint8_t*)ctx->d_bt.p; a.cap_elems = cap; a.cursor = (unsigned long long*)(dout + o_cursor);
        a.offset = (int32_t*)(dout + o_off); a.n_elems = (int32_t*)(dout + o_ne); a.elem_start = (int64_t*)(dout + o_start);
        a.ops = dout + o_ops; a.lens = (int32_t*)(dout + o_lens);
        SW_TRY(cudaEventRecord(ctx->e0, st));
        for (const Piece& pc : pieces) {                 // pieces reuse the scratch: same stream, in order
            a.jobs = (const SwJob*)((const uint8_t*)ctx->d_in.p + i_jobs) + pc.first; a.n_jobs = (int)pc.count;
            switch (pc.cls) {
                case 0: sw_kernel<4><<<(unsigned)pc.count, 32, 0, st>>>(a); break;
                case 1: sw_kernel<8><<<(unsigned)pc.count, 32, 0, st>>>(a); break;
                case 2: sw_kernel<16><<<(unsigned)pc.count, 32, 0, st>>>(a); break;
                default: sw_kernel<32><<<(unsigned)pc.count, 32, 0, st>>>(a); break;
            }
            SW_TRY(cudaGetLastError());
        }
        SW_TRY(cudaEventRecord(ctx->e1, st));
        uint8_t* ho = (uint8_t*)ctx->h_out.p;
        SW_TRY(cudaMemcpyAsync(ho, dout, o_ops, cudaMemcpyDeviceToHost, st));          // cursor + per-alignment scalars
        SW_TRY(cudaStreamSynchronize(st));
        SW_TRY(cudaEventElapsedTime(&r->kernel_ms, ctx->e0, ctx->e1));
        const unsigned long long used = *(const unsigned long long*)(ho + o_cursor);
        if ((int64_t)used > cap) return PHMM_ERR_UNSUPPORTED;            // more CIGAR elements than cap_elems
        if (used) {
            SW_TRY(cudaMemcpyAsync(ho + o_ops, dout + o_ops, (size_t)used, cudaMemcpyDeviceToHost, st));
            SW_TRY(cudaMemcpyAsync(ho + o_lens, dout + o_lens, sizeof(int32_t) * (size_t)used, cudaMemcpyDeviceToHost, st));
            SW_TRY(cudaStreamSynchronize(st));
        }
        const int32_t* h_off = (const int32_t*)(ho + o_off);
        const int32_t* h_ne = (const int32_t*)(ho + o_ne);
        h_start = (const int64_t*)(ho + o_start);
        h_ops = ho + o_ops; h_lens = (const int32_t*)(ho + o_lens);
        for (const SwJob& j : all) { r->offset[j.index] = h_off[j.index]; n_elems[j.index] = h_ne[j.index]; }
    }
    // the caller's compact arrays, in batch order
    int64_t total = 0;
    for (int k = 0; k < n; k++) { r->elem_beg[k] = total; total += n_elems[k]; }
    r->elem_beg[n] = total;
    if (total > r->cap_elems) return PHMM_ERR_UNSUPPORTED;
    for (int k = 0; k < n; k++) {
        const int64_t dst = r->elem_beg[k];
        if (match_len[k]) { r->ops[dst] = 'M'; r->lens[dst] = match_len[k]; continue; }
        std::memcpy(r->ops + dst, h_ops + h_start[k], (size_t)n_elems[k]);
        std::memcpy(r->lens + dst, h_lens + h_start[k], sizeof(int32_t) * (size_t)n_elems[k]);
    }
    return PHMM_OK;
}
