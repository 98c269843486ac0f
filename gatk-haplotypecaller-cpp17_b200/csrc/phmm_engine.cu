// phmm_engine.cu -- host side of libphmm_b200.so: the C ABI of include/phmm.h.
//
// Replaces, for the PairHMM path only, what hc::IntelPairHMM does around its AVX kernels
// (pairhmm/intel_pairhmm.hpp): initNative (:77-113), getData (:154-203), the FP32 -> FP64
// dispatch loop (:115-152) and the log10 conversion (:139,:142).  Differences by design:
//   * many regions per batch, all pairs of a batch in flight at once (the reference walks
//     reads x haplotypes serially, one region at a time, haplotypecaller.hpp:138-152);
//   * per-device memory pool (grow-only pinned + device arenas per pipeline slot) and one stream
//     per slot, so batch N+1 packs/uploads while batch N computes and batch N-1 downloads;
//   * regions are sharded over devices by cell count with no exchange between devices (every
//     pair is independent, intel_pairhmm.hpp:131-147); results are gathered on the host;
//   * raw forward sums come back from the device and log10f/log10 run on the HOST with glibc,
//     because the reference's final value is (double)(log10f(f) - log10f(2^120)) evaluated in
//     float (:142) and device log10f differs from glibc in the last ulp.
// No CPU fallback: every entry point fails with a CUDA error when no device is usable.
#include "../../include/phmm.h"
#include "phmm_launch.h"
#include "phmm_tables.h"

#include <immintrin.h>

#include <algorithm>
#include <atomic>
#include <cfloat>
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

using namespace phmm;

namespace {

constexpr size_t kAlign = 256;
constexpr int kSmemBytesPerWarpBudget = 8 * 1024;    // haplotype stream share of the per-warp shared memory
constexpr int kPerHapTableBytes = 3 * 8 + 4 + 4 + 4; // init_y, two scales (MODE 3), haplotype index, stream position, length
inline size_t align_up(size_t x) { return (x + kAlign - 1) / kAlign * kAlign; }

struct KernelTable {
    // [f64][exact] -> [mode][aligned][shape][list]
    KernelTab fn[2][2];
    KernelTable() {
        register_f32_fast(fn[0][0]);
        register_f32_exact(fn[0][1]);
        register_f64_fast(fn[1][0]);
        register_f64_exact(fn[1][1]);
    }
};
const KernelTable& kernel_table() { static KernelTable t; return t; }

// per-warp dynamic shared memory of one shape: haplotype stream + prior tables + per-haplotype scalars
inline int smem_bytes_per_warp(int stream_cap, int haps_per_job, Shape sh)
{
    return (stream_cap + tables_bytes(sh.K, sh.G) + haps_per_job * kPerHapTableBytes + 127) / 128 * 128;
}

// Shape for a read of R bases against haplotypes of about H bases.  Feasible: K*G >= R + 1 (one
// dummy row on top).  Cost model (warp cycles per scored pair): a step costs ~15 FP32-pipe cycles
// per row plus ~10 of per-step work; a haplotype takes H + G - 1 steps; a warp holds 32/G groups.
// PHMM_FORCE_GROUP=16|32 restricts the choice (benchmarking aid).
inline int pick_shape_uncached(int R, int H);
inline int pick_shape(int R, int H)
{
    // reads of one region share H and mostly R: remember the answers for the current H
    thread_local int cached_h = -1;
    thread_local int16_t cache[kMaxReadLenCompiled + 2];
    if (R < 1 || R > kMaxReadLenCompiled) return pick_shape_uncached(R, H);
    if (H != cached_h) { for (auto& c : cache) c = -2; cached_h = H; }
    if (cache[R] == -2) cache[R] = (int16_t)pick_shape_uncached(R, H);
    return cache[R];
}
inline int pick_shape_uncached(int R, int H)
{
    static const int force = [] { const char* s = getenv("PHMM_FORCE_GROUP"); return s ? atoi(s) : 0; }();
    int best = -1; double best_cost = 0;
    for (int s = 0; s < kFirstPackedShape; s++) {
        const int G = kShapes[s].G, K = kShapes[s].K;
        if (K * G < R + 1) continue;
        if (force && G != force) continue;
        const double cost = (double)G * (15.0 * K + 10.0) * (H + G - 1);
        if (best < 0 || cost < best_cost) { best = s; best_cost = cost; }
    }
    if (best < 0 && force)      // forced width cannot hold this read: fall back to any feasible shape
        for (int s = 0; s < kFirstPackedShape; s++)
            if (kShapes[s].K * kShapes[s].G >= R + 1) { best = s; break; }
    return best;
}

// PACKED shape (free-width lane groups, phmm_kernels.cuh) for reads of exactly R bases, or -1: R = K * nl for
// a compiled K with groups of nl lanes that fill at least 90% of the warp (8, 10, 15 or 16 lanes), except
// where a power-of-two group with one idle lane does as well (150 = 15 x 10 on 16 lanes: ALIGNED).
inline int packed_shape(int R, int* lanes_per_group)
{
    static const bool off = getenv("PHMM_NO_PACKED") != nullptr;
    if (off) return -1;
    for (int s = kNumShapes - 1; s >= kFirstPackedShape; s--) {          // tallest lanes first
        const int K = kShapes[s].K;
        if (R % K) continue;
        const int nl = R / K;
        if (nl != 8 && nl != 10 && nl != 16 && !(nl == 15 && K != 10)) continue;
        if (lanes_per_group) *lanes_per_group = nl;
        return s;
    }
    return -1;
}

// ---- grow-only buffers (the memory pool) ----------------------------------------------------
struct PinnedBuf {
    void* p = nullptr; size_t cap = 0;
    cudaError_t reserve(size_t n) {
        if (n <= cap) return cudaSuccess;
        if (p) cudaFreeHost(p);
        p = nullptr; cap = 0;
        size_t want = std::max(n, (size_t)1 << 20);
        want = want + want / 4;
        cudaError_t e = cudaMallocHost(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
};
// Debug aid standing in for compute-sanitizer (closed on this pool): with PHMM_DEBUG_GUARD=1 every device buffer is
// allocated between two 4 KB guard zones filled with 0xA5 and its body is filled with the byte PHMM_POISON
// (default 0xCD) -- also again on every reserve() that reuses the allocation.  phmm_debug_check() verifies the
// guards (a write past either end of any buffer shows), and running a workload under two different poison bytes
// must give bit-identical results (a read of memory no kernel wrote shows): tests/test_debug_guards.py.
constexpr size_t kGuardBytes = 4096;
inline bool debug_guard_on() { static const bool on = getenv("PHMM_DEBUG_GUARD") != nullptr; return on; }
inline int debug_poison() { static const int v = [] { const char* s = getenv("PHMM_POISON"); return s ? (int)strtol(s, nullptr, 0) & 255 : 0xCD; }(); return v; }

struct DeviceBuf {
    void* p = nullptr; size_t cap = 0;
    void* base = nullptr;                // the allocation itself (== p unless guards are on)
    cudaError_t reserve(size_t n) {
        if (n <= cap) {
            if (debug_guard_on() && p) {                     // stale contents must not matter either
                cudaMemset(p, debug_poison(), cap);          // (null stream: the engine's streams do not wait for it,
                cudaDeviceSynchronize();                     //  so wait here -- debug mode only)
            }
            return cudaSuccess;
        }
        release();
        size_t want = std::max(n, (size_t)1 << 20);
        want = (want + want / 4 + 255) / 256 * 256;
        const size_t g = debug_guard_on() ? kGuardBytes : 0;
        cudaError_t e = cudaMalloc(&base, want + 2 * g);
        if (e != cudaSuccess) { base = nullptr; return e; }
        p = (uint8_t*)base + g; cap = want;
        if (g) {
            cudaMemset(base, 0xA5, g);
            cudaMemset((uint8_t*)p + cap, 0xA5, g);
            cudaMemset(p, debug_poison(), cap);
            cudaDeviceSynchronize();
        }
        return cudaSuccess;
    }
    // number of guard bytes that no longer hold 0xA5 (0 when guards are off)
    int64_t check_guards() const {
        if (!debug_guard_on() || !base) return 0;
        std::vector<uint8_t> h(2 * kGuardBytes);
        if (cudaMemcpy(h.data(), base, kGuardBytes, cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
        if (cudaMemcpy(h.data() + kGuardBytes, (const uint8_t*)p + cap, kGuardBytes, cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
        int64_t bad = 0;
        for (uint8_t b : h) bad += (b != 0xA5);
        return bad;
    }
    void release() { if (base) cudaFree(base); base = nullptr; p = nullptr; cap = 0; }
};

// Copy into pinned staging with NON-TEMPORAL stores: the destination is only ever read by the DMA engine, so pulling
// its lines into the cache first (write-allocate) is a third of the memory traffic of a plain memcpy for nothing --
// and host memory bandwidth is what limits a read-heavy stream on 8 GPUs.
__attribute__((target("avx2"))) static void stream_copy_avx2(uint8_t* dst, const uint8_t* src, size_t n)
{
    const size_t head = (64 - (reinterpret_cast<uintptr_t>(dst) & 63)) & 63;
    if (head) { std::memcpy(dst, src, head); dst += head; src += head; n -= head; }
    const size_t blocks = n / 64;
    for (size_t i = 0; i < blocks; i++) {
        const __m256i a = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(src + 64 * i));
        const __m256i b = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(src + 64 * i + 32));
        _mm256_stream_si256(reinterpret_cast<__m256i*>(dst + 64 * i), a);
        _mm256_stream_si256(reinterpret_cast<__m256i*>(dst + 64 * i + 32), b);
    }
    _mm_sfence();
    if (n % 64) std::memcpy(dst + 64 * blocks, src + 64 * blocks, n % 64);
}
static void staging_copy(uint8_t* dst, const uint8_t* src, size_t n)
{
    static const bool avx2 = __builtin_cpu_supports("avx2") && getenv("PHMM_NO_STREAM_COPY") == nullptr;
    if (avx2 && n >= (size_t)256 << 10) stream_copy_avx2(dst, src, n);
    else std::memcpy(dst, src, n);
}

// ---- host-side parallel_for: planning, packing and the log10 pass of a batch are spread over the
//      engine's host_threads (per device: host_threads - 1 workers + the device's own worker thread) ----
class HostPool {
public:
    explicit HostPool(int n_workers) {
        for (int i = 0; i < n_workers; i++) th_.emplace_back([this] { loop(); });
    }
    ~HostPool() {
        { std::lock_guard<std::mutex> lk(mu_); stop_ = true; }
        cv_.notify_all();
        for (auto& t : th_) t.join();
    }
    int width() const { return (int)th_.size() + 1; }
    // fn(i) for i in [0, n), on the workers and the calling thread; returns when all are done
    void parallel_for(int n, const std::function<void(int)>& fn) {
        if (n <= 0) return;
        if (n == 1 || th_.empty()) { for (int i = 0; i < n; i++) fn(i); return; }
        auto task = std::make_shared<Task>();
        task->fn = fn; task->total = n;
        { std::lock_guard<std::mutex> lk(mu_); cur_ = task; ++generation_; }
        cv_.notify_all();
        run(*task);
        std::unique_lock<std::mutex> lk(task->mu);
        task->cv.wait(lk, [&] { return task->done.load() == task->total; });
    }
private:
    struct Task {
        std::function<void(int)> fn;
        int total = 0;
        std::atomic<int> next{0}, done{0};
        std::mutex mu; std::condition_variable cv;
    };
    static void run(Task& t) {
        for (;;) {
            const int i = t.next.fetch_add(1);
            if (i >= t.total) return;
            t.fn(i);
            if (t.done.fetch_add(1) + 1 == t.total) { std::lock_guard<std::mutex> lk(t.mu); t.cv.notify_all(); }
        }
    }
    void loop() {
        uint64_t seen = 0;
        for (;;) {
            std::shared_ptr<Task> task;
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_.wait(lk, [&] { return stop_ || generation_ != seen; });
                if (stop_) return;
                seen = generation_;
                task = cur_;
            }
            if (task) run(*task);
        }
    }
    std::vector<std::thread> th_;
    std::mutex mu_; std::condition_variable cv_;
    std::shared_ptr<Task> cur_;
    uint64_t generation_ = 0;
    bool stop_ = false;
};

// ---- one device's share of a batch: a contiguous range of regions -----------------------------
struct Part {
    int g0 = 0, g1 = 0;                       // region range in the caller's batch
    int n_regions = 0, n_reads = 0, n_haps = 0;
    int64_t n_pairs = 0, n_cells = 0, out0 = 0;   // out0: offset of this part in the batch output
    int read0 = 0;                            // first read of the part in the caller's batch
    int mode = kModeGeneral;                 // kernel MODE: general / batch-constant gaps / constant with i == d
    uint8_t gap[3] = {0, 0, 0};              // the batch-constant (i, d, c) bytes when mode != general
    int max_H = 0, max_nh = 0;
    int n_long = 0;                          // (long read, haplotype) pairs of the one-warp-per-pair kernel
    int n_jobs = 0;
    int job_beg[2 * kNumShapes + 1] = {0};  // jobs are grouped by kernel slot = shape + kNumShapes * aligned
    int haps_per_job = 1, hap_chunks = 1;
    int haps_per_job64 = 1, hap_chunks64 = 1;    // chunking of the FP64 redo launches (see stage_and_launch)
    bool f64_first = false;                  // FP64-first order (launch_kernels): the previous batch was rescue-dense
    size_t h2d_bytes = 0, d2h_bytes = 0;
    int launches = 0;
    float kernel_ms = 0.f;
    unsigned rescue_count = 0;
};

constexpr int kAuxStreams = 4;      // kernels of different shapes of one batch run side by side

struct Slot {
    cudaStream_t stream = nullptr;
    cudaStream_t aux[kAuxStreams] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t ev_fork = nullptr, ev_join[kAuxStreams] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t ev_k0 = nullptr, ev_k1 = nullptr, ev_done = nullptr;
    cudaEvent_t ev_k32 = nullptr;            // after the FP32 launch of a single-shape batch (dominant-kernel timing)
    bool k32_valid = false;
    PinnedBuf h_in, h_jobs, h_out, h_rescue;
    DeviceBuf d_in, d_jobs, d_out, d_rescue, d_flags, d_work;
    int sm_count = 148;
    bool device_log10 = false;               // d_out = [16 B header | lik32[n_pad] | raw32[n_pad]]; the D2H takes header + lik32
                                             // (device_log10) or everything (host log10f pass)
    float log10_init_f = 0.f;
    int async_rc = PHMM_OK;                  // failure after the submitter was released: reported by phmm_wait
    std::string async_err;
    struct StageCtx {                        // what the pack phase hands to the plan + launch phase
        size_t o_read_off = 0, o_hap_off = 0, o_reg_read = 0, o_reg_hap = 0, o_reg_out = 0, o_bases = 0, o_q = 0,
               o_gi = 0, o_gd = 0, o_gc = 0, o_haps = 0, in_bytes = 0;
        phmm_batch view{};                   // the part as a batch of its own, over the packed copy
        bool general = false, empty = true;
        bool zero_copy = false;              // PHMM_BATCH_PINNED_INPUTS: byte arrays upload from the caller's memory
        int g0 = 0, g1 = 0, read0 = 0;
        int64_t out0 = 0;
        std::chrono::steady_clock::time_point t_begin, t_packed;
    } stage;
    // device-side genotype reduction (phmm_submit_gl): the part's sites and its outputs
    struct GlCtx {
        bool on = false;
        int s0 = 0, s1 = 0;                      // site range of the part in the caller's arrays
        int64_t gl0 = 0, n_gl = 0;               // genotype-likelihood range
        int64_t n_site_reads = 0, n_site_haps = 0;
        // what the submitter hands to the pack phase (valid until the pack phase is over)
        const phmm_sites* sites = nullptr;
        const int64_t* site_hap_off = nullptr; const int64_t* site_read_off = nullptr; const int64_t* gl_off = nullptr;
        size_t o_region = 0, o_alleles = 0, o_hap_off = 0, o_read_off = 0, o_gl_off = 0, o_hap_allele = 0, o_overlap = 0, bytes = 0;
        bool has_overlap = false;
        GenotypeArgs args{};
        size_t out_gl = 0, out_nused = 0, out_keep = 0, out_bytes = 0;      // layout of the output block
    } gl;
    PinnedBuf h_sites, h_gl;
    DeviceBuf d_sites, d_gl_out, d_lik64, d_gl_scratch;
    bool busy = false;
    Part part;
    KernelArgs args{};
    const LongPair* d_long = nullptr;
};

struct phmm_engine_impl;

struct DeviceCtx {
    int ordinal = 0;
    int sm_count = 148;
    float last_rescue_frac = 0.f;        // share of pairs the previous batch redid in FP64
    int fp64_first_opt = 0;              // phmm_options.fp64_first
    bool device_log10 = false;           // phmm_engine::device_log10
    bool scaled_recurrence = true;       // phmm_options.recurrence == 0
    std::unique_ptr<HostPool> pool;      // host_threads - 1 helpers for this device's worker thread
    std::unique_ptr<HostPool> pack_pool; // helpers of the PACKER thread (its copies must not queue behind the worker's
                                         // planning and finalizing in the same pool: a HostPool runs one parallel_for at a time)
    std::vector<Slot> slots;
    int next_slot = 0;
    float* d_ph2pr_f = nullptr; float* d_mm_f = nullptr;
    double* d_ph2pr_d = nullptr; double* d_mm_d = nullptr;
    uint8_t* d_table_blob = nullptr;     // the one allocation the four probability tables live in
    double* d_jacobian = nullptr;        // hc::MathUtils' Jacobian table, uploaded by the first phmm_submit_gl part
    std::once_flag jacobian_once;
    // Two threads per device.  The WORKER plans, launches and finalizes (and serves the staged form); the
    // PACKER only copies a submitter's arrays into pinned staging and starts their upload, so that a
    // submitter is never held up behind the planning or the log10 pass of an earlier batch.
    struct WorkQueue {
        std::thread th;
        std::mutex mu;
        std::condition_variable cv;
        std::deque<std::function<void()>> queue;
        bool stop = false;
        void run(int ordinal) {
            cudaSetDevice(ordinal);
            for (;;) {
                std::function<void()> fn;
                {
                    std::unique_lock<std::mutex> lk(mu);
                    cv.wait(lk, [&] { return stop || !queue.empty(); });
                    if (queue.empty()) return;
                    fn = std::move(queue.front());
                    queue.pop_front();
                }
                fn();
            }
        }
        void post(std::function<void()> fn) {
            { std::lock_guard<std::mutex> lk(mu); queue.push_back(std::move(fn)); }
            cv.notify_one();
        }
        void start(int ordinal) { th = std::thread([this, ordinal] { run(ordinal); }); }
        void finish() {
            { std::lock_guard<std::mutex> lk(mu); stop = true; }
            cv.notify_all();
            if (th.joinable()) th.join();
        }
    };
    WorkQueue worker, packer;
    void post(std::function<void()> fn) { worker.post(std::move(fn)); }
};

struct Latch {
    std::mutex mu; std::condition_variable cv; int remaining;
    explicit Latch(int n) : remaining(n) {}
    void done() { std::lock_guard<std::mutex> lk(mu); if (--remaining == 0) cv.notify_all(); }
    void wait() { std::unique_lock<std::mutex> lk(mu); cv.wait(lk, [&] { return remaining == 0; }); }
};

struct TicketRec {
    std::vector<std::pair<int, int>> parts;   // (device index, slot index)
    int64_t n_pairs = 0;
    bool gl = false;                          // a phmm_submit_gl ticket: waited with phmm_wait_gl
    std::chrono::steady_clock::time_point t0;
};

}  // namespace

struct phmm_engine {
    phmm_options opt{};
    std::vector<std::unique_ptr<DeviceCtx>> devs;
    std::mutex mu;
    std::string last_error;
    std::map<phmm_ticket, TicketRec> tickets;
    phmm_ticket next_ticket = 1;
    int host_threads = 1;
    bool device_log10 = false;           // final float log10 on the device (phmm_finalize.cu); false: host log10f pass

    void set_error(const std::string& s) { std::lock_guard<std::mutex> lk(mu); last_error = s; }
};

struct phmm_staged {
    int dev_index = 0;
    Slot slot;           // private buffers, not part of the pipeline ring
    bool ran = false;
};

namespace {

// Public entry points leave the calling thread's current CUDA device as they found it (a caller that also
// drives torch / its own CUDA code must not be redirected by a library call).
struct DeviceGuard {
    int saved = -1;
    DeviceGuard() { if (cudaGetDevice(&saved) != cudaSuccess) saved = -1; }
    ~DeviceGuard() { if (saved >= 0) cudaSetDevice(saved); }
};

#define CUDA_TRY(expr)                                                                            \
    do {                                                                                          \
        cudaError_t e__ = (expr);                                                                 \
        if (e__ != cudaSuccess) {                                                                 \
            err = std::string(#expr) + ": " + cudaGetErrorString(e__);                            \
            return PHMM_ERR_CUDA;                                                                 \
        }                                                                                         \
    } while (0)

int validate_batch(const phmm_batch* b, std::string& err)
{
    if (!b) { err = "batch is NULL"; return PHMM_ERR_INVALID_ARG; }
    if (b->n_regions < 0 || b->n_reads < 0 || b->n_haps < 0) { err = "negative count"; return PHMM_ERR_INVALID_ARG; }
    if (b->n_regions == 0) return PHMM_OK;
    if (!b->region_read_beg || !b->region_hap_beg || !b->read_off || !b->hap_off) {
        err = "NULL offset array"; return PHMM_ERR_INVALID_ARG;
    }
    if ((b->n_reads && (!b->read_bases || !b->read_q)) || (b->n_haps && !b->hap_bases)) {
        err = "NULL base/quality array"; return PHMM_ERR_INVALID_ARG;
    }
    const bool any_gap = b->read_i || b->read_d || b->read_c;
    if (any_gap && !(b->read_i && b->read_d && b->read_c)) {
        err = "read_i/read_d/read_c must be all set or all NULL"; return PHMM_ERR_INVALID_ARG;
    }
    if (b->region_read_beg[0] != 0 || b->region_hap_beg[0] != 0 ||
        b->region_read_beg[b->n_regions] != b->n_reads || b->region_hap_beg[b->n_regions] != b->n_haps) {
        err = "region ranges must cover [0,n_reads) and [0,n_haps)"; return PHMM_ERR_INVALID_ARG;
    }
    for (int g = 0; g < b->n_regions; g++)
        if (b->region_read_beg[g + 1] < b->region_read_beg[g] || b->region_hap_beg[g + 1] < b->region_hap_beg[g]) {
            err = "region ranges not monotone"; return PHMM_ERR_INVALID_ARG;
        }
    for (int r = 0; r < b->n_reads; r++) {
        int R = b->read_off[r + 1] - b->read_off[r];
        if (R < 1) { err = "empty read"; return PHMM_ERR_INVALID_ARG; }
        if (R > kLongMaxRead) { err = "read longer than " + std::to_string(kLongMaxRead); return PHMM_ERR_UNSUPPORTED; }
    }
    for (int h = 0; h < b->n_haps; h++) {
        int H = b->hap_off[h + 1] - b->hap_off[h];
        if (H < 1) { err = "empty haplotype"; return PHMM_ERR_INVALID_ARG; }
        if (H > PHMM_MAX_HAP_LEN) { err = "haplotype longer than PHMM_MAX_HAP_LEN"; return PHMM_ERR_UNSUPPORTED; }
    }
    return PHMM_OK;
}

int64_t batch_pairs(const phmm_batch* b, int g0, int g1)
{
    int64_t n = 0;
    for (int g = g0; g < g1; g++)
        n += (int64_t)(b->region_read_beg[g + 1] - b->region_read_beg[g]) * (b->region_hap_beg[g + 1] - b->region_hap_beg[g]);
    return n;
}

int64_t region_cells(const phmm_batch* b, int g)
{
    int r0 = b->region_read_beg[g], r1 = b->region_read_beg[g + 1];
    int h0 = b->region_hap_beg[g], h1 = b->region_hap_beg[g + 1];
    return (int64_t)(b->read_off[r1] - b->read_off[r0]) * (b->hap_off[h1] - b->hap_off[h0]);
}

// The units (job, chunk) whose flag byte carries `bit`, as work items for a LIST launch: every such unit is cut
// into `subs` pieces of the consumer's haps_per_job.  One thread per unit of the launch slot.
__global__ void build_work_list(const uint8_t* __restrict__ flags, const int n_units, const unsigned bit, const int subs,
                                uint2* __restrict__ list, unsigned* __restrict__ count)
{
    const int u = blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= n_units || !(flags[u] & bit)) return;
    const unsigned at = atomicAdd(count, (unsigned)subs);
    for (int s = 0; s < subs; ++s) list[at + s] = make_uint2((unsigned)u, (unsigned)s);
}

// Resident warps per SM of a kernel at a given dynamic shared memory size (persistent work-list launches are
// sized to fill the chip once); cached, the occupancy query is not free.
int resident_ctas_per_sm(KernelFn fn, size_t smem)
{
    static std::mutex mu;
    static std::map<std::pair<const void*, size_t>, int> cache;
    std::lock_guard<std::mutex> lk(mu);
    auto key = std::make_pair((const void*)fn, smem);
    auto it = cache.find(key);
    if (it != cache.end()) return it->second;
    int n = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, (const void*)fn, kWarpsPerCta * 32, smem) != cudaSuccess || n < 1) {
        cudaGetLastError();
        n = 8;
    }
    cache[key] = n;
    return n;
}

// Header of the flags buffer: per kernel slot k four counters {count, cursor} of the FP64 work list (filled by
// the FP32 launch) and {count, cursor} of the FP32 work list (filled by an FP64-first launch).
constexpr size_t kWorkHeaderBytes = ((size_t)2 * kNumShapes * 4 * sizeof(unsigned) + 255) / 256 * 256;

// Forward (FP32) and rescue (FP64) kernels of a staged slot.  Every kernel slot (shape x aligned) is an
// independent chain of launches; the chains are spread over a few auxiliary streams forked from / joined to the
// slot's stream so that the small grids of a ragged batch overlap.  ev_k0 / ev_k1 bracket the lot.
// Precision tiers (phmm_kernels.cuh): 1 FP32, 2 its FP64 redo, 3 the flush-exact FP64 redo of pairs that ended
// within reach of the denormal range, 0 the FP64-first pass.  Three orders for the normal pass [1, 2]:
//   FP32 first (default)   FP32 over the grid -> FP64 redo PULLING the (job, haplotype) units the FP32 launch
//                          listed (empty list: the launch is a few idle warps);
//   FP64 first (p.f64_first: the device's previous batch redid most of its pairs)
//                          FP64 over the grid -> FP32 only for the units holding a pair that may not underflow ->
//                          FP64 redo of what that FP32 pass still underflowed; see kCertainUnderflow64;
//   use_double             FP64 over the grid with every flag raised (intel_pairhmm.hpp:135).
// Tier 3 is launched by finalize_part only when the downloaded results show a marked pair (pathological inputs).
int launch_kernels(Slot& s, bool exact, int tier_lo, int tier_hi, std::string& err, bool use_double = false)
{
    Part& p = s.part;
    const KernelArgs& a = s.args;
    const bool normal_pass = tier_lo == 1;
    const bool skip32 = use_double && normal_pass;
    const bool f64_first = p.f64_first && normal_pass && !skip32 && !exact;   // (exact: raw FP32 sums are an output)
    uint8_t* const flags_base = (uint8_t*)s.d_flags.p;
    unsigned* const counters = (unsigned*)flags_base;
    if (normal_pass) {
        p.launches = 0;
        s.k32_valid = false;
        CUDA_TRY(cudaMemsetAsync(s.d_out.p, 0, 16, s.stream));                        // counters
        if (skip32) CUDA_TRY(cudaMemsetAsync(a.raw32, 0, sizeof(float) * (size_t)p.n_pairs, s.stream));   // every raw FP32 sum is 0.0f
        CUDA_TRY(cudaMemsetAsync(flags_base, 0, kWorkHeaderBytes, s.stream));
        CUDA_TRY(cudaMemsetAsync(flags_base + kWorkHeaderBytes, skip32 ? 1 : 0, (size_t)p.n_jobs * p.hap_chunks, s.stream));
        CUDA_TRY(cudaEventRecord(s.ev_k0, s.stream));
    }
    int n_kernels = 0;
    for (int k = 0; k < 2 * kNumShapes; k++) n_kernels += (p.job_beg[k + 1] > p.job_beg[k]);
    const bool fork = n_kernels > 1;
    bool used[kAuxStreams] = {false, false, false, false};
    if (fork) CUDA_TRY(cudaEventRecord(s.ev_fork, s.stream));
    int which = 0;
    const int subs64 = (p.haps_per_job + p.haps_per_job64 - 1) / p.haps_per_job64;    // FP64 items per FP32 unit
    // Launch order: the kernel slot with the most work first.  Chains on different streams run side by side only
    // where the chip has room, so the small grids of a ragged batch end up filling the tail of the big one --
    // launched last (slot order), the big grid drained alone instead (8% of the ragged window stream).
    int order[2 * kNumShapes];
    size_t slot_off[2 * kNumShapes];
    {
        size_t off = 0;                                      // in uint2 items: every slot owns [FP64 list | FP32 list]
        for (int k = 0; k < 2 * kNumShapes; k++) {
            order[k] = k;
            slot_off[k] = off;
            off += (size_t)(p.job_beg[k + 1] - p.job_beg[k]) * p.hap_chunks * (subs64 + 1);
        }
        auto work = [&](int k) { return (int64_t)(p.job_beg[k + 1] - p.job_beg[k]) * kShapes[k % kNumShapes].K * 32; };
        std::stable_sort(order, order + 2 * kNumShapes, [&](int x, int y) { return work(x) > work(y); });
    }
    for (int oi = 0; oi < 2 * kNumShapes; oi++) {
        const int k = order[oi];
        const int n = p.job_beg[k + 1] - p.job_beg[k];
        if (n == 0) continue;
        const size_t work_off = slot_off[k];
        cudaStream_t st = s.stream;
        if (fork) {
            const int ai = which++ % kAuxStreams;
            st = s.aux[ai];
            if (!used[ai]) { CUDA_TRY(cudaStreamWaitEvent(st, s.ev_fork, 0)); used[ai] = true; }
        }
        const int per_warp = smem_bytes_per_warp(a.stream_cap, p.haps_per_job, kShapes[k % kNumShapes]);
        const size_t smem = (size_t)per_warp * kWarpsPerCta;
        const size_t cap64 = (size_t)n * p.hap_chunks * subs64, cap32 = (size_t)n * p.hap_chunks;
        uint2* const list64 = (uint2*)s.d_work.p + work_off;
        uint2* const list32 = list64 + cap64;
        unsigned* const cnt = counters + 4 * k;              // {count64, cursor64, count32, cursor32}
        KernelArgs ak = a;
        ak.jobs = a.jobs + p.job_beg[k];
        ak.n_jobs = n;
        ak.job_flag_base = p.job_beg[k];
        ak.job_flags = flags_base + kWorkHeaderBytes;
        ak.smem_bytes_per_warp = per_warp;
        ak.flag_hpj = p.haps_per_job;
        ak.flag_chunks = p.hap_chunks;
        // one launch of the chain: `tier`, walking the grid or (from_list) pulling the units whose flag carries
        // `bit` from a list that build_work_list() compacts first
        auto launch = [&](int tier, int hpj, int chunks, bool from_list, unsigned bit, uint2* list, unsigned* cnt_cur, size_t cap) -> int {
            KernelArgs al = ak;
            al.tier = tier;
            al.haps_per_job = hpj;
            al.first64 = f64_first ? 1 : 0;
            KernelFn fn = kernel_table().fn[tier != 1][(exact || tier == 3) ? 1 : 0][p.mode][k / kNumShapes][k % kNumShapes][from_list ? 1 : 0];
            if (!fn) { err = "kernel variant not compiled"; return PHMM_ERR_UNSUPPORTED; }
            if (smem > 48 * 1024) CUDA_TRY(cudaFuncSetAttribute((const void*)fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            dim3 grid((n + kWarpsPerCta - 1) / kWarpsPerCta, chunks);
            if (from_list) {
                const int n_units = n * p.hap_chunks, subs = (p.haps_per_job + hpj - 1) / hpj;
                build_work_list<<<(n_units + 255) / 256, 256, 0, st>>>(ak.job_flags + (size_t)ak.job_flag_base * ak.flag_chunks, n_units,
                                                                        bit, subs, list, cnt_cur);
                CUDA_TRY(cudaGetLastError());
                p.launches++;
                al.work_in = list; al.work_in_count = cnt_cur; al.work_in_cursor = cnt_cur + 1;
                const size_t fill = (size_t)s.sm_count * resident_ctas_per_sm(fn, smem);
                grid = dim3((unsigned)std::max<size_t>(1, std::min(fill, cap)), 1);
            }
            fn<<<grid, kWarpsPerCta * 32, smem, st>>>(al);
            CUDA_TRY(cudaGetLastError());
            p.launches++;
            return PHMM_OK;
        };
        const bool lists = !exact;                           // the exact engines walk the grid (no LIST kernels compiled)
        int rc = PHMM_OK;
        if (!normal_pass) {                                  // tier 3 (or an explicit tier range): grid walks
            for (int tier = tier_lo; tier <= tier_hi && !rc; tier++)
                rc = launch(tier, tier == 1 ? p.haps_per_job : p.haps_per_job64, tier == 1 ? p.hap_chunks : p.hap_chunks64, false, 0, nullptr, nullptr, 0);
        } else if (skip32) {
            rc = launch(2, p.haps_per_job64, p.hap_chunks64, false, 0, nullptr, nullptr, 0);
        } else if (f64_first) {
            rc = launch(0, p.haps_per_job, p.hap_chunks, false, 0, nullptr, nullptr, 0);
            if (!rc) rc = launch(1, p.haps_per_job, p.hap_chunks, true, kFlagNeedsF32, list32, cnt + 2, cap32);
            if (!rc && tier_hi >= 2) rc = launch(2, p.haps_per_job64, p.hap_chunks64, true, kFlagRedo64, list64, cnt, cap64);
        } else {
            rc = launch(1, p.haps_per_job, p.hap_chunks, false, 0, nullptr, nullptr, 0);
            if (!rc && !fork) { CUDA_TRY(cudaEventRecord(s.ev_k32, st)); s.k32_valid = true; }
            if (!rc && tier_hi >= 2) rc = launch(2, p.haps_per_job64, p.hap_chunks64, lists, kFlagRedo64, list64, cnt, cap64);
        }
        if (rc) return rc;
    }
    if (fork)
        for (int ai = 0; ai < kAuxStreams; ai++)
            if (used[ai]) {
                CUDA_TRY(cudaEventRecord(s.ev_join[ai], s.aux[ai]));
                CUDA_TRY(cudaStreamWaitEvent(s.stream, s.ev_join[ai], 0));
            }
    if (normal_pass) {
        if (p.n_long) {      // reads beyond one lane-group pass: all precision tiers inside one launch
            launch_long_reads(a, s.d_long, p.n_long, p.mode == kModeGeneral || p.mode == kModeGeneralScaled, exact, use_double, s.stream);
            CUDA_TRY(cudaGetLastError());
            p.launches++;
        }
        if (s.device_log10) {                                // raw sums -> final float log10 values + counters
            launch_finalize(a.raw32, p.n_pairs, s.log10_init_f, (float*)((uint8_t*)s.d_out.p + 16), (unsigned*)s.d_out.p,
                            s.sm_count, s.stream);
            CUDA_TRY(cudaGetLastError());
            p.launches++;
        }
        CUDA_TRY(cudaEventRecord(s.ev_k1, s.stream));
    }
    return PHMM_OK;
}

// Pack regions [g0,g1) of the batch into the slot's pinned block, upload, launch, start D2H.
constexpr int kSlots = 2 * kNumShapes;      // kernel slot = shape + kNumShapes * aligned

// Batch-constant gap penalties?  Always, for the reference's own callers (sam/sam.hpp:30-32); per-base
// arrays are scanned once, and the general kernels run only if they really vary over bytes [rb0, rb1).
int detect_gap_mode(const phmm_batch* b, int rb0, int rb1, uint8_t gap[3])
{
    bool constant = true;
    if (!b->read_i) {
        gap[0] = b->gap_open_i; gap[1] = b->gap_open_d; gap[2] = b->gap_cont_c;
    } else {
        gap[0] = b->read_i[rb0]; gap[1] = b->read_d[rb0]; gap[2] = b->read_c[rb0];
        const uint8_t* end;
        end = b->read_i + rb1; for (const uint8_t* q = b->read_i + rb0; q < end && constant; ++q) constant = (*q == gap[0]);
        end = b->read_d + rb1; for (const uint8_t* q = b->read_d + rb0; q < end && constant; ++q) constant = (*q == gap[1]);
        end = b->read_c + rb1; for (const uint8_t* q = b->read_c + rb0; q < end && constant; ++q) constant = (*q == gap[2]);
    }
    return !constant ? kModeGeneral : (((gap[0] & 127) == (gap[1] & 127)) ? kModeConstShared : kModeConst);
}

// What planning hands to packing: the job lists per kernel slot, the long-read pairs, the output offsets.
struct Plan {
    std::vector<int64_t> out_beg;            // first output index of every region within the part
    std::vector<LongPair> long_pairs;
    std::vector<WarpJob> jobs_k[kSlots];
};

// Pure host logic, no CUDA call: fills `p` (counts, mode, job ranges per slot, chunking) and `plan` for
// regions [g0, g1) of the batch.  Exported for tests and diagnostics as phmm_plan().
int plan_part(const phmm_batch* b, int g0, int g1, int64_t out0, int sm_count, float last_rescue_frac,
              HostPool& pool, Part& p, Plan& plan, std::string& err, int fp64_first_opt = 0)
{
    (void)err;
    static const bool trace_plan = getenv("PHMM_TRACE_PLAN") != nullptr;
    const auto tp0 = std::chrono::steady_clock::now();
    p = Part();
    p.g0 = g0; p.g1 = g1; p.out0 = out0;
    p.n_regions = g1 - g0;
    const int r0 = b->region_read_beg[g0], r1 = b->region_read_beg[g1];
    const int h0 = b->region_hap_beg[g0], h1 = b->region_hap_beg[g1];
    p.n_reads = r1 - r0; p.n_haps = h1 - h0;
    p.n_pairs = batch_pairs(b, g0, g1);
    if (p.n_pairs == 0) return PHMM_OK;
    const int rb0 = b->read_off[r0], rb1 = b->read_off[r1];
    const int hb0 = b->hap_off[h0], hb1 = b->hap_off[h1];
    const size_t hap_bytes = (size_t)(hb1 - hb0);

    // Batch-constant gap penalties?  Always, for the reference's own callers (sam/sam.hpp:30-32);
    // per-base arrays are scanned once here, and the general kernels run only if they really vary.
    p.mode = detect_gap_mode(b, rb0, rb1, p.gap);
    const bool general = p.mode == kModeGeneral;

    // ---- plan: per region, reads sorted by length and cut into warp jobs (2 reads per lane group, 32/G
    //      groups per warp); the job's shape is the one its LONGEST read asks for.  Sorting keeps the
    //      reads of a job within a few bases of each other (few dummy rows) and leaves one partial job per
    //      region instead of one per shape: on the reference's windows, where half of the reads are
    //      clipped to 10..149 bases, that is ~14% fewer jobs than grouping by each read's own shape. ----
    auto is_aligned = [&](int R, int sh) {
        return !general && (R % kShapes[sh].K == 0) && (kShapes[sh].K * kShapes[sh].G - R >= kShapes[sh].K);
    };
    struct PlannedJob { WarpJob job; int shape; bool aligned; };
    // first output index of every region within this part (also uploaded: region_out_beg)
    std::vector<int64_t>& out_beg = plan.out_beg;
    out_beg.assign(p.n_regions + 1, 0);
    for (int g = g0; g < g1; g++)
        out_beg[g - g0 + 1] = out_beg[g - g0] + (int64_t)(b->region_read_beg[g + 1] - b->region_read_beg[g]) *
                                                    (b->region_hap_beg[g + 1] - b->region_hap_beg[g]);
    struct PlanPiece {                       // one contiguous range of regions, planned by one host thread
        std::vector<PlannedJob> planned;
        std::vector<LongPair> long_pairs;
        int64_t n_cells = 0;
        int max_nh = 0;
        int64_t n_jobs_sh[kNumShapes] = {0}, n_aligned_sh[kNumShapes] = {0};
    };
    const int n_pieces = std::max(1, std::min(pool.width(), p.n_reads / 4096));
    std::vector<PlanPiece> pieces(n_pieces);
    std::vector<int> piece_cut(n_pieces + 1, g1);
    piece_cut[0] = g0;
    for (int t = 1, g = g0; t < n_pieces; t++) {           // cut by read count
        const int64_t want = r0 + (int64_t)p.n_reads * t / n_pieces;
        while (g < g1 && b->region_read_beg[g] < want) g++;
        piece_cut[t] = g;
    }
    // PACKED jobs hold reads of ONE length; worth a kernel of their own only when the batch has plenty
    bool packed_on[kMaxReadLenCompiled + 1] = {false};
    if (!general) {
        std::vector<int32_t> n_of_len(kMaxReadLenCompiled + 1, 0);
        for (int r = r0; r < r1; r++) {
            const int R = b->read_off[r + 1] - b->read_off[r];
            if (R <= kMaxReadLenCompiled) n_of_len[R]++;
        }
        for (int R = 1; R <= kMaxReadLenCompiled; R++)
            packed_on[R] = n_of_len[R] >= 256 && n_of_len[R] * 10 >= p.n_reads && packed_shape(R, nullptr) >= 0;
    }
    const auto tp1 = std::chrono::steady_clock::now();
    pool.parallel_for(n_pieces, [&](int t) {
        PlanPiece& pc = pieces[t];
        std::vector<std::pair<int, int>> by_len, sorted;  // (read length, read index within the part)
        pc.planned.reserve((size_t)(b->region_read_beg[piece_cut[t + 1]] - b->region_read_beg[piece_cut[t]]) / 3 + 16);
        for (int g = piece_cut[t]; g < piece_cut[t + 1]; g++) {
            const int nh = b->region_hap_beg[g + 1] - b->region_hap_beg[g];
            pc.max_nh = std::max(pc.max_nh, nh);
            if (nh == 0) continue;
            const int64_t hap_sum = b->hap_off[b->region_hap_beg[g + 1]] - b->hap_off[b->region_hap_beg[g]];
            const int h_avg = (int)(hap_sum / nh);
            by_len.clear();
            for (int r = b->region_read_beg[g]; r < b->region_read_beg[g + 1]; r++) {
                const int R = b->read_off[r + 1] - b->read_off[r];
                pc.n_cells += (int64_t)R * hap_sum;
                if (R > kMaxReadLenCompiled) {   // beyond one lane-group pass: phmm_long.cu, one warp per pair
                    for (int h = 0; h < nh; h++)
                        pc.long_pairs.push_back({r - r0, b->region_hap_beg[g] - h0 + h,
                                                 out_beg[g - g0] + (int64_t)(r - b->region_read_beg[g]) * nh + h});
                    continue;
                }
                by_len.emplace_back(R, r - r0);
            }
            // reads of a PACKED length: jobs of 2 reads per group, floor(32 / nl) groups per warp
            for (size_t i = 0; i < by_len.size();) {
                if (!packed_on[by_len[i].first]) { ++i; continue; }
                const int R = by_len[i].first;
                int nl = 0;
                const int sh = packed_shape(R, &nl);
                const int cap = 2 * std::min(32 / nl, kMaxJobReads / 2);
                PlannedJob pj;
                pj.shape = sh; pj.aligned = true;
                pj.job.region = g - g0;
                int q = 0;
                size_t w = i;                                  // compact the rest of by_len over the reads taken
                for (size_t j = i; j < by_len.size(); j++) {
                    if (by_len[j].first == R && q < cap) pj.job.read[q++] = by_len[j].second;
                    else by_len[w++] = by_len[j];
                }
                by_len.resize(w);
                for (; q < kMaxJobReads; q++) pj.job.read[q] = -1;
                pc.n_jobs_sh[sh]++; pc.n_aligned_sh[sh]++;
                pc.planned.push_back(pj);
            }
            {   // longest first, ties in read order: a stable counting sort over the 255 possible lengths
                int first_of_len[kMaxReadLenCompiled + 2] = {0};
                for (const auto& x : by_len) first_of_len[x.first]++;
                int acc = 0;
                for (int R = kMaxReadLenCompiled; R >= 1; R--) { const int c = first_of_len[R]; first_of_len[R] = acc; acc += c; }
                sorted.resize(by_len.size());
                for (const auto& x : by_len) sorted[first_of_len[x.first]++] = x;
                by_len.swap(sorted);
            }
            for (size_t i = 0; i < by_len.size();) {
                const int sh = pick_shape(by_len[i].first, h_avg);
                const int cap = 2 * (32 / kShapes[sh].G);
                PlannedJob pj;
                pj.shape = sh; pj.aligned = true;
                pj.job.region = g - g0;
                int q = 0;
                for (; q < cap && i < by_len.size(); q++, i++) {
                    pj.job.read[q] = by_len[i].second;
                    pj.aligned = pj.aligned && is_aligned(by_len[i].first, sh);
                }
                for (; q < kMaxJobReads; q++) pj.job.read[q] = -1;
                pc.n_jobs_sh[sh]++; pc.n_aligned_sh[sh] += pj.aligned;
                pc.planned.push_back(pj);
            }
        }
    });
    const auto tp2 = std::chrono::steady_clock::now();
    std::vector<LongPair>& long_pairs = plan.long_pairs;
    long_pairs.clear();
    int64_t n_jobs_sh[kNumShapes] = {0}, n_aligned_sh[kNumShapes] = {0};
    size_t n_planned = 0;
    for (const PlanPiece& pc : pieces) {
        p.n_cells += pc.n_cells;
        p.max_nh = std::max(p.max_nh, pc.max_nh);
        n_planned += pc.planned.size();
        long_pairs.insert(long_pairs.end(), pc.long_pairs.begin(), pc.long_pairs.end());
        for (int sh = 0; sh < kNumShapes; sh++) { n_jobs_sh[sh] += pc.n_jobs_sh[sh]; n_aligned_sh[sh] += pc.n_aligned_sh[sh]; }
    }
    // Jobs whose reads are all a whole number of lanes take the ALIGNED kernels -- provided enough jobs of
    // that shape do (otherwise the split only adds small launches).
    std::vector<WarpJob>* jobs_k = plan.jobs_k;
    for (int k = 0; k < kSlots; k++) jobs_k[k].clear();
    for (const PlanPiece& pc : pieces)
        for (const PlannedJob& pj : pc.planned) {
            const bool use_al = pj.shape >= kFirstPackedShape ||        // PACKED shapes only exist as aligned kernels
                                (pj.aligned && n_aligned_sh[pj.shape] * 10 >= n_jobs_sh[pj.shape] && n_aligned_sh[pj.shape] >= 32);
            jobs_k[pj.shape + (use_al ? kNumShapes : 0)].push_back(pj.job);
        }
    (void)n_planned;
    for (int h = h0; h < h1; h++) p.max_H = std::max(p.max_H, b->hap_off[h + 1] - b->hap_off[h]);
    p.n_jobs = 0;
    for (int k = 0; k < kSlots; k++) { p.job_beg[k] = p.n_jobs; p.n_jobs += (int)jobs_k[k].size(); }
    p.job_beg[kSlots] = p.n_jobs;
    {   // Haplotypes streamed per (job, chunk).  More per chunk: the wavefront fills and drains once per
        // chunk and the per-job setup (prior tables) is paid once.  Fewer: more independent units, and a
        // bounded longest unit -- a launch is over when its longest warp is, and a region with 16 haplotypes
        // in a stream of 2-haplotype regions would otherwise run 8x longer than the rest (measured on the
        // ragged window stream: single-shape kernels took one such job's duration).  Makespan model over the
        // candidates: all steps / resident warps + the longest unit; overhead per chunk ~128 steps.
        const int nhm = std::max(1, p.max_nh);
        const int by_smem = std::max(1, (kSmemBytesPerWarpBudget - 2 * (kSkew * 31 + 3) - 16) / (p.max_H + 2 + kPerHapTableBytes));   // + NEXT (+ NEXT2)
        std::vector<int64_t> jobs_with_nh(nhm + 1, 0);
        for (const PlanPiece& pc : pieces)
            for (const PlannedJob& pj : pc.planned) {
                const int g = g0 + pj.job.region;
                jobs_with_nh[b->region_hap_beg[g + 1] - b->region_hap_beg[g]]++;
            }
        const double h_mean = p.n_haps ? (double)hap_bytes / p.n_haps : 1.0;
        const double resident = (double)sm_count * 12;
        constexpr double kChunkOverheadSteps = 128;   // calibrated on S3: 4 haplotypes per chunk beat 2 by 1%
        static const int force_hpj = [] { const char* s = getenv("PHMM_FORCE_HPJ"); return s ? atoi(s) : 0; }();
        int best_hpj = 1; double best_t = 0;
        for (int hpj = 1; hpj <= std::min(nhm, by_smem); hpj++) {
            double steps = 0;
            for (int n = 1; n <= nhm; n++)
                if (jobs_with_nh[n]) steps += (double)jobs_with_nh[n] * (n * (h_mean + 1) + ((n + hpj - 1) / hpj) * kChunkOverheadSteps);
            const double t = steps / resident + hpj * (double)(p.max_H + 1) + kChunkOverheadSteps;
            if (hpj == 1 || t <= best_t) { best_t = t; best_hpj = hpj; }
        }
        int hpj = force_hpj > 0 ? std::min(force_hpj, std::min(nhm, by_smem)) : best_hpj;
        p.haps_per_job = hpj;
        p.hap_chunks = (nhm + hpj - 1) / hpj;
        // FP64 redo: when few pairs underflow (the normal case), a poorly matching read underflows against
        // every haplotype of its chunk, and one warp redoing them one after the other is the critical path
        // of the whole launch -- so the redo cuts the haplotypes one per warp (warps without work exit on a
        // flag byte).  When most pairs are redone (long reads with low-quality tails) the coarse chunks
        // amortise setup and fill better; the choice follows the device's previous batch.
        const bool dense = last_rescue_frac > 0.05f;
        p.haps_per_job64 = dense ? hpj : 1;
        p.hap_chunks64 = (nhm + p.haps_per_job64 - 1) / p.haps_per_job64;
        // FP64 first pays when more than 2/3 of the pairs end up in FP64 anyway (cost 2 + (1 - x) against 1 + 2x
        // FP32-cell units for a redo share x); phmm_options.fp64_first: 0 auto, 1 never, 2 always
        p.f64_first = fp64_first_opt == 1 ? false : fp64_first_opt == 2 ? true : last_rescue_frac > 0.75f;
    }
    if (trace_plan) {
        const auto tp3 = std::chrono::steady_clock::now();
        auto ms = [](auto a, auto b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
        fprintf(stderr, "phmm plan trace: setup %.3f ms, pieces %.3f ms (%d), merge + chunk model %.3f ms\n",
                ms(tp0, tp1), ms(tp1, tp2), n_pieces, ms(tp2, tp3));
    }

    return PHMM_OK;
}

// Regions [g0,g1) of the batch on one device, in two phases.
//   A (reads the caller's arrays): pack them into the slot's pinned block -- offsets rebased to the part --
//     and start its upload; then `copied()` releases the submitter, who may free or reuse its arrays.
//   B (reads only the packed copy): plan (plan_part on a view of the pinned block), upload the job list,
//     launch, start the download.  Runs while the submitter validates and packs its next batch and while the
//     data upload is in flight; a failure here is kept in the slot and reported by phmm_wait.
// Regions [g0,g1) of the batch on one device, in two phases.
//   stage_pack   (reads the caller's arrays; the device's PACKER thread): pack them into the slot's pinned
//                block -- offsets rebased to the part -- and start its upload.  After it the submitter is
//                released and may free or reuse its arrays.
//   stage_launch (reads only the packed copy; the device's WORKER thread): plan (plan_part on a view of the
//                pinned block), upload the job list, launch, start the download.  Runs while the submitter
//                validates and packs its next batch and while the data upload is in flight; a failure here
//                is kept in the slot and reported by phmm_wait.
int stage_pack(DeviceCtx& dc, Slot& s, const phmm_batch* b, int g0, int g1, int64_t out0, std::string& err)
{
    Slot::StageCtx& c = s.stage;
    c = Slot::StageCtx();
    c.t_begin = std::chrono::steady_clock::now();
    c.g0 = g0; c.g1 = g1; c.out0 = out0;
    Part& p = s.part;
    p = Part();
    p.g0 = g0; p.g1 = g1; p.out0 = out0;
    p.n_regions = g1 - g0;
    HostPool& pool = *dc.pack_pool;
    {   // test hook (tests/test_multi_device_gpu.py): PHMM_FAULT_PACK=<first region> makes the pack phase of the
        // share that starts at that region fail ONCE, to exercise the partial-failure path of phmm_submit
        static std::atomic<int> fault_at{[] { const char* v = getenv("PHMM_FAULT_PACK"); return v ? atoi(v) : -1; }()};
        int want = g0;
        if (g0 > 0 && fault_at.compare_exchange_strong(want, -1)) { err = "injected pack failure (PHMM_FAULT_PACK)"; return PHMM_ERR_CUDA; }
    }
    const int r0 = b->region_read_beg[g0], r1 = b->region_read_beg[g1];
    const int h0 = b->region_hap_beg[g0], h1 = b->region_hap_beg[g1];
    const int n_reads = r1 - r0, n_haps = h1 - h0, n_regions = g1 - g0;
    p.n_reads = n_reads; p.n_haps = n_haps;
    p.read0 = c.read0 = r0;
    p.n_pairs = batch_pairs(b, g0, g1);
    if (p.n_pairs == 0) {
        if (s.gl.on) {                                        // sums over zero reads: finalize_part_gl writes them
            s.gl.gl0 = s.gl.gl_off[s.gl.s0]; s.gl.n_gl = s.gl.gl_off[s.gl.s1] - s.gl.gl0;
            s.gl.sites = nullptr; s.gl.site_hap_off = s.gl.site_read_off = s.gl.gl_off = nullptr;
        }
        return PHMM_OK;
    }
    c.empty = false;
    const int rb0 = b->read_off[r0], rb1 = b->read_off[r1];
    const int hb0 = b->hap_off[h0], hb1 = b->hap_off[h1];
    const size_t read_bytes = (size_t)(rb1 - rb0), hap_bytes = (size_t)(hb1 - hb0);
    uint8_t gap[3];
    const bool general = detect_gap_mode(b, rb0, rb1, gap) == kModeGeneral;
    c.general = general;
    const bool zero_copy = (b->flags & PHMM_BATCH_PINNED_INPUTS) != 0;
    c.zero_copy = zero_copy;

    // ---- phase A: layout of the data block, pack, upload ----
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes); return o; };
    const size_t o_read_off = take(sizeof(int32_t) * (n_reads + 1));
    const size_t o_hap_off  = take(sizeof(int32_t) * (n_haps + 1));
    const size_t o_reg_read = take(sizeof(int32_t) * (n_regions + 1));
    const size_t o_reg_hap  = take(sizeof(int32_t) * (n_regions + 1));
    const size_t o_reg_out  = take(sizeof(int64_t) * (n_regions + 1));
    const size_t o_bases    = take(read_bytes);
    const size_t o_q        = take(read_bytes);
    const size_t o_gi       = general ? take(read_bytes) : 0;
    const size_t o_gd       = general ? take(read_bytes) : 0;
    const size_t o_gc       = general ? take(read_bytes) : 0;
    const size_t o_haps     = take(hap_bytes);
    const size_t in_bytes   = off;
    CUDA_TRY(s.h_in.reserve(in_bytes));
    CUDA_TRY(s.d_in.reserve(in_bytes));
    uint8_t* hp = (uint8_t*)s.h_in.p;
    {
        // the byte arrays are cut into one slice per host thread; the small index arrays ride along
        const int n_slices = std::max(1, std::min(2 * pool.width(), (int)(read_bytes >> 20)));
        auto slice_copy = [&](size_t dst_off, const uint8_t* src, size_t bytes, int t) {
            const size_t lo = bytes * t / n_slices, hi = bytes * (t + 1) / n_slices;
            staging_copy(hp + dst_off + lo, src + lo, hi - lo);
        };
        pool.parallel_for(n_slices + 1, [&](int t) {
            if (t < n_slices) {
                if (zero_copy) return;                      // uploaded straight from the caller's pinned arrays
                slice_copy(o_bases, b->read_bases + rb0, read_bytes, t);
                slice_copy(o_q, b->read_q + rb0, read_bytes, t);
                if (general) {
                    slice_copy(o_gi, b->read_i + rb0, read_bytes, t);
                    slice_copy(o_gd, b->read_d + rb0, read_bytes, t);
                    slice_copy(o_gc, b->read_c + rb0, read_bytes, t);
                }
                slice_copy(o_haps, b->hap_bases + hb0, hap_bytes, t);
                return;
            }
            int32_t* ro = (int32_t*)(hp + o_read_off);
            for (int r = 0; r <= n_reads; r++) ro[r] = b->read_off[r0 + r] - rb0;
            int32_t* ho = (int32_t*)(hp + o_hap_off);
            for (int h = 0; h <= n_haps; h++) ho[h] = b->hap_off[h0 + h] - hb0;
            int32_t* rr = (int32_t*)(hp + o_reg_read);
            int32_t* rh = (int32_t*)(hp + o_reg_hap);
            int64_t* rout = (int64_t*)(hp + o_reg_out);
            int64_t acc = 0;
            for (int g = 0; g <= n_regions; g++) {
                rr[g] = b->region_read_beg[g0 + g] - r0;
                rh[g] = b->region_hap_beg[g0 + g] - h0;
                rout[g] = acc;
                if (g < n_regions)
                    acc += (int64_t)(b->region_read_beg[g0 + g + 1] - b->region_read_beg[g0 + g]) *
                           (b->region_hap_beg[g0 + g + 1] - b->region_hap_beg[g0 + g]);
            }
        });
    }
    c.o_read_off = o_read_off; c.o_hap_off = o_hap_off; c.o_reg_read = o_reg_read; c.o_reg_hap = o_reg_hap;
    c.o_reg_out = o_reg_out; c.o_bases = o_bases; c.o_q = o_q; c.o_gi = o_gi; c.o_gd = o_gd; c.o_gc = o_gc;
    c.o_haps = o_haps; c.in_bytes = in_bytes;
    // the part as a batch of its own, over the packed copy: everything below reads this and not `b`
    phmm_batch& view = c.view;
    view = phmm_batch{};
    view.n_regions = n_regions; view.n_reads = n_reads; view.n_haps = n_haps;
    view.region_read_beg = (const int32_t*)(hp + o_reg_read);
    view.region_hap_beg = (const int32_t*)(hp + o_reg_hap);
    view.read_off = (const int32_t*)(hp + o_read_off);
    // (zero copy: the byte arrays stay the caller's, valid until phmm_wait by contract)
    view.read_bases = zero_copy ? b->read_bases + rb0 : hp + o_bases;
    view.read_q = zero_copy ? b->read_q + rb0 : hp + o_q;
    view.read_i = !general ? nullptr : zero_copy ? b->read_i + rb0 : hp + o_gi;
    view.read_d = !general ? nullptr : zero_copy ? b->read_d + rb0 : hp + o_gd;
    view.read_c = !general ? nullptr : zero_copy ? b->read_c + rb0 : hp + o_gc;
    view.hap_off = (const int32_t*)(hp + o_hap_off);
    view.hap_bases = zero_copy ? b->hap_bases + hb0 : hp + o_haps;
    view.gap_open_i = gap[0]; view.gap_open_d = gap[1]; view.gap_cont_c = gap[2];
    if (s.gl.on) {
        // the part's sites: rebased index arrays + the haplotype -> allele and read-overlap bytes, one pinned block
        Slot::GlCtx& q = s.gl;
        const phmm_sites* S = q.sites;
        const int ns = q.s1 - q.s0;
        const int64_t hap0 = q.site_hap_off[q.s0], read0 = q.site_read_off[q.s0];
        q.n_site_haps = q.site_hap_off[q.s1] - hap0;
        q.n_site_reads = q.site_read_off[q.s1] - read0;
        q.gl0 = q.gl_off[q.s0]; q.n_gl = q.gl_off[q.s1] - q.gl0;
        q.has_overlap = S->read_overlap != nullptr;
        size_t o = 0;
        auto take2 = [&](size_t bytes) { size_t at = o; o = align_up(o + bytes); return at; };
        q.o_region = take2(sizeof(int32_t) * (size_t)ns);
        q.o_alleles = take2(sizeof(int32_t) * (size_t)ns);
        q.o_hap_off = take2(sizeof(int64_t) * (size_t)(ns + 1));
        q.o_read_off = take2(sizeof(int64_t) * (size_t)(ns + 1));
        q.o_gl_off = take2(sizeof(int64_t) * (size_t)(ns + 1));
        q.o_hap_allele = take2((size_t)q.n_site_haps);
        q.o_overlap = take2(q.has_overlap ? (size_t)q.n_site_reads : 0);
        q.bytes = o;
        CUDA_TRY(s.h_sites.reserve(q.bytes + 16));
        CUDA_TRY(s.d_sites.reserve(q.bytes + 16));
        uint8_t* hs = (uint8_t*)s.h_sites.p;
        for (int k = 0; k < ns; k++) {
            ((int32_t*)(hs + q.o_region))[k] = S->site_region[q.s0 + k] - g0;
            ((int32_t*)(hs + q.o_alleles))[k] = S->site_n_alleles[q.s0 + k];
        }
        for (int k = 0; k <= ns; k++) {
            ((int64_t*)(hs + q.o_hap_off))[k] = q.site_hap_off[q.s0 + k] - hap0;
            ((int64_t*)(hs + q.o_read_off))[k] = q.site_read_off[q.s0 + k] - read0;
            ((int64_t*)(hs + q.o_gl_off))[k] = q.gl_off[q.s0 + k] - q.gl0;
        }
        if (q.n_site_haps) std::memcpy(hs + q.o_hap_allele, S->hap_allele + hap0, (size_t)q.n_site_haps);
        if (q.has_overlap && q.n_site_reads) std::memcpy(hs + q.o_overlap, S->read_overlap + read0, (size_t)q.n_site_reads);
        q.sites = nullptr; q.site_hap_off = q.site_read_off = q.gl_off = nullptr;     // the caller's arrays are not touched again
    }
    c.t_packed = std::chrono::steady_clock::now();
    if (s.gl.on && s.gl.bytes) CUDA_TRY(cudaMemcpyAsync(s.d_sites.p, s.h_sites.p, s.gl.bytes, cudaMemcpyHostToDevice, s.stream));
    if (!zero_copy) CUDA_TRY(cudaMemcpyAsync(s.d_in.p, s.h_in.p, in_bytes, cudaMemcpyHostToDevice, s.stream));
    else {
        uint8_t* dpz = (uint8_t*)s.d_in.p;
        CUDA_TRY(cudaMemcpyAsync(dpz, hp, o_bases, cudaMemcpyHostToDevice, s.stream));          // the index arrays
        CUDA_TRY(cudaMemcpyAsync(dpz + o_bases, b->read_bases + rb0, read_bytes, cudaMemcpyHostToDevice, s.stream));
        CUDA_TRY(cudaMemcpyAsync(dpz + o_q, b->read_q + rb0, read_bytes, cudaMemcpyHostToDevice, s.stream));
        if (general) {
            CUDA_TRY(cudaMemcpyAsync(dpz + o_gi, b->read_i + rb0, read_bytes, cudaMemcpyHostToDevice, s.stream));
            CUDA_TRY(cudaMemcpyAsync(dpz + o_gd, b->read_d + rb0, read_bytes, cudaMemcpyHostToDevice, s.stream));
            CUDA_TRY(cudaMemcpyAsync(dpz + o_gc, b->read_c + rb0, read_bytes, cudaMemcpyHostToDevice, s.stream));
        }
        CUDA_TRY(cudaMemcpyAsync(dpz + o_haps, b->hap_bases + hb0, hap_bytes, cudaMemcpyHostToDevice, s.stream));
    }
    return PHMM_OK;                       // the submitter's arrays are no longer needed (zero copy: only the index arrays)
}

int launch_genotype_part(DeviceCtx& dc, Slot& s, std::string& err);

// The scaled recurrence for per-base gap penalties (MODE 4) carries X / pMX_r: stepping from row r-1 to row r multiplies
// X^ by pXX_r pMX_{r-1} / pMX_r, which is harmless as long as the gap-open bytes of a batch do not spread by more than
// the gap-continuation penalty absorbs.  Conservative test over the whole part (bytes & 127, as the kernel reads
// them): continuation penalty >= 10 (pXX <= 0.1) and spread(i) - min(c) <= 10, i.e. that factor is <= 10 everywhere.
bool scaled_general_is_safe(const phmm_batch& v)
{
    if (!v.read_i || !v.read_d || !v.read_c || v.n_reads == 0) return false;
    const size_t n = (size_t)v.read_off[v.n_reads];
    int imin = 127, imax = 0, dmin = 127, dmax = 0, cmin = 127;
    for (size_t k = 0; k < n; k++) {
        const int i = v.read_i[k] & 127, d = v.read_d[k] & 127, c = v.read_c[k] & 127;
        imin = std::min(imin, i); imax = std::max(imax, i); dmin = std::min(dmin, d); dmax = std::max(dmax, d); cmin = std::min(cmin, c);
    }
    // (imin, dmin >= 10: pMM >= 0.8 on every row -- the kernel divides the row's gap weights by it)
    return cmin >= 10 && imin >= 10 && dmin >= 10 && (imax - imin) - cmin <= 10 && (dmax - dmin) - cmin <= 10;
}

int stage_launch(DeviceCtx& dc, Slot& s, bool exact, bool use_double, bool do_launch, std::string& err)
{
    s.device_log10 = dc.device_log10;
    static const bool trace = getenv("PHMM_TRACE") != nullptr;      // development aid: host time per phase
    Slot::StageCtx& c = s.stage;
    if (c.empty) return PHMM_OK;
    Part& p = s.part;
    HostPool& pool = *dc.pool;
    const phmm_batch& view = c.view;
    const int g0 = c.g0, g1 = c.g1, n_regions = c.g1 - c.g0;
    const int64_t out0 = c.out0;
    const bool general = c.general;
    const size_t o_read_off = c.o_read_off, o_hap_off = c.o_hap_off, o_reg_read = c.o_reg_read, o_reg_hap = c.o_reg_hap,
                 o_reg_out = c.o_reg_out, o_bases = c.o_bases, o_q = c.o_q, o_gi = c.o_gi, o_gd = c.o_gd, o_gc = c.o_gc,
                 o_haps = c.o_haps, in_bytes = c.in_bytes;
    const auto t_begin = c.t_begin, t_packed = c.t_packed;
    // ---- phase B: plan on the packed copy, upload the jobs, launch ----
    Plan plan;
    {
        int rcp = plan_part(&view, 0, n_regions, out0, dc.sm_count, dc.last_rescue_frac, pool, p, plan, err, dc.fp64_first_opt);
        if (rcp) return rcp;
        p.g0 = g0; p.g1 = g1; p.out0 = out0; p.read0 = c.read0;
        // constant gap penalties with i == d: every job takes the scaled recurrence (five FP32-pipe
        // instructions per cell) unless the engine was asked for the reference's operation order
        // (not for gap-open penalties beyond Q96: M^ = s M with s = min(1, 64 H pMX) must stay a normal float wherever
        //  M matters, i.e. down to ~1e-30 of the 2^120 scale, which needs H pMX >= 2e-10; the reference's 'I' is 5e-8)
        // (and not below Q10: the scaled kernels fold pMM into the priors and divide the gap weights by it; pMM >= 0.8 there,
        //  while it reaches 0 for gap-open penalties below Q4)
        if (p.mode == kModeConstShared && !exact && dc.scaled_recurrence && (p.gap[0] & 127) <= 96 && (p.gap[0] & 127) >= 10) p.mode = kModeConstScaled;
        if (p.mode == kModeGeneral && !exact && dc.scaled_recurrence && scaled_general_is_safe(view)) p.mode = kModeGeneralScaled;
    }
    const std::vector<LongPair>& long_pairs = plan.long_pairs;
    const std::vector<WarpJob>* jobs_k = plan.jobs_k;
    const auto t_planned = std::chrono::steady_clock::now();
    p.n_long = (int)long_pairs.size();
    const size_t o_long = align_up(sizeof(WarpJob) * (size_t)p.n_jobs);
    const size_t jobs_bytes = o_long + sizeof(LongPair) * long_pairs.size();
    CUDA_TRY(s.h_jobs.reserve(jobs_bytes + 16));
    CUDA_TRY(s.d_jobs.reserve(jobs_bytes + 16));
    const size_t n_pad = ((size_t)p.n_pairs + 3) / 4 * 4;
    const size_t out_bytes_all = 16 + 2 * sizeof(float) * n_pad;
    const size_t out_bytes = s.device_log10 ? 16 + sizeof(float) * (size_t)p.n_pairs : out_bytes_all;   // what the D2H moves
    CUDA_TRY(s.h_out.reserve(out_bytes_all));
    CUDA_TRY(s.d_out.reserve(out_bytes_all));
    CUDA_TRY(s.d_rescue.reserve(sizeof(RescueOut) * (size_t)p.n_pairs));
    CUDA_TRY(s.d_flags.reserve(kWorkHeaderBytes + (size_t)p.n_jobs * p.hap_chunks + 16));
    {   // work lists: per job and FP32 chunk one FP32 item and ceil(hpj / hpj64) FP64 items (launch_kernels)
        const size_t subs64 = (size_t)(p.haps_per_job + p.haps_per_job64 - 1) / p.haps_per_job64;
        CUDA_TRY(s.d_work.reserve(sizeof(uint2) * ((size_t)p.n_jobs * p.hap_chunks * (subs64 + 1) + 16)));
    }
    s.sm_count = dc.sm_count;
    s.log10_init_f = host_tables().log10_init_f;
    {
        uint8_t* hj = (uint8_t*)s.h_jobs.p;
        if (!long_pairs.empty()) std::memcpy(hj + o_long, long_pairs.data(), sizeof(LongPair) * long_pairs.size());
        WarpJob* jd = (WarpJob*)hj;
        for (int k = 0; k < kSlots; k++)
            if (!jobs_k[k].empty()) std::memcpy(jd + p.job_beg[k], jobs_k[k].data(), sizeof(WarpJob) * jobs_k[k].size());
    }

    uint8_t* dp = (uint8_t*)s.d_in.p;
    KernelArgs& a = s.args;
    a.read_off = (const int32_t*)(dp + o_read_off);
    a.hap_off = (const int32_t*)(dp + o_hap_off);
    a.region_read_beg = (const int32_t*)(dp + o_reg_read);
    a.region_hap_beg = (const int32_t*)(dp + o_reg_hap);
    a.region_out_beg = (const int64_t*)(dp + o_reg_out);
    a.read_bases = dp + o_bases;
    a.read_q = dp + o_q;
    a.read_i = general ? dp + o_gi : nullptr;
    a.read_d = general ? dp + o_gd : nullptr;
    a.read_c = general ? dp + o_gc : nullptr;
    if (!general) {
        // transition factors of the one (i,d,c) triple (avx-pairhmm-template.h:114-119), per precision
        const Tables& T = host_tables();
        const int gi = p.gap[0] & 127, gd = p.gap[1] & 127, gc = p.gap[2] & 127;
        const int mx = std::max(gi, gd), mn = std::min(gi, gd);
        const int mmi = ((mx * (mx + 1)) >> 1) + mn;
        a.cg_f[0] = T.mm_f[mmi]; a.cg_f[1] = 1.0f - T.ph2pr_f[gc]; a.cg_f[2] = T.ph2pr_f[gi];
        a.cg_f[3] = T.ph2pr_f[gd]; a.cg_f[4] = T.ph2pr_f[gc];
        a.cg_d[0] = T.mm_d[mmi]; a.cg_d[1] = 1.0 - T.ph2pr_d[gc]; a.cg_d[2] = T.ph2pr_d[gi];
        a.cg_d[3] = T.ph2pr_d[gd]; a.cg_d[4] = T.ph2pr_d[gc];
        a.gs_f = a.cg_f[1] * a.cg_f[2] / a.cg_f[0];          // MODE 3: g' = pGAPM * pMX / pMM (pMM is folded into the priors)
        a.gs_d = a.cg_d[1] * a.cg_d[2] / a.cg_d[0];
    }
    a.hap_bases = dp + o_haps;
    a.ph2pr_f = dc.d_ph2pr_f; a.mm_f = dc.d_mm_f; a.ph2pr_d = dc.d_ph2pr_d; a.mm_d = dc.d_mm_d;
    a.jobs = (const WarpJob*)s.d_jobs.p;
    s.d_long = (const LongPair*)((const uint8_t*)s.d_jobs.p + o_long);
    a.n_jobs = 0;
    a.haps_per_job = p.haps_per_job;
    a.stream_cap = (int32_t)((2 * (kSkew * 31 + 3) + (size_t)p.haps_per_job * (p.max_H + 2) + 127) / 128 * 128);
    a.smem_bytes_per_warp = 0;    // set per shape at launch (the prior tables depend on K and G)
    a.rescue_count = (unsigned*)s.d_out.p;
    a.raw32 = (float*)((uint8_t*)s.d_out.p + 16) + n_pad;
    a.rescue_out = (RescueOut*)s.d_rescue.p;
    a.job_flags = (uint8_t*)s.d_flags.p + kWorkHeaderBytes;
    a.job_flag_base = 0;
    a.work_in = nullptr; a.work_in_count = nullptr; a.work_in_cursor = nullptr; a.first64 = 0;

    if (jobs_bytes) CUDA_TRY(cudaMemcpyAsync(s.d_jobs.p, s.h_jobs.p, jobs_bytes, cudaMemcpyHostToDevice, s.stream));
    p.h2d_bytes = in_bytes + jobs_bytes + (s.gl.on ? s.gl.bytes : 0);
    if (!do_launch) return PHMM_OK;

    int rc = launch_kernels(s, exact, 1, 2, err, use_double);
    if (rc) return rc;
    if (s.gl.on) {
        // genotype reduction on the device: the matrix stays here, only the counters and the per-site vectors go back
        rc = launch_genotype_part(dc, s, err);
        if (rc) return rc;
        CUDA_TRY(cudaMemcpyAsync(s.h_out.p, s.d_out.p, 16, cudaMemcpyDeviceToHost, s.stream));
        p.d2h_bytes = 16 + s.gl.out_bytes;
    } else {
        CUDA_TRY(cudaMemcpyAsync(s.h_out.p, s.d_out.p, out_bytes, cudaMemcpyDeviceToHost, s.stream));
        p.d2h_bytes = out_bytes;
    }
    CUDA_TRY(cudaEventRecord(s.ev_done, s.stream));
    if (trace) {
        const auto t_end = std::chrono::steady_clock::now();
        auto ms = [](auto a, auto b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
        fprintf(stderr, "phmm trace: stage pack %.3f ms | plan %.3f ms, launch %.3f ms (%d jobs, %d launches, %lld pairs)\n",
                ms(t_begin, t_packed), ms(t_packed, t_planned), ms(t_planned, t_end), p.n_jobs, p.launches, (long long)p.n_pairs);
    }
    return PHMM_OK;
}

// Device-side genotype reduction of a staged slot (phmm_genotype.cu) and the download of its (small) outputs.
int launch_genotype_part(DeviceCtx& dc, Slot& s, std::string& err)
{
    Slot::GlCtx& q = s.gl;
    Part& p = s.part;
    const int ns = q.s1 - q.s0;
    std::call_once(dc.jacobian_once, [&] {
        int n = 0;
        const double* t = jacobian_table(&n);
        if (cudaMalloc(&dc.d_jacobian, sizeof(double) * (size_t)n) == cudaSuccess)
            cudaMemcpy(dc.d_jacobian, t, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice);
        else dc.d_jacobian = nullptr;
    });
    if (!dc.d_jacobian) { err = "cannot upload the Jacobian table"; return PHMM_ERR_OOM; }
    CUDA_TRY(s.d_lik64.reserve(sizeof(double) * (size_t)p.n_pairs + 16));
    CUDA_TRY(s.d_gl_scratch.reserve((size_t)q.n_site_reads * (8 * sizeof(double) + 1) + 256));
    size_t o = 0;
    auto take = [&](size_t bytes) { size_t at = o; o = align_up(o + bytes); return at; };
    q.out_gl = take(sizeof(double) * (size_t)q.n_gl);
    q.out_nused = take(sizeof(int32_t) * (size_t)ns);
    q.out_keep = take((size_t)p.n_reads);
    q.out_bytes = o;
    CUDA_TRY(s.d_gl_out.reserve(q.out_bytes + 16));
    CUDA_TRY(s.h_gl.reserve(q.out_bytes + 16));
    const uint8_t* ds = (const uint8_t*)s.d_sites.p;
    uint8_t* dout = (uint8_t*)s.d_gl_out.p;
    GenotypeArgs& g = q.args;
    g.lik64 = (double*)s.d_lik64.p;
    g.n_pairs = p.n_pairs; g.n_regions = p.n_regions; g.n_reads = p.n_reads; g.n_sites = ns;
    g.region_read_beg = s.args.region_read_beg; g.region_hap_beg = s.args.region_hap_beg;
    g.region_out_beg = s.args.region_out_beg; g.read_off = s.args.read_off;
    g.read_keep = dout + q.out_keep;
    g.site_region = (const int32_t*)(ds + q.o_region); g.site_n_alleles = (const int32_t*)(ds + q.o_alleles);
    g.site_hap_off = (const int64_t*)(ds + q.o_hap_off); g.hap_allele = ds + q.o_hap_allele;
    g.site_read_off = (const int64_t*)(ds + q.o_read_off); g.read_overlap = q.has_overlap ? ds + q.o_overlap : nullptr;
    g.gl_off = (const int64_t*)(ds + q.o_gl_off);
    g.scratch_al = (double*)s.d_gl_scratch.p;
    g.scratch_used = (uint8_t*)s.d_gl_scratch.p + (size_t)q.n_site_reads * 8 * sizeof(double);
    g.gl = (double*)(dout + q.out_gl); g.site_n_used = (int32_t*)(dout + q.out_nused);
    g.jacobian = dc.d_jacobian;
    g.inv_step = 1.0 / 0.0001;                               // JacobianLogTable::INV_STEP, math_utils.hpp:23
    g.log10_2 = std::log10(2.0);                             // genotyper.hpp:280, :316 (the host's libm, as the reference's)
    const size_t n_pad = ((size_t)p.n_pairs + 3) / 4 * 4;
    (void)n_pad;
    launch_genotype(g, (const float*)((const uint8_t*)s.d_out.p + 16), (const RescueOut*)s.d_rescue.p, (const unsigned*)s.d_out.p,
                    host_tables().log10_init_d, s.sm_count, s.stream);
    CUDA_TRY(cudaGetLastError());
    p.launches += 4;
    CUDA_TRY(cudaMemcpyAsync(s.h_gl.p, s.d_gl_out.p, q.out_bytes, cudaMemcpyDeviceToHost, s.stream));
    return PHMM_OK;
}

// both phases in the calling thread (the staged form)
int stage_and_launch(DeviceCtx& dc, Slot& s, const phmm_batch* b, int g0, int g1, int64_t out0,
                     bool exact, bool use_double, bool do_launch, std::string& err)
{
    int rc = stage_pack(dc, s, b, g0, g1, out0, err);
    if (rc) return rc;
    return stage_launch(dc, s, exact, use_double, do_launch, err);
}

// Wait for the slot, fetch the rescue list if any, convert raw sums to log10 (intel_pairhmm.hpp:137-143).
int finalize_part(phmm_engine* e, DeviceCtx& dc, Slot& s, phmm_result* r, std::string& err)
{
    static const bool trace = getenv("PHMM_TRACE") != nullptr;
    Part& p = s.part;
    if (s.async_rc) { err = s.async_err; return s.async_rc; }      // planning / launch failed after submit returned
    if (p.n_pairs == 0) return PHMM_OK;
    const auto t_begin = std::chrono::steady_clock::now();
    CUDA_TRY(cudaEventSynchronize(s.ev_done));
    const auto t_synced = std::chrono::steady_clock::now();
    float ms = 0.f;
    CUDA_TRY(cudaEventElapsedTime(&ms, s.ev_k0, s.ev_k1));
    p.kernel_ms = ms;
    const unsigned* header = (const unsigned*)s.h_out.p;     // {rescue_count, marked, underflowed, unscored}
    unsigned count = header[0];
    const size_t n_pad = ((size_t)p.n_pairs + 3) / 4 * 4;
    const float* lik32 = (const float*)((const uint8_t*)s.h_out.p + 16);
    const float* raw32 = lik32 + n_pad;
    if (count > (uint64_t)p.n_pairs) { err = "rescue counter overflow"; return PHMM_ERR_CUDA; }
    const Tables& T = host_tables();
    double* out = r->log10_lik + p.out0;
    float* o32 = r->raw32 ? r->raw32 + p.out0 : nullptr;
    double* o64 = r->raw64 ? r->raw64 + p.out0 : nullptr;
    uint8_t* ores = r->rescued ? r->rescued + p.out0 : nullptr;
    const float log10_init_f = T.log10_init_f;
    std::atomic<int64_t> need_rescue{0}, marked{0}, unscored{0};
    const int nt = (int)std::min<int64_t>(dc.pool->width(), std::max<int64_t>(1, p.n_pairs / 16384));
    if (s.device_log10) {
        // the device took the log10 (phmm_finalize.cu); what is left is widening floats.  Raw FP32 sums come over
        // only when the caller asks for them.
        marked = header[1]; need_rescue = header[2]; unscored = header[3];
        if (o32) {
            CUDA_TRY(cudaMemcpyAsync((void*)raw32, (const uint8_t*)s.d_out.p + 16 + sizeof(float) * n_pad, sizeof(float) * (size_t)p.n_pairs,
                                     cudaMemcpyDeviceToHost, s.stream));
            CUDA_TRY(cudaStreamSynchronize(s.stream));
            p.d2h_bytes += sizeof(float) * (size_t)p.n_pairs;
        }
        auto body = [&](int64_t i0, int64_t i1) {
            for (int64_t i = i0; i < i1; i++) out[i] = (double)lik32[i];        // NaN where the FP64 result goes
            if (o32) for (int64_t i = i0; i < i1; i++) o32[i] = std::fabs(raw32[i]);
            if (o64) std::memset(o64 + i0, 0, sizeof(double) * (size_t)(i1 - i0));
            if (ores) std::memset(ores + i0, 0, (size_t)(i1 - i0));
        };
        dc.pool->parallel_for(nt, [&](int t) { body(p.n_pairs * t / nt, p.n_pairs * (t + 1) / nt); });
    } else {
        auto body = [&](int64_t i0, int64_t i1) {
            int64_t nr = 0, nm = 0, bad = 0;
            for (int64_t i = i0; i < i1; i++) {
                const float f = raw32[i];
                if (f != f) { bad++; out[i] = std::nan(""); }           // an FP64-first pair the FP32 pass never scored: a bug
                else if (f < kMinAccepted) { nr++; nm += std::signbit(f); out[i] = std::nan(""); }
                else out[i] = (double)(log10f(f) - log10_init_f);       // float subtraction, :142
                if (o32) o32[i] = std::fabs(f);     // the sign bit only marks pairs for the flush-exact FP64 tier
                if (o64) o64[i] = 0.0;
                if (ores) ores[i] = 0;
            }
            need_rescue += nr; marked += nm; unscored += bad;
        };
        dc.pool->parallel_for(nt, [&](int t) { body(p.n_pairs * t / nt, p.n_pairs * (t + 1) / nt); });
    }
    if (unscored.load()) { err = std::to_string(unscored.load()) + " pairs left unscored by the FP32 pass"; return PHMM_ERR_CUDA; }
    if (marked.load()) {
        // tier 3: pairs whose FP64 sum came out within reach of the denormal range are redone by the
        // flush-exact FP64 kernels now; they append to the same rescue list
        int rc3 = launch_kernels(s, e->opt.exact_fp32 != 0, 3, 3, err);
        if (rc3) return rc3;
        CUDA_TRY(cudaMemcpyAsync(s.h_out.p, s.d_out.p, 16, cudaMemcpyDeviceToHost, s.stream));
        CUDA_TRY(cudaStreamSynchronize(s.stream));
        count = *(const unsigned*)s.h_out.p;
        p.d2h_bytes += 16;
        if (count > (uint64_t)p.n_pairs) { err = "rescue counter overflow"; return PHMM_ERR_CUDA; }
    }
    p.rescue_count = count;
    dc.last_rescue_frac = (float)((double)count / (double)p.n_pairs);
    if (count) {
        CUDA_TRY(s.h_rescue.reserve(sizeof(RescueOut) * (size_t)count));
        CUDA_TRY(cudaMemcpyAsync(s.h_rescue.p, s.d_rescue.p, sizeof(RescueOut) * (size_t)count, cudaMemcpyDeviceToHost, s.stream));
        CUDA_TRY(cudaStreamSynchronize(s.stream));
        p.d2h_bytes += sizeof(RescueOut) * (size_t)count;
    }
    if ((uint64_t)need_rescue.load() != count) {
        err = "rescue list size " + std::to_string(count) + " != FP32 underflows " + std::to_string(need_rescue.load());
        return PHMM_ERR_CUDA;
    }
    const RescueOut* rl = (const RescueOut*)s.h_rescue.p;
    for (unsigned k = 0; k < count; k++) {
        const int64_t i = rl[k].out_idx;
        double d = rl[k].raw64;
        // x86 FTZ also flushes double denormals (MXCSR, intel_pairhmm.hpp:102-105); the device
        // keeps them, so flush the final sum here.
        if (d < DBL_MIN) d = 0.0;
        out[i] = std::log10(d) - T.log10_init_d;                     // :139
        if (o64) o64[i] = d;
        if (ores) ores[i] = 1;
    }
    if (trace) {
        const auto t_end = std::chrono::steady_clock::now();
        auto ms = [](auto a, auto b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
        fprintf(stderr, "phmm trace: finalize wait %.3f ms, log10 + rescue %.3f ms (kernels %.3f ms)\n",
                ms(t_begin, t_synced), ms(t_synced, t_end), p.kernel_ms);
    }
    return PHMM_OK;
}

// phmm_wait_gl, one part: wait for the slot, run the flush-exact tier if a pair asked for it (and then the reduction
// again), hand the per-site vectors to the caller.
int finalize_part_gl(phmm_engine* e, DeviceCtx& dc, Slot& s, phmm_gl_result* r, std::string& err)
{
    Part& p = s.part;
    Slot::GlCtx& q = s.gl;
    const int ns = q.s1 - q.s0;
    if (s.async_rc) { err = s.async_err; return s.async_rc; }
    if (p.n_pairs == 0) {                                     // no pairs: every sum is over zero reads
        for (int64_t k = 0; k < q.n_gl; k++) r->genotype_lik[q.gl0 + k] = 0.0;
        if (r->site_n_reads) for (int k = 0; k < ns; k++) r->site_n_reads[q.s0 + k] = 0;
        if (r->read_keep) std::memset(r->read_keep + p.read0, 1, (size_t)p.n_reads);
        return PHMM_OK;
    }
    CUDA_TRY(cudaEventSynchronize(s.ev_done));
    float ms = 0.f;
    CUDA_TRY(cudaEventElapsedTime(&ms, s.ev_k0, s.ev_k1));
    p.kernel_ms = ms;
    const unsigned* header = (const unsigned*)s.h_out.p;      // {rescue_count, marked, underflowed, unscored}
    unsigned count = header[0];
    if (header[3]) { err = std::to_string(header[3]) + " pairs left unscored by the FP32 pass"; return PHMM_ERR_CUDA; }
    if (header[1]) {                                          // tier 3, then the reduction again over the patched matrix
        int rc3 = launch_kernels(s, e->opt.exact_fp32 != 0, 3, 3, err);
        if (rc3) return rc3;
        rc3 = launch_genotype_part(dc, s, err);
        if (rc3) return rc3;
        CUDA_TRY(cudaMemcpyAsync(s.h_out.p, s.d_out.p, 16, cudaMemcpyDeviceToHost, s.stream));
        CUDA_TRY(cudaStreamSynchronize(s.stream));
        count = header[0];
        p.d2h_bytes += 16 + q.out_bytes;
    }
    if (count != header[2]) {
        err = "rescue list size " + std::to_string(count) + " != FP32 underflows " + std::to_string(header[2]);
        return PHMM_ERR_CUDA;
    }
    p.rescue_count = count;
    dc.last_rescue_frac = (float)((double)count / (double)p.n_pairs);
    const uint8_t* ho = (const uint8_t*)s.h_gl.p;
    std::memcpy(r->genotype_lik + q.gl0, ho + q.out_gl, sizeof(double) * (size_t)q.n_gl);
    if (r->site_n_reads) std::memcpy(r->site_n_reads + q.s0, ho + q.out_nused, sizeof(int32_t) * (size_t)ns);
    if (r->read_keep) std::memcpy(r->read_keep + p.read0, ho + q.out_keep, (size_t)p.n_reads);
    if (r->capped_lik) {                                      // the matrix after all, on request (tests, hybrid callers)
        CUDA_TRY(cudaMemcpyAsync(r->capped_lik + p.out0, s.d_lik64.p, sizeof(double) * (size_t)p.n_pairs, cudaMemcpyDeviceToHost, s.stream));
        CUDA_TRY(cudaStreamSynchronize(s.stream));
        p.d2h_bytes += sizeof(double) * (size_t)p.n_pairs;
    }
    return PHMM_OK;
}

int init_slot(Slot& s, std::string& err)
{
    CUDA_TRY(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
    for (int i = 0; i < kAuxStreams; i++) {
        CUDA_TRY(cudaStreamCreateWithFlags(&s.aux[i], cudaStreamNonBlocking));
        CUDA_TRY(cudaEventCreateWithFlags(&s.ev_join[i], cudaEventDisableTiming));
    }
    CUDA_TRY(cudaEventCreateWithFlags(&s.ev_fork, cudaEventDisableTiming));
    CUDA_TRY(cudaEventCreate(&s.ev_k0));
    CUDA_TRY(cudaEventCreate(&s.ev_k1));
    CUDA_TRY(cudaEventCreate(&s.ev_k32));
    CUDA_TRY(cudaEventCreateWithFlags(&s.ev_done, cudaEventDisableTiming));
    return PHMM_OK;
}

int init_device(DeviceCtx& dc, int depth, std::string& err)
{
    static const bool trace = getenv("PHMM_TRACE_INIT") != nullptr;
    auto t0 = std::chrono::steady_clock::now();
    auto lap = [&](const char* what) {
        if (!trace) return;
        const auto t1 = std::chrono::steady_clock::now();
        fprintf(stderr, "phmm init trace:   device %d %s %.1f ms\n", dc.ordinal, what, std::chrono::duration<double, std::milli>(t1 - t0).count());
        t0 = t1;
    };
    CUDA_TRY(cudaSetDevice(dc.ordinal));
    CUDA_TRY(cudaFree(nullptr));                              // the primary context is created here
    lap("primary context");
    int major = 0, sms = 0;                                   // (cudaGetDeviceProperties fills ~1 KB of mostly unused fields, slowly)
    CUDA_TRY(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dc.ordinal));
    CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dc.ordinal));
    if (major < 10) { err = "device is not sm_100 or newer"; return PHMM_ERR_NO_DEVICE; }
    dc.sm_count = sms;
    lap("attributes");
    const Tables& T = host_tables();
    lap("host tables");
    // one allocation, one upload for the four probability tables
    const size_t b_pf = sizeof(float) * 128, b_mf = sizeof(float) * kMmEntries, b_pd = sizeof(double) * 128, b_md = sizeof(double) * kMmEntries;
    const size_t o_pd = 0, o_md = align_up(o_pd + b_pd), o_pf = align_up(o_md + b_md), o_mf = align_up(o_pf + b_pf), total = align_up(o_mf + b_mf);
    std::vector<uint8_t> blob(total, 0);
    std::memcpy(blob.data() + o_pd, T.ph2pr_d.data(), b_pd); std::memcpy(blob.data() + o_md, T.mm_d.data(), b_md);
    std::memcpy(blob.data() + o_pf, T.ph2pr_f.data(), b_pf); std::memcpy(blob.data() + o_mf, T.mm_f.data(), b_mf);
    uint8_t* d = nullptr;
    CUDA_TRY(cudaMalloc(&d, total));
    CUDA_TRY(cudaMemcpy(d, blob.data(), total, cudaMemcpyHostToDevice));
    dc.d_table_blob = d;
    dc.d_ph2pr_d = (double*)(d + o_pd); dc.d_mm_d = (double*)(d + o_md);
    dc.d_ph2pr_f = (float*)(d + o_pf); dc.d_mm_f = (float*)(d + o_mf);
    lap("device tables");
    dc.slots.resize(depth);
    for (auto& s : dc.slots) {
        int rc = init_slot(s, err);
        if (rc) return rc;
    }
    lap("slots (streams, events)");
    return PHMM_OK;
}

void free_slot(Slot& s)
{
    s.h_in.release(); s.h_jobs.release(); s.h_out.release(); s.h_rescue.release();
    s.d_in.release(); s.d_jobs.release(); s.d_out.release(); s.d_rescue.release(); s.d_flags.release(); s.d_work.release();
    s.h_sites.release(); s.h_gl.release(); s.d_sites.release(); s.d_gl_out.release(); s.d_lik64.release(); s.d_gl_scratch.release();
    for (int i = 0; i < kAuxStreams; i++) {
        if (s.ev_join[i]) cudaEventDestroy(s.ev_join[i]);
        if (s.aux[i]) cudaStreamDestroy(s.aux[i]);
    }
    if (s.ev_fork) cudaEventDestroy(s.ev_fork);
    if (s.ev_k0) cudaEventDestroy(s.ev_k0);
    if (s.ev_k1) cudaEventDestroy(s.ev_k1);
    if (s.ev_k32) cudaEventDestroy(s.ev_k32);
    if (s.ev_done) cudaEventDestroy(s.ev_done);
    if (s.stream) cudaStreamDestroy(s.stream);
}

// contiguous split of the regions over n devices, balanced by cell count
std::vector<int> split_regions(const phmm_batch* b, int n)
{
    std::vector<int> cut(n + 1, b->n_regions);
    cut[0] = 0;
    if (n == 1) return cut;
    std::vector<int64_t> pre(b->n_regions + 1, 0);
    for (int g = 0; g < b->n_regions; g++) pre[g + 1] = pre[g] + region_cells(b, g);
    const int64_t total = pre[b->n_regions];
    int g = 0;
    for (int d = 1; d < n; d++) {
        const int64_t want = total * d / n;
        while (g < b->n_regions && pre[g + 1] <= want) g++;
        // closer boundary of the two
        if (g < b->n_regions && (want - pre[g]) > (pre[g + 1] - want)) g++;
        cut[d] = std::max(cut[d - 1], g);
    }
    return cut;
}

}  // namespace

// ---- C ABI ---------------------------------------------------------------------------------------

extern "C" {

int phmm_abi_version(void) { return PHMM_ABI_VERSION; }

const char* phmm_strerror(int code)
{
    switch (code) {
        case PHMM_OK: return "ok";
        case PHMM_ERR_INVALID_ARG: return "invalid argument";
        case PHMM_ERR_NO_DEVICE: return "no usable sm_100 CUDA device (this engine has no CPU fallback)";
        case PHMM_ERR_CUDA: return "CUDA error";
        case PHMM_ERR_OOM: return "out of memory";
        case PHMM_ERR_UNSUPPORTED: return "unsupported input size";
        case PHMM_ERR_BAD_TICKET: return "unknown or already waited ticket";
        default: return "unknown error";
    }
}

const char* phmm_last_error(const phmm_engine* e)
{
    // a copy private to the calling thread: the engine's string may be rewritten by another thread's call
    thread_local std::string copy;
    if (!e) return "";
    { std::lock_guard<std::mutex> lk(const_cast<phmm_engine*>(e)->mu); copy = e->last_error; }
    return copy.c_str();
}

int phmm_tables(const float** ph2pr_f32, const float** mm_f32, const double** ph2pr_f64,
                const double** mm_f64, int32_t* mm_entries)
{
    const Tables& T = host_tables();
    if (ph2pr_f32) *ph2pr_f32 = T.ph2pr_f.data();
    if (mm_f32) *mm_f32 = T.mm_f.data();
    if (ph2pr_f64) *ph2pr_f64 = T.ph2pr_d.data();
    if (mm_f64) *mm_f64 = T.mm_d.data();
    if (mm_entries) *mm_entries = kMmEntries;
    return PHMM_OK;
}

int phmm_normalize_filter(double* lik, int32_t n_reads, int32_t n_haps, const int32_t* read_len, uint8_t* keep)
{
    // intel_pairhmm.hpp:24-46; constants :19-23
    if (!lik || !read_len || !keep || n_reads <= 0) return 0;
    if (n_haps <= 0) {                       // no haplotypes: nothing to cap, no evidence against any read
        for (int i = 0; i < n_reads; i++) keep[i] = 1;
        return n_reads;
    }
    int kept = 0;
    for (int i = 0; i < n_reads; i++) {
        double* row = lik + (size_t)i * n_haps;
        double best = *std::max_element(row, row + n_haps);
        const double cap = best + (-4.5);
        for (int j = 0; j < n_haps; j++) if (row[j] < cap) row[j] = cap;
        const double thr = std::min(2.0, std::ceil(read_len[i] * 0.02)) * (-4.0);
        keep[i] = (best < thr) ? 0 : 1;
        kept += keep[i];
    }
    return kept;
}

int phmm_host_register(void* p, size_t bytes)
{
    if (!p || !bytes) return PHMM_ERR_INVALID_ARG;
    return cudaHostRegister(p, bytes, cudaHostRegisterPortable) == cudaSuccess ? PHMM_OK : PHMM_ERR_CUDA;
}

int phmm_host_unregister(void* p)
{
    if (!p) return PHMM_ERR_INVALID_ARG;
    return cudaHostUnregister(p) == cudaSuccess ? PHMM_OK : PHMM_ERR_CUDA;
}

int64_t phmm_debug_check(phmm_engine* e)
{
    // guard zones of every device buffer of every slot (PHMM_DEBUG_GUARD=1); call with nothing in flight
    if (!e) return -1;
    DeviceGuard guard;
    int64_t bad = 0;
    for (auto& dc : e->devs) {
        cudaSetDevice(dc->ordinal);
        cudaDeviceSynchronize();
        for (auto& s : dc->slots)
            for (const DeviceBuf* b : {&s.d_in, &s.d_jobs, &s.d_out, &s.d_rescue, &s.d_flags, &s.d_work, &s.d_sites, &s.d_gl_out, &s.d_lik64, &s.d_gl_scratch}) {
                const int64_t k = b->check_guards();
                if (k < 0) return -1;
                bad += k;
            }
    }
    return bad;
}

int phmm_host_alloc(size_t bytes, void** out)
{
    if (!out || !bytes) return PHMM_ERR_INVALID_ARG;
    *out = nullptr;
    const cudaError_t e = cudaHostAlloc(out, bytes, cudaHostAllocPortable);
    if (e == cudaSuccess) return PHMM_OK;
    cudaGetLastError();
    return e == cudaErrorMemoryAllocation ? PHMM_ERR_OOM : (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver) ? PHMM_ERR_NO_DEVICE : PHMM_ERR_CUDA;
}

int phmm_host_free(void* p)
{
    if (!p) return PHMM_ERR_INVALID_ARG;
    return cudaFreeHost(p) == cudaSuccess ? PHMM_OK : PHMM_ERR_CUDA;
}

int phmm_plan(const phmm_batch* b, int32_t sm_count, int32_t host_threads, phmm_plan_info* info,
              int32_t* jobs_out, int64_t jobs_cap)
{
    if (!b || !info) return PHMM_ERR_INVALID_ARG;
    std::string err;
    int rc = validate_batch(b, err);
    if (rc) return rc;
    phmm_plan_info out{};
    out.struct_size = (int32_t)sizeof(out);
    out.n_shapes = kNumShapes;
    for (int s = 0; s < kNumShapes; s++) { out.shape_g[s] = kShapes[s].G; out.shape_k[s] = kShapes[s].K; }
    if (b->n_regions > 0) {
        HostPool pool(std::max(1, host_threads) - 1);
        Part p; Plan plan;
        rc = plan_part(b, 0, b->n_regions, 0, sm_count > 0 ? sm_count : 148, 0.f, pool, p, plan, err);
        if (rc) return rc;
        out.mode = p.mode; out.n_jobs = p.n_jobs; out.n_long_pairs = (int32_t)plan.long_pairs.size();
        out.haps_per_job = p.haps_per_job; out.hap_chunks = p.hap_chunks;
        out.haps_per_job64 = p.haps_per_job64; out.hap_chunks64 = p.hap_chunks64;
        out.n_pairs = p.n_pairs; out.n_cells = p.n_cells;
        int64_t row = 0;
        for (int k = 0; k < kSlots; k++) {
            (k < kNumShapes ? out.jobs_ragged : out.jobs_aligned)[k % kNumShapes] = (int32_t)plan.jobs_k[k].size();
            for (const WarpJob& j : plan.jobs_k[k]) {
                if (jobs_out && row < jobs_cap) {
                    int32_t* o = jobs_out + row * 10;
                    o[0] = k; o[1] = j.region;
                    for (int q = 0; q < kMaxJobReads; q++) o[2 + q] = j.read[q];
                }
                row++;
            }
        }
    }
    std::memcpy(info, &out, std::min<size_t>(sizeof(out), info->struct_size > 0 ? (size_t)info->struct_size : sizeof(out)));
    return PHMM_OK;
}

// everything a device context owns on the device (also the failure path of phmm_create)
static void teardown_device(DeviceCtx& dc)
{
    cudaSetDevice(dc.ordinal);
    for (auto& s : dc.slots) { if (s.stream) cudaStreamSynchronize(s.stream); free_slot(s); }
    dc.slots.clear();
    cudaFree(dc.d_table_blob); cudaFree(dc.d_jacobian);
    dc.d_table_blob = nullptr; dc.d_ph2pr_f = dc.d_mm_f = nullptr; dc.d_ph2pr_d = dc.d_mm_d = nullptr; dc.d_jacobian = nullptr;
}

int phmm_create(const phmm_options* opt, phmm_engine** out)
{
    if (!out) return PHMM_ERR_INVALID_ARG;
    *out = nullptr;
    DeviceGuard guard;
    std::unique_ptr<phmm_engine> e(new phmm_engine());
    if (opt) std::memcpy(&e->opt, opt, std::min<size_t>(sizeof(phmm_options), opt->struct_size > 0 ? (size_t)opt->struct_size : sizeof(phmm_options)));
    int n_dev = std::max(1, e->opt.n_devices);
    int depth = e->opt.pipeline_depth > 0 ? e->opt.pipeline_depth : 2;
    e->host_threads = std::max(1, e->opt.host_threads);
    // The device takes the final log10 only when the restated glibc algorithm IS this host's libm (self-test);
    // PHMM_HOST_LOG10=1 forces the host pass (A/B measurements, paranoia).
    e->device_log10 = log10_restatement_matches_libm() && getenv("PHMM_HOST_LOG10") == nullptr;
    static const bool trace_init = getenv("PHMM_TRACE_INIT") != nullptr;
    const auto ti0 = std::chrono::steady_clock::now();
    auto since = [&](std::chrono::steady_clock::time_point a) { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - a).count(); };
    int visible = 0;
    if (cudaGetDeviceCount(&visible) != cudaSuccess || visible == 0) return PHMM_ERR_NO_DEVICE;
    if (trace_init) fprintf(stderr, "phmm init trace: cudaGetDeviceCount (driver init) %.1f ms\n", since(ti0));
    std::string err;
    auto fail = [&](int rc) {                 // devices 0..d-1 are already initialised: give their memory back
        for (auto& dc : e->devs) teardown_device(*dc);
        return rc;
    };
    for (int d = 0; d < n_dev; d++) {
        std::unique_ptr<DeviceCtx> dc(new DeviceCtx());
        dc->ordinal = (opt && opt->devices) ? opt->devices[d] : d;
        if (dc->ordinal < 0 || dc->ordinal >= visible) return fail(PHMM_ERR_NO_DEVICE);
        // (an ordinal may be listed more than once: every entry is an independent worker with its own streams,
        //  tables and pools, which lets the sharding / gather path be exercised on a one-GPU box)
        dc->pool.reset(new HostPool(e->host_threads - 1));
        dc->pack_pool.reset(new HostPool(std::max(0, e->host_threads - 1)));
        dc->fp64_first_opt = e->opt.fp64_first;
        dc->device_log10 = e->device_log10;
        dc->scaled_recurrence = e->opt.recurrence == 0 && getenv("PHMM_REFERENCE_ORDER") == nullptr;
        const auto tid = std::chrono::steady_clock::now();
        int rc = init_device(*dc, depth, err);
        if (trace_init) fprintf(stderr, "phmm init trace: device %d context + tables + %d slots %.1f ms\n", dc->ordinal, depth, since(tid));
        if (rc) { fprintf(stderr, "phmm_create: %s\n", err.c_str()); teardown_device(*dc); return fail(rc); }
        e->devs.push_back(std::move(dc));
    }
    for (auto& dc : e->devs) { dc->worker.start(dc->ordinal); dc->packer.start(dc->ordinal); }
    *out = e.release();
    return PHMM_OK;
}

void phmm_destroy(phmm_engine* e)
{
    if (!e) return;
    DeviceGuard guard;
    for (auto& dc : e->devs) {
        dc->packer.finish();      // a pack task may still post its launch task to the worker: packer first
        dc->worker.finish();
        teardown_device(*dc);
    }
    delete e;
}

// Drain the slots of a ticket whose results nobody will fetch (a failed submit, a wait without an output
// buffer): every device that packed has a stage_launch task queued on its worker and copies in flight on the
// slot's stream -- with PHMM_BATCH_PINNED_INPUTS straight out of the caller's memory.  The drain task lands
// BEHIND that launch task on the same worker, synchronises the stream, and only then is the slot reusable.
static void drain_parts(phmm_engine* e, const std::vector<std::pair<int, int>>& parts)
{
    if (parts.empty()) return;
    Latch latch((int)parts.size());
    for (const auto& pr : parts) {
        DeviceCtx* dcp = e->devs[pr.first].get();
        Slot* sp = &dcp->slots[pr.second];
        dcp->post([sp, &latch] {
            if (sp->stream) cudaStreamSynchronize(sp->stream);
            for (int i = 0; i < kAuxStreams; i++) if (sp->aux[i]) cudaStreamSynchronize(sp->aux[i]);
            latch.done();
        });
    }
    latch.wait();
    std::lock_guard<std::mutex> lk(e->mu);
    for (const auto& pr : parts) e->devs[pr.first]->slots[pr.second].busy = false;
}

static int validate_sites(const phmm_batch* b, const phmm_sites* S, std::string& err)
{
    if (!S || S->n_sites < 0) { err = "sites is NULL"; return PHMM_ERR_INVALID_ARG; }
    if (S->n_sites == 0) return PHMM_OK;
    if (!S->site_region || !S->site_n_alleles || !S->hap_allele) { err = "NULL site array"; return PHMM_ERR_INVALID_ARG; }
    int64_t at = 0;
    for (int k = 0; k < S->n_sites; k++) {
        const int g = S->site_region[k], A = S->site_n_alleles[k];
        if (g < 0 || g >= b->n_regions || (k && g < S->site_region[k - 1])) { err = "site_region must be non-decreasing region indices"; return PHMM_ERR_INVALID_ARG; }
        if (A < 1 || A > PHMM_MAX_ALLELES) { err = "site_n_alleles must be 1.." + std::to_string(PHMM_MAX_ALLELES); return PHMM_ERR_INVALID_ARG; }
        const int nh = b->region_hap_beg[g + 1] - b->region_hap_beg[g];
        for (int h = 0; h < nh; h++)
            if (S->hap_allele[at + h] >= A) { err = "hap_allele names an allele the site does not have"; return PHMM_ERR_INVALID_ARG; }
        at += nh;
    }
    return PHMM_OK;
}

static int submit_impl(phmm_engine* e, const phmm_batch* b, const phmm_sites* sites, phmm_ticket* t)
{
    if (!e || !t) return PHMM_ERR_INVALID_ARG;
    DeviceGuard guard;
    std::string err;
    int rc = validate_batch(b, err);
    if (rc) { e->set_error(err); return rc; }
    std::vector<int64_t> site_hap_off, site_read_off, gl_off;      // prefix sums over the sites (live until the pack phase is over)
    if (sites) {
        if (!e->device_log10) { e->set_error("the host's libm is not the glibc this library restates: no device-side reduction"); return PHMM_ERR_UNSUPPORTED; }
        rc = validate_sites(b, sites, err);
        if (rc) { e->set_error(err); return rc; }
        const int ns = sites->n_sites;
        site_hap_off.assign(ns + 1, 0); site_read_off.assign(ns + 1, 0); gl_off.assign(ns + 1, 0);
        for (int k = 0; k < ns; k++) {
            const int g = sites->site_region[k], A = sites->site_n_alleles[k];
            site_hap_off[k + 1] = site_hap_off[k] + (b->region_hap_beg[g + 1] - b->region_hap_beg[g]);
            site_read_off[k + 1] = site_read_off[k] + (b->region_read_beg[g + 1] - b->region_read_beg[g]);
            gl_off[k + 1] = gl_off[k] + A * (A + 1) / 2;
        }
    }
    TicketRec rec;
    rec.gl = sites != nullptr;
    rec.t0 = std::chrono::steady_clock::now();
    rec.n_pairs = b->n_regions ? batch_pairs(b, 0, b->n_regions) : 0;
    const int nd = (int)e->devs.size();
    std::vector<int> cut = b->n_regions ? split_regions(b, nd) : std::vector<int>(nd + 1, 0);
    std::vector<int> rcs(nd, PHMM_OK);
    std::vector<std::string> errs(nd);
    std::vector<int> slot_of(nd, -1);
    int active = 0;
    {   // slot ring state (busy, next_slot) is shared with phmm_wait, which may run on another thread
        std::lock_guard<std::mutex> lk(e->mu);
        for (int d = 0; d < nd; d++) {
            if (cut[d + 1] == cut[d]) continue;
            DeviceCtx& dc = *e->devs[d];
            if (dc.slots[dc.next_slot].busy) {
                e->last_error = "all pipeline slots in flight: call phmm_wait first";
                return PHMM_ERR_INVALID_ARG;
            }
        }
        for (int d = 0; d < nd; d++) {
            if (cut[d + 1] == cut[d]) continue;
            DeviceCtx& dc = *e->devs[d];
            slot_of[d] = dc.next_slot;
            dc.slots[dc.next_slot].busy = true;
            dc.next_slot = (dc.next_slot + 1) % (int)dc.slots.size();
            active++;
        }
    }
    Latch latch(active);
    for (int d = 0; d < nd; d++) {
        if (slot_of[d] < 0) continue;
        DeviceCtx& dc = *e->devs[d];
        Slot& s = dc.slots[slot_of[d]];
        const int64_t out0 = batch_pairs(b, 0, cut[d]);
        const int g0 = cut[d], g1 = cut[d + 1];
        const bool exact = e->opt.exact_fp32 != 0, use_double = e->opt.use_double != 0;
        s.async_rc = PHMM_OK; s.async_err.clear();
        s.gl = Slot::GlCtx();
        if (sites) {                                          // the part's sites: those of regions [g0, g1)
            s.gl.on = true;
            s.gl.s0 = (int)(std::lower_bound(sites->site_region, sites->site_region + sites->n_sites, g0) - sites->site_region);
            s.gl.s1 = (int)(std::lower_bound(sites->site_region, sites->site_region + sites->n_sites, g1) - sites->site_region);
            s.gl.sites = sites;
            s.gl.site_hap_off = site_hap_off.data(); s.gl.site_read_off = site_read_off.data(); s.gl.gl_off = gl_off.data();
        }
        Slot* sp = &s;
        DeviceCtx* dcp = &dc;
        int* rc_out = &rcs[d];
        std::string* err_out = &errs[d];
        Latch* lp = &latch;
        dc.packer.post([=] {
            // pack on the device's packer thread; then the submitter is released -- from there on everything on
            // its stack (rcs, errs, latch, the batch) is gone -- and the plan + launch phase goes to the worker,
            // where a failure is parked in the slot for phmm_wait
            std::string local_err;
            const int rc = stage_pack(*dcp, *sp, b, g0, g1, out0, local_err);
            if (rc) { *rc_out = rc; *err_out = local_err; lp->done(); return; }
            dcp->post([=] {               // queued BEFORE the release: a phmm_wait that follows lands behind it
                std::string e2;
                const int rc2 = stage_launch(*dcp, *sp, exact, use_double, true, e2);
                if (rc2) { sp->async_rc = rc2; sp->async_err = e2; }
            });
            lp->done();
        });
        rec.parts.emplace_back(d, slot_of[d]);
    }
    latch.wait();     // caller's arrays are copied: they may be released now
    for (int d = 0; d < nd; d++)
        if (rcs[d]) {
            // one device failed to pack: the others already have uploads and launches under way.  Wait them out
            // before the slots are handed back and before the caller is told its arrays are its own again.
            drain_parts(e, rec.parts);
            e->set_error(errs[d]);
            return rcs[d];
        }
    std::lock_guard<std::mutex> lk(e->mu);
    *t = e->next_ticket++;
    e->tickets[*t] = std::move(rec);
    return PHMM_OK;
}

int phmm_submit(phmm_engine* e, const phmm_batch* b, phmm_ticket* t) { return submit_impl(e, b, nullptr, t); }

int phmm_submit_gl(phmm_engine* e, const phmm_batch* b, const phmm_sites* sites, phmm_ticket* t)
{
    if (!sites) { if (e) e->set_error("sites is NULL"); return PHMM_ERR_INVALID_ARG; }
    return submit_impl(e, b, sites, t);
}

int phmm_validate(const phmm_batch* b, const phmm_sites* sites)
{
    // the checks phmm_submit / phmm_submit_gl make before anything touches a device (pure host logic)
    std::string err;
    int rc = validate_batch(b, err);
    if (rc || !sites) return rc;
    return validate_sites(b, sites, err);
}

int phmm_jacobian_table(const double** table, int32_t* n)
{
    int k = 0;
    const double* tab = jacobian_table(&k);
    if (table) *table = tab;
    if (n) *n = k;
    return PHMM_OK;
}

int phmm_wait_gl(phmm_engine* e, phmm_ticket t, phmm_gl_result* r)
{
    if (!e || !r) return PHMM_ERR_INVALID_ARG;
    DeviceGuard guard;
    TicketRec rec;
    {
        std::lock_guard<std::mutex> lk(e->mu);
        auto it = e->tickets.find(t);
        if (it == e->tickets.end() || !it->second.gl) return PHMM_ERR_BAD_TICKET;
        rec = std::move(it->second);
        e->tickets.erase(it);
    }
    int64_t n_gl = 0;
    for (auto& pr : rec.parts) n_gl += e->devs[pr.first]->slots[pr.second].gl.n_gl;
    if (n_gl && !r->genotype_lik) {
        drain_parts(e, rec.parts);
        e->set_error("result->genotype_lik is NULL");
        return PHMM_ERR_INVALID_ARG;
    }
    const int np = (int)rec.parts.size();
    std::vector<int> rcs(np, PHMM_OK);
    std::vector<std::string> errs(np);
    Latch latch(np);
    for (int k = 0; k < np; k++) {
        DeviceCtx& dc = *e->devs[rec.parts[k].first];
        dc.post([&, k] {
            rcs[k] = finalize_part_gl(e, *e->devs[rec.parts[k].first], e->devs[rec.parts[k].first]->slots[rec.parts[k].second], r, errs[k]);
            latch.done();
        });
    }
    latch.wait();
    phmm_stats st{};
    st.n_pairs = rec.n_pairs;
    int rc_all = PHMM_OK;
    {
        std::lock_guard<std::mutex> lk(e->mu);
        for (int k = 0; k < np; k++) {
            Slot& s = e->devs[rec.parts[k].first]->slots[rec.parts[k].second];
            const Part& p = s.part;
            st.n_cells += p.n_cells; st.n_rescued += p.rescue_count;
            st.h2d_bytes += (int64_t)p.h2d_bytes; st.d2h_bytes += (int64_t)p.d2h_bytes;
            st.kernel_launches += p.launches;
            st.kernel_ms = std::max(st.kernel_ms, p.kernel_ms);
            st.n_devices_used++;
            s.busy = false;
            if (rcs[k] && !rc_all) { rc_all = rcs[k]; e->last_error = errs[k]; }
        }
    }
    st.total_ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - rec.t0).count();
    r->stats = st;
    return rc_all;
}

int phmm_wait(phmm_engine* e, phmm_ticket t, phmm_result* r)
{
    if (!e || !r) return PHMM_ERR_INVALID_ARG;
    DeviceGuard guard;
    TicketRec rec;
    {
        std::lock_guard<std::mutex> lk(e->mu);
        auto it = e->tickets.find(t);
        if (it == e->tickets.end() || it->second.gl) return PHMM_ERR_BAD_TICKET;
        rec = std::move(it->second);
        e->tickets.erase(it);
    }
    if (rec.n_pairs && !r->log10_lik) {
        drain_parts(e, rec.parts);            // the ticket is consumed: let its work finish, then free the slots
        e->set_error("result->log10_lik is NULL");
        return PHMM_ERR_INVALID_ARG;
    }
    const int np = (int)rec.parts.size();
    std::vector<int> rcs(np, PHMM_OK);
    std::vector<std::string> errs(np);
    Latch latch(np);
    for (int k = 0; k < np; k++) {
        DeviceCtx& dc = *e->devs[rec.parts[k].first];
        dc.post([&, k] {
            rcs[k] = finalize_part(e, *e->devs[rec.parts[k].first], e->devs[rec.parts[k].first]->slots[rec.parts[k].second], r, errs[k]);
            latch.done();
        });
    }
    latch.wait();
    phmm_stats st{};
    st.n_pairs = rec.n_pairs;
    int rc_all = PHMM_OK;
    std::string first_err;
    {
        std::lock_guard<std::mutex> lk(e->mu);
        for (int k = 0; k < np; k++) {
            Slot& s = e->devs[rec.parts[k].first]->slots[rec.parts[k].second];
            const Part& p = s.part;
            st.n_cells += p.n_cells; st.n_rescued += p.rescue_count;
            st.h2d_bytes += (int64_t)p.h2d_bytes; st.d2h_bytes += (int64_t)p.d2h_bytes;
            st.kernel_launches += p.launches;
            st.kernel_ms = std::max(st.kernel_ms, p.kernel_ms);
            st.n_devices_used++;
            s.busy = false;
            if (rcs[k] && !rc_all) { rc_all = rcs[k]; e->last_error = errs[k]; }
        }
    }
    st.total_ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - rec.t0).count();
    r->stats = st;
    return rc_all;
}

int phmm_compute(phmm_engine* e, const phmm_batch* b, phmm_result* r)
{
    phmm_ticket t;
    int rc = phmm_submit(e, b, &t);
    if (rc) return rc;
    return phmm_wait(e, t, r);
}

// ---- device-resident form (bench: inputs already in HBM) -------------------------------------

int phmm_stage(phmm_engine* e, const phmm_batch* b, phmm_staged** out)
{
    if (!e || !out) return PHMM_ERR_INVALID_ARG;
    *out = nullptr;
    std::string err;
    int rc = validate_batch(b, err);
    if (rc) { e->set_error(err); return rc; }
    if (b->n_regions == 0) { e->set_error("empty batch"); return PHMM_ERR_INVALID_ARG; }
    std::unique_ptr<phmm_staged> st(new phmm_staged());
    DeviceCtx& dc = *e->devs[0];
    Latch latch(1);
    dc.post([&] {
        Slot& s = st->slot;
        auto go = [&]() -> int {
            int rc1 = init_slot(s, err);
            if (rc1) return rc1;
            int rc2 = stage_and_launch(dc, s, b, 0, b->n_regions, 0, e->opt.exact_fp32 != 0, e->opt.use_double != 0, false, err);
            if (rc2) return rc2;
            CUDA_TRY(cudaStreamSynchronize(s.stream));
            return PHMM_OK;
        };
        rc = go();
        latch.done();
    });
    latch.wait();
    if (rc) { e->set_error(err); Latch l2(1); dc.post([&] { free_slot(st->slot); l2.done(); }); l2.wait(); return rc; }
    *out = st.release();
    return PHMM_OK;
}

int phmm_run_staged_ex(phmm_engine* e, phmm_staged* st, int32_t iters, float* ms_per_iter, float* fp32_ms_per_iter,
                       int32_t* launches_per_iter)
{
    if (!e || !st || iters < 1) return PHMM_ERR_INVALID_ARG;
    DeviceCtx& dc = *e->devs[0];
    std::string err;
    int rc = PHMM_OK;
    float ms = 0.f, ms32 = 0.f;
    bool have32 = true;
    int launches = 0;
    Latch latch(1);
    dc.post([&] {
        Slot& s = st->slot;
        Part& p = s.part;
        const bool exact = e->opt.exact_fp32 != 0;
        auto go = [&]() -> int {
            static const bool skip_rescue = getenv("PHMM_EXP_SKIP_RESCUE") != nullptr;   // timing experiments only
            // device time of `iters` back-to-back passes: sum of the per-pass [ev_k0, ev_k1] brackets
            for (int it = 0; it < iters; it++) {
                int rc2 = launch_kernels(s, exact, 1, skip_rescue ? 1 : 2, err, e->opt.use_double != 0);
                if (rc2) return rc2;
                CUDA_TRY(cudaEventSynchronize(s.ev_k1));
                float one = 0.f;
                CUDA_TRY(cudaEventElapsedTime(&one, s.ev_k0, s.ev_k1));
                ms += one;
                if (s.k32_valid) { CUDA_TRY(cudaEventElapsedTime(&one, s.ev_k0, s.ev_k32)); ms32 += one; }
                else have32 = false;
                launches = p.launches;
            }
            return PHMM_OK;
        };
        rc = go();
        latch.done();
    });
    latch.wait();
    if (rc) { e->set_error(err); return rc; }
    st->ran = true;
    st->slot.part.launches = launches;
    if (ms_per_iter) *ms_per_iter = ms / iters;
    if (fp32_ms_per_iter) *fp32_ms_per_iter = have32 ? ms32 / iters : -1.f;
    if (launches_per_iter) *launches_per_iter = launches;
    return PHMM_OK;
}

int phmm_run_staged(phmm_engine* e, phmm_staged* st, int32_t iters, float* ms_per_iter, int32_t* launches_per_iter)
{
    return phmm_run_staged_ex(e, st, iters, ms_per_iter, nullptr, launches_per_iter);
}

int phmm_run_staged_pipelined(phmm_engine* e, phmm_staged* const* sts, int32_t n, int32_t steps, float* total_ms,
                              int32_t* launches)
{
    if (!e || !sts || n < 1 || steps < 1) return PHMM_ERR_INVALID_ARG;
    for (int i = 0; i < n; i++) if (!sts[i]) return PHMM_ERR_INVALID_ARG;
    DeviceCtx& dc = *e->devs[0];
    std::string err;
    int rc = PHMM_OK;
    float ms = 0.f;
    int nl = 0;
    Latch latch(1);
    dc.post([&] {
        const bool exact = e->opt.exact_fp32 != 0;
        auto go = [&]() -> int {
            // Step i runs on the stream of staged batch i % n: consecutive steps overlap (one batch's FP64 redo
            // and tail with the next batch's FP32 kernel), a batch's own steps stay in order.  Timed from a
            // start event every stream waits on to an end event that waits on every stream.
            Slot& s0 = sts[0]->slot;
            cudaEvent_t ev_begin = nullptr, ev_end = nullptr;
            CUDA_TRY(cudaEventCreate(&ev_begin));
            CUDA_TRY(cudaEventCreate(&ev_end));
            CUDA_TRY(cudaEventRecord(ev_begin, s0.stream));
            for (int i = 1; i < std::min(n, steps); i++) CUDA_TRY(cudaStreamWaitEvent(sts[i]->slot.stream, ev_begin, 0));
            for (int it = 0; it < steps; it++) {
                Slot& s = sts[it % n]->slot;
                int rc2 = launch_kernels(s, exact, 1, 2, err, e->opt.use_double != 0);
                if (rc2) return rc2;
                nl += s.part.launches;
            }
            for (int i = 1; i < std::min(n, steps); i++) CUDA_TRY(cudaStreamWaitEvent(s0.stream, sts[i]->slot.ev_k1, 0));
            CUDA_TRY(cudaEventRecord(ev_end, s0.stream));
            CUDA_TRY(cudaEventSynchronize(ev_end));
            CUDA_TRY(cudaEventElapsedTime(&ms, ev_begin, ev_end));
            cudaEventDestroy(ev_begin); cudaEventDestroy(ev_end);
            return PHMM_OK;
        };
        rc = go();
        latch.done();
    });
    latch.wait();
    if (rc) { e->set_error(err); return rc; }
    for (int i = 0; i < std::min(n, steps); i++) sts[i]->ran = true;
    if (total_ms) *total_ms = ms;
    if (launches) *launches = nl;
    return PHMM_OK;
}

int phmm_fetch_staged(phmm_engine* e, phmm_staged* st, phmm_result* r)
{
    if (!e || !st || !r || !r->log10_lik) return PHMM_ERR_INVALID_ARG;
    if (!st->ran) { e->set_error("phmm_run_staged has not been called"); return PHMM_ERR_INVALID_ARG; }
    DeviceCtx& dc = *e->devs[0];
    std::string err;
    int rc = PHMM_OK;
    Latch latch(1);
    dc.post([&] {
        Slot& s = st->slot;
        auto go = [&]() -> int {
            const size_t n_pad = ((size_t)s.part.n_pairs + 3) / 4 * 4;
            const size_t out_bytes = s.device_log10 ? 16 + sizeof(float) * (size_t)s.part.n_pairs : 16 + 2 * sizeof(float) * n_pad;
            CUDA_TRY(cudaMemcpyAsync(s.h_out.p, s.d_out.p, out_bytes, cudaMemcpyDeviceToHost, s.stream));
            CUDA_TRY(cudaEventRecord(s.ev_done, s.stream));
            s.part.d2h_bytes = out_bytes;
            return finalize_part(e, dc, s, r, err);
        };
        rc = go();
        latch.done();
    });
    latch.wait();
    if (rc) { e->set_error(err); return rc; }
    const Part& p = st->slot.part;
    phmm_stats stt{};
    stt.n_pairs = p.n_pairs; stt.n_cells = p.n_cells; stt.n_rescued = p.rescue_count;
    stt.h2d_bytes = (int64_t)p.h2d_bytes; stt.d2h_bytes = (int64_t)p.d2h_bytes;
    stt.kernel_launches = p.launches; stt.n_devices_used = 1; stt.kernel_ms = p.kernel_ms;
    r->stats = stt;
    return PHMM_OK;
}

void phmm_free_staged(phmm_engine* e, phmm_staged* st)
{
    if (!e || !st) return;
    DeviceCtx& dc = *e->devs[0];
    Latch latch(1);
    dc.post([&] { if (st->slot.stream) cudaStreamSynchronize(st->slot.stream); free_slot(st->slot); latch.done(); });
    latch.wait();
    delete st;
}

}  // extern "C"
