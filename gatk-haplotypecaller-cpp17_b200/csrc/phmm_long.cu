// phmm_long.cu -- forward kernel for reads LONGER than one lane-group pass (256 .. PHMM_MAX_READ_LEN bases).
//
// The register-tiled kernels of phmm_kernels.cuh hold K <= 10 rows per lane in registers and stop at 255
// rows.  Longer reads are outside the reference's own envelope (its constant gap-penalty strings are 200
// long, sam/sam.hpp:30-32, so it reads out of bounds beyond that) but inside the C ABI's, so they get a
// simple, self-contained path instead of an error: ONE WARP PER (read, haplotype) PAIR, lane l owns the
// K = ceil(R / 32) consecutive rows l*K .., state and per-row parameters in per-thread local-memory
// arrays (dynamic row index; interleaved by lane, so every access is one coalesced L1 line), the same
// column wavefront with the bottom row handed down by shuffle.  All three precision tiers run inside the
// one launch: FP32 (ftz); if the sum is below 1e-28f the warp redoes the pair in FP64
// (intel_pairhmm.hpp:137); and if that ends within reach of the denormal range, flush-exact FP64 (see
// kFlushDanger in phmm_kernels.cuh).  Roughly 10x slower per cell than the register kernels; it only ever
// sees the reads they cannot take.
//
// Arithmetic follows avx-pairhmm-template.h:83-128 (row parameters), :161-175 (column 0), :188,:194,:197
// (recurrence) and :328-343 (two running sums): fused in the fast engine, unfused in the reference's
// operation order with exact_fp32 (bit-identical raw sums, tested).
#include "phmm_launch.h"

namespace phmm {

namespace {

constexpr int kLongK = kLongMaxRead / 32;        // rows per lane at the longest supported read

template <class S> struct LongOps;
template <> struct LongOps<float> {
    // flush-to-zero stated in the instruction (intel_pairhmm.hpp:102-105), and never contracted by ptxas
    __device__ static __forceinline__ float mul(float a, float b) { float r; asm("mul.rn.ftz.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b)); return r; }
    __device__ static __forceinline__ float add(float a, float b) { float r; asm("add.rn.ftz.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b)); return r; }
    __device__ static __forceinline__ float fma(float a, float b, float c) { float r; asm("fma.rn.ftz.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }
    __device__ static __forceinline__ float flush(float r) { return r; }          // ftz already
    __device__ static __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
    __device__ static __forceinline__ float div(float a, float b) { return __fdiv_rn(a, b); }
    __device__ static __forceinline__ float init_const() { return 1.329227995784916e+36f; }   // 2^120, Context.h:149
    __device__ static __forceinline__ const float* ph2pr(const KernelArgs& a) { return a.ph2pr_f; }
    __device__ static __forceinline__ const float* mm(const KernelArgs& a) { return a.mm_f; }
    __device__ static __forceinline__ float cg(const KernelArgs& a, int i) { return a.cg_f[i]; }
};
template <> struct LongOps<double> {
    __device__ static __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
    __device__ static __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
    __device__ static __forceinline__ double fma(double a, double b, double c) { return __fma_rn(a, b, c); }
    __device__ static __forceinline__ double flush(double r) { return (__double2hiint(r) < 0x00100000) ? 0.0 : r; }
    __device__ static __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }
    __device__ static __forceinline__ double div(double a, double b) { return __ddiv_rn(a, b); }
    __device__ static __forceinline__ double init_const() { return 1.1235582092889474e+307; }  // 2^1020, Context.h:109
    __device__ static __forceinline__ const double* ph2pr(const KernelArgs& a) { return a.ph2pr_d; }
    __device__ static __forceinline__ const double* mm(const KernelArgs& a) { return a.mm_d; }
    __device__ static __forceinline__ double cg(const KernelArgs& a, int i) { return a.cg_d[i]; }
};

// One pair, one warp.  EXACT: unfused, reference order; FLUSH: every product flushed (FP64 under the
// reference's MXCSR flush-to-zero).  Returns the raw sum in every lane.
template <class S, bool EXACT, bool FLUSH>
__device__ __noinline__ S long_forward(const KernelArgs& args, const int ro, const int R, const int ho, const int H,
                                       const bool general, const int lane)
{
    using O = LongOps<S>;
    S M[kLongK], X[kLongK], Y[kLongK];
    S mat[kLongK], mis[kLongK], fMM[kLongK], fGAPM[kLongK], fMX[kLongK], fMY[kLongK], fYY[kLongK];
    uint8_t rcode[kLongK];

    const int K = (R + 31) / 32;                 // rows per lane
    const int L = (R + K - 1) / K;               // lanes that own rows; the last one may own fewer than K
    const int row0 = lane * K;
    const int nrow = max(0, min(K, R - row0));
    const S* __restrict__ ph2pr = O::ph2pr(args);
    const S* __restrict__ mmtab = O::mm(args);
    const S one = (S)1, three = (S)3;

    // per-row parameters (avx-pairhmm-template.h:83-128)
    for (int k = 0; k < nrow; ++k) {
        const int ri = ro + row0 + k;
        rcode[k] = (uint8_t)base_code(args.read_bases[ri]);
        const S dist = ph2pr[args.read_q[ri] & 127];
        mat[k] = O::sub(one, dist);
        mis[k] = O::div(dist, three);
        if (general) {
            const int gi = args.read_i[ri] & 127, gd = args.read_d[ri] & 127, gc = args.read_c[ri] & 127;
            const int mx = max(gi, gd), mn = min(gi, gd);
            fMM[k] = mmtab[((mx * (mx + 1)) >> 1) + mn];
            fGAPM[k] = O::sub(one, ph2pr[gc]);
            fMX[k] = ph2pr[gi]; fMY[k] = ph2pr[gd]; fYY[k] = ph2pr[gc];
        } else {
            fMM[k] = O::cg(args, 0); fGAPM[k] = O::cg(args, 1); fMX[k] = O::cg(args, 2);
            fMY[k] = O::cg(args, 3); fYY[k] = O::cg(args, 4);
        }
        M[k] = 0; X[k] = 0; Y[k] = 0;            // column 0 (:161-175)
    }
    auto MUL = [](S a, S b) { return FLUSH ? O::flush(O::mul(a, b)) : O::mul(a, b); };

    const S init_y = O::div(O::init_const(), (S)H);     // row 0: M = X = 0, Y = INITIAL_CONSTANT / haplen (:86-92)
    S inM = 0, inX = 0, inY = (lane == 0) ? init_y : (S)0;
    S dgM = inM, dgX = inX, dgY = inY;
    S sumM = 0, sumX = 0;
    const int steps = H + L - 1;
    for (int t = 0; t < steps; ++t) {
        const int c = t - lane + 1;              // this lane's column (1-based) at this step
        if (lane < L && c >= 1 && c <= H) {
            const int hb = base_code(args.hap_bases[ho + c - 1]);
            S dM = dgM, dX = dgX, dY = dgY;      // (row-1, c-1)
            S uM = inM, uX = inX;                // (row-1, c)
            for (int k = 0; k < nrow; ++k) {
                const S oM = M[k], oX = X[k], oY = Y[k];                 // (row, c-1)
                const S prior = (rcode[k] == 4 || hb == 4 || rcode[k] == hb) ? mat[k] : mis[k];
                S nM, nX, nY;
                if (EXACT) {
                    nM = MUL(O::add(O::add(MUL(dM, fMM[k]), MUL(dX, fGAPM[k])), MUL(dY, fGAPM[k])), prior);   // :188
                    nX = O::add(MUL(uM, fMX[k]), MUL(uX, fYY[k]));                                            // :194
                    nY = O::add(MUL(oM, fMY[k]), MUL(oY, fYY[k]));                                            // :197
                } else {
                    nM = O::mul(O::fma(dY, fGAPM[k], O::fma(dX, fGAPM[k], O::mul(dM, fMM[k]))), prior);
                    nX = O::fma(uX, fYY[k], O::mul(uM, fMX[k]));
                    nY = O::fma(oY, fYY[k], O::mul(oM, fMY[k]));
                }
                M[k] = nM; X[k] = nX; Y[k] = nY;
                dM = oM; dX = oX; dY = oY;
                uM = nM; uX = nX;
            }
            if (lane == L - 1) {                 // last read row: running sums in column order (:328-343)
                sumM = O::add(sumM, M[nrow - 1]);
                sumX = O::add(sumX, X[nrow - 1]);
            }
        }
        // hand the bottom row to the lane below; lane 0 keeps row 0 above it
        const S bM = nrow ? M[nrow - 1] : (S)0, bX = nrow ? X[nrow - 1] : (S)0, bY = nrow ? Y[nrow - 1] : (S)0;
        dgM = inM; dgX = inX; dgY = inY;
        const S rM = __shfl_up_sync(0xffffffffu, bM, 1), rX = __shfl_up_sync(0xffffffffu, bX, 1),
                rY = __shfl_up_sync(0xffffffffu, bY, 1);
        if (lane != 0) { inM = rM; inX = rX; inY = rY; }
    }
    const S res = O::add(sumM, sumX);
    return __shfl_sync(0xffffffffu, res, L - 1);
}

template <bool EXACT>
__global__ void __launch_bounds__(kLongWarpsPerCta * 32)
long_read_kernel(const KernelArgs args, const LongPair* __restrict__ pairs, const int n_pairs, const int general, const int use_double)
{
    const int lane = threadIdx.x & 31;
    const int idx = blockIdx.x * kLongWarpsPerCta + (threadIdx.x >> 5);
    if (idx >= n_pairs) return;
    const LongPair p = pairs[idx];
    const int ro = args.read_off[p.read], R = args.read_off[p.read + 1] - ro;
    const int ho = args.hap_off[p.hap], H = args.hap_off[p.hap + 1] - ho;
    // use_double: intel_pairhmm.hpp:135 `g_use_double ? 0.0f : compute_float(&tc)`
    const float f = use_double ? 0.0f : long_forward<float, EXACT, false>(args, ro, R, ho, H, general != 0, lane);
    if (lane == 0) args.raw32[p.out_idx] = f;
    if (f < kMinAccepted) {                      // intel_pairhmm.hpp:137
        double d = long_forward<double, EXACT, EXACT>(args, ro, R, ho, H, general != 0, lane);
        if (!EXACT && d < kFlushDanger) d = long_forward<double, true, true>(args, ro, R, ho, H, general != 0, lane);
        if (lane == 0) {
            const unsigned slot = atomicAdd(args.rescue_count, 1u);
            args.rescue_out[slot].out_idx = p.out_idx;
            args.rescue_out[slot].raw64 = d;
        }
    }
}

}  // namespace

void launch_long_reads(const KernelArgs& args, const LongPair* pairs, int n_pairs, bool general, bool exact, bool use_double, cudaStream_t st)
{
    if (n_pairs <= 0) return;
    const int grid = (n_pairs + kLongWarpsPerCta - 1) / kLongWarpsPerCta;
    if (exact) long_read_kernel<true><<<grid, kLongWarpsPerCta * 32, 0, st>>>(args, pairs, n_pairs, general ? 1 : 0, use_double ? 1 : 0);
    else long_read_kernel<false><<<grid, kLongWarpsPerCta * 32, 0, st>>>(args, pairs, n_pairs, general ? 1 : 0, use_double ? 1 : 0);
}

}  // namespace phmm
