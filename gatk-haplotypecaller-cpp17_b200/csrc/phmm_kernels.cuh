// phmm_kernels.cuh -- sm_100a PairHMM forward kernels (FP32 packed f32x2, FP64 rescue).
//
// Replaces the reference's per-pair AVX kernels compute_full_prob_avxs<float> /
// compute_full_prob_avxd<double> (pairhmm/native/avx-pairhmm-template.h:210-346) and their
// per-pair setup (initializeVectors :83-128, precompute_masks :3-35).  Not a translation: the
// reference stripes 8 (4) rows over AVX lanes and walks anti-diagonals with lane shifts
// (avx-vector-shift.h); here
//
//   * a LANE GROUP of G lanes owns one column wavefront; lane l owns K consecutive read rows whose
//     M/X/Y state, priors and transition factors live in registers, so one step (= one haplotype
//     column per lane) is K cell updates against 3 shuffled values (the bottom row handed to lane
//     l+1), instead of 3 lane shifts per cell;
//   * the FP32 kernel packs TWO reads (same haplotype) into the two halves of f32x2 registers
//     (FFMA2/FMUL2/FADD2 on sm_100): 7-8 packed FP32-pipe instructions update two cells, half the
//     instruction stream of the scalar form for the same FP32 throughput (measured on B200: FFMA and
//     FFMA2 both saturate at ~125 lane-FMAs/clk/SM -- profiles/r01_microbench_pipes.txt), so shuffles,
//     loads and loop overhead are amortised over twice the work;
//   * the match/mismatch prior (avx-pairhmm-template.h:3-35,70-75,152-158) is NOT selected with
//     integer/select instructions: ncu showed every LOP3/FSEL paying 1-2 extra dispatch-stall cycles
//     behind the FFMA2s (register-file read ports), ~30% of the kernel.  Instead each read pair gets
//     a shared-memory table prior[hap base A,C,T,G,N][row] (both packed reads side by side), built
//     once per job; a step fetches its K priors with LDS.128 from the sub-table its haplotype base
//     names (the haplotype is staged as one byte per column holding that sub-table's offset).  The
//     LSU path is otherwise idle, the loads are conflict-free (odd 16-byte lane stride, 128-byte
//     aligned sub-tables) and the 2K prior registers of the select scheme are gone;
//   * gap penalties: the reference only ever passes the constant strings 'I','I','+' (sam/sam.hpp:30-32),
//     so MODE 1/2 take ONE (i,d,c) triple per batch and keep the five transition factors in
//     warp-uniform operands.  That matters beyond register count: measured on B200, FFMA2 with three
//     distinct vector register pairs issues every ~3.1 cycles, with one uniform/constant operand
//     every ~2.2 (register-file read ports: ~2 32-bit reads/lane/clk; profiles/r01_microbench_pipes.txt).
//     MODE 2 (i == d) additionally shares the product M*p between X of the row below and Y of the
//     next column (7 instead of 8 FP32-pipe instructions per cell, bit-identical results).
//     MODE 0 is the general per-base path (per-row factors in registers).  MODE 3 / MODE 4 (round 2, the fast
//     engines' default) are the SCALED forms of MODE 2 / MODE 0: X / pMX and Y / pMY are carried instead of X and Y,
//     which removes both M * p products -- six FP32-pipe instructions per cell; see the comment at the MODE enum.
//   * reads are right-aligned in the K*G row block: missing rows at the top are DUMMY rows that
//     reproduce row 0 of the reference (M = X = 0, Y = INITIAL_CONSTANT/haplen,
//     avx-pairhmm-template.h:86-92,161-175) exactly (priors 0, Y self-transition 1), so the last
//     read row is always the last row of the last lane and the final sum (:328-343) is a running
//     sum in that lane, in the reference's column order.  Two faster layouts exist for reads that are a
//     whole number of lanes (ALIGNED: dummy rows fill whole lanes; PACKED: lane groups of any width,
//     no dummy rows at all) -- see the kernel's comment;
//   * three precision tiers: FP32 (ftz), FP64 redo of FP32 underflows, flush-exact FP64 redo of the
//     pairs that end near the double denormal range (kFlushDanger).  Reads beyond 255 bases take
//     phmm_long.cu.
//
// Arithmetic: FMA-contracted, in the reference's operation order (MODE 0/1/2: 8 / 8 / 7 FP32-pipe instructions per
// cell; 8 is the roofline unit of SURVEY.md section 8d) or as the scaled recurrence (MODE 3/4: 6); EXACT=true uses
// unfused mul/add in the reference's operation order and is bit-identical to the reference's raw results.  FP32 is
// flush-to-zero (intel_pairhmm.hpp:102-105).
#pragma once
#include <cstdint>
#include <type_traits>
#include <cuda_runtime.h>

namespace phmm {

#ifndef PHMM_WARPS_PER_CTA
#define PHMM_WARPS_PER_CTA 1
#endif
// One warp per CTA: warps never cooperate, and single-warp CTAs let the block scheduler fill every SM
// sub-partition evenly at any register count (measured +9% on S3 over 4-warp CTAs at 186 registers).
constexpr int kWarpsPerCta = PHMM_WARPS_PER_CTA;
constexpr int kMaxJobReads = 8;          // reads per warp job: 2 per lane group, up to 4 groups (G = 8)
constexpr float kMinAccepted = 1e-28f;   // pairhmm/native/pairhmm_common.h:16
// Precision tiers.  1: FP32 (ftz).  2: FP64 redo of the pairs whose FP32 sum is below kMinAccepted
// (intel_pairhmm.hpp:137).  3: the reference's FP64 pass also runs flush-to-zero, so products that would
// be denormal vanish there, while the device keeps them.  That can only show at the 1e-9 level when the
// final FP64 sum is itself within ~15 orders of magnitude of DBL_MIN (at most 8*R*H flushed products of
// < 2.2e-308 each, carried forward with weight <= 1): such pairs (log10 likelihood below about -590) are
// redone by the EXACT FP64 kernel, whose products are flushed by hand.  The FP64 kernel marks them by
// setting the sign bit of their raw FP32 sum and the (job, chunk) flag to 2.
constexpr double kFlushDanger = 1e-285;
// FP64-FIRST order (tier 0), used when the previous batch redid most of its pairs in FP64: every pair is scored
// in FP64 first, and the FP32 pass -- whose only purpose would be to find out that the pair underflows -- is
// skipped wherever the FP64 sum PROVES that it would.  The FP32 kernel computes the same positive recurrence
// with factors that differ from the FP64 ones by <= 1.3e-6 relative (separately generated tables,
// native/Context.h:141-174) and rounds every operation to 2^-24 relative; along any path of R + H <= 8447
// cells that compounds to < 2% (flush-to-zero only makes the FP32 value smaller).  So an FP64 sum below HALF of
// the rescue threshold -- rescaled from 2^1020 to 2^120, i.e. 0.5e-28 * 2^900 -- guarantees raw FP32 < 1e-28f
// (intel_pairhmm.hpp:137) with a margin of 2x against that 2%: the pair goes straight to the rescue list with
// raw FP32 reported as 0.0f.  Every other pair is marked kNeedsF32 and gets the ordinary FP32 pass (and, if that
// does underflow after all, the ordinary FP64 redo): identical decisions, identical log10 values.
constexpr double kCertainUnderflow64 = 0.5e-28 * 8.452712498170644e+270;   // 0.5e-28 * 2^900
constexpr uint32_t kNeedsF32 = 0x7fc00001u;                                // a NaN payload no computation produces
// (job, chunk) flag byte: which later launches have work in this unit
enum : unsigned { kFlagRedo64 = 1u, kFlagFlush64 = 2u, kFlagNeedsF32 = 4u };
// OR into a flag byte (several warps / lanes may raise different bits of one byte: word-wide atomic)
__device__ __forceinline__ void flag_or(uint8_t* flag, unsigned bits) {
    const uintptr_t a = reinterpret_cast<uintptr_t>(flag);
    atomicOr(reinterpret_cast<unsigned*>(a & ~(uintptr_t)3), bits << (8u * (unsigned)(a & 3)));
}
// Lane l+1 runs kSkew columns behind lane l.  With 2, the bottom row a lane shuffles down at the end
// of a step is not consumed until a whole step later (shuffle latency off the X-chain critical path).
// Measured on the final kernel: no gain (the other warps hide that latency) while the straddling
// zone between two haplotypes doubles; 1 is 2-7% faster on every shape, so 1 is the default.
#ifndef PHMM_SKEW
#define PHMM_SKEW 1
#endif
constexpr int kSkew = PHMM_SKEW;

struct WarpJob {
    int32_t region;
    int32_t read[kMaxJobReads];          // global read indices, -1 = none
};

struct RescueOut {                       // one rescued pair (FP64 redo), compacted for D2H
    int64_t out_idx;
    double  raw64;
};

struct KernelArgs {
    // batch (device pointers)
    const int32_t* read_off;
    const uint8_t* read_bases;
    const uint8_t* read_q;
    const uint8_t* read_i;
    const uint8_t* read_d;
    const uint8_t* read_c;
    const int32_t* hap_off;
    const uint8_t* hap_bases;
    const int32_t* region_read_beg;
    const int32_t* region_hap_beg;
    const int64_t* region_out_beg;
    // probability tables (host-built, native/Context.h semantics)
    const float*  ph2pr_f;
    const float*  mm_f;
    const double* ph2pr_d;
    const double* mm_d;
    // MODE 1/2: batch-constant transition factors {pMM, pGAPM, pMX, pMY, pXX(=pYY)} per precision
    float  cg_f[5];
    double cg_d[5];
    // MODE 3 (scaled recurrence): pGAPM * pMX, rounded once on the host in each precision
    float  gs_f;
    double gs_d;
    // work list
    const WarpJob* jobs;
    int32_t n_jobs;
    int32_t haps_per_job;                // haplotypes streamed per (job, chunk)
    int32_t stream_cap;                  // bytes of haplotype stream per warp (multiple of 16)
    int32_t smem_bytes_per_warp;         // stream_cap + per-haplotype tables
    // results
    float*     raw32;                    // [n_pairs]
    RescueOut* rescue_out;               // [n_pairs] capacity
    unsigned*  rescue_count;
    uint8_t*   job_flags;                // [n_jobs_total * hap_chunks]: 1 = some pair of this (job, chunk)
                                         // underflowed in FP32, so the FP64 kernel has work there
    int32_t    job_flag_base;            // index of this launch's first job in job_flags
    int32_t    flag_hpj, flag_chunks;    // chunking of the FLAGS = the FP32 launch's haps_per_job and grid.y; an
                                         // FP64 launch may cut the haplotypes finer (its own haps_per_job)
    int32_t    tier;                     // FP64 launches: 2 = redo of FP32 underflows, 3 = flush-exact redo of the
                                         // pairs whose FP64 sum came out within reach of the denormal range;
                                         // 0 = FP64 FIRST (every pair; see kCertainUnderflow64)
    // Work list (LIST instantiations).  A launch either walks the grid (blockIdx.x = job, blockIdx.y = haplotype
    // chunk) or PULLS units from a compact list that build_work_list() made of the flag bytes an earlier launch
    // of the same stream raised: a fixed, small grid of warps, each taking the next item with an atomic cursor
    // until the list is exhausted.  An empty list costs a few microseconds instead of one CTA launch per
    // (job, chunk) -- the FP64 redo of a batch without underflows used to cost 7% of the ragged window stream
    // that way -- and a dense one is balanced dynamically.  Item = {unit, sub}: unit = job * flag_chunks + chunk
    // in the flag chunking (flag_hpj haplotypes per chunk), sub = which haps_per_job-sized piece of that chunk.
    const uint2* work_in;
    const unsigned* work_in_count;
    unsigned*  work_in_cursor;
    int32_t    first64;                  // this chain started FP64 first: a raw FP32 sum of exactly 0 is a pair that
                                         // is already on the rescue list (proven underflow), not one to redo
};

// ---- precision policies --------------------------------------------------------------------

struct PolicyF32x2 {
    using S = float;
    using V = float2;
    static constexpr int NH = 2;         // reads packed per lane group
    static constexpr bool kIsF32 = true;
    __device__ static __forceinline__ V splat(S s) { return make_float2(s, s); }
    __device__ static __forceinline__ S get(const V& v, int h) { return h ? v.y : v.x; }
    __device__ static __forceinline__ void set(V& v, int h, S s) { if (h) v.y = s; else v.x = s; }
    // Inline PTX with explicit .rn.ftz: (1) ptxas never contracts mul.rn + add.rn into an fma, which
    // the EXACT kernels rely on (the __fmul2_rn/__fadd2_rn intrinsics DID get fused under -fmad);
    // (2) flush-to-zero is stated in the instruction, not left to a compile flag.
    __device__ static __forceinline__ unsigned long long u64(V a) { return *reinterpret_cast<unsigned long long*>(&a); }
    __device__ static __forceinline__ V f2(unsigned long long a) { return *reinterpret_cast<V*>(&a); }
    __device__ static __forceinline__ V mul(V a, V b) {
        unsigned long long r; asm("mul.rn.ftz.f32x2 %0, %1, %2;" : "=l"(r) : "l"(u64(a)), "l"(u64(b))); return f2(r);
    }
    __device__ static __forceinline__ V mulx(V a, V b) { return mul(a, b); }   // EXACT product: ftz is in the instruction
    __device__ static __forceinline__ V add(V a, V b) {
        unsigned long long r; asm("add.rn.ftz.f32x2 %0, %1, %2;" : "=l"(r) : "l"(u64(a)), "l"(u64(b))); return f2(r);
    }
    __device__ static __forceinline__ V fma(V a, V b, V c) {
        unsigned long long r; asm("fma.rn.ftz.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(u64(a)), "l"(u64(b)), "l"(u64(c))); return f2(r);
    }
    // add that ptxas cannot contract with a preceding packed mul: ptxas 12.9 fuses mul.rn.f32x2 +
    // add.rn.f32x2 into FFMA2 even with -fmad=false (unlike the scalar forms), so the EXACT kernels
    // add the two halves with scalar add.rn (FMUL2 + 2 FADD, never an FMA).
    __device__ static __forceinline__ V addx(V a, V b) { return make_float2(__fadd_rn(a.x, b.x), __fadd_rn(a.y, b.y)); }
    __device__ static __forceinline__ V sel(bool p0, bool p1, V a, V b) {
        return make_float2(p0 ? a.x : b.x, p1 ? a.y : b.y);
    }
    __device__ static __forceinline__ V shfl_up(V v, int width) {
#ifdef PHMM_EXP_NOSHFL   // timing experiment only (wrong results): how much do the shuffles cost?
        return v;
#else
        return make_float2(__shfl_up_sync(0xffffffffu, v.x, 1, width),
                           __shfl_up_sync(0xffffffffu, v.y, 1, width));
#endif
    }
    __device__ static __forceinline__ S one() { return 1.0f; }
    __device__ static __forceinline__ S three() { return 3.0f; }
    __device__ static __forceinline__ S init_const() { return 1.329227995784916e+36f; }   // 2^120, Context.h:149
    __device__ static __forceinline__ S ssub(S a, S b) { return __fsub_rn(a, b); }
    __device__ static __forceinline__ S sdiv(S a, S b) { return __fdiv_rn(a, b); }
    __device__ static __forceinline__ S sadd(S a, S b) { return __fadd_rn(a, b); }
    __device__ static __forceinline__ const S* ph2pr(const KernelArgs& a) { return a.ph2pr_f; }
    __device__ static __forceinline__ const S* mm(const KernelArgs& a) { return a.mm_f; }
    __device__ static __forceinline__ S cg(const KernelArgs& a, int i) { return a.cg_f[i]; }
    __device__ static __forceinline__ S gs(const KernelArgs& a) { return a.gs_f; }
    __device__ static __forceinline__ S kScaleHeadroom() { return 0.015625f; }           // 2^-6: row 0 of the scaled recurrence <= 2^126
    __device__ static __forceinline__ S smul(S a, S b) { return __fmul_rn(a, b); }
};

struct PolicyF64 {
    using S = double;
    using V = double;
    static constexpr int NH = 1;
    static constexpr bool kIsF32 = false;
    __device__ static __forceinline__ V splat(S s) { return s; }
    __device__ static __forceinline__ S get(const V& v, int) { return v; }
    __device__ static __forceinline__ void set(V& v, int, S s) { v = s; }
    __device__ static __forceinline__ V mul(V a, V b) { return __dmul_rn(a, b); }
    // EXACT product: the reference runs with MXCSR flush-to-zero (intel_pairhmm.hpp:102-105), which also
    // flushes DOUBLE denormal results; only products can come out denormal here (every value is >= 0,
    // so a sum is at least its larger term), and the device has no ftz for FP64, so flush by hand.
    __device__ static __forceinline__ V mulx(V a, V b) {
        const double r = __dmul_rn(a, b);
        return (__double2hiint(r) < 0x00100000) ? 0.0 : r;
    }
    __device__ static __forceinline__ V add(V a, V b) { return __dadd_rn(a, b); }
    __device__ static __forceinline__ V fma(V a, V b, V c) { return __fma_rn(a, b, c); }
    __device__ static __forceinline__ V addx(V a, V b) { return __dadd_rn(a, b); }
    __device__ static __forceinline__ V sel(bool p0, bool, V a, V b) { return p0 ? a : b; }
    __device__ static __forceinline__ V shfl_up(V v, int width) { return __shfl_up_sync(0xffffffffu, v, 1, width); }
    __device__ static __forceinline__ S one() { return 1.0; }
    __device__ static __forceinline__ S three() { return 3.0; }
    __device__ static __forceinline__ S init_const() { return 1.1235582092889474e+307; }   // 2^1020, Context.h:109
    __device__ static __forceinline__ S ssub(S a, S b) { return __dsub_rn(a, b); }
    __device__ static __forceinline__ S sdiv(S a, S b) { return __ddiv_rn(a, b); }
    __device__ static __forceinline__ S sadd(S a, S b) { return __dadd_rn(a, b); }
    __device__ static __forceinline__ const S* ph2pr(const KernelArgs& a) { return a.ph2pr_d; }
    __device__ static __forceinline__ const S* mm(const KernelArgs& a) { return a.mm_d; }
    __device__ static __forceinline__ S cg(const KernelArgs& a, int i) { return a.cg_d[i]; }
    __device__ static __forceinline__ S gs(const KernelArgs& a) { return a.gs_d; }
    __device__ static __forceinline__ S kScaleHeadroom() { return 0.25; }                // 2^-2: row 0 <= 2^1022
    __device__ static __forceinline__ S smul(S a, S b) { return __dmul_rn(a, b); }
};

// Shared-memory loads by 32-bit shared address (keeps each cursor in ONE register; the generic
// pointer form made ptxas rebuild the shared window base every step).
__device__ __forceinline__ uint32_t lds_u8(uint32_t saddr) {
    uint32_t v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(saddr));
    return v;
}
template <class V>
__device__ __forceinline__ void lds_2v(uint32_t saddr, V& a, V& b) {   // 16 bytes = two 8-byte V
    unsigned long long x, y;
    asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(x), "=l"(y) : "r"(saddr));
    a = *reinterpret_cast<V*>(&x);
    b = *reinterpret_cast<V*>(&y);
}

// base byte -> table index; everything that is not A,C,T,G,N is 'A' (pairhmm_common.h:26-44)
__device__ __forceinline__ int base_code(uint8_t b) {
    int n = 0;
    n = (b == 'C') ? 1 : n;
    n = (b == 'T') ? 2 : n;
    n = (b == 'G') ? 3 : n;
    n = (b == 'N') ? 4 : n;
    return n;
}

// ---- shared-memory geometry of one shape (host and device agree through these) ----------------
// Prior table of one lane group: 5 sub-tables (haplotype base A,C,T,G,N), each holding for every
// lane its K priors (V = 8 bytes: both packed reads) at a lane stride of an ODD number of 16-byte
// units, so the 8 lanes of an LDS.128 quarter-warp hit 8 distinct bank groups.
__host__ __device__ constexpr int lane_units(int K) { return (((K + 1) / 2) % 2) ? (K + 1) / 2 : (K + 1) / 2 + 1; }
__host__ __device__ constexpr int subtable_bytes(int K, int G) { return (G * lane_units(K) * 16 + 127) / 128 * 128; }
constexpr int kZeroRowBytes = 128;      // a row of zero priors behind the tables (the folded recurrence's zero steps read it)
__host__ __device__ constexpr int tables_bytes(int K, int G) { return (32 / G) * 5 * subtable_bytes(K, G) + kZeroRowBytes; }
constexpr int kStreamNext = 0x80, kStreamIdle = 0x81, kStreamNext2 = 0x82;   // stream bytes >= 0x80 are not columns

// ---- the forward kernel ----------------------------------------------------------------------
//
// grid.x : warp jobs (kWarpsPerCta per CTA);  grid.y : haplotype chunks of args.haps_per_job
// One warp = 32/G lane groups.  FP32: group g scores reads job.read[2g], job.read[2g+1] together.
// FP64 (rescue): group g redoes job.read[2g] and then job.read[2g+1], each only for the haplotypes
// whose raw FP32 result is below 1e-28f (intel_pairhmm.hpp:137); it reads that decision straight
// from args.raw32 (and a per-(job,chunk) flag byte), so no work list is built between the kernels.
// MODE 3, the SCALED recurrence (constant gap penalties with i == d, fast engines only, every layout -- so that a
// pair's bits do not depend on which kernel its batch happened to put it in): carry X^ = X / pMX and Y^ = Y / pMY
// instead of X and Y (in fact (s M, X / s2, Y / s2) with s s2 = pMX, the split chosen per haplotype so that nothing
// overflows or eats the underflow guard: see where s_inity is filled).  Then
//     M[r][c]  = prior * (pMM * M[r-1][c-1] + g * (X^[r-1][c-1] + Y^[r-1][c-1])),   g = pGAPM * pMX  (= pGAPM * pMY)
//     X^[r][c] = M[r-1][c] + pXX * X^[r-1][c]          Y^[r][c] = M[r][c-1] + pYY * Y^[r][c-1]
// -- the products M * pMX and M * pMY are gone: six FP32-pipe instructions per cell (mul, fma, fma, mul, fma, fma)
// instead of the eight of the reference's expression (seven in MODE 2), the M * p array of MODE 2 and its K register
// pairs too; the final sum is sum(M) + pMX * sum(X^).  With pMM folded into the prior table on top (FOLD, below:
// M = prior' * (M_diag + g' * (X^_diag + Y^_diag))) it is FIVE: add, fma, mul, fma, fma.  Same recurrence in exact arithmetic; in floating point it
// rounds differently from the reference's operation order (as the FMA-contracted MODE 2 already does), well inside
// the 1e-4 bar on the final log10 -- measured, with the number of rescue decisions it moves at the threshold, in
// DESIGN.md section 4.1.  Row 0 of the reference (Y = INITIAL_CONSTANT / haplen) becomes Y^ = that / pMY.
// MODE 4 is the same idea for PER-BASE gap penalties: row r carries X^ = X / pMX_r and Y^ = Y / pMY_r, its five
// register-resident factors become pMM_r (folded into the row's priors: FOLD), pGAPM_r pMX_{r-1}, pGAPM_r pMY_{r-1} (both
// divided by pMM_r), pXX_r pMX_{r-1} / pMX_r and pYY_r, and
// the final sum is sum(M) + pMX_R sum(X^).  M itself is unscaled and the reference's row 0 keeps its own Y
// (its "pMY" is taken as 1), so nothing moves towards overflow or underflow as long as neighbouring rows' gap-open
// penalties do not differ by more than the engine checks for (phmm_engine.cu: scaled_general_is_safe).
enum : int { kModeGeneral = 0, kModeConst = 1, kModeConstShared = 2, kModeConstScaled = 3, kModeGeneralScaled = 4 };

#ifndef PHMM_MIN_WARPS
#define PHMM_MIN_WARPS 16      // resident warps per SM the small-K constant-gap FP32 kernels are held to
#endif
// ALIGNED: every read of the job has a length that is a multiple of K (and leaves >= K dummy rows), so
// dummy rows fill whole lanes.  Then no row needs its own Y self-transition: all rows use the uniform
// factor (one register-file read less per cell, K register pairs less per lane), dummy lanes may
// compute garbage Y, and the only Y that matters -- the one handed to the first real lane -- is
// overridden with init_Y as it is shuffled out.  Standard read lengths (150 = 15 lanes x 10 rows) qualify.
//
// PACKED (with ALIGNED, G == 32): lane groups of ANY width.  Every read of the job has the same length
// R = K * nl; the warp holds floor(32 / nl) groups of nl consecutive lanes back to back (100 bases: three
// groups of 10 lanes, 30 of 32 lanes busy, where the power-of-two groups manage 100 of 112 slots with a
// private Y factor per row).  No dummy rows or lanes at all: the first lane of a group takes the
// reference's row 0 -- (0, 0, init_Y) -- instead of what the shuffle hands it from the group above:
// zeroed factors annihilate the M and X it received, a select replaces the Y.  Lanes behind the last
// group shadow group 0 with zero priors and never emit.
// LIST: the launch pulls its units from a work list (KernelArgs::work_in) instead of walking the grid; the FP32
// LIST kernel is the selective pass of the FP64-first order (scores only the pairs marked kNeedsF32).
template <class P, int K, int G, int MODE, bool EXACT, bool ALIGNED, bool PACKED = false, bool LIST = false>
__global__ void __launch_bounds__(kWarpsPerCta * 32, (P::kIsF32 && K <= 5 && MODE != kModeGeneral) ? PHMM_MIN_WARPS / kWarpsPerCta : 1)
forward_kernel(const KernelArgs args)
{
    static_assert(!ALIGNED || (MODE != kModeGeneral && MODE != kModeGeneralScaled), "ALIGNED needs batch-constant gap penalties");
    static_assert(!PACKED || (ALIGNED && G == 32), "PACKED is a variant of ALIGNED on the whole warp");
    // ZRESET (the lane-aligned layouts): a lane does not RESET its registers between two haplotypes, it runs
    // two ordinary cell steps with its transition factors zeroed (stream bytes NEXT, NEXT2).  The recurrence
    // is linear in the state: with pMM = pGAPM = 0 the first step gives M = Pm = 0 (and X = 0, because the
    // lane above sent zeros one step earlier) and leaves Y = old Pm, the second, with pYY = 0 too, clears Y.
    // The straddling steps between haplotypes then cost one cell step each instead of a divergent ~150-
    // instruction reset block per lane (ncu: those 3% of the steps were 7% of the launch); idle bytes are
    // zero steps as well, so the straddling loop has no branch but the last lane's hand-over of its sums.
    constexpr bool ZRESET = ALIGNED;
    using S = typename P::S;
    using V = typename P::V;
    constexpr int NH = P::NH;
    constexpr int NG = 32 / G;
    constexpr bool CONSTG = MODE != kModeGeneral && MODE != kModeGeneralScaled;
    constexpr bool SHARED = MODE == kModeConstShared;
    constexpr bool SCALEDC = MODE == kModeConstScaled;        // scaled recurrence, constant gap penalties
    constexpr bool SCALEDG = MODE == kModeGeneralScaled;      // scaled recurrence, per-base gap penalties
    constexpr bool SCALED = SCALEDC || SCALEDG;
    // FOLD (the scaled modes): pMM is folded into the prior table (prior' = prior * pMM, rounded once per row and base) and
    // the gap weights are divided by it, so that M = prior' * (M_diag + g' * (X^_diag + Y^_diag)) -- add, fma, mul: FIVE
    // FP32-pipe instructions per cell with the two fma of X^ and Y^ (per-base gaps: fma, fma, mul, no add).  The engine
    // selects these modes only where pMM >= 0.8 (gap-open penalties >= Q10).
    constexpr bool FOLD = SCALED;
    static_assert(!SCALED || !EXACT, "the scaled recurrence exists for the fast kernels (exact = the reference's operation order)");
    constexpr int KP = CONSTG ? 1 : K;    // per-row factor arrays collapse to one warp-uniform entry
    constexpr int SUBT = subtable_bytes(K, G);
    constexpr int LANE_B = lane_units(K) * 16;
    static_assert(K >= 1 && K <= 16, "K rows per lane");
    static_assert(NG * 2 <= kMaxJobReads, "job too small for this group width");
    static_assert(4 * (SUBT / 128) < 128, "sub-table offset must fit the stream byte");
    static_assert(sizeof(V) == 8, "prior table entries are 8 bytes");

    extern __shared__ __align__(16) uint8_t smem[];
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    int grp = lane / G;                   // lane group and lane within it (PACKED: set once the job is known)
    int l   = lane % G;
    const unsigned n_items = LIST ? *args.work_in_count : 0u;   // final: the list was built earlier in this stream
#pragma unroll 1
    for (;;) {                            // LIST: one unit per round; grid walk: a single round
    int job_idx, h_first, h_cap;
    if constexpr (LIST) {
        unsigned i = 0;
        if (lane == 0) i = atomicAdd(args.work_in_cursor, 1u);
        i = __shfl_sync(0xffffffffu, i, 0);
        if (i >= n_items) return;
        const uint2 item = args.work_in[i];
        const int chunk = (int)(item.x % (unsigned)args.flag_chunks);
        job_idx = (int)(item.x / (unsigned)args.flag_chunks);
        h_first = chunk * args.flag_hpj + (int)item.y * args.haps_per_job;
        h_cap = min((chunk + 1) * args.flag_hpj, h_first + args.haps_per_job);
        grp = lane / G; l = lane % G;
        __syncwarp();                     // the previous unit's shared memory is free
    } else {
        job_idx = blockIdx.x * kWarpsPerCta + warp;
        if (job_idx >= args.n_jobs) return;
        h_first = blockIdx.y * args.haps_per_job;
        h_cap = h_first + args.haps_per_job;
    }

    uint8_t* const my_flag = args.job_flags + ((size_t)(args.job_flag_base + job_idx) * args.flag_chunks + h_first / args.flag_hpj);
    // grid walk of an FP64 redo: nothing to redo here -> one byte read, done
    if (!P::kIsF32 && !LIST && args.tier != 0 && !(*my_flag & (args.tier == 3 ? kFlagFlush64 : kFlagRedo64))) return;
    // FP64: which pairs this launch redoes, told from their raw FP32 sum (tier 0, FP64 first: every pair)
    auto needs_redo = [&](const float raw) {
        return args.tier == 0 ? true : args.tier == 3 ? (__float_as_uint(raw) >> 31) != 0u
                                                      : (raw < kMinAccepted && !(args.first64 && raw == 0.0f));
    };
    const WarpJob job = args.jobs[job_idx];
    const int hap_beg = args.region_hap_beg[job.region];
    const int nh      = args.region_hap_beg[job.region + 1] - hap_beg;
    if (h_first >= nh) { if constexpr (LIST) continue; else return; }
    const int h_last  = min(nh, h_cap);
    const int rd_beg  = args.region_read_beg[job.region];
    const int64_t out_base = args.region_out_beg[job.region];
    int nl = G;                           // lanes per group
    bool lane_live = true;                // PACKED: false behind the last group
    if (PACKED) {
        nl = (args.read_off[job.read[0] + 1] - args.read_off[job.read[0]]) / K;    // same for every read of the job
        const int n_grp = min(32 / nl, kMaxJobReads / 2);
        grp = lane / nl; l = lane - grp * nl;
        lane_live = grp < n_grp;
        if (!lane_live) { l = min(lane - n_grp * nl, nl - 1); grp = 0; }
    }

    const S* __restrict__ ph2pr = P::ph2pr(args);
    const S* __restrict__ mmtab = P::mm(args);

    // per-warp shared memory: [haplotype stream | prior tables | per-haplotype scalars]
    uint8_t* sb   = smem + (size_t)warp * args.smem_bytes_per_warp;
    uint8_t* stab = sb + args.stream_cap;
    S*   s_inity  = reinterpret_cast<S*>(stab + tables_bytes(K, G));
    S*   s_s2     = s_inity + args.haps_per_job;       // SCALED: per-haplotype scale of X^, Y^ and 1 / (scale of M^)
    S*   s_invs   = s_s2 + args.haps_per_job;
    int* s_hidx   = reinterpret_cast<int*>(s_invs + args.haps_per_job);
    int* s_apos   = s_hidx + args.haps_per_job;
    int* s_alen   = s_apos + args.haps_per_job;
    V* const my_tab = reinterpret_cast<V*>(stab + (PACKED ? 0 : grp) * 5 * SUBT + (PACKED ? lane : l) * LANE_B);   // + b * SUBT, [k]

    constexpr int NSUB = 2 / NH;         // FP64: the two reads of a group one after the other
#pragma unroll 1
    for (int sub = 0; sub < NSUB; ++sub) {
        int  rd[NH];
        bool valid[NH];
#pragma unroll
        for (int hf = 0; hf < NH; ++hf) {
            rd[hf] = job.read[2 * grp + sub * NH + hf];
            valid[hf] = rd[hf] >= 0 && lane_live;
        }
        if (!P::kIsF32) {
            // rescue kernel: skip this read unless some haplotype of the chunk needs the redo
            bool need = false;
            if (valid[0])
                for (int h = h_first + l; h < h_last; h += nl)
                    need |= needs_redo(args.raw32[out_base + (int64_t)(rd[0] - rd_beg) * nh + h]);
            if (!__any_sync(0xffffffffu, need)) continue;
        }
        const bool group_live = valid[0];
        if (!valid[0]) rd[0] = job.read[0];                 // idle group: shadow read 0, no output
#pragma unroll
        for (int hf = 1; hf < NH; ++hf) if (!valid[hf]) rd[hf] = rd[0];

        // ---- per-row setup (avx-pairhmm-template.h:83-128): transition factors in registers,
        //      priors (1 - dist on a match, dist / 3 otherwise; 0 on dummy rows) into the table ----
        V pYY[ALIGNED ? 1 : K];           // Y self-transition (1 on dummy rows).  In MODE 0 it is
                                          // also the row's X self-transition (pXX == pYY, :117,:119)
        V pMM[KP], pGAPM[KP], pMX[KP], pMY[KP];
        V pXXc = P::splat(0);             // MODE 1/2: X self-transition of every row
        int pad[NH];
        if (CONSTG) {
            pMM[0] = P::splat(P::cg(args, 0)); pGAPM[0] = P::splat(SCALEDC ? P::gs(args) : P::cg(args, 1));
            pMX[0] = P::splat(P::cg(args, 2)); pMY[0] = P::splat(P::cg(args, 3));
            pXXc   = P::splat(P::cg(args, 4));
        }
        __syncwarp();                     // previous sub / previous user of this warp's tables is done
        if (FOLD && ZRESET) reinterpret_cast<uint32_t*>(stab + tables_bytes(K, G) - kZeroRowBytes)[lane] = 0u;
        {
#pragma unroll
            for (int hf = 0; hf < NH; ++hf) {
                const int r  = rd[hf];
                const int ro = args.read_off[r];
                const int R  = args.read_off[r + 1] - ro;
                pad[hf] = PACKED ? 0 : K * G - R;      // >= 1 by construction of the plan (PACKED: R == K * nl)
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    const int ri = l * K + k - pad[hf];
                    S mat = 0, mis = 0, yy = P::one();
                    S mm_ = 0, gapm = 0, mx_ = 0, my_ = 0;
                    int rc = 4;
                    if (ri >= 0 && (!PACKED || lane_live)) {
                        rc = base_code(args.read_bases[ro + ri]);
                        const S dist = ph2pr[args.read_q[ro + ri] & 127];
                        mat = P::ssub(P::one(), dist);
                        mis = P::sdiv(dist, P::three());
                        if (!CONSTG) {
                            const int gi = args.read_i[ro + ri] & 127;
                            const int gd = args.read_d[ro + ri] & 127;
                            const int gc = args.read_c[ro + ri] & 127;
                            const int mx = max(gi, gd), mn = min(gi, gd);
                            mm_  = mmtab[((mx * (mx + 1)) >> 1) + mn];
                            gapm = P::ssub(P::one(), ph2pr[gc]);
                            mx_  = ph2pr[gi];
                            my_  = ph2pr[gd];
                            yy   = ph2pr[gc];
                            if (SCALEDG) {
                                // the row above: the read's previous base, or the reference's row 0 (factors taken as 1)
                                const S pmx_up = ri > 0 ? ph2pr[args.read_i[ro + ri - 1] & 127] : P::one();
                                const S pmy_up = ri > 0 ? ph2pr[args.read_d[ro + ri - 1] & 127] : P::one();
                                const S own_mx = mx_;
                                mx_  = P::smul(gapm, pmy_up);                         // slot of pMX: weight of Y^ above-left
                                my_  = P::sdiv(P::smul(yy, pmx_up), own_mx);          // slot of pMY: X^ self-transition
                                gapm = P::smul(gapm, pmx_up);                         // weight of X^ above-left
                                if (FOLD) { mx_ = P::sdiv(mx_, mm_); gapm = P::sdiv(gapm, mm_); }
                            }
                        } else {
                            yy = P::cg(args, 4);
                        }
                        if (FOLD) {                   // prior' = prior * pMM of the row
                            const S mmv = CONSTG ? P::cg(args, 0) : mm_;
                            mat = P::smul(mat, mmv); mis = P::smul(mis, mmv);
                        }
                    }
                    // prior table entry [haplotype base b][row]: half hf of an 8-byte V
#pragma unroll
                    for (int b = 0; b < 5; ++b)   // N on either side matches (:11,:21-26)
                        reinterpret_cast<S*>(reinterpret_cast<uint8_t*>(my_tab) + b * SUBT + k * 8)[hf] =
                            (rc == 4 || b == 4 || rc == b) ? mat : mis;
                    // odd K: the last LDS.128 of a column also fetches entry K of the lane's sub-table row -- never used,
                    // but it is a shared-memory read, so it reads a written value
                    if ((K & 1) && k == K - 1) {
#pragma unroll
                        for (int b = 0; b < 5; ++b)
                            reinterpret_cast<S*>(reinterpret_cast<uint8_t*>(my_tab) + b * SUBT + K * 8)[hf] = (S)0;
                    }
                    if (!ALIGNED) P::set(pYY[k], hf, yy);
                    if (!CONSTG) {
                        P::set(pMM[k], hf, mm_);
                        P::set(pGAPM[k], hf, gapm);
                        P::set(pMX[k], hf, mx_);
                        P::set(pMY[k], hf, my_);
                    }
                }
            }
        }
        // row 0 of the lane takes its "cell above" from the shuffle; lane 0 receives its own
        // bottom row back (shfl_up at the group edge), which these two zeros annihilate.
        // PACKED: the first lane of EVERY group stands under the reference's row 0, whatever the lane above
        // (the last lane of another read's group) shuffles down: two more zeroed factors for the diagonal.
        const bool top = (l == 0);
        const V pMX0 = top ? P::splat(0) : pMX[0];
        const V pXX0 = top ? P::splat(0) : (CONSTG ? pXXc : SCALEDG ? pMY[0] : pYY[0]);
        const V pMM0   = (PACKED && top) ? P::splat(0) : pMM[0];
        const V pGAPX0 = (PACKED && top) ? P::splat(0) : pGAPM[0];
        bool dummy_lane[NH];              // ALIGNED: this lane holds only dummy rows of packed read hf
#pragma unroll
        for (int hf = 0; hf < NH; ++hf) dummy_lane[hf] = l * K < pad[hf];

        // ---- haplotypes of this chunk, streamed back to back through ONE wavefront ----
        // Shared memory (per warp): a byte stream
        //     [LEAD x IDLE] hap_0 columns [NEXT] hap_1 columns [NEXT] ... hap_{n-1} columns [NEXT] [LEAD+2 x IDLE]
        // one byte per column = 128-byte offset of the prior sub-table of that column's base.
        // Lane l reads position t + kSkew (G-1-l) at step t, so lanes drop into the next haplotype
        // one after the other while the lanes behind them are still finishing the previous one: the
        // fill/drain bubbles of a wavefront are paid once per chunk, not once per pair.  A NEXT byte
        // makes the lane hand over its result (last lane only) and reset to the column-0 state of
        // the next haplotype; the bottom row it then shuffles down is exactly the (0, 0, y0)
        // boundary the lane below needs.
        constexpr int LEAD = kSkew * (G - 1) + 1;          // idle bytes before / after the haplotypes
        int n = 0, pos = LEAD;
#pragma unroll 1
        for (int h = h_first; h < h_last; ++h) {
            if (!P::kIsF32) {
                const bool w = group_live && needs_redo(args.raw32[out_base + (int64_t)(rd[0] - rd_beg) * nh + h]);
                if (!__any_sync(0xffffffffu, w)) continue;     // nobody in this warp redoes this haplotype
            }
            const int ho = args.hap_off[hap_beg + h];
            const int H  = args.hap_off[hap_beg + h + 1] - ho;
            for (int j = lane; j < H; j += 32)
                sb[pos + j] = (uint8_t)(base_code(args.hap_bases[ho + j]) * (SUBT / 128));
            if (lane == 0) {
                sb[pos + H] = (uint8_t)kStreamNext;
                if (ZRESET) sb[pos + H + 1] = (uint8_t)kStreamNext2;
                s_inity[n] = P::sdiv(P::init_const(), (S)H);   // avx-pairhmm-template.h:86
                if (SCALEDC) {
                    // The scaled state is (M^, X^, Y^) = (s M, X / s2, Y / s2) with s s2 = pMX: the recurrence is
                    // homogeneous, so the split only shows in row 0 (Y^ = Y / s2) and in the final sum.  s2 = pMX would
                    // put row 0 at 2^120 / (H pMX) -- beyond FLT_MAX for the reference's 'I' -- so s2 is the smallest
                    // scale that keeps row 0 at or below 2^126 (2^1022 in FP64): s2 = max(pMX, 1 / (64 H)); M^ then sits
                    // at most ~3 decades lower than the reference's M, well inside the 10 decades between the rescue
                    // threshold and the smallest normal float.
                    const S s2 = max(P::cg(args, 2), P::sdiv(P::kScaleHeadroom(), (S)H));
                    s_inity[n] = P::sdiv(s_inity[n], s2);
                    s_s2[n] = s2;
                    s_invs[n] = P::sdiv(s2, P::cg(args, 2));   // 1 / s
                }
                s_hidx[n] = h; s_apos[n] = pos; s_alen[n] = H;
            }
            pos += H + 1 + (ZRESET ? 1 : 0); ++n;
        }
        if (n == 0) continue;
        for (int j = lane; j < LEAD; j += 32) sb[j] = (uint8_t)kStreamIdle;
        for (int j = lane; j <= LEAD + 1; j += 32) sb[pos + j] = (uint8_t)kStreamIdle;
        __syncwarp();
        asm volatile("" ::: "memory");
        const int p_end = pos - 1;                          // the NEXT byte closing the last haplotype

        // column-0 state (avx-pairhmm-template.h:161-175): M = X = 0; Y = init_Y on row 0 (dummy rows)
        V M[K], X[K], Y[K];
        V Pm[SHARED ? K : 1];             // MODE 2: M * pMX (== M * pMY) of the previous column
        V sumM = P::splat(0), sumX = P::splat(0);
        int jcur = 0;                     // haplotype (stream slot) this lane is in
        auto reset_state = [&](const S init_y) {
#pragma unroll
            for (int k = 0; k < K; ++k) {
                M[k] = P::splat(0); X[k] = P::splat(0);
                if (SHARED) Pm[k] = P::splat(0);
#pragma unroll
                for (int hf = 0; hf < NH; ++hf)
                    P::set(Y[k], hf, (l * K + k < pad[hf]) ? init_y : (S)0);
            }
            sumM = P::splat(0); sumX = P::splat(0);
        };
        S inity_cur = s_inity[0];
        reset_state(inity_cur);
        // values of the row above this lane's first row: at the current column (in*) and at the
        // previous column (dg*).  Lane 0 never uses them (its row 0 is a dummy row with zero
        // priors, pMX0 = pXX0 = 0 and pYY = 1).
        V inM = P::splat(0), inX = P::splat(0), inY = P::shfl_up(Y[K - 1], G);
        if (PACKED) inY = top ? P::splat(inity_cur) : P::splat(0);
        V dgM = inM, dgX = inX, dgY = inY;
        V qM = inM, qX = inX, qY = inY;   // kSkew == 2: bottom row in flight (sent last step, used next step)

        const uint32_t tab_addr = (uint32_t)__cvta_generic_to_shared(my_tab);
        const uint32_t zero_addr = (uint32_t)__cvta_generic_to_shared(stab + tables_bytes(K, G) - kZeroRowBytes);
        (void)zero_addr;
        // products: EXACT ones honour the reference's flush-to-zero in both precisions (P::mulx)
        auto MUL = [](const V a, const V b) { return EXACT ? P::mulx(a, b) : P::mul(a, b); };
        // K cell updates of this lane's current column; `cb` is the column's stream byte
        auto cells = [&](const uint32_t pa, const V fMM, const V fG, const V fYY, const V fMM0, const V fGX0) {
            // priors of this column: K entries of the sub-table the haplotype base names.  Issued
            // first; phase A below (4K FP32-pipe instructions) covers the LDS latency.
            V prior[K + 1];                   // pa: shared address of the lane's row in the sub-table of the column's base
#pragma unroll
            for (int k = 0; k < K; k += 2) lds_2v<V>(pa + 8u * k, prior[k], prior[k + 1]);
            // Phase A: everything that reads the previous column's state (so every old value is dead
            // before it is overwritten: no register copies at the loop back-edge).
            V t0[K];
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const int kk = CONSTG ? 0 : k;
                const V dM = k ? M[k - 1] : dgM;            // (row-1, c-1)
                const V dX = k ? X[k - 1] : dgX;
                const V dY = k ? Y[k - 1] : dgY;
                // ZRESET: the factors come as arguments (uniform in the steady loop, zeroed per lane at the
                // NEXT / NEXT2 / idle bytes of the straddling loop)
                const V cMM = (PACKED && k == 0) ? fMM0 : (ZRESET ? fMM : pMM[kk]);
                const V cGX = (PACKED && k == 0) ? fGX0 : (ZRESET ? fG : pGAPM[kk]);
                const V cGY = ZRESET ? fG : SCALEDG ? pMX[kk] : pGAPM[kk];
                if (EXACT) {
                    // reference operation order, unfused (avx-pairhmm-template.h:188)
                    t0[k] = P::addx(P::addx(MUL(dM, cMM), MUL(dX, cGX)), MUL(dY, cGY));
                } else if (FOLD && SCALEDC) {
                    // (a PACKED group's first lane receives zeros for M and X: see rotate)
                    t0[k] = P::fma(P::add(dX, dY), cGY, dM);
                } else if (FOLD) {
                    t0[k] = P::fma(dY, cGY, P::fma(dX, cGX, dM));
                } else {
                    t0[k] = P::fma(dY, cGY, P::fma(dX, cGX, MUL(dM, cMM)));
                }
            }
            // Y from the left neighbour (:197); needs M of the previous column
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const int kk = CONSTG ? 0 : k;
                const V yv = SCALED ? M[k] : SHARED ? Pm[k] : MUL(M[k], pMY[kk]);     // SCALED: Y^ = M(c-1) + pYY Y^(c-1)
                const V cYY = ALIGNED ? fYY : pYY[ALIGNED ? 0 : k];
                Y[k] = EXACT ? P::addx(yv, MUL(Y[k], cYY)) : P::fma(Y[k], cYY, yv);
            }
            // Phase B: M = t0 * prior (:152-158, :188)
#pragma unroll
            for (int k = 0; k < K; ++k) {
                M[k] = MUL(t0[k], prior[k]);
                if (SHARED) Pm[k] = MUL(M[k], pMX[0]);
            }
            // Phase C: X runs down the column (cell above, :194)
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const int kk = CONSTG ? 0 : k;
                const V cXX = (k == 0) ? pXX0 : (CONSTG ? pXXc : SCALEDG ? pMY[kk] : pYY[ALIGNED ? 0 : k]);
                const V uX = k ? X[k - 1] : inX;            // (row-1, c)
                V um;                                       // M(row-1, c) * pMX(row); SCALED: M(row-1, c) itself
                if (k == 0) um = SCALED ? inM : MUL(inM, pMX0);
                else um = SCALED ? M[k - 1] : SHARED ? Pm[k - 1] : MUL(M[k - 1], pMX[kk]);
                X[k] = EXACT ? P::addx(um, MUL(uX, cXX)) : P::fma(uX, cXX, um);
            }
            // last row of the last lane is the last read row: running sums (:328-343)
            sumM = EXACT ? P::addx(sumM, M[K - 1]) : P::add(sumM, M[K - 1]);
            sumX = EXACT ? P::addx(sumX, X[K - 1]) : P::add(sumX, X[K - 1]);
        };
        // hand the bottom row to the lane below
        auto rotate = [&]() {
            dgM = inM; dgX = inX; dgY = inY;
            // ALIGNED: a dummy lane's Y is garbage; what the lane below must see is row 0's init_Y
            // PACKED: the lane above a group's first lane belongs to another read and runs nl-1 columns
            // behind, so the RECEIVER substitutes its own haplotype's init_Y
            const V outY = (ALIGNED && !PACKED) ? P::sel(dummy_lane[0], dummy_lane[NH - 1], P::splat(inity_cur), Y[K - 1]) : Y[K - 1];
            V rM = P::shfl_up(M[K - 1], G), rX = P::shfl_up(X[K - 1], G), rY = P::shfl_up(outY, G);
            if (PACKED) rY = P::sel(top, top, P::splat(inity_cur), rY);
            // SCALED: there is no product with pMX0 = 0 left to annihilate what the top lane receives (another group's
            // last row, or -- the shuffle has no source for lane 0 -- its own)
            if (SCALED && (PACKED || !ALIGNED)) rM = P::sel(top, top, P::splat(0), rM);
            if (FOLD && PACKED) rX = P::sel(top, top, P::splat(0), rX);      // no zeroed factor left for it either
            if (kSkew == 2) {
                inM = qM; inX = qX; inY = qY;
                qM = rM; qX = rX; qY = rY;
            } else {
                inM = rM; inX = rX; inY = rY;
            }
        };
        // a haplotype ends for this lane
        auto emit = [&]() {
            {
                const int h = s_hidx[jcur];
#pragma unroll
                for (int hf = 0; hf < NH; ++hf) {
                    const int64_t oi = out_base + (int64_t)(rd[hf] - rd_beg) * nh + h;
                    bool w = valid[hf] && group_live;
                    if (!P::kIsF32) w = w && needs_redo(args.raw32[oi]);
                    if (P::kIsF32 && LIST) w = w && __float_as_uint(args.raw32[oi]) == kNeedsF32;    // selective pass
                    if (!w) continue;
                    S res;
                    if (SCALEDC) res = P::sadd(P::smul(P::get(sumM, hf), s_invs[jcur]), P::smul(P::get(sumX, hf), s_s2[jcur]));   // sum(M^) / s + s2 sum(X^)
                    else if (SCALEDG) res = P::sadd(P::get(sumM, hf), P::smul(P::get(sumX, hf), ph2pr[args.read_i[args.read_off[rd[hf] + 1] - 1] & 127]));
                    else res = P::sadd(P::get(sumM, hf), P::get(sumX, hf));
                    if (P::kIsF32) {
                        args.raw32[oi] = (float)res;
                        if ((float)res < kMinAccepted) { if (LIST) flag_or(my_flag, kFlagRedo64); else *my_flag = (uint8_t)kFlagRedo64; }
                    } else if (args.tier == 0 && !((double)res < kCertainUnderflow64)) {
                        // FP64 first: the FP32 pass may NOT underflow here -- it has to be run to find out
                        args.raw32[oi] = __uint_as_float(kNeedsF32);
                        flag_or(my_flag, kFlagNeedsF32);
                    } else if (!EXACT && (double)res < kFlushDanger) {
                        // tier 3 will redo this pair with flushed products (FP64 first: its FP32 sum is a proven 0)
                        args.raw32[oi] = __uint_as_float((args.tier == 0 ? 0u : __float_as_uint(args.raw32[oi])) | 0x80000000u);
                        flag_or(my_flag, kFlagFlush64);
                    } else {
                        if (args.tier == 0) args.raw32[oi] = 0.0f;     // proven < 1e-28f, never computed
                        const unsigned slot = atomicAdd(args.rescue_count, 1u);
                        args.rescue_out[slot].out_idx = oi;
                        args.rescue_out[slot].raw64 = (double)res;
                    }
                }
            }
        };
        auto boundary = [&]() {
            if (l == nl - 1) emit();
            ++jcur;
            inity_cur = s_inity[min(jcur, n - 1)];
            reset_state(inity_cur);
            // PACKED: a group's first lane enters the next haplotype -- row 0 above it changes with it
            if (PACKED && top) { inY = P::splat(inity_cur); if (kSkew == 2) qY = inY; }
        };

        // shared address of this lane's byte at step t is bp + t
        const uint32_t bp = (uint32_t)__cvta_generic_to_shared(sb) + (uint32_t)(kSkew * (nl - 1 - l));
        int t = PACKED ? LEAD - 1 - kSkew * (nl - 1) : 0;   // PACKED: skip the steps in which no lane has work yet
        uint32_t b_next = lds_u8(bp + (uint32_t)t);
#pragma unroll 1
        for (int j = 0; j <= n; ++j) {
            // [t_a, t_s): every lane of the group is inside haplotype j -> no tests at all
            const int t_a = (j < n) ? s_apos[j] : p_end + 1;
            const int t_s = (j < n) ? t_a + s_alen[j] - kSkew * (nl - 1) : t_a;
            // lanes straddle two haplotypes (or the ends of the stream)
#pragma unroll 1
            for (; t < t_a; ++t) {
                const uint32_t bcur = b_next;
                b_next = lds_u8(bp + (uint32_t)(t + 1));
                if (ZRESET) {
                    const bool zs = bcur >= 0x80u;                  // NEXT, NEXT2 or idle: a zero step
                    if (bcur == (uint32_t)kStreamNext) {            // this lane has seen the last column of its haplotype
                        if (l == nl - 1) emit();
                        ++jcur;
                        inity_cur = s_inity[min(jcur, n - 1)];
                        sumM = P::splat(0); sumX = P::splat(0);
                        if (PACKED && top) { inY = P::splat(inity_cur); if (kSkew == 2) qY = inY; }
                    }
                    const V z = P::splat(0);
                    // a zero step reads the lane's own 'A' sub-table: any FINITE priors do (0 * prior must be 0)
                    // FOLD: M = prior' * (...) has no factor left to zero -- the zero step reads a row of zero priors
                    const uint32_t pa = (FOLD && zs) ? zero_addr : tab_addr + ((zs ? 0u : bcur) << 7);
                    cells(pa, zs ? z : pMM[0], zs ? z : pGAPM[0], zs ? z : pXXc, zs ? z : pMM0, zs ? z : pGAPX0);
                } else {
                    if (bcur < 0x80u) cells(tab_addr + (bcur << 7), pMM[0], pGAPM[0], pXXc, pMM0, pGAPX0);
                    else if (bcur == (uint32_t)kStreamNext) boundary();
                }
                rotate();
            }
#pragma unroll 4
            for (; t < t_s; ++t) {
                const uint32_t bcur = b_next;
                b_next = lds_u8(bp + (uint32_t)(t + 1));
                cells(tab_addr + (bcur << 7), pMM[0], pGAPM[0], pXXc, pMM0, pGAPX0);
                rotate();
            }
        }
    }
    if constexpr (!LIST) return;
    }   // for (;;)
}

}  // namespace phmm
