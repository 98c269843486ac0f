// phmm_launch.h -- the compiled forward_kernel instantiations and how the planner names them.
#pragma once
#include <utility>
#include "phmm_kernels.cuh"

namespace phmm {

using KernelFn = void (*)(const KernelArgs);

// A lane-group shape: G lanes per wavefront, K read rows per lane; scores reads of up to K*G-1
// bases in one pass (one dummy row on top is mandatory, see phmm_kernels.cuh).
struct Shape { int G, K; };

// Shapes compiled.  Narrow groups (G = 16) put more rows in a lane -- fewer fill/drain steps, more FP32 work per
// shuffle and loop instruction -- and win the planner's cost model G (15 K + 10) (H + G - 1) for every read they can
// hold (R <= 159); the wide groups (G = 32) are only ever picked for reads of 160..255 bases, so only K = 6, 7, 8 of
// them exist (the five smaller ones were a quarter of the fatbin and never launched: shape histogram over S2..S5 and
// the chrM-like contig).  The last three are the PACKED shapes (free-width lane groups on the whole warp,
// phmm_kernels.cuh): they only exist as "aligned" kernels of the constant-gap modes and are never picked by read
// length alone.
constexpr int kNumShapes = 16;
constexpr int kFirstPackedShape = 13;
constexpr Shape kShapes[kNumShapes] = {
    {16, 1}, {16, 2}, {16, 3}, {16, 4}, {16, 5}, {16, 6}, {16, 7}, {16, 8}, {16, 9}, {16, 10},
    {32, 6}, {32, 7}, {32, 8},
    {32, 8}, {32, 9}, {32, 10}};
constexpr int kMaxReadLenCompiled = 32 * 8 - 1;   // 255

// Reads longer than kMaxReadLenCompiled take the one-warp-per-pair kernel of phmm_long.cu.
constexpr int kLongMaxRead = 2048;           // == PHMM_MAX_READ_LEN (include/phmm.h)
constexpr int kLongWarpsPerCta = 2;
struct LongPair { int32_t read, hap; int64_t out_idx; };   // read / haplotype: indices into the part's offset arrays
void launch_long_reads(const KernelArgs& args, const LongPair* pairs, int n_pairs, bool general, bool exact, bool use_double, cudaStream_t st);

// phmm_finalize.cu: raw FP32 sums -> final float log10 values + counters, on the device (glibc's log10f restated)
void launch_finalize(const float* raw32, int64_t n_pairs, float log10_init_f, float* lik32, unsigned* header, int sm_count,
                     cudaStream_t st);

// phmm_genotype.cu: cap / filter + per-site genotype likelihoods on the device; device pointers
struct GenotypeArgs {
    double* lik64;                       // [n_pairs] the capped double matrix (scratch; downloadable)
    int64_t n_pairs;
    int32_t n_regions, n_reads, n_sites;
    const int32_t* region_read_beg; const int32_t* region_hap_beg; const int64_t* region_out_beg; const int32_t* read_off;
    uint8_t* read_keep;                  // [n_reads] out: 0 = poorly modelled (intel_pairhmm.hpp:35-38)
    const int32_t* site_region; const int32_t* site_n_alleles;
    const int64_t* site_hap_off; const uint8_t* hap_allele;          // haplotype -> allele, per site
    const int64_t* site_read_off; const uint8_t* read_overlap;       // read overlaps the site, per site (nullptr: all do)
    const int64_t* gl_off;
    double* scratch_al; uint8_t* scratch_used;                        // [site_read_off[n_sites]] x 8 doubles / bytes
    double* gl; int32_t* site_n_used;    // out
    const double* jacobian; double inv_step, log10_2;
};
void launch_genotype(const GenotypeArgs& g, const float* lik32, const RescueOut* rescue, const unsigned* rescue_count,
                     double log10_init_d, int sm_count, cudaStream_t st);

constexpr int kNumModes = 5;      // 3 = the scaled recurrence (fast engines; the exact kernels of that slot are the mode-2 ones)
// tab[mode][aligned][shape][list]; aligned = every read length of the job is a multiple of K (constant-gap
// modes only; the general mode has no aligned variant and its [1] row repeats [0]); list = the launch pulls its
// units from a work list (phmm_kernels.cuh: LIST) -- compiled for the fast engines only (nullptr otherwise: the
// exact engines walk the grid).
using KernelTab = KernelFn[kNumModes][2][kNumShapes][2];

// One translation unit per (precision, exact) keeps nvcc parallel; each fills its slice.
void register_f32_fast(KernelTab& tab);
void register_f32_exact(KernelTab& tab);
void register_f64_fast(KernelTab& tab);
void register_f64_exact(KernelTab& tab);

template <class P, bool EXACT, bool LIST, int S>
inline void register_shape(KernelTab& tab)
{
    constexpr int G = kShapes[S].G, K = kShapes[S].K;
    constexpr int L = LIST ? 1 : 0;
    if constexpr (S >= kFirstPackedShape) {          // PACKED: lane-aligned kernels of the constant-gap modes only
        tab[0][0][S][L] = nullptr; tab[0][1][S][L] = nullptr; tab[1][0][S][L] = nullptr; tab[2][0][S][L] = nullptr; tab[3][0][S][L] = nullptr;
        tab[4][0][S][L] = nullptr; tab[4][1][S][L] = nullptr;
        tab[1][1][S][L] = forward_kernel<P, K, G, 1, EXACT, true, true, LIST>;
        tab[2][1][S][L] = forward_kernel<P, K, G, 2, EXACT, true, true, LIST>;
        if constexpr (EXACT) tab[3][1][S][L] = tab[2][1][S][L];
        else tab[3][1][S][L] = forward_kernel<P, K, G, 3, EXACT, true, true, LIST>;
    } else {
        if constexpr (EXACT) {                       // (the exact engines never select modes 3 / 4; tier 3 of such a part does)
            tab[3][0][S][L] = forward_kernel<P, K, G, 2, EXACT, false, false, LIST>;
            tab[3][1][S][L] = forward_kernel<P, K, G, 2, EXACT, true, false, LIST>;
            tab[4][0][S][L] = tab[4][1][S][L] = forward_kernel<P, K, G, 0, EXACT, false, false, LIST>;
        } else {
            tab[3][0][S][L] = forward_kernel<P, K, G, 3, EXACT, false, false, LIST>;
            tab[3][1][S][L] = forward_kernel<P, K, G, 3, EXACT, true, false, LIST>;
            tab[4][0][S][L] = tab[4][1][S][L] = forward_kernel<P, K, G, 4, EXACT, false, false, LIST>;
        }
        tab[0][0][S][L] = tab[0][1][S][L] = forward_kernel<P, K, G, 0, EXACT, false, false, LIST>;
        tab[1][0][S][L] = forward_kernel<P, K, G, 1, EXACT, false, false, LIST>;
        tab[1][1][S][L] = forward_kernel<P, K, G, 1, EXACT, true, false, LIST>;
        tab[2][0][S][L] = forward_kernel<P, K, G, 2, EXACT, false, false, LIST>;
        tab[2][1][S][L] = forward_kernel<P, K, G, 2, EXACT, true, false, LIST>;
    }
}
template <class P, bool EXACT, bool LIST, int... S>
inline void register_shapes(KernelTab& tab, std::integer_sequence<int, S...>) { (register_shape<P, EXACT, LIST, S>(tab), ...); }

// every Shape x MODE x aligned of one (precision, exact); WITH_LIST adds the work-list variants
template <class P, bool EXACT, bool WITH_LIST>
inline void register_all(KernelTab& tab)
{
    for (int m = 0; m < kNumModes; m++)
        for (int a = 0; a < 2; a++)
            for (int s = 0; s < kNumShapes; s++) tab[m][a][s][0] = tab[m][a][s][1] = nullptr;
    register_shapes<P, EXACT, false>(tab, std::make_integer_sequence<int, kNumShapes>{});
    if constexpr (WITH_LIST) register_shapes<P, EXACT, true>(tab, std::make_integer_sequence<int, kNumShapes>{});
}

}  // namespace phmm
