// phmm_launch.h -- the compiled forward_kernel instantiations and how the planner names them.
#pragma once
#include "phmm_kernels.cuh"

namespace phmm {

using KernelFn = void (*)(const KernelArgs);

// A lane-group shape: G lanes per wavefront, K read rows per lane; scores reads of up to K*G-1
// bases in one pass (one dummy row on top is mandatory, see phmm_kernels.cuh).
struct Shape { int G, K; };

// Shapes compiled in this round.  Wide groups (G=32) keep registers low; narrow groups (G=16) put
// more rows in a lane: fewer fill/drain steps and more FP32 work per shuffle/loop instruction.
// The last three are the PACKED shapes (free-width lane groups on the whole warp, phmm_kernels.cuh): they
// only exist as "aligned" kernels of the constant-gap modes and are never picked by read length alone.
constexpr int kNumShapes = 21;
constexpr int kFirstPackedShape = 18;
constexpr Shape kShapes[kNumShapes] = {
    {32, 1}, {32, 2}, {32, 3}, {32, 4}, {32, 5}, {32, 6}, {32, 7}, {32, 8},
    {16, 1}, {16, 2}, {16, 3}, {16, 4}, {16, 5}, {16, 6}, {16, 7}, {16, 8}, {16, 9}, {16, 10},
    {32, 8}, {32, 9}, {32, 10}};
constexpr int kMaxReadLenCompiled = 32 * 8 - 1;   // 255

// Reads longer than kMaxReadLenCompiled take the one-warp-per-pair kernel of phmm_long.cu.
constexpr int kLongMaxRead = 2048;           // == PHMM_MAX_READ_LEN (include/phmm.h)
constexpr int kLongWarpsPerCta = 2;
struct LongPair { int32_t read, hap; int64_t out_idx; };   // read / haplotype: indices into the part's offset arrays
void launch_long_reads(const KernelArgs& args, const LongPair* pairs, int n_pairs, bool general, bool exact, cudaStream_t st);

constexpr int kNumModes = 3;
// tab[mode][aligned][shape]; aligned = every read length of the job is a multiple of K (constant-gap
// modes only; the general mode has no aligned variant and its [1] row repeats [0])
using KernelTab = KernelFn[kNumModes][2][kNumShapes];

// One translation unit per (precision, exact) keeps nvcc parallel; each fills its slice.
void register_f32_fast(KernelTab& tab);
void register_f32_exact(KernelTab& tab);
void register_f64_fast(KernelTab& tab);
void register_f64_exact(KernelTab& tab);

#define PHMM_REGISTER_ALL(POLICY, EXACT)                                                         \
    do {                                                                                         \
        for (int m_ = 0; m_ < kNumModes; m_++)                                                   \
            for (int a_ = 0; a_ < 2; a_++)                                                       \
                for (int s_ = kFirstPackedShape; s_ < kNumShapes; s_++) tab[m_][a_][s_] = nullptr; \
        tab[1][1][18] = forward_kernel<POLICY, 8, 32, 1, EXACT, true, true>;                     \
        tab[1][1][19] = forward_kernel<POLICY, 9, 32, 1, EXACT, true, true>;                     \
        tab[1][1][20] = forward_kernel<POLICY, 10, 32, 1, EXACT, true, true>;                    \
        tab[2][1][18] = forward_kernel<POLICY, 8, 32, 2, EXACT, true, true>;                     \
        tab[2][1][19] = forward_kernel<POLICY, 9, 32, 2, EXACT, true, true>;                     \
        tab[2][1][20] = forward_kernel<POLICY, 10, 32, 2, EXACT, true, true>;                    \
        tab[0][0][0] = forward_kernel<POLICY, 1, 32, 0, EXACT, false>;                           \
        tab[0][0][1] = forward_kernel<POLICY, 2, 32, 0, EXACT, false>;                           \
        tab[0][0][2] = forward_kernel<POLICY, 3, 32, 0, EXACT, false>;                           \
        tab[0][0][3] = forward_kernel<POLICY, 4, 32, 0, EXACT, false>;                           \
        tab[0][0][4] = forward_kernel<POLICY, 5, 32, 0, EXACT, false>;                           \
        tab[0][0][5] = forward_kernel<POLICY, 6, 32, 0, EXACT, false>;                           \
        tab[0][0][6] = forward_kernel<POLICY, 7, 32, 0, EXACT, false>;                           \
        tab[0][0][7] = forward_kernel<POLICY, 8, 32, 0, EXACT, false>;                           \
        tab[0][0][8] = forward_kernel<POLICY, 1, 16, 0, EXACT, false>;                           \
        tab[0][0][9] = forward_kernel<POLICY, 2, 16, 0, EXACT, false>;                           \
        tab[0][0][10] = forward_kernel<POLICY, 3, 16, 0, EXACT, false>;                          \
        tab[0][0][11] = forward_kernel<POLICY, 4, 16, 0, EXACT, false>;                          \
        tab[0][0][12] = forward_kernel<POLICY, 5, 16, 0, EXACT, false>;                          \
        tab[0][0][13] = forward_kernel<POLICY, 6, 16, 0, EXACT, false>;                          \
        tab[0][0][14] = forward_kernel<POLICY, 7, 16, 0, EXACT, false>;                          \
        tab[0][0][15] = forward_kernel<POLICY, 8, 16, 0, EXACT, false>;                          \
        tab[0][0][16] = forward_kernel<POLICY, 9, 16, 0, EXACT, false>;                          \
        tab[0][0][17] = forward_kernel<POLICY, 10, 16, 0, EXACT, false>;                         \
        tab[0][1][0] = forward_kernel<POLICY, 1, 32, 0, EXACT, false>;                           \
        tab[0][1][1] = forward_kernel<POLICY, 2, 32, 0, EXACT, false>;                           \
        tab[0][1][2] = forward_kernel<POLICY, 3, 32, 0, EXACT, false>;                           \
        tab[0][1][3] = forward_kernel<POLICY, 4, 32, 0, EXACT, false>;                           \
        tab[0][1][4] = forward_kernel<POLICY, 5, 32, 0, EXACT, false>;                           \
        tab[0][1][5] = forward_kernel<POLICY, 6, 32, 0, EXACT, false>;                           \
        tab[0][1][6] = forward_kernel<POLICY, 7, 32, 0, EXACT, false>;                           \
        tab[0][1][7] = forward_kernel<POLICY, 8, 32, 0, EXACT, false>;                           \
        tab[0][1][8] = forward_kernel<POLICY, 1, 16, 0, EXACT, false>;                           \
        tab[0][1][9] = forward_kernel<POLICY, 2, 16, 0, EXACT, false>;                           \
        tab[0][1][10] = forward_kernel<POLICY, 3, 16, 0, EXACT, false>;                          \
        tab[0][1][11] = forward_kernel<POLICY, 4, 16, 0, EXACT, false>;                          \
        tab[0][1][12] = forward_kernel<POLICY, 5, 16, 0, EXACT, false>;                          \
        tab[0][1][13] = forward_kernel<POLICY, 6, 16, 0, EXACT, false>;                          \
        tab[0][1][14] = forward_kernel<POLICY, 7, 16, 0, EXACT, false>;                          \
        tab[0][1][15] = forward_kernel<POLICY, 8, 16, 0, EXACT, false>;                          \
        tab[0][1][16] = forward_kernel<POLICY, 9, 16, 0, EXACT, false>;                          \
        tab[0][1][17] = forward_kernel<POLICY, 10, 16, 0, EXACT, false>;                         \
        tab[1][0][0] = forward_kernel<POLICY, 1, 32, 1, EXACT, false>;                           \
        tab[1][0][1] = forward_kernel<POLICY, 2, 32, 1, EXACT, false>;                           \
        tab[1][0][2] = forward_kernel<POLICY, 3, 32, 1, EXACT, false>;                           \
        tab[1][0][3] = forward_kernel<POLICY, 4, 32, 1, EXACT, false>;                           \
        tab[1][0][4] = forward_kernel<POLICY, 5, 32, 1, EXACT, false>;                           \
        tab[1][0][5] = forward_kernel<POLICY, 6, 32, 1, EXACT, false>;                           \
        tab[1][0][6] = forward_kernel<POLICY, 7, 32, 1, EXACT, false>;                           \
        tab[1][0][7] = forward_kernel<POLICY, 8, 32, 1, EXACT, false>;                           \
        tab[1][0][8] = forward_kernel<POLICY, 1, 16, 1, EXACT, false>;                           \
        tab[1][0][9] = forward_kernel<POLICY, 2, 16, 1, EXACT, false>;                           \
        tab[1][0][10] = forward_kernel<POLICY, 3, 16, 1, EXACT, false>;                          \
        tab[1][0][11] = forward_kernel<POLICY, 4, 16, 1, EXACT, false>;                          \
        tab[1][0][12] = forward_kernel<POLICY, 5, 16, 1, EXACT, false>;                          \
        tab[1][0][13] = forward_kernel<POLICY, 6, 16, 1, EXACT, false>;                          \
        tab[1][0][14] = forward_kernel<POLICY, 7, 16, 1, EXACT, false>;                          \
        tab[1][0][15] = forward_kernel<POLICY, 8, 16, 1, EXACT, false>;                          \
        tab[1][0][16] = forward_kernel<POLICY, 9, 16, 1, EXACT, false>;                          \
        tab[1][0][17] = forward_kernel<POLICY, 10, 16, 1, EXACT, false>;                         \
        tab[1][1][0] = forward_kernel<POLICY, 1, 32, 1, EXACT, true>;                            \
        tab[1][1][1] = forward_kernel<POLICY, 2, 32, 1, EXACT, true>;                            \
        tab[1][1][2] = forward_kernel<POLICY, 3, 32, 1, EXACT, true>;                            \
        tab[1][1][3] = forward_kernel<POLICY, 4, 32, 1, EXACT, true>;                            \
        tab[1][1][4] = forward_kernel<POLICY, 5, 32, 1, EXACT, true>;                            \
        tab[1][1][5] = forward_kernel<POLICY, 6, 32, 1, EXACT, true>;                            \
        tab[1][1][6] = forward_kernel<POLICY, 7, 32, 1, EXACT, true>;                            \
        tab[1][1][7] = forward_kernel<POLICY, 8, 32, 1, EXACT, true>;                            \
        tab[1][1][8] = forward_kernel<POLICY, 1, 16, 1, EXACT, true>;                            \
        tab[1][1][9] = forward_kernel<POLICY, 2, 16, 1, EXACT, true>;                            \
        tab[1][1][10] = forward_kernel<POLICY, 3, 16, 1, EXACT, true>;                           \
        tab[1][1][11] = forward_kernel<POLICY, 4, 16, 1, EXACT, true>;                           \
        tab[1][1][12] = forward_kernel<POLICY, 5, 16, 1, EXACT, true>;                           \
        tab[1][1][13] = forward_kernel<POLICY, 6, 16, 1, EXACT, true>;                           \
        tab[1][1][14] = forward_kernel<POLICY, 7, 16, 1, EXACT, true>;                           \
        tab[1][1][15] = forward_kernel<POLICY, 8, 16, 1, EXACT, true>;                           \
        tab[1][1][16] = forward_kernel<POLICY, 9, 16, 1, EXACT, true>;                           \
        tab[1][1][17] = forward_kernel<POLICY, 10, 16, 1, EXACT, true>;                          \
        tab[2][0][0] = forward_kernel<POLICY, 1, 32, 2, EXACT, false>;                           \
        tab[2][0][1] = forward_kernel<POLICY, 2, 32, 2, EXACT, false>;                           \
        tab[2][0][2] = forward_kernel<POLICY, 3, 32, 2, EXACT, false>;                           \
        tab[2][0][3] = forward_kernel<POLICY, 4, 32, 2, EXACT, false>;                           \
        tab[2][0][4] = forward_kernel<POLICY, 5, 32, 2, EXACT, false>;                           \
        tab[2][0][5] = forward_kernel<POLICY, 6, 32, 2, EXACT, false>;                           \
        tab[2][0][6] = forward_kernel<POLICY, 7, 32, 2, EXACT, false>;                           \
        tab[2][0][7] = forward_kernel<POLICY, 8, 32, 2, EXACT, false>;                           \
        tab[2][0][8] = forward_kernel<POLICY, 1, 16, 2, EXACT, false>;                           \
        tab[2][0][9] = forward_kernel<POLICY, 2, 16, 2, EXACT, false>;                           \
        tab[2][0][10] = forward_kernel<POLICY, 3, 16, 2, EXACT, false>;                          \
        tab[2][0][11] = forward_kernel<POLICY, 4, 16, 2, EXACT, false>;                          \
        tab[2][0][12] = forward_kernel<POLICY, 5, 16, 2, EXACT, false>;                          \
        tab[2][0][13] = forward_kernel<POLICY, 6, 16, 2, EXACT, false>;                          \
        tab[2][0][14] = forward_kernel<POLICY, 7, 16, 2, EXACT, false>;                          \
        tab[2][0][15] = forward_kernel<POLICY, 8, 16, 2, EXACT, false>;                          \
        tab[2][0][16] = forward_kernel<POLICY, 9, 16, 2, EXACT, false>;                          \
        tab[2][0][17] = forward_kernel<POLICY, 10, 16, 2, EXACT, false>;                         \
        tab[2][1][0] = forward_kernel<POLICY, 1, 32, 2, EXACT, true>;                            \
        tab[2][1][1] = forward_kernel<POLICY, 2, 32, 2, EXACT, true>;                            \
        tab[2][1][2] = forward_kernel<POLICY, 3, 32, 2, EXACT, true>;                            \
        tab[2][1][3] = forward_kernel<POLICY, 4, 32, 2, EXACT, true>;                            \
        tab[2][1][4] = forward_kernel<POLICY, 5, 32, 2, EXACT, true>;                            \
        tab[2][1][5] = forward_kernel<POLICY, 6, 32, 2, EXACT, true>;                            \
        tab[2][1][6] = forward_kernel<POLICY, 7, 32, 2, EXACT, true>;                            \
        tab[2][1][7] = forward_kernel<POLICY, 8, 32, 2, EXACT, true>;                            \
        tab[2][1][8] = forward_kernel<POLICY, 1, 16, 2, EXACT, true>;                            \
        tab[2][1][9] = forward_kernel<POLICY, 2, 16, 2, EXACT, true>;                            \
        tab[2][1][10] = forward_kernel<POLICY, 3, 16, 2, EXACT, true>;                           \
        tab[2][1][11] = forward_kernel<POLICY, 4, 16, 2, EXACT, true>;                           \
        tab[2][1][12] = forward_kernel<POLICY, 5, 16, 2, EXACT, true>;                           \
        tab[2][1][13] = forward_kernel<POLICY, 6, 16, 2, EXACT, true>;                           \
        tab[2][1][14] = forward_kernel<POLICY, 7, 16, 2, EXACT, true>;                           \
        tab[2][1][15] = forward_kernel<POLICY, 8, 16, 2, EXACT, true>;                           \
        tab[2][1][16] = forward_kernel<POLICY, 9, 16, 2, EXACT, true>;                           \
        tab[2][1][17] = forward_kernel<POLICY, 10, 16, 2, EXACT, true>;                          \
    } while (0)

}  // namespace phmm
