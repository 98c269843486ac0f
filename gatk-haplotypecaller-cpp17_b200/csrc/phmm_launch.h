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
constexpr int kNumShapes = 18;
constexpr Shape kShapes[kNumShapes] = {
    {32, 1}, {32, 2}, {32, 3}, {32, 4}, {32, 5}, {32, 6}, {32, 7}, {32, 8},
    {16, 1}, {16, 2}, {16, 3}, {16, 4}, {16, 5}, {16, 6}, {16, 7}, {16, 8}, {16, 9}, {16, 10}};
constexpr int kMaxReadLenCompiled = 32 * 8 - 1;   // 255

constexpr int kNumModes = 3;
// tab[mode][shape]
using KernelTab = KernelFn[kNumModes][kNumShapes];

// One translation unit per (precision, exact) keeps nvcc parallel; each fills its slice.
void register_f32_fast(KernelTab& tab);
void register_f32_exact(KernelTab& tab);
void register_f64_fast(KernelTab& tab);
void register_f64_exact(KernelTab& tab);

#define PHMM_REGISTER_MODE(POLICY, EXACT, MODE)                                                  \
    do {                                                                                         \
        tab[MODE][0]  = forward_kernel<POLICY, 1, 32, MODE, EXACT>;                              \
        tab[MODE][1]  = forward_kernel<POLICY, 2, 32, MODE, EXACT>;                              \
        tab[MODE][2]  = forward_kernel<POLICY, 3, 32, MODE, EXACT>;                              \
        tab[MODE][3]  = forward_kernel<POLICY, 4, 32, MODE, EXACT>;                              \
        tab[MODE][4]  = forward_kernel<POLICY, 5, 32, MODE, EXACT>;                              \
        tab[MODE][5]  = forward_kernel<POLICY, 6, 32, MODE, EXACT>;                              \
        tab[MODE][6]  = forward_kernel<POLICY, 7, 32, MODE, EXACT>;                              \
        tab[MODE][7]  = forward_kernel<POLICY, 8, 32, MODE, EXACT>;                              \
        tab[MODE][8]  = forward_kernel<POLICY, 1, 16, MODE, EXACT>;                              \
        tab[MODE][9]  = forward_kernel<POLICY, 2, 16, MODE, EXACT>;                              \
        tab[MODE][10] = forward_kernel<POLICY, 3, 16, MODE, EXACT>;                              \
        tab[MODE][11] = forward_kernel<POLICY, 4, 16, MODE, EXACT>;                              \
        tab[MODE][12] = forward_kernel<POLICY, 5, 16, MODE, EXACT>;                              \
        tab[MODE][13] = forward_kernel<POLICY, 6, 16, MODE, EXACT>;                              \
        tab[MODE][14] = forward_kernel<POLICY, 7, 16, MODE, EXACT>;                              \
        tab[MODE][15] = forward_kernel<POLICY, 8, 16, MODE, EXACT>;                              \
        tab[MODE][16] = forward_kernel<POLICY, 9, 16, MODE, EXACT>;                              \
        tab[MODE][17] = forward_kernel<POLICY, 10, 16, MODE, EXACT>;                             \
    } while (0)
#define PHMM_REGISTER_ALL(POLICY, EXACT)                                                         \
    do {                                                                                         \
        PHMM_REGISTER_MODE(POLICY, EXACT, 0);                                                    \
        PHMM_REGISTER_MODE(POLICY, EXACT, 1);                                                    \
        PHMM_REGISTER_MODE(POLICY, EXACT, 2);                                                    \
    } while (0)

}  // namespace phmm
