// phmm_launch.h -- lookup of the compiled forward_kernel instantiations.
#pragma once
#include "phmm_kernels.cuh"

namespace phmm {

using KernelFn = void (*)(const KernelArgs);

constexpr int kMaxRowsPerLane = 8;     // K = 1..8
constexpr int kGroupWidth = 32;        // G (lanes per pair wavefront) compiled in this round

// One translation unit per (precision, exact) keeps nvcc parallel; each fills its slice.
void register_f32_fast(KernelFn (*tab)[kMaxRowsPerLane]);
void register_f32_exact(KernelFn (*tab)[kMaxRowsPerLane]);
void register_f64_fast(KernelFn (*tab)[kMaxRowsPerLane]);
void register_f64_exact(KernelFn (*tab)[kMaxRowsPerLane]);

// tab[mode][K-1]   (mode: 0 general per-base gaps, 1 batch-constant gaps, 2 constant with i == d)
#define PHMM_REGISTER_MODE(POLICY, EXACT, MODE)                                                  \
    do {                                                                                         \
        tab[MODE][0] = forward_kernel<POLICY, 1, kGroupWidth, MODE, EXACT>;                      \
        tab[MODE][1] = forward_kernel<POLICY, 2, kGroupWidth, MODE, EXACT>;                      \
        tab[MODE][2] = forward_kernel<POLICY, 3, kGroupWidth, MODE, EXACT>;                      \
        tab[MODE][3] = forward_kernel<POLICY, 4, kGroupWidth, MODE, EXACT>;                      \
        tab[MODE][4] = forward_kernel<POLICY, 5, kGroupWidth, MODE, EXACT>;                      \
        tab[MODE][5] = forward_kernel<POLICY, 6, kGroupWidth, MODE, EXACT>;                      \
        tab[MODE][6] = forward_kernel<POLICY, 7, kGroupWidth, MODE, EXACT>;                      \
        tab[MODE][7] = forward_kernel<POLICY, 8, kGroupWidth, MODE, EXACT>;                      \
    } while (0)
#define PHMM_REGISTER_ALL(POLICY, EXACT)                                                         \
    do {                                                                                         \
        PHMM_REGISTER_MODE(POLICY, EXACT, 0);                                                    \
        PHMM_REGISTER_MODE(POLICY, EXACT, 1);                                                    \
        PHMM_REGISTER_MODE(POLICY, EXACT, 2);                                                    \
    } while (0)

}  // namespace phmm
