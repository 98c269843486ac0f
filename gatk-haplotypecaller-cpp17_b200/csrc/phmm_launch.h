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

// tab[uniform][K-1]
#define PHMM_REGISTER_ALL(POLICY, EXACT)                                                         \
    do {                                                                                         \
        tab[0][0] = forward_kernel<POLICY, 1, kGroupWidth, false, EXACT>;                        \
        tab[0][1] = forward_kernel<POLICY, 2, kGroupWidth, false, EXACT>;                        \
        tab[0][2] = forward_kernel<POLICY, 3, kGroupWidth, false, EXACT>;                        \
        tab[0][3] = forward_kernel<POLICY, 4, kGroupWidth, false, EXACT>;                        \
        tab[0][4] = forward_kernel<POLICY, 5, kGroupWidth, false, EXACT>;                        \
        tab[0][5] = forward_kernel<POLICY, 6, kGroupWidth, false, EXACT>;                        \
        tab[0][6] = forward_kernel<POLICY, 7, kGroupWidth, false, EXACT>;                        \
        tab[0][7] = forward_kernel<POLICY, 8, kGroupWidth, false, EXACT>;                        \
        tab[1][0] = forward_kernel<POLICY, 1, kGroupWidth, true, EXACT>;                         \
        tab[1][1] = forward_kernel<POLICY, 2, kGroupWidth, true, EXACT>;                         \
        tab[1][2] = forward_kernel<POLICY, 3, kGroupWidth, true, EXACT>;                         \
        tab[1][3] = forward_kernel<POLICY, 4, kGroupWidth, true, EXACT>;                         \
        tab[1][4] = forward_kernel<POLICY, 5, kGroupWidth, true, EXACT>;                         \
        tab[1][5] = forward_kernel<POLICY, 6, kGroupWidth, true, EXACT>;                         \
        tab[1][6] = forward_kernel<POLICY, 7, kGroupWidth, true, EXACT>;                         \
        tab[1][7] = forward_kernel<POLICY, 8, kGroupWidth, true, EXACT>;                         \
    } while (0)

}  // namespace phmm
