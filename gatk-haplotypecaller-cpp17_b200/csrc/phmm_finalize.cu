// phmm_finalize.cu -- what happens to the raw forward sums ON THE DEVICE.
//
//   finalize_kernel      raw FP32 sum -> (float)(log10f(raw) - log10f(2^120)), the reference's own float
//                        arithmetic (intel_pairhmm.hpp:142) with glibc's log10f restated operation for operation
//                        (phmm_log10.h: bit-identical, checked exhaustively), plus the counters the host used to
//                        take from a pass over all raw sums (pairs below MIN_ACCEPTED, pairs marked for the
//                        flush-exact tier, pairs left unscored).  The host then only widens floats.
//
// The genotype-likelihood reduction builds on the same values; see phmm_genotype.cu.
#include "phmm_launch.h"
#include "phmm_log10.h"

namespace phmm {

namespace {

__global__ void __launch_bounds__(256)
finalize_kernel(const float* __restrict__ raw32, const int64_t n_pairs, const float log10_init_f,
                float* __restrict__ lik32, unsigned* __restrict__ header)
{
    // header: {rescue_count (written by the FP64 kernels), marked, underflowed, unscored}
    unsigned marked = 0, under = 0, unscored = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_pairs; i += (int64_t)gridDim.x * blockDim.x) {
        const float f = raw32[i];
        float out;
        if (f != f) { unscored++; out = f; }
        else if (f < kMinAccepted) {                      // intel_pairhmm.hpp:137: the FP64 result replaces it (host side)
            under++; marked += __float_as_uint(f) >> 31;
            out = __uint_as_float(0x7fc00000u);
        } else {
            out = __fsub_rn(glibc_log10f(f), log10_init_f);   // float subtraction, :142
        }
        lik32[i] = out;
    }
    // block-wide sums, one atomic per counter and block (most blocks add nothing)
    __shared__ unsigned sh[3];
    if (threadIdx.x < 3) sh[threadIdx.x] = 0;
    __syncthreads();
    if (marked) atomicAdd(&sh[0], marked);
    if (under) atomicAdd(&sh[1], under);
    if (unscored) atomicAdd(&sh[2], unscored);
    __syncthreads();
    if (threadIdx.x < 3 && sh[threadIdx.x]) atomicAdd(&header[1 + threadIdx.x], sh[threadIdx.x]);
}

}  // namespace

void launch_finalize(const float* raw32, int64_t n_pairs, float log10_init_f, float* lik32, unsigned* header, int sm_count,
                     cudaStream_t st)
{
    if (n_pairs <= 0) return;
    const int64_t want = (n_pairs + 255) / 256;
    const int grid = (int)std::min<int64_t>(want, (int64_t)sm_count * 8);
    finalize_kernel<<<grid, 256, 0, st>>>(raw32, n_pairs, log10_init_f, lik32, header);
}

}  // namespace phmm
