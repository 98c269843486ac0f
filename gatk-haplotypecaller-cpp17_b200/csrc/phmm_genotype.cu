// phmm_genotype.cu -- the consumer of the likelihood matrix, on the device.
//
// After the forward kernels and the device log10 (phmm_finalize.cu) the matrix of a region never has to leave
// the GPU: what hc::Genetyper::assign_genotype_likelihoods needs from it, per variant site, is the vector of
// diploid genotype likelihoods (genotyper/genotyper.hpp:389-390) -- A (A + 1) / 2 doubles instead of
// reads x haplotypes.  Everything between the matrix and that vector is done here, bit for bit:
//
//   widen_kernel / rescue_kernel   the double matrix: (double) of the float log10 values (intel_pairhmm.hpp:142), and
//                                  log10(raw64) - log10(2^1020) for FP64-rescued pairs (:139) with glibc's log10
//                                  restated (phmm_log10.h), raw sums below DBL_MIN flushed as x86 FTZ does;
//   row_kernel                     normalize_likelihoods_and_filter_poorly_modeled_reads (intel_pairhmm.hpp:24-46):
//                                  cap every row at best - 4.5, flag reads whose best is below
//                                  min(2, ceil(0.02 len)) * -4 (the reference erases them);
//   site_kernel                    marginal_likelihoods (genotyper.hpp:245-264: max over the haplotypes of an
//                                  allele, lowest() where none), the per-read genotype terms (:276-309: L + log10 2,
//                                  or approximate_log10_sum_log10 with the Jacobian table, utils/math_utils.hpp:11-30),
//                                  and the sum over the reads IN READ ORDER from 0.0 minus n log10 2 (:311-320).
//
// What stays on the host is what does not depend on the likelihoods (events, alleles, haplotype -> allele maps:
// genotyper.hpp:111-233, passed in as phmm_sites) and what consumes the vector (genotype quality, the call:
// :329-368).  Every double operation is an explicit round-to-nearest intrinsic: no contraction into FMA.
#include "phmm_launch.h"
#include "phmm_log10.h"

#include <cfloat>

namespace phmm {

namespace {

__global__ void __launch_bounds__(256)
widen_kernel(const float* __restrict__ lik32, const int64_t n_pairs, double* __restrict__ lik64)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_pairs; i += (int64_t)gridDim.x * blockDim.x)
        lik64[i] = (double)lik32[i];                       // NaN where an FP64 result is still to come
}

__global__ void __launch_bounds__(128)
rescue_kernel(const RescueOut* __restrict__ rescue, const unsigned* __restrict__ count, const double log10_init_d,
              double* __restrict__ lik64)
{
    const unsigned n = *count;
    for (unsigned k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
        double d = rescue[k].raw64;
        if (d < DBL_MIN) d = 0.0;                          // MXCSR flush-to-zero reaches doubles too (intel_pairhmm.hpp:102-105)
        lik64[rescue[k].out_idx] = __dsub_rn(glibc_log10(d), log10_init_d);      // :139
    }
}

// One thread per read: cap its row in place, set keep[read].
__global__ void __launch_bounds__(128)
row_kernel(const GenotypeArgs g)
{
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= g.n_reads) return;
    int lo = 0, hi = g.n_regions;                           // region of read r: last region whose first read is <= r
    while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (g.region_read_beg[mid] <= r) lo = mid; else hi = mid; }
    const int nh = g.region_hap_beg[lo + 1] - g.region_hap_beg[lo];
    if (nh == 0) { g.read_keep[r] = 1; return; }
    double* row = g.lik64 + g.region_out_beg[lo] + (int64_t)(r - g.region_read_beg[lo]) * nh;
    double best = row[0];
    for (int j = 1; j < nh; j++) if (best < row[j]) best = row[j];                 // *std::max_element, :28
    const double cap = __dadd_rn(best, -4.5);                                      // :29
    for (int j = 0; j < nh; j++) if (row[j] < cap) row[j] = cap;                   // :30-33
    const int len = g.read_off[r + 1] - g.read_off[r];
    const double thr = __dmul_rn(fmin(2.0, ceil(__dmul_rn((double)len, 0.02))), -4.0);   // :35-36
    g.read_keep[r] = (best < thr) ? 0 : 1;
}

__device__ __forceinline__ double approx_log10_sum_log10(double a, double b, const double* __restrict__ jac, const double inv_step)
{
    if (a > b) { const double t = a; a = b; b = t; }       // math_utils.hpp:13
    const double diff = __dsub_rn(b, a);
    return __dadd_rn(b, diff < 8.0 ? jac[(size_t)round(__dmul_rn(diff, inv_step))] : 0.0);
}

// One block per site.
__global__ void __launch_bounds__(128)
site_kernel(const GenotypeArgs g)
{
    const int s = blockIdx.x;
    const int region = g.site_region[s];
    const int A = g.site_n_alleles[s];
    const int r0 = g.region_read_beg[region], nr = g.region_read_beg[region + 1] - r0;
    const int nh = g.region_hap_beg[region + 1] - g.region_hap_beg[region];
    const double* __restrict__ lik = g.lik64 + g.region_out_beg[region];
    const uint8_t* __restrict__ hap_allele = g.hap_allele + g.site_hap_off[s];
    const uint8_t* __restrict__ overlap = g.read_overlap ? g.read_overlap + g.site_read_off[s] : nullptr;
    double* __restrict__ al = g.scratch_al + g.site_read_off[s] * 8;               // [read][allele], 8 doubles per read
    uint8_t* __restrict__ used = g.scratch_used + g.site_read_off[s];
    // marginal_likelihoods over the reads that survived the filter and overlap the site (genotyper.hpp:235-264)
    for (int r = threadIdx.x; r < nr; r += blockDim.x) {
        const bool use = g.read_keep[r0 + r] && (!overlap || overlap[r]);
        used[r] = use ? 1 : 0;
        if (!use) continue;
        double m[8];
#pragma unroll
        for (int a = 0; a < 8; a++) m[a] = -DBL_MAX;                               // numeric_limits<double>::lowest()
        for (int h = 0; h < nh; h++) {
            const double v = lik[(int64_t)r * nh + h];
            const int a = hap_allele[h];
#pragma unroll
            for (int q = 0; q < 8; q++) if (q == a && v > m[q]) m[q] = v;          // (no dynamic register indexing)
        }
#pragma unroll
        for (int a = 0; a < 8; a++) al[(int64_t)r * 8 + a] = m[a];
    }
    __syncthreads();
    // one thread per genotype (a1 <= a2, a1 outer: genotyper.hpp:297-307); reads in order
    const int n_gt = A * (A + 1) / 2;
    if ((int)threadIdx.x < n_gt) {
        int a1 = 0, a2 = 0, k = (int)threadIdx.x;
        for (a1 = 0; a1 < A; a1++) { if (k < A - a1) { a2 = a1 + k; break; } k -= A - a1; }
        double sum = 0.0;
        int n_used = 0;
        for (int r = 0; r < nr; r++) {
            if (!used[r]) continue;
            const double x = al[(int64_t)r * 8 + a1], y = al[(int64_t)r * 8 + a2];
            const double v = (a1 == a2) ? __dadd_rn(x, g.log10_2) : approx_log10_sum_log10(x, y, g.jacobian, g.inv_step);
            sum = __dadd_rn(sum, v);                                               // std::accumulate from 0.0, :318
            n_used++;
        }
        g.gl[g.gl_off[s] + threadIdx.x] = __dsub_rn(sum, __dmul_rn((double)(unsigned long long)n_used, g.log10_2));   // :316-319
        if (threadIdx.x == 0 && g.site_n_used) g.site_n_used[s] = n_used;
    }
}

}  // namespace

void launch_genotype(const GenotypeArgs& g, const float* lik32, const RescueOut* rescue, const unsigned* rescue_count,
                     double log10_init_d, int sm_count, cudaStream_t st)
{
    if (g.n_pairs > 0) {
        const int grid = (int)std::min<int64_t>((g.n_pairs + 255) / 256, (int64_t)sm_count * 8);
        widen_kernel<<<grid, 256, 0, st>>>(lik32, g.n_pairs, g.lik64);
        rescue_kernel<<<std::max(1, sm_count), 128, 0, st>>>(rescue, rescue_count, log10_init_d, g.lik64);
    }
    if (g.n_reads > 0) row_kernel<<<(g.n_reads + 127) / 128, 128, 0, st>>>(g);
    if (g.n_sites > 0) site_kernel<<<g.n_sites, 128, 0, st>>>(g);
}

}  // namespace phmm
