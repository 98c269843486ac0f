// ref_harness.cpp -- thin extern "C" shim around the REFERENCE's own PairHMM code.
//
// TEST INFRASTRUCTURE ONLY.  Compiled (oracle/Makefile) from the sources where they lie under
// /root/reference (nothing is copied into this repo) into oracle/_ref/libref_pairhmm.so, which is
// git-ignored.  It is used to (1) validate oracle/pairhmm_oracle.c bit-for-bit, (2) generate
// tests/golden/*.json, (3) serve as bench.py's cpu_baseline / --impl reference arm
// (cpu_baseline.kind == "reference").  The product never loads it.
//
// Two levels are exposed:
//   kernel level        compute_full_prob_avxs<float> / compute_full_prob_avxd<double>
//                       (native/avx-pairhmm-template.h:210) plus the private 15-line dispatch loop
//                       of intel_pairhmm.hpp:131-147 restated here (it cannot be called directly);
//   call-surface level  hc::IntelPairHMM::compute_likelihoods (intel_pairhmm.hpp:48-56), built with
//                       the empty boost/serialization stubs under oracle/stub/.
#include <cstdint>
#include <algorithm>
#include <cassert>
#include <limits>
#include <map>
#include <set>
#include <vector>
#include <numeric>
#include <iostream>
#include <memory>
#include <string>
#include <cmath>
#include <cstring>

#include "pairhmm/intel_pairhmm.hpp"   // pulls native/avx-pairhmm.h, sam.hpp, haplotype.hpp
#define private public                 // Genetyper's marginalisation / genotype-likelihood helpers are private members
#include "genotyper/genotyper.hpp"     // (the consumer of the matrix; pins oracle/genotype_oracle.c)
#undef private
#include "smithwaterman/intel_smithwaterman.hpp"   // the reference's SW aligner

namespace {
Context<float>*  g_f = nullptr;
Context<double>* g_d = nullptr;
inline void ftz_on() { _MM_SET_FLUSH_ZERO_MODE(_MM_FLUSH_ZERO_ON); }   // intel_pairhmm.hpp:102-105
inline testcase make_tc(const uint8_t* rs, const uint8_t* q, const uint8_t* i, const uint8_t* d,
                        const uint8_t* c, int R, const uint8_t* hap, int H)
{
    testcase tc;
    tc.rslen = R; tc.haplen = H;
    tc.rs = (const char*)rs; tc.q = (const char*)q; tc.i = (const char*)i;
    tc.d = (const char*)d; tc.c = (const char*)c; tc.hap = (const char*)hap;
    return tc;
}
}

extern "C" {

void ref_init(void)
{
    if (!g_f) { g_f = new Context<float>(); g_d = new Context<double>(); ConvertChar::init(); }
    ftz_on();
}

const float*  ref_ph2pr_f32(void) { ref_init(); return ContextBase<float>::ph2pr; }
const double* ref_ph2pr_f64(void) { ref_init(); return ContextBase<double>::ph2pr; }
const float*  ref_mm_f32(void)    { ref_init(); return ContextBase<float>::matchToMatchProb; }
const double* ref_mm_f64(void)    { ref_init(); return ContextBase<double>::matchToMatchProb; }
float  ref_log10_init_f32(void)   { ref_init(); return ContextBase<float>::LOG10_INITIAL_CONSTANT; }
double ref_log10_init_f64(void)   { ref_init(); return ContextBase<double>::LOG10_INITIAL_CONSTANT; }

float ref_forward_f32(const uint8_t* rs, const uint8_t* q, const uint8_t* i, const uint8_t* d,
                      const uint8_t* c, int R, const uint8_t* hap, int H)
{
    ref_init();
    testcase tc = make_tc(rs, q, i, d, c, R, hap, H);
    return compute_full_prob_avxs<float>(&tc);
}

double ref_forward_f64(const uint8_t* rs, const uint8_t* q, const uint8_t* i, const uint8_t* d,
                       const uint8_t* c, int R, const uint8_t* hap, int H)
{
    ref_init();
    testcase tc = make_tc(rs, q, i, d, c, R, hap, H);
    return compute_full_prob_avxd<double>(&tc);
}

// dispatch of one pair: intel_pairhmm.hpp:131-147
double ref_pair(const uint8_t* rs, const uint8_t* q, const uint8_t* i, const uint8_t* d,
                const uint8_t* c, int R, const uint8_t* hap, int H,
                float* raw32, double* raw64, uint8_t* rescued)
{
    testcase tc = make_tc(rs, q, i, d, c, R, hap, H);
    double result_final;
    float result_float = compute_full_prob_avxs<float>(&tc);
    if (raw32) *raw32 = result_float;
    if (result_float < MIN_ACCEPTED) {
        double result_double = compute_full_prob_avxd<double>(&tc);
        result_final = log10(result_double) - g_d->LOG10_INITIAL_CONSTANT;
        if (raw64) *raw64 = result_double;
        if (rescued) *rescued = 1;
    } else {
        result_final = (double)(log10f(result_float) - g_f->LOG10_INITIAL_CONSTANT);
        if (raw64) *raw64 = 0.0;
        if (rescued) *rescued = 0;
    }
    return result_final;
}

// Same batch layout and output convention as oracle_batch() / include/phmm.h.
int ref_batch(int n_regions, const int32_t* region_read_beg, const int32_t* region_hap_beg,
              const int32_t* read_off, const uint8_t* read_bases, const uint8_t* read_q,
              const uint8_t* read_i, const uint8_t* read_d, const uint8_t* read_c,
              const int32_t* hap_off, const uint8_t* hap_bases,
              double* log10_out, float* raw32, double* raw64, uint8_t* rescued, int threads)
{
    ref_init();
    std::vector<int64_t> out_beg(n_regions + 1, 0);
    for (int g = 0; g < n_regions; g++)
        out_beg[g + 1] = out_beg[g] + (int64_t)(region_read_beg[g + 1] - region_read_beg[g]) *
                                          (region_hap_beg[g + 1] - region_hap_beg[g]);
    if (threads < 1) threads = 1;
    for (int g = 0; g < n_regions; g++) {
        int r0 = region_read_beg[g], r1 = region_read_beg[g + 1];
        int h0 = region_hap_beg[g], h1 = region_hap_beg[g + 1];
        int nh = h1 - h0;
        // intel_pairhmm.hpp:128-130: parallel over reads, dynamic schedule (inert as shipped)
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads)
        for (int r = r0; r < r1; r++) {
            ftz_on();
            int ro = read_off[r], R = read_off[r + 1] - ro;
            for (int h = h0; h < h1; h++) {
                int ho = hap_off[h], H = hap_off[h + 1] - ho;
                int64_t o = out_beg[g] + (int64_t)(r - r0) * nh + (h - h0);
                float f; double d; uint8_t resc;
                log10_out[o] = ref_pair(read_bases + ro, read_q + ro, read_i + ro, read_d + ro,
                                        read_c + ro, R, hap_bases + ho, H, &f, &d, &resc);
                if (raw32) raw32[o] = f;
                if (raw64) raw64[o] = d;
                if (rescued) rescued[o] = resc;
            }
        }
    }
    return 0;
}

// Call-surface level: one region through hc::IntelPairHMM::compute_likelihoods.
// Reads carry only SEQ/QUAL (GOP/GCP are the reference's constant strings, sam.hpp:30-32,47-49,
// so read lengths must be <= 200).  keep[r] = 1 for reads that survive the filter;
// lik_out receives the returned [kept][n_haps] matrix row-major.  Returns the number kept.
int ref_compute_likelihoods(int n_reads, const int32_t* read_off, const uint8_t* read_bases,
                            const uint8_t* read_q, int n_haps, const int32_t* hap_off,
                            const uint8_t* hap_bases, double* lik_out, uint8_t* keep)
{
    std::vector<hc::SAMRecord> reads(n_reads);
    for (int r = 0; r < n_reads; r++) {
        reads[r].QNAME = std::to_string(r);
        reads[r].SEQ.assign((const char*)read_bases + read_off[r], read_off[r + 1] - read_off[r]);
        reads[r].QUAL.assign((const char*)read_q + read_off[r], read_off[r + 1] - read_off[r]);
        reads[r].MAPQ = 60;
    }
    std::vector<hc::Haplotype> haps(n_haps);
    for (int h = 0; h < n_haps; h++)
        haps[h].bases.assign((const char*)hap_bases + hap_off[h], hap_off[h + 1] - hap_off[h]);
    hc::IntelPairHMM engine;                     // fresh per region: haplotypecaller.hpp:90
    auto lik = engine.compute_likelihoods(haps, reads);
    std::memset(keep, 0, n_reads);
    for (size_t k = 0; k < reads.size(); k++) {
        keep[std::stoi(reads[k].QNAME)] = 1;
        for (int h = 0; h < n_haps; h++) lik_out[k * n_haps + h] = lik[k][h];
    }
    return (int)reads.size();
}

// hc::IntelSWAligner::align (smithwaterman/intel_smithwaterman.hpp:29-44) with explicit weights; returns
// the alignment offset and writes the CIGAR string.
int ref_sw_align(const uint8_t* ref, int nref, const uint8_t* alt, int nalt,
                 int w_match, int w_mismatch, int w_open, int w_extend, char* cigar, int cap)
{
    hc::IntelSWAligner aligner;
    hc::IntelSWAligner::SWParameters prm{w_match, w_mismatch, w_open, w_extend};
    auto [off, cg] = aligner.align(std::string_view((const char*)ref, (size_t)nref),
                                   std::string_view((const char*)alt, (size_t)nalt), prm);
    std::snprintf(cigar, (size_t)cap, "%s", cg.to_string().c_str());
    return (int)off;
}


// the reference's own marginal_likelihoods (genotyper.hpp:245-264) + calculate_genotype_likelihoods
// (:311-327) for one site.  lik [n_reads][n_haps] is the matrix the genotyper receives (capped, erased rows already
// gone); keep_idx are the indices get_read_indices_to_keep would return.  out: n_alleles (n_alleles + 1) / 2 doubles.
int ref_genotype_likelihoods(const double* lik, int n_reads, int n_haps, const int32_t* keep_idx, int n_keep,
                             int n_alleles, const uint8_t* hap_allele, double* out)
{
    hc::Genetyper g;
    std::vector<std::vector<double>> m(n_reads, std::vector<double>(n_haps));
    for (int r = 0; r < n_reads; r++) for (int h = 0; h < n_haps; h++) m[r][h] = lik[(size_t)r * n_haps + h];
    std::vector<std::size_t> mapper(n_haps), idx(n_keep);
    for (int h = 0; h < n_haps; h++) mapper[h] = hap_allele[h];
    for (int k = 0; k < n_keep; k++) idx[k] = (std::size_t)keep_idx[k];
    auto al = g.marginal_likelihoods((std::size_t)n_alleles, mapper, idx, m);
    auto gl = g.calculate_genotype_likelihoods(al, (std::size_t)n_alleles);
    for (std::size_t k = 0; k < gl.size(); k++) out[k] = gl[k];
    return (int)gl.size();
}

// The reference's Jacobian table as compiled into ITS binary (math_utils.hpp:24-28, evaluated by GCC at compile time)
const double* ref_jacobian_table(int* n)
{
    if (n) *n = (int)hc::MathUtils::JacobianLogTable::cache.size();
    return hc::MathUtils::JacobianLogTable::cache.data();
}

} // extern "C"
