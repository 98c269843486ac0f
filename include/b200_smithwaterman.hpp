// b200_smithwaterman.hpp -- C++17 host-side mirror of the reference's Smith-Waterman aligner call surface
// (SURVEY.md section 8f-4).
//
//   hc::B200SWAligner aligner;
//   auto [alignment_begin, cigar] = aligner.align(ref, h.bases);          // assembler/graph_wrapper.hpp:232-239
//
// replaces hc::IntelSWAligner (smithwaterman/intel_smithwaterman.hpp:9-59) one for one: same parameter
// struct and named parameter sets, same default (NEW_SW_PARAMETERS), same result: the alignment offset into
// `ref` and the CIGAR -- here as its string form, which hc::Cigar accepts by assignment and construction
// (sam/cigar.hpp:76-84), so `h.cigar = std::move(cigar)` compiles unchanged.  Same errors: empty sequences
// throw std::invalid_argument (:33-34).  All device work goes through phmm_sw_align (include/phmm.h); there
// is no CPU fallback (std::runtime_error with the library's message when no device is usable).
//
// One alignment per call is all latency on a GPU; align_batch() scores every haplotype of a region (or of
// many regions) in one launch and is what a batching caller should use.
#pragma once

#include <cstdint>
#include <stdexcept>
#include <string>
#include <string_view>
#include <utility>
#include <vector>

#include "phmm.h"

namespace hc
{

struct B200SWAligner
{
    struct SWParameters
    {
        int w_match;
        int w_mismatch;
        int w_open;
        int w_extend;
    };
    // intel_smithwaterman.hpp:21-25
    static constexpr SWParameters ORIGINAL_DEFAULT{3, -1, -4, -3};
    static constexpr SWParameters STANDARD_NGS{25, -50, -110, -6};
    static constexpr SWParameters NEW_SW_PARAMETERS{200, -150, -260, -11};
    static constexpr SWParameters ALIGNMENT_TO_BEST_HAPLOTYPE_SW_PARAMETERS{10, -15, -30, -5};

    int device = 0;

    // offset, CIGAR string -- intel_smithwaterman.hpp:29-44
    std::pair<std::size_t, std::string> align(std::string_view ref, std::string_view alt,
                                              const SWParameters& params = NEW_SW_PARAMETERS) const
    {
        return std::move(align_batch({{ref, alt}}, params)[0]);
    }

    std::vector<std::pair<std::size_t, std::string>>
    align_batch(const std::vector<std::pair<std::string_view, std::string_view>>& pairs,
                const SWParameters& params = NEW_SW_PARAMETERS) const
    {
        const std::size_t n = pairs.size();
        std::vector<std::pair<std::size_t, std::string>> out(n);
        if (n == 0) return out;
        std::vector<int32_t> ref_off(n + 1, 0), alt_off(n + 1, 0);
        std::vector<uint8_t> ref_bases, alt_bases;
        for (std::size_t k = 0; k < n; k++) {
            if (pairs[k].first.empty() || pairs[k].second.empty())
                throw std::invalid_argument("Non-null sequences are required for the SW aligner");
            ref_bases.insert(ref_bases.end(), pairs[k].first.begin(), pairs[k].first.end());
            alt_bases.insert(alt_bases.end(), pairs[k].second.begin(), pairs[k].second.end());
            ref_off[k + 1] = (int32_t)ref_bases.size();
            alt_off[k + 1] = (int32_t)alt_bases.size();
        }
        // room for every CIGAR: nref + nalt + 2 elements per alignment always suffice
        const int64_t cap = (int64_t)ref_bases.size() + (int64_t)alt_bases.size() + 2 * (int64_t)n;
        std::vector<int32_t> offset(n), lens((std::size_t)cap);
        std::vector<int64_t> elem_beg(n + 1);
        std::vector<uint8_t> ops((std::size_t)cap);
        phmm_sw_batch b{(int32_t)n, ref_off.data(), ref_bases.data(), alt_off.data(), alt_bases.data(),
                        params.w_match, params.w_mismatch, params.w_open, params.w_extend};
        phmm_sw_result r{offset.data(), elem_beg.data(), cap, ops.data(), lens.data(), 0.f};
        const int rc = phmm_sw_align(device, &b, &r);
        if (rc != PHMM_OK) throw std::runtime_error(std::string("phmm_sw_align: ") + phmm_strerror(rc));
        last_kernel_ms = r.kernel_ms;
        for (std::size_t k = 0; k < n; k++) {
            out[k].first = (std::size_t)offset[k];
            for (int64_t e = elem_beg[k]; e < elem_beg[k + 1]; e++) {
                out[k].second += std::to_string(lens[(std::size_t)e]);
                out[k].second += (char)ops[(std::size_t)e];
            }
        }
        return out;
    }

    mutable float last_kernel_ms = 0.f;
};

} // namespace hc
