// b200_pairhmm.hpp -- C++17 host-side mirror of the reference's likelihood-engine call surface.
//
//   hc::B200PairHMM pairhmm;
//   auto likelihoods = pairhmm.compute_likelihoods(haplotypes, reads);      // haplotypecaller.hpp:103
//
// replaces hc::IntelPairHMM (pairhmm/intel_pairhmm.hpp:16-56) one for one: same name and argument
// meaning, same result type (std::vector<std::vector<double>> indexed [kept read][haplotype]), same
// side effect (poorly modelled reads are ERASED from the caller's `reads` vector, :40-45), same cap
// at best - 4.5 (:29-33).  Swapping it in is the include at haplotypecaller.hpp:15 and the type at :90.
//
// It is a header-only template over the reference's own value types: it touches only
// Haplotype::bases, SAMRecord::SEQ, SAMRecord::QUAL and SAMRecord::size(), exactly the fields
// IntelPairHMM::getData reads (:154-203), so it compiles against hc::Haplotype / hc::SAMRecord
// unchanged (and against any struct with those members, which is how tests/cpp exercises it here,
// where Boost is absent).  Gap penalties are the reference's constants: insertionGOP/deletionGOP =
// 'I' repeated, overallGCP = '+' repeated (sam/sam.hpp:30-32,47-49); unlike the reference there is no
// 200-base limit on the read (its constant strings are 200 long).
//
// All device work goes through the C ABI of include/phmm.h (libphmm_b200.so).  The engine (streams,
// memory pool, tables) is created ONCE per process and shared, because the reference constructs its
// engine object per region (haplotypecaller.hpp:90) and a CUDA context must not be.
// Errors: the reference has no error path here; this class throws std::runtime_error with the
// library's message (never falls back to a CPU implementation).
#pragma once

#include <cstdint>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

#include "phmm.h"

namespace hc
{

class B200Engine
{
public:
    // process-wide engine; `devices` empty = CUDA device 0
    static phmm_engine* get(const std::vector<int32_t>& devices = {})
    {
        static std::mutex mu;
        static phmm_engine* eng = nullptr;
        std::lock_guard<std::mutex> lk(mu);
        if (!eng) {
            phmm_options opt{};
            opt.struct_size = (int32_t)sizeof(opt);
            opt.n_devices = devices.empty() ? 1 : (int32_t)devices.size();
            opt.devices = devices.empty() ? nullptr : devices.data();
            opt.pipeline_depth = 2;
            opt.host_threads = 1;
            int rc = phmm_create(&opt, &eng);
            if (rc != PHMM_OK) throw std::runtime_error(std::string("phmm_create: ") + phmm_strerror(rc));
        }
        return eng;
    }
};

struct B200PairHMM
{
    static constexpr char GAP_OPEN = 'I';      // sam/sam.hpp:31
    static constexpr char GAP_CONT = '+';      // sam/sam.hpp:32

    // Packs one region into the SoA batch of include/phmm.h.
    template <class HaplotypeT, class ReadT>
    struct Packed
    {
        std::vector<int32_t> region_read_beg, region_hap_beg, read_off, hap_off, read_len;
        std::vector<uint8_t> read_bases, read_q, hap_bases;
        phmm_batch batch{};

        Packed(const std::vector<HaplotypeT>& haps, const std::vector<ReadT>& reads)
        {
            region_read_beg = {0, (int32_t)reads.size()};
            region_hap_beg  = {0, (int32_t)haps.size()};
            read_off.push_back(0);
            for (const auto& r : reads) {
                if (r.SEQ.size() != r.QUAL.size()) throw std::runtime_error("B200PairHMM: SEQ and QUAL lengths differ");
                read_bases.insert(read_bases.end(), r.SEQ.begin(), r.SEQ.end());
                read_q.insert(read_q.end(), r.QUAL.begin(), r.QUAL.end());
                read_off.push_back((int32_t)read_bases.size());
                read_len.push_back((int32_t)r.SEQ.size());
            }
            hap_off.push_back(0);
            for (const auto& h : haps) {
                hap_bases.insert(hap_bases.end(), h.bases.begin(), h.bases.end());
                hap_off.push_back((int32_t)hap_bases.size());
            }
            batch.n_regions = 1;
            batch.n_reads = (int32_t)reads.size();
            batch.n_haps = (int32_t)haps.size();
            batch.region_read_beg = region_read_beg.data();
            batch.region_hap_beg = region_hap_beg.data();
            batch.read_off = read_off.data();
            batch.read_bases = read_bases.data();
            batch.read_q = read_q.data();
            batch.read_i = batch.read_d = batch.read_c = nullptr;       // constant strings of sam.hpp
            batch.gap_open_i = (uint8_t)GAP_OPEN; batch.gap_open_d = (uint8_t)GAP_OPEN; batch.gap_cont_c = (uint8_t)GAP_CONT;
            batch.hap_off = hap_off.data();
            batch.hap_bases = hap_bases.data();
        }
    };

    // intel_pairhmm.hpp:48-56
    template <class HaplotypeT, class ReadT>
    std::vector<std::vector<double>> compute_likelihoods(const std::vector<HaplotypeT>& haplotypeDataArray,
                                                         std::vector<ReadT>& readDataArray)
    {
        const std::size_t n_reads = readDataArray.size(), n_haps = haplotypeDataArray.size();
        std::vector<std::vector<double>> likelihoodArray(n_reads, std::vector<double>(n_haps));
        if (n_reads == 0 || n_haps == 0) return likelihoodArray;
        Packed<HaplotypeT, ReadT> p(haplotypeDataArray, readDataArray);
        std::vector<double> flat(n_reads * n_haps);
        phmm_result res{};
        res.log10_lik = flat.data();
        phmm_engine* eng = B200Engine::get();
        int rc = phmm_compute(eng, &p.batch, &res);
        if (rc != PHMM_OK)
            throw std::runtime_error(std::string("phmm_compute: ") + phmm_strerror(rc) + ": " + phmm_last_error(eng));
        last_stats = res.stats;
        // normalize_likelihoods_and_filter_poorly_modeled_reads (:24-46), host side
        std::vector<uint8_t> keep(n_reads);
        phmm_normalize_filter(flat.data(), (int32_t)n_reads, (int32_t)n_haps, p.read_len.data(), keep.data());
        std::size_t w = 0;
        for (std::size_t r = 0; r < n_reads; r++) {
            if (!keep[r]) continue;
            likelihoodArray[w].assign(flat.begin() + r * n_haps, flat.begin() + (r + 1) * n_haps);
            if (w != r) readDataArray[w] = std::move(readDataArray[r]);
            ++w;
        }
        likelihoodArray.resize(w);
        readDataArray.erase(readDataArray.begin() + w, readDataArray.end());
        return likelihoodArray;
    }

    phmm_stats last_stats{};
};

} // namespace hc
