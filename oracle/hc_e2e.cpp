// hc_e2e.cpp -- chrM-style END-TO-END harness: the reference's whole driver
// hc::HaplotypeCaller::do_work (haplotypecaller.hpp:112-154: FASTA + SAM load, 245 bp windows, read
// filters, clipping, local assembly, PairHMM, genotyping, VCF) compiled UNMODIFIED from /root/reference
// around either likelihood engine.  TEST INFRASTRUCTURE ONLY; output: oracle/_ref/hc_e2e_{ref,b200}.
//
//   hc_e2e_ref   -I reads.sam -R ref.fa -O out.vcf      hc::IntelPairHMM  (the reference as shipped, CPU)
//   hc_e2e_b200  -I reads.sam -R ref.fa -O out.vcf      hc::B200PairHMM   (this repo, GPU through the C ABI)
//   hc_e2e_b200_batched  ... [-T threads]                hc::B200RegionBatcher: the caller side of SURVEY 8f-2.
//       The reference's window loop restructured into passes -- (1) per window, in parallel over host
//       threads: filters, clipping, local assembly (haplotypecaller.hpp:86-100, every function the
//       reference's own); (2) every region handed to the batcher, a few asynchronous batches for the whole
//       contig instead of one synchronous call per window; (3) per window, in order: matrix, genotyper,
//       VCF lines (:104-106).  Windows are independent in the reference (fresh Assembler / engine /
//       Genetyper per region, :88-90), so the VCF must be and is byte-identical.
//
// The engine swap is the one-line type change of INTEGRATION.md, made here without touching the
// reference's file: intel_pairhmm.hpp is included first (its `#pragma once` then keeps
// haplotypecaller.hpp from including it again), and -DHC_USE_B200 renames the token IntelPairHMM to
// B200PairHMM for the rest of the translation unit, i.e. inside call_region (:90).
// Boost.Graph comes from the shim oracle/stub/boost/graph (Boost is not installed here);
// boost::program_options (main.cpp) is replaced by the three flags parsed below.
// Two hazards of the reference itself are neutralised by the INPUT, not by patching it:
//   * select_one_read draws with std::random_device when several reads share a start (:44-50) ->
//     the synthetic SAM has at most one read per start position;
//   * the window loop reads reads_map[begin] before its bounds check (:142) past the end of the vector
//     in the last windows -> mallopt keeps that vector on the brk heap, where the stray read lands in
//     mapped memory (it is never dereferenced: .empty() is evaluated on garbage and `begin < size` fails).
#include <malloc.h>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <random>
#include <sstream>
#include <string>
#include <string_view>
#include <thread>

#include "pairhmm/intel_pairhmm.hpp"
#include "b200_pairhmm.hpp"
#ifdef HC_USE_B200
#define IntelPairHMM B200PairHMM
#endif
#ifdef HC_USE_B200_SW       // hc_e2e_b200_sw: the haplotype -> reference aligner swapped too, same trick:
#include "smithwaterman/intel_smithwaterman.hpp"      // the reference's header first (#pragma once), then the
#include "b200_smithwaterman.hpp"                     // token IntelSWAligner names hc::B200SWAligner inside
#define IntelSWAligner B200SWAligner                  // assembler/graph_wrapper.hpp:232-239
#endif
#ifdef HC_BATCHED
#define private public      // the batched driver below calls the reference's own private helpers
#endif                      // (load_all_reads, select_one_read, filter_reads, hard_clip_reads)
#include "haplotypecaller.hpp"
#ifdef HC_BATCHED
#undef private

namespace {

#ifdef HC_DEVICE_GL
// Genetyper's constants (genotyper.hpp:16-19: declared before its first access label, i.e. private by class default,
// out of reach of the `#define private public` above)
constexpr std::size_t kAlleleExtension = 2, kMinHeterozygosityQuality = 50, kMaxAlleleCount = 7;
// What Genetyper::assign_genotype_likelihoods (genotyper.hpp:369-397) knows about a site BEFORE it looks at a
// likelihood -- its loop cut in two at the marginalize() call (:389).
struct SitePlan {
    std::vector<std::string> alleles;
    hc::Interval alleles_loc;
    std::size_t allele_count = 0;
};
#endif

struct Window {
    hc::Interval origin, padded;
    std::vector<hc::SAMRecord> reads;
    std::vector<hc::Haplotype> haplotypes;
    std::string_view ref;
    int region_id = -1;
#ifdef HC_DEVICE_GL
    std::vector<SitePlan> plans;
    std::vector<hc::B200Site> sites;
#endif
};

#ifdef HC_DEVICE_GL
// First half of assign_genotype_likelihoods: events, alleles, haplotype -> allele map, overlapping reads
// (genotyper.hpp:376-389, every call the reference's own).
void plan_sites(Window& w)
{
    hc::Genetyper g;
    auto events_begins = g.set_events_for_haplotypes(w.haplotypes, w.ref, w.padded);
    const auto& [contig, origin_begin, origin_end] = w.origin;
    for (auto begin : events_begins) {
        if (begin < origin_begin || begin >= origin_end) continue;
        auto events = g.get_events_from_haplotypes(begin, w.haplotypes);
        g.replace_span_dels(events, w.ref[begin - w.padded.begin], begin);
        auto [alleles, alleles_loc] = g.get_compatible_alleles(events);
        auto allele_count = alleles.size();
        if (allele_count > kMaxAlleleCount) continue;
        auto allele_mapper = g.get_allele_mapper(alleles, begin, w.haplotypes);
        auto haplotype_mapper = g.get_haplotype_mapper(allele_mapper, w.haplotypes.size());
        const auto overlap = alleles_loc.expand_within_contig(kAlleleExtension);
        hc::B200Site site;
        site.n_alleles = (int32_t)allele_count;
        for (auto a : haplotype_mapper) site.hap_allele.push_back((uint8_t)a);
        for (const auto& r : w.reads) site.read_overlap.push_back(r.get_interval().overlaps(overlap) ? 1 : 0);   // :235-244
        w.sites.push_back(std::move(site));
        w.plans.push_back({std::move(alleles), alleles_loc, allele_count});
    }
}
#endif

// hc::HaplotypeCaller::do_work (haplotypecaller.hpp:112-154) in passes; see the header comment.
void do_work_batched(hc::HaplotypeCaller& caller, int n_threads, std::size_t region_size = 245, std::size_t padding_size = 85)
{
    std::ifstream ifs(caller.ref_path);
    if (!ifs) throw std::runtime_error("cannot open " + caller.ref_path);
    auto fasta = hc::Fasta{};
    ifs >> fasta;
    ifs.close();
    std::transform(fasta.seq.begin(), fasta.seq.end(), fasta.seq.begin(), ::toupper);       // :122
    const auto ref = std::string_view{fasta.seq};

    const auto windows_number = (ref.size() + region_size - 1) / region_size;
    auto reads_map = caller.load_all_reads(ref.size());

    // pass 0 (serial): the reference's window walk and read selection (:127-143,152-155)
    std::vector<Window> windows(windows_number);
    {
        auto origin = hc::Interval{fasta.name, 0, region_size};
        auto padded = origin;
        padded.end += padding_size;
        for (auto& w : windows) {
            w.origin = origin; w.padded = padded;
            for (auto begin = padded.begin; begin != padded.end; begin++)
                if (begin < reads_map.size() && !reads_map[begin].empty())
                    w.reads.emplace_back(caller.select_one_read(reads_map[begin]));
            w.ref = ref.substr(padded.begin, padded.size());
            origin.begin += region_size; origin.end += region_size;
            padded.begin = origin.begin - padding_size; padded.end = origin.end + padding_size;
        }
    }
    // pass 1 (parallel over windows): call_region up to the assembler (:92-100)
    {
        std::atomic<std::size_t> next{0};
        auto work = [&] {
            for (std::size_t i; (i = next.fetch_add(1)) < windows.size();) {
                Window& w = windows[i];
                if (w.reads.empty()) continue;
                caller.filter_reads(w.reads);
                caller.hard_clip_reads(w.reads, w.padded);
                if (w.reads.empty()) continue;
                hc::Assembler assembler;
                w.haplotypes = assembler.assemble(w.reads, w.ref);
            }
        };
        std::vector<std::thread> pool;
        for (int t = 1; t < n_threads; t++) pool.emplace_back(work);
        work();
        for (auto& t : pool) t.join();
    }
    // pass 2: every region with more than one haplotype goes to the device, batched across windows
#ifdef HC_DEVICE_GL
    hc::B200RegionBatcher batcher((int64_t)1.6e10, 4096, 3, /*device_gl=*/true);
    for (auto& w : windows)
        if (!w.reads.empty() && w.haplotypes.size() > 1) { plan_sites(w); w.region_id = batcher.add_region(w.haplotypes, w.reads, w.sites); }
#else
    hc::B200RegionBatcher batcher;
    for (auto& w : windows)
        if (!w.reads.empty() && w.haplotypes.size() > 1) w.region_id = batcher.add_region(w.haplotypes, w.reads);
#endif
    batcher.flush();
    // pass 3 (window order): matrix, genotyper, VCF (:103-106)
    auto ofs = std::ofstream{caller.out_path};
    if (!ofs) throw std::runtime_error("cannot open " + caller.out_path);
    ofs << "##fileformat=VCFv4.2\n";
    ofs << "##FORMAT=<ID=GQ,Number=1,Type=Integer,Description=\"Genotype Quality\">\n";
    ofs << "##FORMAT=<ID=GT,Number=1,Type=String,Description=\"Genotype\">\n";
    ofs << "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\tNA12878\n";
    for (auto& w : windows) {
        if (w.region_id < 0) continue;
#ifdef HC_DEVICE_GL
        // second half of assign_genotype_likelihoods (genotyper.hpp:390-396) over the vectors the device computed
        auto gl = batcher.take_gl(w.region_id);
        hc::Genetyper genetyper;
        std::vector<hc::Variant> variants;
        for (std::size_t k = 0; k < w.plans.size(); k++) {
            auto& plan = w.plans[k];
            const auto& genotype_likelihoods = gl.site_genotype_likelihoods[k];
            auto [genotype_index, genotype_quality] = genetyper.get_genotype_quality_and_max_genotype_index(genotype_likelihoods);
            if (genotype_index == 0) continue;
            auto genotype = genetyper.get_genotype(plan.allele_count, genotype_index);
            if (genotype.first == 0 && genotype_quality < kMinHeterozygosityQuality) continue;
            variants.emplace_back(std::move(plan.alleles_loc), std::move(plan.alleles), genotype, genotype_quality);
        }
#else
        auto likelihoods = batcher.take(w.region_id, w.reads);
        hc::Genetyper genetyper;
        auto variants = genetyper.assign_genotype_likelihoods(w.reads, w.haplotypes, likelihoods, w.ref, w.padded, w.origin);
#endif
        for (const auto& variant : variants) variant.print(ofs);
    }
    std::fprintf(stderr, "hc_e2e: batched: %zu windows, %d batches, %lld pairs, %.3e cells, device kernels %.2f ms\n",
                 windows.size(), batcher.batches_submitted, (long long)batcher.total_stats.n_pairs,
                 (double)batcher.total_stats.n_cells, batcher.total_stats.kernel_ms);
}

}  // namespace
#endif

int main(int argc, char** argv)
{
    mallopt(M_MMAP_THRESHOLD, 1 << 30);
    mallopt(M_TOP_PAD, 1 << 20);
    std::string in, out, ref;
    int threads = (int)std::max(1u, std::thread::hardware_concurrency());
    for (int i = 1; i + 1 < argc; i += 2) {
        if (!std::strcmp(argv[i], "-I")) in = argv[i + 1];
        else if (!std::strcmp(argv[i], "-O")) out = argv[i + 1];
        else if (!std::strcmp(argv[i], "-R")) ref = argv[i + 1];
        else if (!std::strcmp(argv[i], "-T")) threads = std::max(1, std::atoi(argv[i + 1]));
    }
    (void)threads;
    if (in.empty() || out.empty() || ref.empty()) { std::fprintf(stderr, "usage: %s -I in.sam -R ref.fa -O out.vcf\n", argv[0]); return 2; }
    const auto t0 = std::chrono::steady_clock::now();
    try {
#ifdef HC_USE_B200
        hc::B200Engine::get();                 // CUDA context + tables once, outside the region loop
#endif
        const auto t1 = std::chrono::steady_clock::now();
#ifdef HC_BATCHED
        std::cout.setstate(std::ios::failbit);           // the assembler's progress lines, from many threads
        hc::HaplotypeCaller caller{in, out, ref};
        do_work_batched(caller, threads);
#else
        hc::HaplotypeCaller{in, out, ref}.do_work();      // main.cpp:26
#endif
        const auto t2 = std::chrono::steady_clock::now();
        std::fprintf(stderr, "hc_e2e: engine=%s init_s=%.3f do_work_s=%.3f\n",
#if defined(HC_USE_B200_SW)
                     "b200+sw",
#elif defined(HC_DEVICE_GL)
                     "b200-batched+device-gl",
#elif defined(HC_BATCHED)
                     "b200-batched",
#elif defined(HC_USE_B200)
                     "b200",
#else
                     "ref",
#endif
                     std::chrono::duration<double>(t1 - t0).count(), std::chrono::duration<double>(t2 - t1).count());
    } catch (const std::exception& e) {
        std::fprintf(stderr, "hc_e2e: error: %s\n", e.what());
        return 1;
    }
    return 0;
}
