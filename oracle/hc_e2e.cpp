// hc_e2e.cpp -- chrM-style END-TO-END harness: the reference's whole driver
// hc::HaplotypeCaller::do_work (haplotypecaller.hpp:112-154: FASTA + SAM load, 245 bp windows, read
// filters, clipping, local assembly, PairHMM, genotyping, VCF) compiled UNMODIFIED from /root/reference
// around either likelihood engine.  TEST INFRASTRUCTURE ONLY; output: oracle/_ref/hc_e2e_{ref,b200}.
//
//   hc_e2e_ref   -I reads.sam -R ref.fa -O out.vcf      hc::IntelPairHMM  (the reference as shipped, CPU)
//   hc_e2e_b200  -I reads.sam -R ref.fa -O out.vcf      hc::B200PairHMM   (this repo, GPU through the C ABI)
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
#include <chrono>
#include <cstdio>
#include <cstring>
#include <string>

#include "pairhmm/intel_pairhmm.hpp"
#include "b200_pairhmm.hpp"
#ifdef HC_USE_B200
#define IntelPairHMM B200PairHMM
#endif
#include "haplotypecaller.hpp"

int main(int argc, char** argv)
{
    mallopt(M_MMAP_THRESHOLD, 1 << 30);
    mallopt(M_TOP_PAD, 1 << 20);
    std::string in, out, ref;
    for (int i = 1; i + 1 < argc; i += 2) {
        if (!std::strcmp(argv[i], "-I")) in = argv[i + 1];
        else if (!std::strcmp(argv[i], "-O")) out = argv[i + 1];
        else if (!std::strcmp(argv[i], "-R")) ref = argv[i + 1];
    }
    if (in.empty() || out.empty() || ref.empty()) { std::fprintf(stderr, "usage: %s -I in.sam -R ref.fa -O out.vcf\n", argv[0]); return 2; }
    const auto t0 = std::chrono::steady_clock::now();
    try {
#ifdef HC_USE_B200
        hc::B200Engine::get();                 // CUDA context + tables once, outside the region loop
#endif
        const auto t1 = std::chrono::steady_clock::now();
        hc::HaplotypeCaller{in, out, ref}.do_work();      // main.cpp:26
        const auto t2 = std::chrono::steady_clock::now();
        std::fprintf(stderr, "hc_e2e: engine=%s init_s=%.3f do_work_s=%.3f\n",
#ifdef HC_USE_B200
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
