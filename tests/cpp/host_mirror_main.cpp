// Exercises hc::B200PairHMM (include/b200_pairhmm.hpp) with stand-ins for hc::Haplotype/hc::SAMRecord
// that carry the members the engine touches.  Prints the kept read indices and the matrix (%.17g) so
// tests/test_cpp_mirror.py can compare it with the Python face and the golden fixtures.
//   usage: host_mirror_main <region.txt>     lines: "H <bases>" | "R <seq> <qual>"
//          host_mirror_main --sw <ref> <alt> [<alt> ...]      hc::B200SWAligner::align / align_batch: "offset cigar" lines
//          host_mirror_main --batched <r1.txt> <r2.txt> ...   every region through hc::B200RegionBatcher (tiny
//          flush threshold, so several asynchronous batches), taken in REVERSE order; prints "region <i>"
//          before each region's block
#include <cstdio>
#include <fstream>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

#include "b200_pairhmm.hpp"
#include "b200_smithwaterman.hpp"

struct Haplotype { std::string bases; };
struct SAMRecord { std::string QNAME, SEQ, QUAL; std::size_t size() const { return SEQ.size(); } };

static void load(const char* path, std::vector<Haplotype>& haps, std::vector<SAMRecord>& reads)
{
    std::ifstream in(path);
    std::string line;
    while (std::getline(in, line)) {
        std::istringstream is(line); std::string tag; is >> tag;
        if (tag == "H") { Haplotype h; is >> h.bases; haps.push_back(h); }
        else if (tag == "R") { SAMRecord r; is >> r.SEQ >> r.QUAL; r.QNAME = std::to_string(reads.size()); reads.push_back(r); }
    }
}

static void print(const std::vector<SAMRecord>& reads, const std::vector<std::vector<double>>& lik)
{
    std::printf("kept");
    for (auto& r : reads) std::printf(" %s", r.QNAME.c_str());
    std::printf("\n");
    for (auto& row : lik) { for (double v : row) std::printf("%.17g ", v); std::printf("\n"); }
}

int main(int argc, char** argv)
{
    if (argc < 2) { std::fprintf(stderr, "usage: %s [--batched] region.txt ...\n", argv[0]); return 2; }
    if (std::string(argv[1]) == "--sw") {             // host_mirror_main --sw <ref> <alt> [<alt> ...]: hc::B200SWAligner
        try {
            hc::B200SWAligner aligner;                 // assembler/graph_wrapper.hpp:232
            std::vector<std::pair<std::string_view, std::string_view>> pairs;
            for (int i = 3; i < argc; i++) pairs.emplace_back(argv[2], argv[i]);
            if (pairs.size() == 1) {
                auto [alignment_begin, cigar] = aligner.align(argv[2], argv[3]);       // :235
                std::printf("%zu %s\n", alignment_begin, cigar.c_str());
            } else {
                for (auto& r : aligner.align_batch(pairs)) std::printf("%zu %s\n", r.first, r.second.c_str());
            }
        } catch (const std::exception& e) {
            std::fprintf(stderr, "error: %s\n", e.what());
            return 1;
        }
        return 0;
    }
    if (std::string(argv[1]) == "--batched") {
        const int n = argc - 2;
        std::vector<std::vector<Haplotype>> haps(n);
        std::vector<std::vector<SAMRecord>> reads(n);
        std::vector<int> ids(n);
        try {
            hc::B200RegionBatcher batcher(/*flush_cells=*/20000, /*flush_regions=*/3, /*max_in_flight=*/2);
            for (int i = 0; i < n; i++) { load(argv[2 + i], haps[i], reads[i]); ids[i] = batcher.add_region(haps[i], reads[i]); }
            for (int i = n - 1; i >= 0; i--) {
                auto lik = batcher.take(ids[i], reads[i]);
                std::printf("region %d\n", i);
                print(reads[i], lik);
            }
            std::fprintf(stderr, "batches %d\n", batcher.batches_submitted);
            // every region is handed out once (its batch's storage is released with the last one)
            bool refused = false;
            try { if (n) (void)batcher.take(ids[0], reads[0]); } catch (const std::runtime_error&) { refused = true; }
            if (n && !refused) { std::fprintf(stderr, "error: a second take of region 0 was not refused\n"); return 2; }
        } catch (const std::exception& e) {
            std::fprintf(stderr, "error: %s\n", e.what());
            return 1;
        }
        return 0;
    }
    std::vector<Haplotype> haps; std::vector<SAMRecord> reads;
    load(argv[1], haps, reads);
    try {
        hc::B200PairHMM pairhmm;                                  // haplotypecaller.hpp:90
        auto lik = pairhmm.compute_likelihoods(haps, reads);      // :103
        print(reads, lik);
    } catch (const std::exception& e) {
        std::fprintf(stderr, "error: %s\n", e.what());
        return 1;
    }
    return 0;
}
