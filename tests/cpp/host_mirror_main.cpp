// Exercises hc::B200PairHMM (include/b200_pairhmm.hpp) with stand-ins for hc::Haplotype/hc::SAMRecord
// that carry the members the engine touches.  Prints the kept read indices and the matrix (%.17g) so
// tests/test_cpp_mirror.py can compare it with the Python face and the golden fixtures.
//   usage: host_mirror_main <region.txt>     lines: "H <bases>" | "R <seq> <qual>"
#include <cstdio>
#include <fstream>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

#include "b200_pairhmm.hpp"

struct Haplotype { std::string bases; };
struct SAMRecord { std::string QNAME, SEQ, QUAL; std::size_t size() const { return SEQ.size(); } };

int main(int argc, char** argv)
{
    if (argc < 2) { std::fprintf(stderr, "usage: %s region.txt\n", argv[0]); return 2; }
    std::ifstream in(argv[1]);
    std::vector<Haplotype> haps; std::vector<SAMRecord> reads;
    std::string line;
    while (std::getline(in, line)) {
        std::istringstream is(line); std::string tag; is >> tag;
        if (tag == "H") { Haplotype h; is >> h.bases; haps.push_back(h); }
        else if (tag == "R") { SAMRecord r; is >> r.SEQ >> r.QUAL; r.QNAME = std::to_string(reads.size()); reads.push_back(r); }
    }
    try {
        hc::B200PairHMM pairhmm;                                  // haplotypecaller.hpp:90
        auto lik = pairhmm.compute_likelihoods(haps, reads);      // :103
        std::printf("kept");
        for (auto& r : reads) std::printf(" %s", r.QNAME.c_str());
        std::printf("\n");
        for (auto& row : lik) { for (double v : row) std::printf("%.17g ", v); std::printf("\n"); }
    } catch (const std::exception& e) {
        std::fprintf(stderr, "error: %s\n", e.what());
        return 1;
    }
    return 0;
}
