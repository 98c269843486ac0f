// vcf_slice.cpp -- VCF-level oracle: the reference's own post-assembly stages around EITHER engine.
//
// TEST INFRASTRUCTURE ONLY (compiled into oracle/_ref/vcf_slice from the reference's headers where
// they lie under /root/reference; git-ignored; travels to the GPU box with the snapshot).
//
// the full program cannot be built here (Boost.Graph / program_options are
// absent), but everything AFTER the assembler can.  For each region of the input file this harness
//   1. SW-aligns every candidate haplotype to the reference window with hc::IntelSWAligner, filling
//      cigar / alignment_begin_wrt_ref exactly like assembler/graph_wrapper.hpp:232-239,
//   2. scores reads x haplotypes with `--engine ref`  -> hc::IntelPairHMM   (the reference, CPU)
//                                  or `--engine b200` -> hc::B200PairHMM    (this repo, GPU, C ABI),
//      both through compute_likelihoods(haplotypes, reads) as haplotypecaller.hpp:103 does,
//   3. genotypes with hc::Genetyper::assign_genotype_likelihoods (haplotypecaller.hpp:104),
//   4. prints every variant with hc::Variant::print (the VCF body lines, :105-106).
// tests/test_vcf_slice.py diffs the two outputs byte for byte.  With --dump FILE the per-pair
// likelihood matrix (post cap/filter, %.17g) is written too.
//
// Input format (one or more regions):
//   REGION <contig> <padded_begin> <padded_end> <origin_begin> <origin_end>      (0-based, half open)
//   REF <bases of the padded window>
//   H <haplotype bases>                       (first haplotype = reference path)
//   R <POS 1-based> <CIGAR> <SEQ> <QUAL>
//   END
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

#include "pairhmm/intel_pairhmm.hpp"
#include "smithwaterman/intel_smithwaterman.hpp"
#include "genotyper/genotyper.hpp"
#include "b200_pairhmm.hpp"

template <class Engine>
static int run(std::istream& in, std::ostream& vcf, std::ostream* dump)
{
    std::string line, contig, ref;
    std::size_t pb = 0, pe = 0, ob = 0, oe = 0;
    std::vector<hc::Haplotype> haps;
    std::vector<hc::SAMRecord> reads;
    int n_regions = 0;
    while (std::getline(in, line)) {
        std::istringstream is(line);
        std::string tag; is >> tag;
        if (tag == "REGION") { is >> contig >> pb >> pe >> ob >> oe; haps.clear(); reads.clear(); ref.clear(); }
        else if (tag == "REF") is >> ref;
        else if (tag == "H") { hc::Haplotype h; is >> h.bases; haps.push_back(std::move(h)); }
        else if (tag == "R") {
            hc::SAMRecord r; std::string cigar;
            is >> r.POS >> cigar >> r.SEQ >> r.QUAL;
            r.QNAME = "r" + std::to_string(reads.size()); r.FLAG = 0; r.RNAME = contig; r.MAPQ = 60;
            r.CIGAR = cigar; r.RNEXT = "="; r.PNEXT = 0; r.TLEN = 0;
            reads.push_back(std::move(r));
        } else if (tag == "END") {
            hc::Interval padded(contig, pb, pe), origin(contig, ob, oe);
            hc::IntelSWAligner aligner;                              // graph_wrapper.hpp:232-239
            for (auto& h : haps) {
                auto [alignment_begin, cigar] = aligner.align(ref, h.bases);
                h.alignment_begin_wrt_ref = alignment_begin;
                h.cigar = std::move(cigar);
            }
            Engine pairhmm;                                          // haplotypecaller.hpp:90
            hc::Genetyper genotyper;
            auto likelihoods = pairhmm.compute_likelihoods(haps, reads);                          // :103
            if (dump) {
                *dump << "REGION " << n_regions << " kept " << reads.size() << "\n";
                for (std::size_t r = 0; r < likelihoods.size(); r++) {
                    *dump << reads[r].QNAME;
                    char buf[40];
                    for (double v : likelihoods[r]) { std::snprintf(buf, sizeof buf, " %.17g", v); *dump << buf; }
                    *dump << "\n";
                }
            }
            auto variants = genotyper.assign_genotype_likelihoods(reads, haps, likelihoods, ref, padded, origin);  // :104
            for (const auto& v : variants) v.print(vcf);                                          // :105-106
            ++n_regions;
        }
    }
    return n_regions;
}

int main(int argc, char** argv)
{
    std::string engine = "ref", input, dump_path;
    for (int i = 1; i < argc; i++) {
        if (!std::strcmp(argv[i], "--engine") && i + 1 < argc) engine = argv[++i];
        else if (!std::strcmp(argv[i], "--dump") && i + 1 < argc) dump_path = argv[++i];
        else input = argv[i];
    }
    if (input.empty()) { std::fprintf(stderr, "usage: vcf_slice [--engine ref|b200] [--dump file] regions.txt\n"); return 2; }
    std::ifstream in(input);
    if (!in) { std::fprintf(stderr, "cannot open %s\n", input.c_str()); return 2; }
    std::ofstream dump;
    if (!dump_path.empty()) dump.open(dump_path);
    try {
        int n = engine == "b200" ? run<hc::B200PairHMM>(in, std::cout, dump_path.empty() ? nullptr : &dump)
                                 : run<hc::IntelPairHMM>(in, std::cout, dump_path.empty() ? nullptr : &dump);
        std::fprintf(stderr, "vcf_slice: %d regions through engine %s\n", n, engine.c_str());
    } catch (const std::exception& e) {
        std::fprintf(stderr, "vcf_slice: error: %s\n", e.what());
        return 1;
    }
    return 0;
}
