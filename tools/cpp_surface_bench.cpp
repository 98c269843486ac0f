// cpp_surface_bench.cpp -- end-to-end throughput at the REFERENCE-FACING C++ surface (include/b200_pairhmm.hpp), from
// the data model the reference's caller holds: one std::string SEQ / QUAL per read, one std::string per haplotype
// (sam/sam.hpp:14-28, haplotype/haplotype.hpp).  S3 regions (150 bp x 500 bp, 256 reads x 16 haplotypes, synthetic,
// std::mt19937_64) go
//   (a) through hc::B200RegionBatcher: add_region() per region, take() per region (cap, filter, rows erased) --
//       what oracle/hc_e2e.cpp -DHC_BATCHED does; timed from the first add_region to the last take;
//   (b) through hc::B200PairHMM::compute_likelihoods, one synchronous call per region (the one-line swap).
// Prints one JSON object.  bench.py runs it (N = 1) and reports it as e2e_cpp_surface.
//   g++ -std=c++17 -O2 -Iinclude tools/cpp_surface_bench.cpp -o tools/cpp_surface_bench -L<pkg> -lphmm_b200 -Wl,-rpath,<pkg> -pthread
#include <chrono>
#include <cstdio>
#include <random>
#include <string>
#include <vector>

#include "b200_pairhmm.hpp"

struct Haplotype { std::string bases; };
struct SAMRecord { std::string SEQ, QUAL; std::size_t size() const { return SEQ.size(); } };
struct Region { std::vector<Haplotype> haps; std::vector<SAMRecord> reads; };

static Region make_region(std::mt19937_64& rng, int n_reads = 256, int n_haps = 16, int R = 150, int H = 500)
{
    static const char acgt[] = "ACGT";
    Region g;
    std::string backbone(H, 'A');
    for (auto& c : backbone) c = acgt[rng() & 3];
    for (int h = 0; h < n_haps; h++) {
        Haplotype hp{backbone};
        for (int s = 0; s < 3; s++) { const size_t p = rng() % H; hp.bases[p] = acgt[(rng() & 3)]; }
        g.haps.push_back(std::move(hp));
    }
    for (int r = 0; r < n_reads; r++) {
        const std::string& src = g.haps[rng() % n_haps].bases;
        const size_t o = rng() % (H - R + 1);
        SAMRecord rec{src.substr(o, R), std::string(R, 'I')};
        for (int i = 0; i < R; i++) {
            if (rng() % 100 == 0) rec.SEQ[i] = acgt[rng() & 3];
            rec.QUAL[i] = (char)(33 + 20 + rng() % 21);
        }
        g.reads.push_back(std::move(rec));
    }
    return g;
}

int main(int argc, char** argv)
{
    const int n_regions = argc > 1 ? std::atoi(argv[1]) : 1024;
    const int per_window_regions = argc > 2 ? std::atoi(argv[2]) : 64;
    try {
        std::mt19937_64 rng(1003);
        std::vector<Region> regions;
        for (int i = 0; i < n_regions; i++) regions.push_back(make_region(rng));
        const double cells = (double)n_regions * 256.0 * 16 * 150 * 500;
        hc::B200Engine::get();
        double best = 1e30;
        long long kept = 0;
        for (int rep = 0; rep < 3; rep++) {                      // (rep 0 also warms the engine's pools)
            std::vector<Region> work = regions;                   // take() erases reads: a fresh copy per repetition (untimed)
            const auto t0 = std::chrono::steady_clock::now();
            hc::B200RegionBatcher batcher;
            std::vector<int> ids(work.size());
            for (size_t i = 0; i < work.size(); i++) ids[i] = batcher.add_region(work[i].haps, work[i].reads);
            kept = 0;
            for (size_t i = 0; i < work.size(); i++) kept += (long long)batcher.take(ids[i], work[i].reads).size();
            const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            if (rep) best = std::min(best, dt);
        }
        double best_pw = 1e30;
        for (int rep = 0; rep < 2; rep++) {
            std::vector<Region> work(regions.begin(), regions.begin() + std::min(per_window_regions, n_regions));
            const auto t0 = std::chrono::steady_clock::now();
            for (auto& g : work) { hc::B200PairHMM pairhmm; auto lik = pairhmm.compute_likelihoods(g.haps, g.reads); kept += (long long)lik.size(); }
            best_pw = std::min(best_pw, std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());
        }
        const double cells_pw = cells * std::min(per_window_regions, n_regions) / n_regions;
        std::printf("{\"regions\": %d, \"cells\": %.4e, \"batcher_gcups\": %.1f, \"batcher_s\": %.4f, \"per_region_call_gcups\": %.1f, "
                    "\"per_region_call_ms\": %.3f, \"reads_kept_last\": %lld, "
                    "\"what\": \"S3 regions held as one std::string per read / haplotype; (a) hc::B200RegionBatcher add_region + take "
                    "(gather into page-locked slabs, cross-region batches of 1.6e10 cells, 3 in flight, cap + filter + row erase on the way out), "
                    "(b) hc::B200PairHMM::compute_likelihoods, one synchronous call per region\"}\n",
                    n_regions, cells, cells / best / 1e9, best, cells_pw / best_pw / 1e9, 1e3 * best_pw / std::min(per_window_regions, n_regions), kept);
    } catch (const std::exception& e) {
        std::fprintf(stderr, "error: %s\n", e.what());
        return 1;
    }
    return 0;
}
