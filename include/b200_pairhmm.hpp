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

#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

#include "phmm.h"

namespace hc
{

// Process-wide engine (streams, memory pool, tables: created once, because the reference constructs its engine
// object per region, haplotypecaller.hpp:90, and a CUDA context must not be).
//
// Configuration, in order of precedence: B200Engine::configure() before the first use; else the environment:
//   PHMM_DEVICES=0,1,..    CUDA ordinals to shard regions over (default: device 0)
//   PHMM_EXACT=1           unfused FP32 arithmetic in the reference's operation order: raw FP32 sums bit-identical
//                          to compute_full_prob_avxs.  THIS is the mode that carries the bit-identical-VCF
//                          guarantee; the default (FMA-contracted) mode agrees to <= 1e-4 log10 and can in
//                          principle flip a `raw < 1e-28f` rescue decision or a GQ rounding.
//   PHMM_USE_DOUBLE=1      the reference's g_use_double switch (intel_pairhmm.hpp:58,71,135)
//   PHMM_REFERENCE_ORDER=1 (read by the library) the default engine's arithmetic in the reference's operation order,
//                          FMA-contracted, instead of the scaled recurrence (include/phmm.h: phmm_options.recurrence)
//   PHMM_HOST_THREADS=n    host threads per device for planning / packing / finalizing (default 4)
class B200Engine
{
public:
    struct Config
    {
        std::vector<int32_t> devices;      // empty = CUDA device 0
        bool exact = false;
        bool use_double = false;
        int pipeline_depth = 4;            // B200RegionBatcher keeps up to 3 batches in flight
        int host_threads = 4;

        bool operator==(const Config& o) const
        { return devices == o.devices && exact == o.exact && use_double == o.use_double &&
                 pipeline_depth == o.pipeline_depth && host_threads == o.host_threads; }

        static Config from_env()
        {
            Config c;
            if (const char* d = std::getenv("PHMM_DEVICES"))
                for (const char* q = d; *q;) {
                    char* end = nullptr;
                    const long v = std::strtol(q, &end, 10);
                    if (end == q) break;
                    c.devices.push_back((int32_t)v);
                    q = (*end == ',') ? end + 1 : end;
                }
            auto on = [](const char* name) { const char* v = std::getenv(name); return v && *v && *v != '0'; };
            c.exact = on("PHMM_EXACT");
            c.use_double = on("PHMM_USE_DOUBLE");
            if (const char* t = std::getenv("PHMM_HOST_THREADS")) c.host_threads = std::max(1, std::atoi(t));
            return c;
        }
    };

    // Must come before the first get(); a second, different configuration is an error, not a silent no-op.
    static void configure(const Config& c)
    {
        State& st = state();
        std::lock_guard<std::mutex> lk(st.mu);
        if (st.eng && !(st.cfg == c)) throw std::runtime_error("B200Engine::configure: the engine already exists with a different configuration");
        st.cfg = c; st.configured = true;
    }

    // The engine; `devices` non-empty must name the devices it was (or is now) created with.
    static phmm_engine* get(const std::vector<int32_t>& devices = {})
    {
        State& st = state();
        std::lock_guard<std::mutex> lk(st.mu);
        if (!st.eng) {
            if (!st.configured) { st.cfg = Config::from_env(); st.configured = true; }
            if (!devices.empty()) st.cfg.devices = devices;
            phmm_options opt{};
            opt.struct_size = (int32_t)sizeof(opt);
            opt.n_devices = st.cfg.devices.empty() ? 1 : (int32_t)st.cfg.devices.size();
            opt.devices = st.cfg.devices.empty() ? nullptr : st.cfg.devices.data();
            opt.pipeline_depth = st.cfg.pipeline_depth;
            opt.host_threads = st.cfg.host_threads;
            opt.exact_fp32 = st.cfg.exact ? 1 : 0;
            opt.use_double = st.cfg.use_double ? 1 : 0;
            int rc = phmm_create(&opt, &st.eng);
            if (rc != PHMM_OK) { st.eng = nullptr; throw std::runtime_error(std::string("phmm_create: ") + phmm_strerror(rc)); }
            // registered after the CUDA runtime's own exit handler (first touched inside phmm_create), so it
            // runs before it: streams, pinned memory and worker threads go away while the context is alive
            std::atexit([] { B200Engine::shutdown(); });
        } else if (!devices.empty() && devices != effective_devices(st.cfg)) {
            throw std::runtime_error("B200Engine::get: the engine already runs on a different device list");
        }
        return st.eng;
    }

    static const Config& config() { return state().cfg; }

    // Destroys the engine (all batchers must be drained and gone).  The next get() creates a new one.
    static void shutdown()
    {
        State& st = state();
        std::lock_guard<std::mutex> lk(st.mu);
        if (st.eng) { phmm_destroy(st.eng); st.eng = nullptr; }
    }

private:
    struct State { std::mutex mu; phmm_engine* eng = nullptr; Config cfg; bool configured = false; };
    static State& state() { static State* s = new State(); return *s; }     // never destructed: no static-order races at exit
    static std::vector<int32_t> effective_devices(const Config& c) { return c.devices.empty() ? std::vector<int32_t>{0} : c.devices; }
};

// Page-locked byte slab from the library (phmm_host_alloc): SoA arrays are gathered straight into it and
// uploaded from it (PHMM_BATCH_PINNED_INPUTS) -- written once by the caller, read once by the DMA engine.
class B200PinnedBytes
{
public:
    B200PinnedBytes() = default;
    B200PinnedBytes(const B200PinnedBytes&) = delete;
    B200PinnedBytes& operator=(const B200PinnedBytes&) = delete;
    ~B200PinnedBytes() { if (p_) phmm_host_free(p_); }
    void reserve(std::size_t n)
    {
        if (n <= cap_) return;
        // page-locking is slow (of the order of a millisecond per megabyte) and growing means doing it again: start at
        // 4 MB (a 1.6e10-cell batch of 150-base reads is about that) and double
        std::size_t want = std::max<std::size_t>(std::max(n, 2 * cap_), (std::size_t)4 << 20);
        void* q = nullptr;
        int rc = phmm_host_alloc(want, &q);
        if (rc != PHMM_OK) throw std::runtime_error(std::string("phmm_host_alloc: ") + phmm_strerror(rc));
        if (size_) std::memcpy(q, p_, size_);
        if (p_) phmm_host_free(p_);
        p_ = (uint8_t*)q; cap_ = want;
    }
    void append(const void* src, std::size_t n) { reserve(size_ + n); std::memcpy(p_ + size_, src, n); size_ += n; }
    void clear() { size_ = 0; }
    std::size_t size() const { return size_; }
    const uint8_t* data() const { return p_; }
private:
    uint8_t* p_ = nullptr;
    std::size_t cap_ = 0, size_ = 0;
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

// B200RegionBatcher -- cross-window batching for the caller side of the path.
//
// The reference's window loop (haplotypecaller.hpp:138-152) scores one small region at a time; on a
// B200 a region is well under a millisecond of device time, so one synchronous call per window is all
// launch and copy latency.  The batcher lets the caller hand over regions as the assembler produces
// them and collect the matrices later:
//
//   hc::B200RegionBatcher batcher;
//   for (window w) { ...filter, clip, assemble...;  ids[w] = batcher.add_region(haplotypes[w], reads[w]); }
//   for (window w) { auto likelihoods = batcher.take(ids[w], reads[w]);  genotyper...(reads[w], ..., likelihoods, ...); }
//
// add_region() copies the bytes the engine needs into the batch under construction and, once that
// holds `flush_cells` DP cells (or `flush_regions` regions), submits it asynchronously (phmm_submit);
// up to `max_in_flight` batches overlap upload, kernels and download while the caller keeps
// assembling.  take() returns exactly what B200PairHMM::compute_likelihoods would have returned for
// that region -- capped rows, poorly modelled reads erased from the caller's vector
// (intel_pairhmm.hpp:24-46) -- so the genotyper consumes it unchanged.  Regions may be taken in any
// order, each ONCE (a second take throws): the batch's result storage is released with its last region.
// One thread calls add_region()/take() (the engine has one submitter).
// One variant site of a region for the DEVICE-SIDE genotype reduction (phmm_submit_gl):
// what Genetyper::assign_genotype_likelihoods knows about the site before it looks at a likelihood
// (genotyper/genotyper.hpp:381-388): the allele count, the allele every haplotype carries (get_haplotype_mapper)
// and which reads overlap the site's interval (get_read_indices_to_keep).  INTEGRATION.md shows where the
// reference's loop is cut in two around it.
struct B200Site
{
    int32_t n_alleles = 0;
    std::vector<uint8_t> hap_allele;       // [haplotypes of the region]
    std::vector<uint8_t> read_overlap;     // [reads of the region as handed to add_region], 1 = overlaps
};

// What the device hands back for a region submitted with sites: per site the diploid genotype likelihoods
// (calculate_genotype_likelihoods, genotyper.hpp:322-327, bit for bit) in allele_index_cache order.
struct B200RegionGL
{
    std::vector<std::vector<double>> site_genotype_likelihoods;
    std::vector<int32_t> site_n_reads;     // reads that entered the sums
    std::vector<uint8_t> read_keep;        // 0 = poorly modelled read (the reference erases it, intel_pairhmm.hpp:35-45)
};

class B200RegionBatcher
{
public:
    // device_gl: regions carry their variant sites (add_region(haps, reads, sites)) and come back as genotype
    // likelihoods (take_gl) -- the reads x haplotypes matrix never leaves the GPU.
    explicit B200RegionBatcher(int64_t flush_cells = (int64_t)1.6e10, int32_t flush_regions = 4096, int max_in_flight = 3,
                               bool device_gl = false)
        : flush_cells_(flush_cells), flush_regions_(flush_regions), max_in_flight_(max_in_flight), device_gl_(device_gl),
          eng_(B200Engine::get()) {}
    ~B200RegionBatcher()
    {
        try { drain(); } catch (...) {}
        // page-locked slabs outlive the batcher: the next one (a new contig, a new chunk of windows) starts warm
        std::lock_guard<std::mutex> lk(slab_pool_mu());
        for (auto& sl : free_slabs_) if (sl && slab_pool().size() < 8) slab_pool().push_back(std::move(sl));
    }
    B200RegionBatcher(const B200RegionBatcher&) = delete;
    B200RegionBatcher& operator=(const B200RegionBatcher&) = delete;

    template <class HaplotypeT, class ReadT>
    int add_region(const std::vector<HaplotypeT>& haps, const std::vector<ReadT>& reads, const std::vector<B200Site>& sites)
    {
        if (!device_gl_) throw std::runtime_error("B200RegionBatcher: sites need a batcher constructed with device_gl = true");
        for (const auto& st : sites)
            if (st.hap_allele.size() != haps.size() || st.read_overlap.size() != reads.size() || st.n_alleles < 1 || st.n_alleles > PHMM_MAX_ALLELES)
                throw std::runtime_error("B200RegionBatcher: site arrays do not match the region");
        pending_sites_ = &sites;
        const int id = add_region(haps, reads);
        pending_sites_ = nullptr;
        return id;
    }

    template <class HaplotypeT, class ReadT>
    int add_region(const std::vector<HaplotypeT>& haps, const std::vector<ReadT>& reads)
    {
        if (device_gl_ && !pending_sites_) throw std::runtime_error("B200RegionBatcher: a device_gl batcher takes regions with their sites");
        for (const auto& r : reads)           // checked before anything is appended: a refused region leaves the batch as it was
            if (r.SEQ.size() != r.QUAL.size()) throw std::runtime_error("B200RegionBatcher: SEQ and QUAL lengths differ");
        if (!cur_) {
            cur_ = std::make_unique<Pending>();
            if (!free_slabs_.empty()) { cur_->slab = std::move(free_slabs_.back()); free_slabs_.pop_back(); }
            else {
                std::lock_guard<std::mutex> lk(slab_pool_mu());
                if (!slab_pool().empty()) { cur_->slab = std::move(slab_pool().back()); slab_pool().pop_back(); }
            }
            if (!cur_->slab) cur_->slab = std::make_unique<Slabs>();
        }
        Pending& b = *cur_;
        Slabs& sl = *b.slab;
        if (b.region_read_beg.empty()) { b.gl_off.assign(1, 0); b.region_site_beg.assign(1, 0); }
        if (b.region_read_beg.empty()) { b.region_read_beg.push_back(0); b.region_hap_beg.push_back(0); b.read_off.push_back(0); b.hap_off.push_back(0); b.out_beg.push_back(0); }
        int64_t read_bytes = 0, hap_bytes = 0;
        for (const auto& r : reads) {
            sl.read_bases.append(r.SEQ.data(), r.SEQ.size());       // gathered straight into page-locked memory
            sl.read_q.append(r.QUAL.data(), r.QUAL.size());
            b.read_off.push_back((int32_t)sl.read_bases.size());
            read_bytes += (int64_t)r.SEQ.size();
        }
        for (const auto& h : haps) {
            sl.hap_bases.append(h.bases.data(), h.bases.size());
            b.hap_off.push_back((int32_t)sl.hap_bases.size());
            hap_bytes += (int64_t)h.bases.size();
        }
        b.region_read_beg.push_back((int32_t)(b.read_off.size() - 1));
        b.region_hap_beg.push_back((int32_t)(b.hap_off.size() - 1));
        b.out_beg.push_back(b.out_beg.back() + (int64_t)reads.size() * (int64_t)haps.size());
        b.cells += read_bytes * hap_bytes;
        if (device_gl_) {
            const int32_t region = (int32_t)(b.region_read_beg.size() - 2);
            for (const auto& st : *pending_sites_) {
                b.site_region.push_back(region);
                b.site_n_alleles.push_back(st.n_alleles);
                b.hap_allele.insert(b.hap_allele.end(), st.hap_allele.begin(), st.hap_allele.end());
                b.read_overlap.insert(b.read_overlap.end(), st.read_overlap.begin(), st.read_overlap.end());
                b.gl_off.push_back(b.gl_off.back() + (int64_t)st.n_alleles * (st.n_alleles + 1) / 2);
            }
            b.region_site_beg.push_back((int32_t)b.site_region.size());
        }
        const int id = (int)where_.size();
        where_.push_back({next_batch_id_, (int32_t)(b.region_read_beg.size() - 2)});
        if (b.cells >= flush_cells_ || (int32_t)(b.region_read_beg.size() - 1) >= flush_regions_) flush();
        return id;
    }

    // submit the batch under construction now (called automatically by add_region / take / drain)
    void flush()
    {
        if (!cur_ || cur_->region_read_beg.size() <= 1) return;
        while ((int)in_flight_ >= max_in_flight_) wait_oldest();
        Pending& b = *cur_;
        phmm_batch pb{};
        pb.n_regions = (int32_t)(b.region_read_beg.size() - 1);
        pb.n_reads = (int32_t)(b.read_off.size() - 1);
        pb.n_haps = (int32_t)(b.hap_off.size() - 1);
        pb.region_read_beg = b.region_read_beg.data();
        pb.region_hap_beg = b.region_hap_beg.data();
        pb.read_off = b.read_off.data();
        pb.read_bases = b.slab->read_bases.data();
        pb.read_q = b.slab->read_q.data();
        pb.flags = PHMM_BATCH_PINNED_INPUTS;                            // uploaded in place; the slab lives until the wait
        pb.read_i = pb.read_d = pb.read_c = nullptr;                   // constant strings of sam/sam.hpp:30-32
        pb.gap_open_i = (uint8_t)B200PairHMM::GAP_OPEN; pb.gap_open_d = (uint8_t)B200PairHMM::GAP_OPEN;
        pb.gap_cont_c = (uint8_t)B200PairHMM::GAP_CONT;
        pb.hap_off = b.hap_off.data();
        pb.hap_bases = b.slab->hap_bases.data();
        int rc;
        if (device_gl_) {
            phmm_sites ps{};
            ps.n_sites = (int32_t)b.site_region.size();
            ps.site_region = b.site_region.data(); ps.site_n_alleles = b.site_n_alleles.data();
            ps.hap_allele = b.hap_allele.data(); ps.read_overlap = b.read_overlap.data();
            b.gl.resize((size_t)b.gl_off.back()); b.site_n_reads.resize(b.site_region.size()); b.read_keep.resize((size_t)pb.n_reads);
            rc = phmm_submit_gl(eng_, &pb, &ps, &b.ticket);
        } else {
            b.lik.resize((size_t)b.out_beg.back());
            rc = phmm_submit(eng_, &pb, &b.ticket);
        }
        if (rc != PHMM_OK) throw std::runtime_error(std::string("phmm_submit: ") + phmm_strerror(rc) + ": " + phmm_last_error(eng_));
        b.submitted = true;
        ++in_flight_; ++batches_submitted;
        batches_.push_back(std::move(cur_));
        ++next_batch_id_;
    }

    template <class ReadT>
    std::vector<std::vector<double>> take(int region_id, std::vector<ReadT>& reads)
    {
        if (region_id < 0 || region_id >= (int)where_.size()) throw std::runtime_error("B200RegionBatcher: unknown region id");
        const Where w = where_[region_id];
        if (w.batch == next_batch_id_) flush();                          // still under construction
        Pending& b = *batches_.at((size_t)w.batch);
        if (b.region_read_beg.empty()) throw std::runtime_error("B200RegionBatcher: region taken twice");
        const int32_t r0 = b.region_read_beg[w.region], r1 = b.region_read_beg[w.region + 1];
        const int32_t h0 = b.region_hap_beg[w.region], h1 = b.region_hap_beg[w.region + 1];
        const std::size_t n_reads = (std::size_t)(r1 - r0), n_haps = (std::size_t)(h1 - h0);
        if (n_reads != reads.size()) throw std::runtime_error("B200RegionBatcher: reads vector differs from the one added");
        mark_taken(b, w.region);
        while (!b.done) wait_oldest();
        Releaser release_when_done{&b};
        if (n_reads == 0 || n_haps == 0) return std::vector<std::vector<double>>(n_reads, std::vector<double>(n_haps));
        double* flat = b.lik.data() + b.out_beg[w.region];
        std::vector<int32_t> read_len(n_reads);
        for (std::size_t r = 0; r < n_reads; r++) read_len[r] = b.read_off[r0 + r + 1] - b.read_off[r0 + r];
        std::vector<uint8_t> keep(n_reads);
        phmm_normalize_filter(flat, (int32_t)n_reads, (int32_t)n_haps, read_len.data(), keep.data());   // :24-46
        std::vector<std::vector<double>> out;
        out.reserve(n_reads);
        std::size_t k = 0;
        for (std::size_t r = 0; r < n_reads; r++) {
            if (!keep[r]) continue;
            out.emplace_back(flat + r * n_haps, flat + (r + 1) * n_haps);       // one allocation per surviving row
            if (k != r) reads[k] = std::move(reads[r]);
            ++k;
        }
        reads.erase(reads.begin() + k, reads.end());
        return out;
    }

    // The genotype likelihoods of a region added with its sites (device_gl batchers), each region once.
    B200RegionGL take_gl(int region_id)
    {
        if (!device_gl_) throw std::runtime_error("B200RegionBatcher: take_gl needs device_gl = true");
        if (region_id < 0 || region_id >= (int)where_.size()) throw std::runtime_error("B200RegionBatcher: unknown region id");
        const Where w = where_[region_id];
        if (w.batch == next_batch_id_) flush();
        Pending& b = *batches_.at((size_t)w.batch);
        mark_taken(b, w.region);
        while (!b.done) wait_oldest();
        Releaser release_when_done{&b};
        B200RegionGL out;
        const int32_t s0 = b.region_site_beg[w.region], s1 = b.region_site_beg[w.region + 1];
        for (int32_t k = s0; k < s1; k++) {
            out.site_genotype_likelihoods.emplace_back(b.gl.begin() + b.gl_off[k], b.gl.begin() + b.gl_off[k + 1]);
            out.site_n_reads.push_back(b.site_n_reads[k]);
        }
        out.read_keep.assign(b.read_keep.begin() + b.region_read_beg[w.region], b.read_keep.begin() + b.region_read_beg[w.region + 1]);
        return out;
    }

    // wait for everything submitted so far
    void drain() { flush(); while (in_flight_) wait_oldest(); }

    int batches_submitted = 0;
    phmm_stats total_stats{};          // summed over finished batches (kernel_ms: sum, not max)

private:
    struct Slabs {
        B200PinnedBytes read_bases, read_q, hap_bases;
        void clear() { read_bases.clear(); read_q.clear(); hap_bases.clear(); }
    };
    // process-wide pool of idle slabs (never destructed: page-locked memory must not be freed after the CUDA runtime)
    static std::mutex& slab_pool_mu() { static std::mutex* m = new std::mutex(); return *m; }
    static std::vector<std::unique_ptr<Slabs>>& slab_pool() { static auto* v = new std::vector<std::unique_ptr<Slabs>>(); return *v; }
    struct Pending {
        std::vector<int32_t> region_read_beg, region_hap_beg, read_off, hap_off;
        std::vector<int64_t> out_beg;
        std::unique_ptr<Slabs> slab;          // page-locked byte arrays; back to the free list once the batch is done
        std::vector<double> lik;
        // device_gl: the sites of the batch and what comes back for them
        std::vector<int32_t> site_region, site_n_alleles, region_site_beg, site_n_reads;
        std::vector<uint8_t> hap_allele, read_overlap, read_keep;
        std::vector<int64_t> gl_off;
        std::vector<double> gl;
        std::vector<uint8_t> taken;           // per region: handed out already
        int32_t n_taken = 0;
        int64_t cells = 0;
        phmm_ticket ticket = 0;
        bool submitted = false, done = false;
    };
    struct Where { int batch; int32_t region; };

    // every region is taken exactly once; the batch's storage goes when the last one has been
    static void mark_taken(Pending& b, int32_t region)
    {
        if (b.region_read_beg.empty()) throw std::runtime_error("B200RegionBatcher: region taken twice");   // released batch
        const std::size_t n = b.region_read_beg.size() - 1;
        if (b.taken.size() != n) b.taken.assign(n, 0);
        if (b.taken[(std::size_t)region]) throw std::runtime_error("B200RegionBatcher: region taken twice");
        b.taken[(std::size_t)region] = 1;
        ++b.n_taken;
    }
    struct Releaser {                          // runs when take()/take_gl() return: a whole-genome run must not keep every matrix
        Pending* b;
        ~Releaser()
        {
            if (b->n_taken != (int32_t)(b->region_read_beg.size() - 1)) return;
            Pending fresh;
            fresh.submitted = b->submitted; fresh.done = b->done; fresh.ticket = b->ticket;
            *b = std::move(fresh);             // vectors freed; the shell stays so that batch ids keep indexing batches_
        }
    };

    void wait_oldest()
    {
        for (auto& pb : batches_) {
            Pending& b = *pb;
            if (!b.submitted || b.done) continue;
            phmm_result res{};
            int rc;
            if (device_gl_) {
                phmm_gl_result gr{};
                gr.genotype_lik = b.gl.data(); gr.site_n_reads = b.site_n_reads.data(); gr.read_keep = b.read_keep.data();
                rc = phmm_wait_gl(eng_, b.ticket, &gr);
                res.stats = gr.stats;
            } else {
                res.log10_lik = b.lik.data();
                rc = phmm_wait(eng_, b.ticket, &res);
            }
            if (rc != PHMM_OK) throw std::runtime_error(std::string("phmm_wait: ") + phmm_strerror(rc) + ": " + phmm_last_error(eng_));
            b.done = true; --in_flight_;
            b.slab->clear(); free_slabs_.push_back(std::move(b.slab));
            total_stats.n_pairs += res.stats.n_pairs; total_stats.n_cells += res.stats.n_cells;
            total_stats.n_rescued += res.stats.n_rescued; total_stats.h2d_bytes += res.stats.h2d_bytes;
            total_stats.d2h_bytes += res.stats.d2h_bytes; total_stats.kernel_launches += res.stats.kernel_launches;
            total_stats.kernel_ms += res.stats.kernel_ms;
            return;
        }
        throw std::runtime_error("B200RegionBatcher: nothing in flight");
    }

    int64_t flush_cells_;
    int32_t flush_regions_;
    int max_in_flight_;
    bool device_gl_;
    const std::vector<B200Site>* pending_sites_ = nullptr;
    phmm_engine* eng_;
    std::unique_ptr<Pending> cur_;
    std::vector<std::unique_ptr<Slabs>> free_slabs_;
    std::deque<std::unique_ptr<Pending>> batches_;
    std::vector<Where> where_;
    int next_batch_id_ = 0;
    int in_flight_ = 0;
};

} // namespace hc
