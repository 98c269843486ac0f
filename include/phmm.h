/*
 * phmm.h -- C ABI of the B200-native PairHMM forward engine (libphmm_b200.so).
 *
 * This is the drop-in boundary for ONE path of avis9ditiu/gatk-haplotypecaller-cpp17: the
 * read x haplotype log10-likelihood step.  Reference citations are relative to
 * /root/reference/src/haplotypecaller/.
 *
 * What each entry point replaces in the reference:
 *   phmm_create / phmm_destroy   hc::IntelPairHMM construction + initNative()
 *                                (pairhmm/intel_pairhmm.hpp:16, :77-113): table build
 *                                (native/Context.h:101-115,141-155), base-code table
 *                                (native/pairhmm_common.h:26-44).  Here: once per process,
 *                                not once per region (haplotypecaller.hpp:90).
 *   phmm_batch                   getData()'s testcase[read][hap] pointer grid
 *                                (intel_pairhmm.hpp:154-203, native/pairhmm_common.h:20-24),
 *                                laid out as the SoA form of the accelerator batch the reference
 *                                declares but never calls (native/shacc_pairhmm.h:12-35), and
 *                                widened to many regions per batch.
 *   phmm_compute                 computeLikelihoodsNative() (intel_pairhmm.hpp:115-152): FP32
 *                                forward per pair, FP64 redo iff raw < 1e-28f, log10 minus the
 *                                scaling constant.  The kernels replace compute_full_prob_avxs /
 *                                compute_full_prob_avxd (native/avx-pairhmm-template.h:210-346).
 *   phmm_submit / phmm_wait      no reference analogue (its window loop is serial,
 *                                haplotypecaller.hpp:138-152): asynchronous form of phmm_compute
 *                                so batch N+1 uploads while batch N computes.
 *   phmm_normalize_filter        normalize_likelihoods_and_filter_poorly_modeled_reads()
 *                                (intel_pairhmm.hpp:24-46), host-side, unchanged semantics.
 *
 * Semantics that are deliberately the reference's:
 *   - quality, gap-open and gap-continuation bytes are consumed RAW (ASCII, & 127), no Phred+33
 *     subtraction (native/avx-pairhmm-template.h:110-112,125);
 *   - bases map A,C,T,G,N -> codes, every other byte (incl. lower case) -> 'A'
 *     (native/pairhmm_common.h:26-44); N on either side matches;
 *   - FP32 with flush-to-zero, result scaled by 2^120; FP64 scaled by 2^1020.
 *
 * There is no CPU fallback: every compute entry point fails with PHMM_ERR_CUDA /
 * PHMM_ERR_NO_DEVICE when no sm_100 device is usable.
 */
#ifndef PHMM_H
#define PHMM_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PHMM_ABI_VERSION 1

/* status codes (reference has none on this path: bad input there is UB) */
enum {
    PHMM_OK              = 0,
    PHMM_ERR_INVALID_ARG = 1,
    PHMM_ERR_NO_DEVICE   = 2,
    PHMM_ERR_CUDA        = 3,
    PHMM_ERR_OOM         = 4,
    PHMM_ERR_UNSUPPORTED = 5,   /* e.g. read longer than PHMM_MAX_READ_LEN */
    PHMM_ERR_BAD_TICKET  = 6
};

#define PHMM_MAX_READ_LEN 2048    /* reads of up to 255 bases take the register-tiled kernels, longer ones a
                                     slower one-warp-per-pair kernel; beyond this: PHMM_ERR_UNSUPPORTED       */
#define PHMM_MAX_HAP_LEN  8192    /* columns staged in shared memory per lane group            */

typedef struct phmm_engine phmm_engine;
typedef int64_t phmm_ticket;

typedef struct phmm_options {
    int32_t struct_size;        /* sizeof(phmm_options), for forward compatibility             */
    int32_t n_devices;          /* 0 or 1: one device; >1: shard regions over devices[]        */
    const int32_t* devices;     /* CUDA ordinals; NULL = 0..n_devices-1                        */
    int32_t pipeline_depth;     /* in-flight batches per device (streams + buffers); 0 = 2     */
    int32_t exact_fp32;         /* 1: unfused mul/add, raw FP32 bit-identical to the reference */
    int32_t host_threads;       /* threads for planning, packing and finalizing a batch; 0 = 1 */
    int32_t use_double;         /* 1: the reference's g_use_double switch (intel_pairhmm.hpp:58,71,135): the FP32
                                   pass is skipped (its raw result is 0.0f) and EVERY pair is scored in FP64  */
    int32_t fp64_first;         /* order of the precision passes.  0 (auto): FP32 first; FP64 FIRST for a batch
                                   when the device's previous batch redid > 75% of its pairs in FP64 -- every
                                   pair is then scored in FP64 and the FP32 pass only runs where the FP64 sum
                                   does not prove the underflow (identical log10 values and rescue decisions;
                                   the optional raw32 output of a proven underflow is 0.0f).  1: never.  2: always.
                                   Ignored (never) with exact_fp32, whose raw FP32 sums are part of the contract. */
    int32_t recurrence;         /* arithmetic of the default (non-exact) FP32 / FP64 kernels for constant gap penalties
                                   with i == d -- the reference's only case.  0: the SCALED recurrence where the layout
                                   allows it (carries X / pMX and Y / pMY and folds pMM into the priors: five FP32-pipe
                                   instructions per cell instead of seven; same mathematics, different rounding, well
                                   inside 1e-4 on the log10; gap-open penalties Q10..Q96, the reference order otherwise).
                                   1: the reference's operation order, FMA-contracted (round 1's kernels).
                                   exact_fp32 overrides both.  Environment: PHMM_REFERENCE_ORDER=1 forces 1. */
} phmm_options;

/*
 * One batch = many active regions; every read of a region is scored against every haplotype of
 * that region.  All arrays are caller-owned and may be released after phmm_submit returns.
 *   region g: reads  [region_read_beg[g], region_read_beg[g+1])
 *             haps   [region_hap_beg[g],  region_hap_beg[g+1])
 *   read r:   bytes  [read_off[r], read_off[r+1]) of read_bases/read_q/read_i/read_d/read_c
 *   hap h:    bytes  [hap_off[h], hap_off[h+1]) of hap_bases
 * read_i/read_d/read_c may all three be NULL: then gap_open_i/gap_open_d/gap_cont_c apply to
 * every base (the reference's constant 'I' / 'I' / '+' strings, sam/sam.hpp:30-32,47-49).
 * Output order: region g occupies a row-major [reads_g][haps_g] block starting at
 * sum_{g'<g} reads_g' * haps_g'.
 */
typedef struct phmm_batch {
    int32_t n_regions, n_reads, n_haps;
    const int32_t* region_read_beg;   /* [n_regions+1] */
    const int32_t* region_hap_beg;    /* [n_regions+1] */
    const int32_t* read_off;          /* [n_reads+1]   */
    const uint8_t* read_bases;
    const uint8_t* read_q;
    const uint8_t* read_i;            /* insertion gap-open, per base, or NULL */
    const uint8_t* read_d;            /* deletion gap-open, per base, or NULL  */
    const uint8_t* read_c;            /* gap continuation, per base, or NULL   */
    const int32_t* hap_off;           /* [n_haps+1]    */
    const uint8_t* hap_bases;
    uint8_t gap_open_i, gap_open_d, gap_cont_c;               /* used when read_i == NULL */
    uint8_t flags;                                           /* PHMM_BATCH_* */
} phmm_batch;

/* phmm_batch.flags.  PINNED_INPUTS: read_bases, read_q, (read_i, read_d, read_c) and hap_bases are page-locked
 * (cudaMallocHost / cudaHostRegister, e.g. phmm_host_register below) and stay valid and unchanged until
 * phmm_wait returns: the engine then uploads them straight from the caller's memory instead of copying them
 * into its own pinned staging first (the index arrays are still copied and may be released at once).  Worth
 * it for read-heavy streams on many GPUs, where that copy is the host-memory-bandwidth limit. */
#define PHMM_BATCH_PINNED_INPUTS 1
int  phmm_host_register(void* p, size_t bytes);      /* cudaHostRegister / cudaHostUnregister for callers   */
int  phmm_host_unregister(void* p);                  /* without a CUDA toolchain of their own               */
/* Page-locked host memory handed out by the library (cudaHostAlloc, portable across devices): a caller that
 * GATHERS its reads into the SoA arrays anyway (the reference's reads are one std::string each,
 * sam/sam.hpp:14-28) gathers straight into such a slab and submits with PHMM_BATCH_PINNED_INPUTS -- the bytes
 * are then written once by the caller and read once by the DMA engine, with no staging copy in between.
 * hc::B200RegionBatcher (include/b200_pairhmm.hpp) does exactly that. */
int  phmm_host_alloc(size_t bytes, void** out);
int  phmm_host_free(void* p);

typedef struct phmm_stats {
    int64_t n_pairs, n_cells, n_rescued;
    int64_t h2d_bytes, d2h_bytes;
    int32_t kernel_launches;
    int32_t n_devices_used;
    float   kernel_ms;      /* device time of the forward kernels (CUDA events), max over devices */
    float   total_ms;       /* host wall time submit -> results written                          */
} phmm_stats;

typedef struct phmm_result {
    double*  log10_lik;     /* [n_pairs] required: final log10 likelihoods (pre cap/filter)     */
    float*   raw32;         /* [n_pairs] optional: raw FP32 forward sums (scaled by 2^120)      */
    double*  raw64;         /* [n_pairs] optional: raw FP64 sums of rescued pairs, 0 elsewhere  */
    uint8_t* rescued;       /* [n_pairs] optional: 1 where the FP64 redo was taken              */
    phmm_stats stats;       /* out */
} phmm_result;

/*
 * Threading contract.  phmm_create / phmm_destroy: one thread, nothing else in flight.  phmm_submit (and
 * phmm_compute, phmm_stage, phmm_run_staged*) : ONE submitting thread at a time per engine.  phmm_wait may run
 * on a different thread than phmm_submit (producer / consumer): the slot ring, the ticket table and the error
 * string are guarded by the engine's mutex; tickets may be waited in any order, each once.
 * phmm_last_error returns a copy private to the calling thread, valid until that thread's next call of it.
 * Every entry point leaves the calling thread's current CUDA device unchanged.
 */
int  phmm_create(const phmm_options* opt, phmm_engine** out);
void phmm_destroy(phmm_engine* e);
int  phmm_compute(phmm_engine* e, const phmm_batch* b, phmm_result* r);
int  phmm_submit(phmm_engine* e, const phmm_batch* b, phmm_ticket* t);
int  phmm_wait(phmm_engine* e, phmm_ticket t, phmm_result* r);
const char* phmm_strerror(int code);
const char* phmm_last_error(const phmm_engine* e);
int  phmm_abi_version(void);

/*
 * Device-side consumer of the matrix: per variant site, the diploid genotype likelihoods
 * hc::Genetyper computes from the capped / filtered matrix (genotyper/genotyper.hpp:245-328: marginal_likelihoods,
 * calculate_genotype_likelihoods, with utils/math_utils.hpp:11-30) -- bit for bit, so that the reads x haplotypes
 * matrix never has to leave the GPU: the download shrinks to A (A + 1) / 2 doubles per site.  The host keeps what
 * does not depend on the likelihoods (events, alleles, haplotype -> allele maps: genotyper.hpp:111-233, handed in
 * as phmm_sites) and what consumes the vector (genotype quality and the call, :329-368).
 *   site s belongs to region site_region[s] (non-decreasing) and has site_n_alleles[s] alleles (allele 0 = REF);
 *   hap_allele: for every site in order, one byte per haplotype OF ITS REGION = the allele that haplotype carries
 *               (Genetyper::get_haplotype_mapper, genotyper.hpp:224-232);
 *   read_overlap: for every site in order, one byte per read OF ITS REGION, 1 = the read overlaps the site's
 *               interval (get_read_indices_to_keep, :235-244), or NULL when every read does.
 * Output: genotype_lik holds for every site in order its A (A + 1) / 2 likelihoods in the reference's genotype
 * order ((a1, a2), a1 <= a2, a1 outer: genotyper.hpp:22-33).  Reads the poorly-modelled filter drops
 * (intel_pairhmm.hpp:35-38) do not enter the sums, exactly as the reference erases them before genotyping.
 * Needs a host whose libm is the glibc this library restates (phmm_log10.h; checked at phmm_create), otherwise
 * PHMM_ERR_UNSUPPORTED: the host path (phmm_wait + the reference's own Genetyper) always works.
 */
#define PHMM_MAX_ALLELES 7      /* Genetyper::MAX_ALLELE_COUNT, genotyper.hpp:19 */
typedef struct phmm_sites {
    int32_t n_sites;
    const int32_t* site_region;       /* [n_sites]                                       */
    const int32_t* site_n_alleles;    /* [n_sites], 1..PHMM_MAX_ALLELES                  */
    const uint8_t* hap_allele;        /* [sum_s n_haps(site_region[s])]                  */
    const uint8_t* read_overlap;      /* [sum_s n_reads(site_region[s])] or NULL         */
} phmm_sites;
typedef struct phmm_gl_result {
    double*  genotype_lik;            /* [sum_s A_s (A_s + 1) / 2] required                            */
    int32_t* site_n_reads;            /* [n_sites] optional: reads that entered the sums               */
    uint8_t* read_keep;               /* [n_reads] optional: 0 = poorly modelled read (erased)         */
    double*  capped_lik;              /* [n_pairs] optional: the capped double matrix, row-major per region
                                         (what compute_likelihoods returns before rows are erased)     */
    phmm_stats stats;                 /* out */
} phmm_gl_result;
int  phmm_submit_gl(phmm_engine* e, const phmm_batch* b, const phmm_sites* sites, phmm_ticket* t);
int  phmm_wait_gl(phmm_engine* e, phmm_ticket t, phmm_gl_result* r);
/* The argument checks of phmm_submit (sites == NULL) / phmm_submit_gl, on their own: pure host logic, no device. */
int  phmm_validate(const phmm_batch* b, const phmm_sites* sites);
/* The Jacobian-logarithm table the reduction indexes (utils/math_utils.hpp:17-29), for tests. */
int  phmm_jacobian_table(const double** table, int32_t* n);

/* Host-side cap + filter of one region's matrix, in place (intel_pairhmm.hpp:24-46).
 * keep[r] = 0 for reads to erase; returns the number kept.  Rows are not compacted. */
int  phmm_normalize_filter(double* lik, int32_t n_reads, int32_t n_haps,
                           const int32_t* read_len, uint8_t* keep);

/* Debug aid (compute-sanitizer is not available everywhere): with PHMM_DEBUG_GUARD=1 in the environment every
 * device buffer of the engine sits between guard zones and starts out filled with the byte PHMM_POISON; this
 * returns the number of guard bytes overwritten so far (0 = no out-of-bounds write; -1 = CUDA error; always 0
 * without the variable).  Call with no ticket in flight.  See tests/test_debug_guards.py. */
int64_t phmm_debug_check(phmm_engine* e);

/* Read-only views of the host-built probability tables (native/Context.h:17-24), for tests. */
int  phmm_tables(const float** ph2pr_f32, const float** mm_f32,
                 const double** ph2pr_f64, const double** mm_f64, int32_t* mm_entries);

/*
 * The planner on its own -- pure host logic, no device needed (tests, diagnostics): how a batch would be
 * cut into warp jobs (reads of a region sorted by length, 2 per lane group, 32/G groups per warp, the
 * job's lane-group shape (G lanes x K rows) taken from its longest read), which jobs take the lane-aligned
 * kernels, how many haplotypes a (job, chunk) unit streams, and which reads go to the long-read kernel.
 * jobs_out (optional): jobs_cap rows of 10 int32 {slot, region, read[8] (-1 = none)}, slot = shape index
 * (+ n_shapes for a lane-aligned job), read indices within the batch.  info->struct_size must be set.
 */
#define PHMM_PLAN_MAX_SHAPES 32
typedef struct phmm_plan_info {
    int32_t struct_size;
    int32_t mode;                     /* 0 per-base gap penalties, 1 batch-constant, 2 constant with i == d */
    int32_t n_jobs, n_long_pairs;
    int32_t haps_per_job, hap_chunks;       /* FP32 launches: haplotypes per unit, grid.y                */
    int32_t haps_per_job64, hap_chunks64;   /* FP64 redo launches                                        */
    int64_t n_pairs, n_cells;
    int32_t n_shapes;
    int32_t shape_g[PHMM_PLAN_MAX_SHAPES], shape_k[PHMM_PLAN_MAX_SHAPES];
    int32_t jobs_ragged[PHMM_PLAN_MAX_SHAPES], jobs_aligned[PHMM_PLAN_MAX_SHAPES];
} phmm_plan_info;
int  phmm_plan(const phmm_batch* b, int32_t sm_count, int32_t host_threads, phmm_plan_info* info,
               int32_t* jobs_out, int64_t jobs_cap);

/*
 * Device-resident form, used by bench.py for the "inputs already in HBM" number:
 * phmm_stage uploads and plans once, phmm_run_staged launches the forward + rescue kernels
 * `iters` times back to back and returns the mean device time per iteration (CUDA events on the
 * launching stream), phmm_fetch_staged brings the results of the last run back.  The device-resident form lives on
 * the engine's FIRST device (devices[0]): it is the measuring instrument of one GPU, the multi-device scheduler is
 * phmm_submit / phmm_wait.
 */
typedef struct phmm_staged phmm_staged;
int  phmm_stage(phmm_engine* e, const phmm_batch* b, phmm_staged** out);
int  phmm_run_staged(phmm_engine* e, phmm_staged* s, int32_t iters, float* ms_per_iter,
                     int32_t* launches_per_iter);
/* As phmm_run_staged; additionally the mean device time of the FP32 forward launch alone (the dominant
 * kernel, bracketed by events on its stream), or -1 when the batch needs several shapes (forked streams). */
int  phmm_run_staged_ex(phmm_engine* e, phmm_staged* s, int32_t iters, float* ms_per_iter,
                        float* fp32_ms_per_iter, int32_t* launches_per_iter);
/* `steps` passes over n staged batches, step i on batch i % n, each batch on its own stream so that
 * consecutive steps overlap (one batch's FP64 redo and kernel tail with the next batch's FP32 kernel), as
 * they do behind phmm_submit with several tickets in flight.  total_ms: device time from a start event
 * every stream waits on to an end event that waits on every stream; launches: kernels launched. */
int  phmm_run_staged_pipelined(phmm_engine* e, phmm_staged* const* staged, int32_t n, int32_t steps,
                               float* total_ms, int32_t* launches);
int  phmm_fetch_staged(phmm_engine* e, phmm_staged* s, phmm_result* r);
void phmm_free_staged(phmm_engine* e, phmm_staged* s);

/*
 * Smith-Waterman haplotype -> reference alignment with back-track: a batch of
 * (reference window, haplotype) pairs, each scored exactly like hc::IntelSWAligner::align
 * (smithwaterman/intel_smithwaterman.hpp:29-44): all-match shortcut (equal length, <= 2 mismatches ->
 * offset 0, "<len>M"), otherwise runSWOnePairBT_avx2 (native/PairWiseSW.h:41-447) with the SOFTCLIP
 * overhang strategy.  Integer arithmetic, bit-exact: same CIGAR, same offset.  Independent of phmm_engine
 * (takes a CUDA device ordinal); synchronous.  Sequences longer than PHMM_SW_MAX_LEN (the reference's
 * MAX_SEQ_LEN, where its own arrays end) -> PHMM_ERR_UNSUPPORTED; empty ones -> PHMM_ERR_INVALID_ARG (the
 * reference throws); no device -> PHMM_ERR_NO_DEVICE (there is no CPU fallback).
 *   alignment k: reference bytes [ref_off[k], ref_off[k+1]) of ref_bases, haplotype bytes
 *   [alt_off[k], alt_off[k+1]) of alt_bases.  Output: offset[k] (alignment begin in the reference) and the
 *   CIGAR elements [elem_beg[k], elem_beg[k+1]) of the compact arrays ops / lens, ops in {'M','I','D','S'},
 *   in CIGAR order.  More than cap_elems elements in total -> PHMM_ERR_UNSUPPORTED (retry with more room;
 *   sum(nref + nalt) always suffices).
 */
#define PHMM_SW_MAX_LEN 1023     /* the reference indexes E[MAX_SEQ_LEN - nrow - 1] and 2*MAX_SEQ_LEN^2 back-track words */
typedef struct phmm_sw_batch {
    int32_t n;
    const int32_t* ref_off;           /* [n+1] */
    const uint8_t* ref_bases;
    const int32_t* alt_off;           /* [n+1] */
    const uint8_t* alt_bases;
    int32_t w_match, w_mismatch, w_open, w_extend;   /* IntelSWAligner::SWParameters; NEW_SW_PARAMETERS = 200,-150,-260,-11 */
} phmm_sw_batch;
typedef struct phmm_sw_result {
    int32_t* offset;                  /* [n]                                                   */
    int64_t* elem_beg;                /* [n+1] out                                             */
    int64_t  cap_elems;               /* capacity of ops / lens, >= n                          */
    uint8_t* ops;                     /* [cap_elems]                                           */
    int32_t* lens;                    /* [cap_elems]                                           */
    float    kernel_ms;               /* out: device time of the alignment kernels (CUDA events) */
} phmm_sw_result;
int  phmm_sw_align(int32_t device, const phmm_sw_batch* b, phmm_sw_result* r);
/* The aligner keeps one context per device (stream, pinned staging, device buffers, grow-only, shared by all
 * threads; calls on one device are serialised).  phmm_sw_release frees them all. */
void phmm_sw_release(void);

#ifdef __cplusplus
}
#endif
#endif /* PHMM_H */
