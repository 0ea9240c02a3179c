/* figbird_cuda.h — C ABI of libfigbird_cuda.so: the figbird EM hot path of GapProber on B200 (sm_100a).
 *
 * The reference has no library/FFI boundary: figbird.cpp is one translation unit whose main() drives
 * `class GapFiller` (figbird.cpp:1517-2532) gap by gap (figbird.cpp:2819-2855).  This header is the
 * boundary a maintainer would bind instead: each entry point replaces the GapFiller method(s) cited
 * next to it.  Plain C types only; the caller owns every buffer it passes (inputs are copied before
 * the call returns) and every buffer it receives results in.  No exceptions and no exit() cross the
 * ABI: every function returns FB_OK (0) or a negative fb_status.  One host thread drives a context;
 * a context owns one CUDA device, one stream and its device memory.  There is NO CPU fallback: with
 * no usable CUDA device fb_create fails with FB_ECUDA.
 */
#ifndef FIGBIRD_CUDA_H
#define FIGBIRD_CUDA_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FB_ABI_VERSION 1

typedef enum fb_status {
    FB_OK = 0,
    FB_EINVAL = -1,   /* bad argument / model outside supported bounds                               */
    FB_EDOMAIN = -2,  /* gap or record outside the territory where the reference is defined: */
                      /* gap closer than max_distance to a scaffold end, read longer than max_distance.  */
    FB_ECUDA = -3,    /* CUDA runtime error or no device; fb_last_error() has the text                 */
    FB_ENOMEM = -4,
    FB_ESTATE = -5,   /* call out of order (e.g. fb_result before fb_run)                              */
    FB_EUNSUPPORTED = -6 /* f_mode other than 0 / 1; f_mode 1 with the record-sharded partition; a host libm whose log() */
                         /* is not monotone around the finalize cut-off (see fb_create)                               */
} fb_status;

/* Constant inputs of the hot path = outputs of the reference's preamble (figbird.cpp:475-823,
 * 2753-2794).  Defaults of the compile-time constants: max_distance 200 (figbird.cpp:103),
 * num_iter 5 (figbird.cpp:2466), f_mode 0 (what GapProber.py:44 passes). */
typedef struct fb_model {
    int max_read_len;          /* maxReadLength                                   */
    const double* error_pos;   /* errorPosDist[max_read_len]     figbird.cpp:511  */
    const double* in_pos;      /* inPosDist[max_read_len]        figbird.cpp:518  */
    const double* del_pos;     /* delPosDist[max_read_len]       figbird.cpp:539  */
    double error_type[5][5];   /* errorTypeProbs                 figbird.cpp:480  */
    const double* ild;         /* insertLengthDistSmoothed[ild_len]  figbird.cpp:624 */
    int ild_len;               /* maxInsertSize                                   */
    int thr_min, thr_max;      /* insertThresholdMin/Max         figbird.cpp:788  */
    int gap_prob_cutoff;       /* gapProbCutOff                  figbird.cpp:2778 */
    int max_distance;
    int num_iter;
    int f_mode;                /* argv[2] of figbird: 0 linear-domain scores (what GapProber.py passes), 1 log-domain   */
                               /* (what run_EM_eval.sh passes): log() per term, pow(10, score) into the counts          */
} fb_model;

/* One (anchor, unmapped mate) line pair of Gaps/gaps_<n>.sam, as placeReads parses it
 * (figbird.cpp:1726-1801). */
typedef struct fb_record {
    int pos1;          /* anchor POS column, unadjusted                         */
    int len1;          /* strlen(anchor SEQ)                                    */
    int strand1;       /* (anchor FLAG & 16) >> 4                               */
    const char* seq2;  /* mate SEQ as stored in the file (not NUL-terminated ok) */
    int len2;          /* strlen(mate SEQ)                                      */
} fb_record;

typedef struct fb_gap_result {
    int gap_min, gap_max;     /* scanned estimate range           figbird.cpp:2452-2453 */
    int best_estimate;        /* maxGapEstimate                   figbird.cpp:2485      */
    double best_lik;          /* maxLikelihood                    figbird.cpp:2484      */
    int total_reads;          /* finalize's Total_Reads           figbird.cpp:2382      */
    int discarded;            /* finalize's Discard                                     */
    int filled_len;           /* getConcensus() length            figbird.cpp:2050      */
    long likelihoods;         /* (record, placement) scores evaluated by the EM scan of this gap */
} fb_gap_result;

typedef struct fb_stats {
    long n_gaps, n_units, n_records;
    long likelihoods;           /* EM scan: sum over units, iterations, records of in-window placements */
    long finalize_likelihoods;  /* hard-assignment pass                                               */
    long base_terms;            /* likelihoods x mate length                                          */
    double em_kernel_ms;        /* CUDA-event time of the EM kernel launches of the last fb_run_device */
    double finalize_kernel_ms;
    double step_ms;             /* CUDA-event time of the whole last fb_run_device (kernels, copies, host phases) */
    int em_launches, finalize_launches;
    long h2d_bytes, d2h_bytes;  /* of the last fb_upload + fb_run_device                               */
    long algorithmic_bytes;     /* SURVEY §8(d) HBM byte model summed over (unit, iteration)            */
    int aux_launches;           /* other kernels of the pass: record-table staging, record-shard closing step */
} fb_stats;

typedef struct fb_ctx fb_ctx;

/* Create a context on CUDA device `device` and copy the model.  Replaces the globals the reference's
 * main() leaves behind for GapFiller (figbird.cpp:2718-2794). */
int fb_create(const fb_model* model, int device, fb_ctx** out);
/* Optional: bring up the CUDA context of `device` and load the kernels ahead of fb_create, e.g. on a second host thread
 * while the caller is still learning the model (a fresh process spends several hundred milliseconds here). */
int fb_warmup(int device);
void fb_destroy(fb_ctx* ctx);

/* Register one gap = GapFiller::allocate + initGapFiller (figbird.cpp:1543-1566, 1610-1618) plus the
 * parse of its gap file (figbird.cpp:1718-1812, done by the caller).  `contig` is the upper-cased
 * scaffold (figbird.cpp:2652-2658); only the max_distance bases either side of the gap are read.
 * Returns the gap index (>= 0) or a negative fb_status. */
int fb_add_gap(fb_ctx* ctx, int contig_no, long gap_start, int orig_gap, const char* contig, long contig_len,
               const fb_record* recs, int nrec);

/* Same, parsing the gap file in the library exactly as placeReads does on every call (figbird.cpp:1680-1812):
 * fgets line pairs of at most 1023 bytes, '@' lines skipped, tab-separated columns
 * qname flag rname pos cigar tlen seq.  A missing file is FB_EINVAL. */
int fb_add_gap_file(fb_ctx* ctx, int contig_no, long gap_start, int orig_gap, const char* contig, long contig_len,
                    const char* gap_file);

/* Bulk form: parse and pack n gap files on several host threads (threads = 0: one per hardware thread, at most 16), then
 * register them in the order given.  handles[i] receives the gap index or a negative fb_status for gap i; the call
 * returns FB_OK only if every gap was accepted. */
typedef struct fb_gap_desc {
    int contig_no;
    long gap_start;
    int orig_gap;
    const char* contig;
    long contig_len;
    const char* gap_file;
} fb_gap_desc;
int fb_add_gap_files(fb_ctx* ctx, const fb_gap_desc* descs, int n, int* handles, int threads);

/* Restrict this context to its share of the work when several processes (one per GPU) split one gap
 * set: gaps are dealt to `world` ranks by descending cost (longest-processing-time first); this
 * context runs those of `rank`.  Results of other ranks' gaps are absent (fb_result -> FB_ESTATE).
 * Call after the last fb_add_gap and before fb_upload/fb_run.  No data-path collective is involved. */
int fb_set_partition(fb_ctx* ctx, int rank, int world);
/* Estimate-level variant, for gap sets dominated by few gaps: the independent units are (gap, estimate) pairs
 * (figbird.cpp:2468-2501 scans estimates without any cross-estimate dependence until the final arg-max), so the
 * units are dealt instead of the gaps.  Every rank registers every gap, then runs
 *     fb_upload; fb_run_scan;  [exchange: fb_scan_liks -> all-gather -> fb_merge_liks];  fb_run_finalize
 * and the rank that scanned a gap's winning estimate finalizes that gap (fb_result elsewhere: FB_ESTATE).
 * The exchange is a few doubles per gap and is the caller's (gapprober_b200/multigpu.py: torch.distributed). */
int fb_set_partition_units(fb_ctx* ctx, int rank, int world);
int fb_run_scan(fb_ctx* ctx);        /* EM scan of the owned units + their likelihoods                       */
int fb_run_finalize(fb_ctx* ctx);    /* selection + finalize; needs every estimate's likelihood (own or merged) */
/* Likelihood per estimate after fb_run_scan; NaN where another rank scans the estimate.  Returns the count. */
int fb_scan_liks(fb_ctx* ctx, int gap, double* liks, int cap);
/* Fill in likelihoods scanned elsewhere (NaN entries of `liks` are ignored; values computed here are kept). */
int fb_merge_liks(fb_ctx* ctx, int gap, const double* liks, int n);
/* 1 if gap `gap` is run by this context under the current partition. */
int fb_owns_gap(fb_ctx* ctx, int gap);

/* Record-sharded scan — north_star's "by read where one gap dominates" and BASELINE config 3.  Every rank
 * scans every (gap, estimate) unit, but only over its contiguous block of each gap's records (figbird.cpp:1718 reads
 * them in file order; rank r takes records [r*R/world, (r+1)*R/world)).  The M-step sums of figbird.cpp:1874-1881 /
 * 1940-1947 are then partial: after every one of the num_iter iterations the partial counts of all ranks are added, in
 * rank order, by the kernel that starts the next iteration (it reads the peers' buffers directly: own memory,
 * peer-mapped memory of the other GPUs over NVLink, or a gathered copy — the library only sees device pointers).
 * The summation order inside a rank is the reference's; across ranks it is ((p0 + p1) + p2) + ...: counts and
 * probabilities agree with the single-GPU path to rounding (1e-5 relative is the bar, observed ~1e-15), filled
 * sequences are identical unless a cell ties to the last bit.  Call order, on every rank:
 *     fb_set_partition_records; fb_upload;
 *     for it in 0 .. num_iter-1:  fb_shard_scan_iter(it, peers' buffers of slot (it-1)&1, own buffer of slot it&1);  barrier
 *     fb_shard_scan_finish(peers' buffers of slot (num_iter-1)&1, peers' maxima);  barrier;  fb_run_finalize
 * fb_run_finalize then fills the gaps whose index is congruent to `rank` modulo `world` (fb_owns_gap tells).        */
int fb_set_partition_records(fb_ctx* ctx, int rank, int world);
/* Sizes (in doubles) of one partial-count buffer and of the per-record maxima buffer, valid after fb_upload. */
int fb_shard_sizes(fb_ctx* ctx, size_t* counts_doubles, size_t* maxprob_doubles);
/* This context's own exchange buffers (device memory from cudaMalloc, exportable with fb_ipc_export). */
int fb_shard_buffers(fb_ctx* ctx, double** counts_slot0, double** counts_slot1, double** maxprob, unsigned long long** flags);
/* One EM iteration over the local record blocks.  peer_counts: `world` device pointers (rank order, own included) to
 * the previous iteration's partial counts, ignored for it == 0.  out_counts: where this iteration's partial counts
 * go.  Returns when the iteration has finished on the device. */
int fb_shard_scan_iter(fb_ctx* ctx, int it, const double* const* peer_counts, double* out_counts);
/* Sum the last iteration's partial counts (rank order) -> consensus per unit; merge the per-record maxima of the
 * ranks -> likelihood per estimate (identical on every rank).  Leaves the context as fb_run_scan does. */
int fb_shard_scan_finish(fb_ctx* ctx, const double* const* peer_counts, const double* const* peer_maxprob);
/* The whole scan in one call with a device-side barrier (one GPU per rank, all buffers peer-mapped): every array holds
 * `world` device pointers in rank order, entry `rank` being this context's own buffer from fb_shard_buffers.  All
 * iterations, the waits between them (flags the ranks store into each other's memory over NVLink) and the closing
 * reduction are enqueued at once; the host synchronises once at the end.  Every rank must call it the same number of
 * times.  Equivalent to the fb_shard_scan_iter / barrier / fb_shard_scan_finish sequence. */
int fb_shard_scan_all(fb_ctx* ctx, const double* const* peer_counts_slot0, const double* const* peer_counts_slot1,
                      const double* const* peer_maxprob, unsigned long long* const* peer_flags);
/* A wait of the device-side barrier gives up after FB_BARRIER_TIMEOUT_MS (environment, default 20000) and the call
 * returns FB_ECUDA.  The failure is passed on: the rank that gave up poisons its flag in every peer's memory, a peer
 * that reads the poison fails the same way instead of adding up stale counts, and the launches still queued behind a
 * failed barrier do nothing.  A rank that had already finished its pass may still return FB_OK, so callers reduce the
 * status over the ranks before using results (RecordShardSession.scan does) and then call fb_shard_reset on EVERY rank
 * (flags, error word and pass counter start afresh; ranks must be quiescent, e.g. between two host barriers). */
int fb_shard_reset(fb_ctx* ctx);
/* Parity tap: summed counts of the gap rows of one unit after fb_shard_scan_finish, laid out [estimate][5]. */
int fb_shard_counts(fb_ctx* ctx, int gap, int estimate, double* counts);
/* CUDA IPC plumbing for one-process-per-GPU runs: export a cudaMalloc'ed buffer as a 64-byte handle, map a peer's
 * handle into this process (reads then go over NVLink), unmap it. */
int fb_ipc_export(fb_ctx* ctx, const void* dev_ptr, unsigned char handle[64]);
int fb_ipc_open(fb_ctx* ctx, const unsigned char handle[64], void** dev_ptr);
int fb_ipc_close(fb_ctx* ctx, void* dev_ptr);
/* The dealing rule itself, as pure host code (no GPU needed): items by descending cost, each to the least loaded
 * rank, ties to the lower rank; owner[i] receives the rank of item i.  fb_set_partition applies it to fb_gap_cost. */
int fb_partition_plan(const double* cost, int n, int world, int* owner);
/* Work estimate of a registered gap: over its estimates, in-window placements x mate length (+ a constant per estimate). */
double fb_gap_cost(fb_ctx* ctx, int gap);

/* GapFiller::fillGap for every registered gap (figbird.cpp:2436-2510): estimate scan, num_iter EM
 * iterations per estimate, best-estimate selection, finalize.  fb_run = fb_upload + fb_run_device. */
int fb_run(fb_ctx* ctx);
int fb_upload(fb_ctx* ctx);      /* pack + host->device copy of every registered gap */
int fb_run_device(fb_ctx* ctx);  /* everything after the upload; may be called repeatedly (benchmark steps) */
/* The gap loop of figbird's main() (figbird.cpp:2819-2855) in one call, gap files in, results out: equivalent to
 * fb_add_gap_files + fb_run on a context without gaps, but the gap set goes to the device in two batches — the first
 * eighth is parsed, packed, uploaded and started on a second context of the same model, and the host reads, parses, packs
 * and plans the rest while that batch's EM scan runs.  Gaps are independent, so every result is the same, bit for bit.
 * handles[i] indexes the getters below (the second batch's gaps come first); a refused gap (FB_EDOMAIN) gets a negative
 * handle, the others are still filled and the call returns FB_EDOMAIN.  Afterwards the context can be read, run again
 * (fb_run_device) or cleared (fb_clear_gaps); adding gaps or changing the partition needs fb_clear_gaps first.
 * Sets of fewer than 128 gaps (FB_PIPE_MIN) and partitioned contexts take the plain route. */
int fb_fill_gap_files(fb_ctx* ctx, const fb_gap_desc* descs, int n, int* handles, int threads);

/* GapFiller::getConcensus (figbird.cpp:2050-2063) + the numbers fillGap/finalize print. */
int fb_result(fb_ctx* ctx, int gap, fb_gap_result* out, char* filled, int filled_cap);

/* Parity taps (the reference prints none of this; tests compare with oracle/ dumps). */
int fb_estimate_liks(fb_ctx* ctx, int gap, double* liks, int cap);   /* final-iteration likelihood per estimate */
int fb_best_string(fb_ctx* ctx, int gap, char* buf, int cap);         /* bestString handed to finalize        */
/* Re-runs estimate `estimate` of gap `gap` alone and copies its final profile:
 * probs/counts are [(estimate + 2*max_distance) * 5] doubles, row-major like probsGap/countsGap. */
int fb_dump_profile(fb_ctx* ctx, int gap, int estimate, double* probs, double* counts);

int fb_get_stats(fb_ctx* ctx, fb_stats* out);
/* Parity tap for the C-library functions on the path: the device's restatement of the host glibc's log (what = 0) and
 * pow(10, .) (what = 1) applied to n doubles.  The reference calls log() on every record's best score
 * (figbird.cpp:1951-1955) and, in f=1 mode, log() per base term and pow(10, s) per placement (figbird.cpp:1856, 1878);
 * tests compare these results with the host libm bit for bit. */
int fb_devmath_eval(int device, int what, const double* in, double* out, long n);
/* Measurement aid (no counterpart in the reference): shared-memory read bandwidth of `device` in GB/s, measured with
 * conflict-free 128-bit loads on every SM — the roof that binds the E-step; bench.py reports against it. */
int fb_measure_smem_peak(int device, double* gbs);
int fb_clear_gaps(fb_ctx* ctx);                 /* drop all gaps, keep the model */
/* Estimate range of fillGap (figbird.cpp:2438-2453), float arithmetic included. */
void fb_gap_range(int orig_gap, int* gap_min, int* gap_max);
const char* fb_last_error(fb_ctx* ctx);
const char* fb_status_str(int status);
int fb_abi_version(void);

#ifdef __cplusplus
}
#endif
#endif
