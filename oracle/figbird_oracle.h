/* figbird_oracle.h — CPU restatement of GapProber's figbird EM hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
 * The product path (gapprober_b200/, include/) never links, imports or executes this code.
 *
 * Parity status: PINNED.  tests/test_oracle_golden.py checks this restatement against outputs of
 * the unmodified reference (and of a build of the reference instrumented at build time by
 * oracle/patch_ref.py) committed under tests/golden/, and — when /root/reference is present —
 * against the freshly built oracle/_ref binaries.
 *
 * Every function cites the reference lines it restates (/root/reference/figbird.cpp).
 */
#ifndef FIGBIRD_ORACLE_H
#define FIGBIRD_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

/* Constant inputs of the hot path, produced by the reference's preamble (figbird.cpp:475-823, 2753-2794). */
typedef struct orc_model {
    int max_read_len;            /* maxReadLength (stat.txt)                         */
    const double* error_pos;     /* errorPosDist[max_read_len]     figbird.cpp:511   */
    const double* in_pos;        /* inPosDist[max_read_len]        figbird.cpp:518   */
    const double* del_pos;       /* delPosDist[max_read_len]       figbird.cpp:539   */
    double error_type[5][5];     /* errorTypeProbs                 figbird.cpp:480   */
    const double* ild;           /* insertLengthDistSmoothed[ild_len] figbird.cpp:624 */
    int ild_len;                 /* maxInsertSize                                    */
    int thr_min, thr_max;        /* insertThresholdMin/Max         figbird.cpp:788   */
    int gap_prob_cutoff;         /* gapProbCutOff                  figbird.cpp:2778  */
    int max_distance;            /* maxDistance = 200              figbird.cpp:103   */
    int num_iter;                /* num_itr = 5                    figbird.cpp:2466  */
    int f_mode;                  /* argv[2]: 0 linear domain, 1 log domain (figbird.cpp:2578) */
} orc_model;

/* One (anchor, unmapped mate) line pair of Gaps/gaps_<n>.sam as placeReads parses it
 * (figbird.cpp:1726-1801): anchor flag/pos/sequence length, and the mate's sequence as stored. */
typedef struct orc_record {
    int pos1;            /* anchor POS (1-based, compared unadjusted with the 0-based gapStart) */
    int len1;            /* strlen(anchor SEQ) */
    int strand1;         /* (flag1 & 16) >> 4 */
    const char* seq2;    /* mate SEQ, NUL terminated, file orientation */
} orc_record;

typedef struct orc_result {
    int gap_min, gap_max;        /* scanned estimate range (figbird.cpp:2452-2453) */
    int best_estimate;           /* maxGapEstimate */
    double best_lik;             /* maxLikelihood */
    int total_reads, discarded;  /* finalize's "Total_Reads / Discard" line (figbird.cpp:2382) */
    int filled_len;              /* getConcensus length */
    char* filled;                /* caller buffer, >= 2*orig_gap + 2 bytes */
    /* optional taps (NULL to skip) */
    double* lik_per_estimate;    /* [gap_max-gap_min+1] final-iteration likelihood of every estimate */
    int dump_estimate;           /* estimate whose final profile is copied to probs_dump (or <=0)    */
    double* probs_dump;          /* [(dump_estimate + 2*max_distance) * 5] probsGap after the last iteration */
    double* counts_dump;         /* same shape, countsGap after the last iteration */
    char* best_string;           /* bestString handed to finalize; caller buffer >= 2*orig_gap + 2 */
} orc_result;

/* GapFiller::fillGap for one gap (figbird.cpp:2436-2510), incl. finalize and getConcensus. */
int orc_fill_gap(const orc_model* m, const char* contig, long contig_len, long gap_start, int orig_gap,
                 const orc_record* recs, int nrec, orc_result* out);

/* The same with the M-step sum of placeReads grouped the way the record-sharded multi-GPU scan groups it: the records
 * are cut into `shards` contiguous blocks [s*R/S, (s+1)*R/S), each block is accumulated in the reference's order into
 * its own array, and the arrays are added in block order, ((p0 + p1) + p2) + ...  With shards <= 1 this is orc_fill_gap.
 * This is NOT the reference's arithmetic (it differs from it in the last bits); it pins, bit for bit, what
 * fb_set_partition_records computes with `shards` ranks. */
int orc_fill_gap_sharded(const orc_model* m, const char* contig, long contig_len, long gap_start, int orig_gap,
                         const orc_record* recs, int nrec, int shards, orc_result* out);

/* Test aid for gaps whose full scan takes minutes on one core (the single-gap stress corners): scan[gp - gap_min] == 0
 * leaves estimate gp out of the scan (its likelihood reads NaN); the '>=' selection and finalize then run over the
 * scanned estimates only.  scan == NULL scans everything.  Each scanned estimate is computed exactly as in the full
 * scan (estimates are independent, figbird.cpp:2468-2501). */
int orc_fill_gap_subset(const orc_model* m, const char* contig, long contig_len, long gap_start, int orig_gap,
                        const orc_record* recs, int nrec, int shards, const unsigned char* scan, orc_result* out);

/* Estimate range only (figbird.cpp:2438-2453), float arithmetic included. */
void orc_gap_range(int orig_gap, int* gap_min, int* gap_max);

/* Number of (record, placement) likelihoods one placeReads call evaluates for estimate gp
 * (the window test of figbird.cpp:1826-1833 / 1889-1896). */
long orc_count_likelihoods(const orc_model* m, long gap_start, int gp, const orc_record* recs, int nrec);

/* Parse Gaps/gaps_<n>.sam the way placeReads does (figbird.cpp:1718-1801).  Returns the number of
 * records; *recs_out and the strings are malloc'd (free with orc_free_records). */
int orc_read_gap_file(const char* path, orc_record** recs_out);
void orc_free_records(orc_record* recs, int n);

#ifdef __cplusplus
}
#endif
#endif
