// fb_device.h — device-side data layout and launch interface of the figbird EM kernels (sm_100a).
// Shared between fb_kernels.cu (device code) and fb_context.cpp (host orchestration).
#pragma once
#include <cstddef>
#include <cstdint>
#include <cuda_runtime.h>

#define FB_MAX_READ 128          // longest mate the packed layout holds (2-bit codes in 2 x uint4)
#define FB_THREADS 512           // CTA size of the EM kernel (1024 where only one CTA fits an SM)
#define FB_FIN_THREADS 256       // CTA size of the finalize kernel
#define FB_MAX_SHARDS 16         // ranks of a record-sharded scan

// Per-record metadata, relative to the gap's profile window (row 0 = scaffold position gapStart - D).
// u = row of the mate's first base for a placement.  Insert size of a placement:
//   anchor left of the gap  (figbird.cpp:1824-1828): t = u + tb
//   anchor right of the gap (figbird.cpp:1887-1891): t = tb - u
// [a_lo, a_hi] is the u-range that passes the insert-window test, before clamping to the estimate.
struct FbRecMeta {
    int a_lo, a_hi;
    int tb;
    int flags;   // bits 0-7 mate length, bit 8 isReverse (anchor forward), bit 9 anchor left of gap,
                 // bits 10-31: first chain of this record in the left-flank prefix table (set at upload; see FbPrefixParams)
};
#define FB_PRE_SHIFT 10
#define FB_PRE_MAX_CHAINS (1u << 22)

struct FbGapMeta {
    int rec_off, nrec;
    int orig;    // original N-run length
    int lmax;    // longest mate of this gap
};

struct FbUnit {
    int gap;         // index into gap meta
    int gp;          // gap-length estimate
    long long mp_off;    // offset of this unit's per-record maxima in unit_maxprob
    long long cons_off;  // offset of this unit's consensus string in unit_cons
    int hole_lo;     // profile rows [hole_lo, hole_lo + hole_len) are touched by no placement of this unit and are left
    int hole_len;    // out of the tables; hole_lo - D is a multiple of 32, the hole lies inside the gap rows; 0 = none
};

struct FbEmParams {
    // model
    double2 eo_tab[2 * FB_MAX_READ];   // {1 - e[k], e[k]} per read position: lives in the kernel-parameter constant bank, so
                                   // the E-step fetches it through the constant cache instead of shared-memory bandwidth.
                                   // Twice: [k] forwards, and [FB_MAX_READ + i] = entry FB_MAX_READ-1-i, so that a
                                   // reverse-complemented mate of length L (read position L-1-j for base j) walks the table
                                   // upwards too, from 2*FB_MAX_READ - L: one E loop for both orientations
    const double* ild;      // insertLengthDistSmoothed[thr_min .. thr_max]
    double T[25];           // errorTypeProbs
    int thr_min, thr_max, D, num_iter, max_read_len;
    // gaps / records (struct of arrays in HBM)
    const FbGapMeta* gaps;
    const uint8_t* flank;   // [n_gaps][2*D] base codes: D left of the gap, D right of it
    const uint4* codes;     // [n_records][2]  2-bit bases, 16 per 32-bit word, already in scoring orientation
    const uint4* nmask;     // [n_records]     bit j set: base j is not ACGT
    const FbRecMeta* rmeta; // [n_records]
    // per-record tables built by fb_stage_kernel (the EM kernel bulk-copies them chunk by chunk into shared memory)
    const uint32_t* st_mask;   // [n_records][20]
    const uint32_t* st_cls4;   // [n_records][16]
    const uint8_t* st_jl;      // [n_records][FB_MAX_READ + 16]
    const uint8_t* st_cum;     // [n_records][32]
    const double* pre;         // left-flank prefix table (fb_prefix_kernel); nullptr: every chain starts at ILD[t]
    // First-iteration terms.  initialize() gives every gap row the same counts (N = 1, figbird.cpp:1665-1673), so in the
    // first iteration the term of a base against ANY gap row depends only on its class c and read position k:
    // t0_tab[c][k] = P*(1-e[k]) + e[k]*Q of that row (its log in f=1 mode), computed once on the device by the E-step's own
    // operations (fb_t0_kernel).  Where all placements of a task are on gap rows the E-step multiplies by t0_tab[c][k]
    // — a uniform operand from the constant bank — instead of loading {P,Q} per lane.  use_t0 = 0 switches it off.
    double t0_tab[5][FB_MAX_READ];
    int use_t0;
    int e_rec_major;           // experiment knob (FB_E_REC_MAJOR=1): E tasks of a round in (record, block) order instead of (block, record)
    // work list
    const FbUnit* units;
    int n_units;
    int* work_counter;
    // outputs
    double* unit_maxprob;   // per (unit, record) max placement score of the last iteration
    double* unit_lik;       // per unit: the likelihood placeReads returns in the last iteration (nullptr: not wanted)
    char* unit_cons;        // per unit consensus string (computeSequence(0))
    double* dump_probs;     // optional: full profile of unit dump_unit, [(gp + 2D) * 5]
    double* dump_counts;
    int dump_unit;
    // shared-memory geometry of this launch
    int nr_stride;          // rows of the P/Q table
    int gp_stride;          // rows of the counts table
    int p_stride;           // placements per record slot in the score buffer
    int chunk;              // records per E/M round
    // global-memory tables for gaps too long for shared memory (one slice per CTA)
    double* gws;
    size_t gws_stride;      // doubles per CTA
    // record-sharded scan (one launch per EM iteration, num_iter = 1): a unit's partial counts live at
    // 5 * cons_off in every rank's exchange buffer, class-major [5][gp]
    const double* parts_in[FB_MAX_SHARDS];   // last iteration's partial counts of every rank (own, peer-mapped or gathered)
    int n_parts;                             // 0: first iteration, start from initialize()
    double* part_out;                        // this rank's partial counts of this iteration (nullptr: not sharded)
    const int* abort;                        // record-sharded scan: non-zero once a barrier has failed -> the launch does nothing
    // two-pass scan (one-call pass over many units): pass 1 leaves the M-step of the last iteration and the consensus out,
    // pass 2 runs the winners of the selection again in full; its unit count is only known on the device
    int skip_last_m;
    const int* n_units_ptr;                  // non-null: units to run = min(n_units, *n_units_ptr)
};

struct FbStageParams {
    const uint4* codes;
    const uint4* nmask;
    const FbRecMeta* rmeta;
    int n_records;
    uint32_t* st_mask;
    uint32_t* st_cls4;
    uint8_t* st_jl;
    uint8_t* st_cum;
};

// Left-flank prefixes.  A placement u < D multiplies its first D - u terms against rows of the left flank, and those
// rows never change: not between the iterations, not between the estimates of a gap (figbird.cpp:1657-1664 writes them
// once, placeReads only zeroes and refills the gap rows, :1708-1714).  The insert size of a placement does not depend on
// the estimate either (pack_gap).  So the running score of placement u after its first flank terms is the same number in
// every E-step of every unit of the gap.  A record's placements are dealt to lanes in chains of 32 (slot k = u - lo,
// lo = max(a_lo, D - L2 + 1)); for chain c every lane has at least js = D - (lo + 32c + 31) flank terms.  When js >= 4 the
// table holds, per lane, the score after the first (js & ~3) terms — computed by the same operations in the same order
// (ILD[t], then one rounded term and one rounded product per base) — and the E-step starts the chain from there.
// A record owns n = min((D - lo - 3) / 32, chains of [lo, a_hi]) consecutive chains of 32 doubles from the chain index
// in its flags; slots past a_hi hold 0.
struct FbPrefixParams {
    double2 eo_tab[FB_MAX_READ];
    double T[25];
    const double* ild;         // the E-step's start values (ILD[t], or log(ILD[t]) in f=1 mode), [thr_min .. thr_max]
    int thr_min, D;
    const FbGapMeta* gaps;
    int n_gaps;
    const uint8_t* flank;
    const FbRecMeta* rmeta;
    const uint32_t* st_cls4;   // built by fb_stage_kernel earlier in the stream
    double* pre;
};
// t0[c * FB_MAX_READ + k], see FbEmParams::t0_tab; eo/T/out are device pointers
cudaError_t fb_launch_t0(const double2* eo, const double* T, double* out, bool f1, cudaStream_t s);
cudaError_t fb_launch_prefix(const FbPrefixParams& p, int slices, bool f1, cudaStream_t s);   // CTAs per gap = slices

// Closing step of a record-sharded scan: rank-ordered sum of the partial counts -> computeSequence(0); per-record
// maxima merged over the ranks.
struct FbShardFinParams {
    const FbUnit* units;
    int n_units;
    const int* unit_nrec;                    // records of the unit's gap (all ranks' records)
    const double* parts_in[FB_MAX_SHARDS];
    const double* maxprob_in[FB_MAX_SHARDS]; // per rank: unit_maxprob in the global record layout, 0 where not owned
    int n_parts;
    double* sum_counts;                      // same layout as the partial counts
    double* maxprob_out;
    char* unit_cons;
    const int* abort;                        // non-zero once a barrier of this pass has failed: skip the reduction
};

struct FbFinParams {
    const double* ild;
    const double* epos;     // e[k]
    const double* match;    // 1 - e[k] - in[k] - del[k]
    double T[25];
    int thr_min, thr_max, D;
    const FbGapMeta* gaps;
    const uint8_t* flank;
    const uint4* codes;
    const uint4* nmask;
    const FbRecMeta* rmeta;
    const int* gap_list;    // gaps to finalize in this launch
    const int* best_gp;     // [n_gaps] winning estimate
    const long long* best_cons_off;  // [n_gaps] offset of the winning consensus in unit_cons; -1: no winner, empty bestString
    const char* unit_cons;
    int gmax_stride;        // smem bytes for the gap string
    int gb_stride;          // smem rows of the per-CTA counts (largest winning estimate of the launch)
    double theta;           // smallest double x with -log(x) <= gapProbCutOff (host libm): the cut-off test is x >= theta
    int* fin_counts;        // global counts of the gaps finalized by several CTAs, [.][5], zeroed before the launch
    const int* fin_counts_off;   // [n_gaps] offset of a gap's counts in fin_counts
    int* tickets;           // [n_gaps] CTAs of the gap that have added their counts, zeroed before the launch
    int* discarded;         // [n_gaps] out: records that failed the cut-off test (zeroed before the launch)
    char* filled;           // out: the filled sequences of the launch's gaps, back to back
    const long long* filled_off;   // [n_gaps] offset of gap g's sequence in `filled`
};

// fillGap's selection among the estimates of a gap (figbird.cpp:2479-2490) on the device: thread per gap, estimates in
// ascending order, 'likelihood >= maxLikelihood' (the last one wins ties; NaN and -inf never win).
struct FbSelParams {
    int n_gaps;
    const int* sel_off;          // [n_gaps + 1] first entry of gap g in sel_unit (no entries: gap not run by this context)
    const int* sel_unit;         // per (gap, estimate in ascending order): index of the unit in the launch order
    const int* gmin;             // [n_gaps] first estimate of the scan
    const double* unit_lik;
    const FbUnit* units;
    int* best_gp;                // [n_gaps] out: maxGapEstimate
    long long* best_cons_off;    // [n_gaps] out: where the winner's consensus lies, -1 when no estimate won
    double* best_lik;            // [n_gaps] out: maxLikelihood
    // two-pass scan: the winners, appended to the work list of their size bin (nullptr: single pass)
    FbUnit* win_units;           // laid out like `units`: bin b's winners from bin_first[b] on
    int* win_count;              // [n_bins] zeroed before the launch
    const int* bin_first;        // [n_bins] first unit of every bin in the launch order
    int n_bins;
};
cudaError_t fb_launch_select(const FbSelParams& p, cudaStream_t s);

size_t fb_em_smem_bytes(int nr_stride, int gp_stride, int p_stride, int chunk, int pwin, bool global_tables);
// threads: 512 or 1024; f1: the reference's log-domain mode (argv[2] = 1)
cudaError_t fb_launch_em(const FbEmParams& p, int grid, int threads, bool global_tables, bool f1, size_t smem, cudaStream_t s);
cudaError_t fb_em_ctas_per_sm(int threads, bool global_tables, bool f1, size_t smem, int* n);   // occupancy of the plain variant, static smem included
// cluster variant: n_clusters thread-block clusters of `cluster` CTAs, each cluster running one unit at a time
cudaError_t fb_launch_em_cluster(const FbEmParams& p, int n_clusters, int cluster, int threads, bool f1, size_t smem, cudaStream_t s);
cudaError_t fb_em_max_clusters(int cluster, int threads, bool f1, size_t smem, int* n);     // co-resident clusters of that size on this device
// n gaps x `slices` CTAs per gap (a CTA takes every slices-th group of 8 records)
cudaError_t fb_launch_finalize(const FbFinParams& p, int n, int slices, size_t smem, cudaStream_t s);
// Flags of the record-sharded scan's device-side barrier (one 64-bit word per rank, in every rank's exchange memory).
struct FbFlagParams {
    unsigned long long* peer_flags[FB_MAX_SHARDS];   // signal: every rank's flag array (own included)
    const unsigned long long* own_flags;             // wait: this rank's array, written by the peers
    unsigned long long token;
    long long timeout_cycles;                        // wait: give up after this many SM clock cycles
    int rank, world;
    int* error;                                      // set to 1 when a peer did not arrive in time, or reported that one did not
};
#define FB_FLAG_POISON (~0ull)                       // stored instead of a token by a rank whose pass has failed
cudaError_t fb_launch_flag_signal(const FbFlagParams& p, cudaStream_t s);
cudaError_t fb_launch_flag_wait(const FbFlagParams& p, cudaStream_t s);
cudaError_t fb_launch_stage(const FbStageParams& p, cudaStream_t s);
cudaError_t fb_launch_shard_finish(const FbShardFinParams& p, cudaStream_t s);
cudaError_t fb_kernels_init(int* max_smem_optin, int* sm_count);
cudaError_t fb_launch_devmath(const double* in, double* out, long n, int what, cudaStream_t s);   // what: 0 log, 1 pow(10, .)
cudaError_t fb_microbench_smem(double* gbs, int reps);     // fb_microbench.cu: measured shared-memory read bandwidth (LDS.128, conflict-free)
