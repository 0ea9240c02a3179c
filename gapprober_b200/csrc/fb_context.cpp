// fb_context.cpp — host side of libfigbird_cuda.so: the C ABI of include/figbird_cuda.h.
//
// Orchestrates GapFiller::fillGap (figbird.cpp:2436-2510) for a whole gap set on one GPU:
//   pack records once (the reference re-parses the gap file 5*E times, figbird.cpp:1680-1812)
//   -> EM scan of every (gap, estimate) unit on the device (fb_kernels.cu)
//   -> per-unit likelihood = sum of glibc log() over per-record maxima, in record order, on the host
//      (figbird.cpp:1951-1955; the host libm is the reference's libm, so the '>=' selection of
//      figbird.cpp:2479 sees the same bits)
//   -> finalize scoring on the device, cut-off test / counting / computeSequence(1) on the host.
// There is no CPU implementation of the EM in here: without a CUDA device fb_create fails.

#include "../../include/figbird_cuda.h"
#include "fb_device.h"

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <unistd.h>
#include <string>
#include <vector>

namespace {

// FB_TRACE=1: phase wall times on stderr
struct Trace {
    bool on = getenv("FB_TRACE") != nullptr;
    std::chrono::steady_clock::time_point t = std::chrono::steady_clock::now();
    void lap(const char* what) {
        if (!on) return;
        auto n = std::chrono::steady_clock::now();
        fprintf(stderr, "[fb_trace] %-28s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(n - t).count());
        t = n;
    }
};

inline int base_code(char ch) {   // charCodes, figbird.cpp:2753-2776
    switch (ch) {
        case 'A': return 0;
        case 'C': return 1;
        case 'G': return 2;
        case 'T': return 3;
        default: return 4;
    }
}

inline int revcomp_code(char ch) {   // reverse(), figbird.cpp:1382-1403, then charCodes
    switch (ch) {
        case 'A': case 'a': return 3;
        case 'C': case 'c': return 2;
        case 'G': case 'g': return 1;
        case 'T': case 't': return 0;
        default: return 4;
    }
}

struct HostGap {
    int contig_no;
    long gap_start;
    int orig, gmin, gmax;
    int rec_off, nrec, lmax;
    int scan_off = 0, scan_n = -1;      // record-sharded scan: this context's block of the gap's records (-1 = all)
    bool owned = true;
    std::vector<char> est_owned;        // estimate-level partition: which estimates this context scans (empty = all)
    std::vector<long long> cons_off;    // per estimate: offset of its consensus on this device, -1 when scanned elsewhere
    double cost = 0;
    // results
    bool done = false;
    bool no_winner = false;   // no estimate's likelihood passed 'likelihood >= maxLikelihood' (all -inf / NaN; f=1 only): bestString stays empty
    int best_est = 0;
    double best_lik = 0;
    int total_reads = 0, discarded = 0;
    long likelihoods = 0;
    std::string filled;
    std::vector<double> liks;
};

template <typename T>
struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    cudaError_t ensure(size_t want) {
        if (want <= n && p) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        n = 0;
        if (want == 0) want = 1;
        cudaError_t e = cudaMalloc((void**)&p, want * sizeof(T));
        if (e == cudaSuccess) n = want;
        return e;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        n = 0;
    }
};

template <typename T>
struct PinBuf {
    T* p = nullptr;
    size_t n = 0;
    cudaError_t ensure(size_t want) {
        if (want <= n && p) return cudaSuccess;
        if (p) cudaFreeHost(p);
        p = nullptr;
        n = 0;
        if (want == 0) want = 1;
        cudaError_t e = cudaMallocHost((void**)&p, want * sizeof(T));
        if (e == cudaSuccess) n = want;
        return e;
    }
    void release() {
        if (p) cudaFreeHost(p);
        p = nullptr;
        n = 0;
    }
};

// doubles of global workspace per CTA when the tables do not fit shared memory: pq (5*nrs + 32 pad rows of double2) + counts
inline size_t gws_stride_of(int nr_stride, int gp_stride) { return (size_t)10 * nr_stride + 64 + 5 * (size_t)gp_stride + 16; }

struct Bin {
    int first_unit, n_units;      // range in the sorted unit array
    int gpmax, lmax;   // gpmax: most gap rows the tables of a unit hold (estimate minus the rows nothing touches)
    int gp_full;       // largest estimate
    int nr_stride, gp_stride, p_stride, chunk;
    bool global_tables;
    int ctas_per_sm;
    int threads;      // per CTA: 512, or 1024 when only one CTA fits an SM
    int cluster;      // CTAs per unit (thread-block cluster) when the launch has fewer units than the GPU has CTA slots; 1 = one CTA per unit
    bool one_cta_few; // few units + wide insert window: tables sized for one CTA per SM on purpose (see plan_bins)
    size_t smem;
    int grid;         // CTAs, or clusters when cluster > 1
    size_t gws_off;   // doubles: this bin's slice of the global workspace (bins run concurrently)
};

}  // namespace

void host_pool_warm();   // creates the library's host thread pool

struct fb_ctx {
    int device = 0;
    cudaStream_t stream = nullptr, copy_stream = nullptr;
    std::vector<cudaStream_t> bin_streams;            // one per EM launch (size bin)
    std::vector<cudaEvent_t> bin_done, bin_copied;   // per EM launch: kernel finished / its results are on the host
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev2 = nullptr, ev3 = nullptr, evs = nullptr, eve = nullptr, ev_fork = nullptr;
    int max_smem = 0, sm_count = 0;
    std::string err;

    // model (host copies)
    int max_read_len = 0, thr_min = 0, thr_max = 0, cutoff = 0, D = 200, num_iter = 5;
    bool f1 = false;           // the reference's log-domain mode (argv[2] = 1, figbird.cpp:1814-1815)
    double theta = 0;          // smallest double that passes finalize's cut-off test (cutoff_threshold)
    bool host_log = false;     // FB_HOST_LOG=1: per-record maxima go home and the host libm sums their logs (the round-1 path)
    double T[25];
    std::vector<double> epos, match;
    std::vector<double2> h_eo;   // {1 - e[k], e[k]}, copied into every EM launch's parameters
    std::vector<double> h_t0;    // FbEmParams::t0_tab, computed on the device at fb_create
    // model (device)
    DevBuf<double> d_ild, d_epos, d_match;
    DevBuf<double> d_ild_em;   // what the E-step starts a score from: ILD[t], or log(ILD[t]) from the host libm in f=1 mode

    // gaps + records (host staging)
    std::vector<HostGap> gaps;
    std::vector<FbGapMeta> h_gapmeta;
    std::vector<uint8_t> h_flank;
    std::vector<uint4> h_codes, h_nmask;
    std::vector<FbRecMeta> h_rmeta;
    std::vector<uint8_t> h_bytes;     // per record FB_MAX_READ scoring-orientation codes (host finalize)
    int part_rank = 0, part_world = 1;

    // units
    std::vector<FbUnit> h_units;
    std::vector<int> unit_nrec;
    std::vector<Bin> bins;
    bool uploaded = false, ran = false, scanned = false;
    bool unit_mode = false;   // partition by (gap, estimate) instead of by gap
    bool rec_mode = false;    // record-sharded scan: every unit on every rank, records split in contiguous blocks
    size_t shard_counts = 0, shard_maxprob = 0;   // doubles per partial-count buffer / per maxima buffer

    // device
    DevBuf<FbGapMeta> d_gapmeta;
    DevBuf<uint8_t> d_flank;
    DevBuf<uint4> d_codes, d_nmask;
    DevBuf<FbRecMeta> d_rmeta;
    DevBuf<uint32_t> d_st_mask, d_st_cls4;           // per-record tables built on the device by fb_stage_kernel
    DevBuf<uint8_t> d_st_jl, d_st_cum;
    DevBuf<double> d_pre;                            // left-flank prefix table (fb_prefix_kernel), 32 doubles per chain
    size_t pre_chains = 0;                           // 0: no table (nothing to tabulate, FB_NO_PREFIX, or more chains than the flags can index)
    int pre_slices = 1;
    DevBuf<FbUnit> d_units, d_units_scan;          // _scan: the units / gap metadata a record-sharded scan launches with
    DevBuf<FbGapMeta> d_gapmeta_scan;
    DevBuf<double> d_parts[2], d_sumcounts, d_maxmerge;
    DevBuf<int> d_unit_nrec, d_shard_err;
    DevBuf<unsigned long long> d_flags;   // device-side barrier of the record-sharded scan: one word per rank, written by the peers
    unsigned long long shard_epoch = 0;   // passes run with the device-side barrier (the same on every rank)
    bool shard_poisoned = false;          // a barrier failed: flags hold FB_FLAG_POISON until the next upload resets them
    DevBuf<int> d_counters;
    DevBuf<double> d_maxprob, d_gws, d_dump;
    DevBuf<char> d_cons;
    DevBuf<int> d_gaplist;
    DevBuf<long long> d_filled_off;
    DevBuf<double> d_unit_lik;
    // per-gap outputs of selection and finalize, one allocation (one copy home): best_lik f64 | best_off i64 | best_gp i32 |
    // discarded i32 | tickets i32, each [n_gaps]
    DevBuf<unsigned char> d_gapout;
    PinBuf<unsigned char> h_gapout;
    size_t gapout_n = 0;
    double* go_lik(unsigned char* b) const { return reinterpret_cast<double*>(b); }
    long long* go_off(unsigned char* b) const { return reinterpret_cast<long long*>(b + 8 * gapout_n); }
    int* go_gp(unsigned char* b) const { return reinterpret_cast<int*>(b + 16 * gapout_n); }
    int* go_disc(unsigned char* b) const { return reinterpret_cast<int*>(b + 20 * gapout_n); }
    int* go_tick(unsigned char* b) const { return reinterpret_cast<int*>(b + 24 * gapout_n); }
    DevBuf<int> d_sel_off, d_sel_unit, d_gmin, d_bin_first;
    DevBuf<FbUnit> d_win_units;
    bool two_pass = false;       // many units per CTA slot: first pass without the last M-step, winners run again (see run_fused)
    std::vector<int> h_sel_off, h_sel_unit, h_gmin;
    // one-call pass (fb_run_device without a partition that needs a host exchange): everything is enqueued at once —
    // scan, selection, finalize, copies — and the host waits once.  The finalize geometry is fixed at upload from the
    // largest estimates, not from the winners.
    bool fused_ok = false;
    cudaGraphExec_t graph_exec = nullptr;   // the one-call pass of the current upload, captured
    bool graph_failed = false;
    int runs_since_upload = 0;
    bool fused_dirty = false;    // fb_run_finalize has overwritten the gap list / offsets the one-call pass keeps on the device
    std::vector<int> fin_gaplist, fin_counts_off;
    std::vector<long long> fin_filled_off;
    int fin_slices = 1, fin_rows = 0, fin_gbmax = 1;
    size_t fin_smem = 0, fin_counts_total = 0, fin_fill_total = 0;
    DevBuf<int> d_fin_counts, d_fin_counts_off;
    DevBuf<char> d_filled;
    PinBuf<double> h_maxprob, h_unit_lik;
    PinBuf<char> h_filled;
    PinBuf<unsigned char> h_stage;   // pinned staging for the host->device copies

    // fb_fill_gap_files: the model as fb_create received it (a second context is created from it on first use), the
    // context that runs the first batch, and whether the gaps of the last fill are spread over both
    std::vector<double> m_error_pos, m_in_pos, m_del_pos, m_ild;
    fb_model m_model;
    fb_ctx* child = nullptr;
    bool piped = false;

    fb_stats stats;
    double t_parse_ms = 0, t_pack_ms = 0;   // FB_TRACE: time spent in fb_add_gap_file since the last upload

    int fail(int code, const char* fmt, ...) {
        char buf[512];
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(buf, sizeof buf, fmt, ap);
        va_end(ap);
        err = buf;
        return code;
    }
};

// finalize keeps a record when -log(maxProb) <= gapProbCutOff (figbird.cpp:2293).  log() is monotone, so that is a
// threshold on maxProb: theta = the smallest positive double that passes, found by bisection over the bit patterns with
// the reference's own libm, then checked exhaustively for 4096 doubles either side (a libm whose log wiggled there
// would make the threshold form wrong: refuse rather than guess).  maxProb = 0 never passes (-log(0) = inf); when no
// double passes, theta = +inf.
static bool cutoff_threshold(int cutoff, double* theta) {
    auto pass = [&](double x) { return -log(x) <= cutoff; };
    auto as_d = [](uint64_t u) { double x; memcpy(&x, &u, 8); return x; };
    uint64_t lo = 1, hi = 0x7fefffffffffffffull;           // denorm_min .. DBL_MAX
    if (pass(as_d(lo))) { *theta = as_d(lo); return true; }
    if (!pass(as_d(hi))) { *theta = INFINITY; return true; }
    while (hi - lo > 1) {                                    // invariant: !pass(lo) && pass(hi)
        const uint64_t mid = lo + (hi - lo) / 2;
        if (pass(as_d(mid))) hi = mid;
        else lo = mid;
    }
    for (int64_t k = -4096; k <= 4096; k++) {
        const uint64_t u = hi + (uint64_t)k;
        if (u < 1 || u > 0x7fefffffffffffffull) continue;
        if (pass(as_d(u)) != (u >= hi)) return false;
    }
    *theta = as_d(hi);
    return true;
}

#define FB_CUDA(ctx, call)                                                                             \
    do {                                                                                               \
        cudaError_t e__ = (call);                                                                      \
        if (e__ != cudaSuccess) return (ctx)->fail(FB_ECUDA, "%s: %s", #call, cudaGetErrorString(e__)); \
    } while (0)

extern "C" {

int fb_abi_version(void) { return FB_ABI_VERSION; }

const char* fb_status_str(int s) {
    switch (s) {
        case FB_OK: return "ok";
        case FB_EINVAL: return "invalid argument";
        case FB_EDOMAIN: return "input outside the reference's defined territory";
        case FB_ECUDA: return "CUDA error";
        case FB_ENOMEM: return "out of memory";
        case FB_ESTATE: return "call out of order";
        case FB_EUNSUPPORTED: return "unsupported mode";
        default: return "unknown status";
    }
}

const char* fb_last_error(fb_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

void fb_gap_range(int orig_gap, int* gap_min, int* gap_max) {
    // figbird.cpp:2438-2453: float fractions, int*float evaluated in float, truncation on assignment.
    float lo, hi;
    if (orig_gap > 300) lo = .95;
    else if (orig_gap < 100) lo = 0.1;
    else lo = 0.85;
    if (orig_gap > 200) hi = 1.05;
    else hi = 1.5;
    *gap_min = orig_gap * lo < 10 ? 1 : orig_gap * lo;
    *gap_max = orig_gap * hi;
}

int fb_warmup(int device) {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1 || device < 0 || device >= ndev) return FB_ECUDA;
    if (cudaSetDevice(device) != cudaSuccess || cudaFree(nullptr) != cudaSuccess) return FB_ECUDA;
    int smem = 0, sms = 0;
    if (fb_kernels_init(&smem, &sms) != cudaSuccess) return FB_ECUDA;
    void* pin = nullptr;                       // the first pinned allocation sets up the host allocator
    if (cudaMallocHost(&pin, 1 << 16) == cudaSuccess) cudaFreeHost(pin);
    host_pool_warm();
    return FB_OK;
}

int fb_create(const fb_model* m, int device, fb_ctx** out) {
    if (!m || !out) return FB_EINVAL;
    *out = nullptr;
    if (m->f_mode != 0 && m->f_mode != 1) return FB_EUNSUPPORTED;
    if (m->max_read_len < 1 || m->max_distance < 1 || m->num_iter < 1 || !m->error_pos || !m->in_pos || !m->del_pos || !m->ild)
        return FB_EINVAL;
    if (m->thr_min < 0 || m->thr_max < m->thr_min || m->thr_max >= m->ild_len) return FB_EINVAL;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1 || device < 0 || device >= ndev) return FB_ECUDA;
    fb_ctx* c = new fb_ctx();
    c->device = device;
    memset(&c->stats, 0, sizeof c->stats);
    auto bail = [&](int code) {
        fb_destroy(c);
        return code;
    };
    if (cudaSetDevice(device) != cudaSuccess) return bail(FB_ECUDA);
    if (fb_kernels_init(&c->max_smem, &c->sm_count) != cudaSuccess) return bail(FB_ECUDA);
    if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) return bail(FB_ECUDA);
    if (cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking) != cudaSuccess) return bail(FB_ECUDA);
    cudaEventCreate(&c->ev0);
    cudaEventCreate(&c->ev1);
    cudaEventCreate(&c->ev2);
    cudaEventCreate(&c->ev3);
    cudaEventCreate(&c->evs);
    cudaEventCreate(&c->eve);
    cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming);
    c->max_read_len = m->max_read_len;
    c->thr_min = m->thr_min;
    c->thr_max = m->thr_max;
    c->cutoff = m->gap_prob_cutoff;
    c->D = m->max_distance;
    c->num_iter = m->num_iter;
    for (int i = 0; i < 5; i++)
        for (int j = 0; j < 5; j++) c->T[i * 5 + j] = m->error_type[i][j];
    const int n = std::min(m->max_read_len, FB_MAX_READ);
    std::vector<double2>& eo = c->h_eo;
    eo.assign(FB_MAX_READ, make_double2(0, 0));
    c->epos.assign(FB_MAX_READ, 0.0);
    c->match.assign(FB_MAX_READ, 0.0);
    for (int k = 0; k < n; k++) {
        eo[k] = make_double2(1 - m->error_pos[k], m->error_pos[k]);                 // figbird.cpp:1855
        c->epos[k] = m->error_pos[k];
        c->match[k] = 1 - m->error_pos[k] - m->in_pos[k] - m->del_pos[k];           // figbird.cpp:2272
    }
    c->m_error_pos.assign(m->error_pos, m->error_pos + m->max_read_len);
    c->m_in_pos.assign(m->in_pos, m->in_pos + m->max_read_len);
    c->m_del_pos.assign(m->del_pos, m->del_pos + m->max_read_len);
    c->m_ild.assign(m->ild, m->ild + m->ild_len);
    c->m_model = *m;
    c->m_model.error_pos = c->m_model.in_pos = c->m_model.del_pos = c->m_model.ild = nullptr;   // set when a copy is handed out
    c->f1 = m->f_mode == 1;
    c->host_log = getenv("FB_HOST_LOG") && atoi(getenv("FB_HOST_LOG")) != 0;
    if (!cutoff_threshold(m->gap_prob_cutoff, &c->theta)) {
        c->err = "the host libm's log() is not monotone around exp(-gapProbCutOff): cannot turn the cut-off test into a threshold";
        fprintf(stderr, "figbird_cuda: %s\n", c->err.c_str());
        return bail(FB_EUNSUPPORTED);
    }
    const int pwin = m->thr_max - m->thr_min + 1;
    if (c->d_ild.ensure(pwin) || c->d_ild_em.ensure(pwin) || c->d_epos.ensure(FB_MAX_READ) || c->d_match.ensure(FB_MAX_READ))
        return bail(FB_ECUDA);
    cudaMemcpy(c->d_ild.p, m->ild + m->thr_min, sizeof(double) * pwin, cudaMemcpyHostToDevice);
    {
        // f=1: tempProb starts from log(ILD[t]) (figbird.cpp:1835): taken with the host libm, once per t
        std::vector<double> start(m->ild + m->thr_min, m->ild + m->thr_min + pwin);
        if (c->f1)
            for (double& v : start) v = log(v);
        cudaMemcpy(c->d_ild_em.p, start.data(), sizeof(double) * pwin, cudaMemcpyHostToDevice);
    }
    cudaMemcpy(c->d_epos.p, c->epos.data(), sizeof(double) * FB_MAX_READ, cudaMemcpyHostToDevice);
    cudaMemcpy(c->d_match.p, c->match.data(), sizeof(double) * FB_MAX_READ, cudaMemcpyHostToDevice);
    {
        // the first iteration's gap-row terms (FbEmParams::t0_tab): by the device's own operations, then kept on the host
        DevBuf<double2> d_eo;
        DevBuf<double> d_T, d_t0;
        c->h_t0.assign(5 * FB_MAX_READ, 0.0);
        if (d_eo.ensure(FB_MAX_READ) || d_T.ensure(25) || d_t0.ensure(5 * FB_MAX_READ)) return bail(FB_ECUDA);
        cudaMemcpy(d_eo.p, eo.data(), sizeof(double2) * FB_MAX_READ, cudaMemcpyHostToDevice);
        cudaMemcpy(d_T.p, c->T, sizeof c->T, cudaMemcpyHostToDevice);
        cudaError_t e = fb_launch_t0(d_eo.p, d_T.p, d_t0.p, c->f1, nullptr);
        if (e == cudaSuccess) e = cudaMemcpy(c->h_t0.data(), d_t0.p, sizeof(double) * 5 * FB_MAX_READ, cudaMemcpyDeviceToHost);
        d_eo.release(); d_T.release(); d_t0.release();
        if (e != cudaSuccess) return bail(FB_ECUDA);
    }
    if (cudaDeviceSynchronize() != cudaSuccess) return bail(FB_ECUDA);
    *out = c;
    return FB_OK;
}

void fb_destroy(fb_ctx* c) {
    if (!c) return;
    if (c->child) fb_destroy(c->child);
    c->child = nullptr;
    cudaSetDevice(c->device);
    if (c->graph_exec) cudaGraphExecDestroy(c->graph_exec);
    c->graph_exec = nullptr;
    c->d_ild.release(); c->d_epos.release(); c->d_match.release();
    c->d_gapmeta.release(); c->d_flank.release(); c->d_codes.release(); c->d_nmask.release(); c->d_rmeta.release();
    c->d_units.release(); c->d_counters.release(); c->d_maxprob.release(); c->d_gws.release();
    c->d_dump.release(); c->d_cons.release(); c->d_gaplist.release(); c->d_gapout.release(); c->h_gapout.release();
    c->d_st_mask.release(); c->d_st_cls4.release(); c->d_st_jl.release(); c->d_st_cum.release(); c->d_pre.release();
    c->d_units_scan.release(); c->d_gapmeta_scan.release(); c->d_parts[0].release(); c->d_parts[1].release();
    c->d_sumcounts.release(); c->d_maxmerge.release(); c->d_unit_nrec.release(); c->d_shard_err.release(); c->d_flags.release();
    c->h_maxprob.release(); c->h_stage.release();
    c->d_ild_em.release(); c->d_filled_off.release(); c->d_unit_lik.release(); c->d_fin_counts.release(); c->d_fin_counts_off.release();
    c->d_sel_off.release(); c->d_sel_unit.release(); c->d_gmin.release(); c->d_bin_first.release(); c->d_win_units.release();
    c->d_filled.release(); c->h_unit_lik.release(); c->h_filled.release();
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    if (c->ev2) cudaEventDestroy(c->ev2);
    if (c->ev3) cudaEventDestroy(c->ev3);
    if (c->evs) cudaEventDestroy(c->evs);
    if (c->eve) cudaEventDestroy(c->eve);
    if (c->ev_fork) cudaEventDestroy(c->ev_fork);
    for (cudaStream_t t : c->bin_streams) cudaStreamDestroy(t);
    for (cudaEvent_t e : c->bin_done) cudaEventDestroy(e);
    for (cudaEvent_t e : c->bin_copied) cudaEventDestroy(e);
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

int fb_clear_gaps(fb_ctx* c) {
    if (!c) return FB_EINVAL;
    if (c->child) fb_clear_gaps(c->child);
    c->piped = false;
    c->gaps.clear(); c->h_gapmeta.clear(); c->h_flank.clear(); c->h_codes.clear(); c->h_nmask.clear();
    c->h_rmeta.clear(); c->h_bytes.clear(); c->h_units.clear(); c->unit_nrec.clear(); c->bins.clear();
    c->uploaded = c->ran = c->scanned = false;
    return FB_OK;
}

}  // extern "C"

namespace {

// One gap packed for the device, built without touching the context (so gaps can be packed by several threads).
struct PackedGap {
    HostGap g;
    std::vector<uint8_t> flank, bytes;
    std::vector<uint4> codes, nmask;
    std::vector<FbRecMeta> rmeta;
    int status = FB_OK;
    std::string err;
};

struct PackConsts {
    int D, thr_min, thr_max, max_read_len;
};

void pack_fail(PackedGap& out, int code, const char* fmt, ...) {
    char buf[256];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    out.status = code;
    out.err = buf;
}

void pack_gap(const PackConsts& k, int contig_no, long gap_start, int orig_gap, const char* contig, long contig_len,
              const fb_record* recs, int nrec, PackedGap& out) {
    const int D = k.D;
    // Outside these bounds the reference indexes past its arrays: refuse.
    if (orig_gap < 1 || gap_start < D || gap_start + orig_gap + D > contig_len)
        return pack_fail(out, FB_EDOMAIN, "gap at %ld (+%d) is closer than %d to a scaffold end", gap_start, orig_gap, D);
    for (int r = 0; r < nrec; r++) {
        if (recs[r].len2 < 1 || recs[r].len2 > FB_MAX_READ || recs[r].len2 > k.max_read_len || recs[r].len2 > D || !recs[r].seq2)
            return pack_fail(out, FB_EDOMAIN, "record %d: mate length %d unsupported", r, recs[r].len2);
    }
    HostGap& g = out.g;
    g.contig_no = contig_no;
    g.gap_start = gap_start;
    g.orig = orig_gap;
    fb_gap_range(orig_gap, &g.gmin, &g.gmax);
    if (g.gmax < g.gmin) return pack_fail(out, FB_EDOMAIN, "empty estimate range for gap length %d", orig_gap);
    g.rec_off = 0;
    g.nrec = nrec;
    g.lmax = 1;
    // flanks: D bases before the gap, D bases after the ORIGINAL gap (figbird.cpp:1657-1670)
    out.flank.resize(2 * (size_t)D);
    for (int i = 0; i < D; i++) {
        out.flank[i] = (uint8_t)base_code(contig[gap_start - D + i]);
        out.flank[D + i] = (uint8_t)base_code(contig[gap_start + orig_gap + i]);
    }
    const long sp = gap_start - D;
    out.bytes.assign((size_t)nrec * FB_MAX_READ, 5);
    out.codes.reserve((size_t)nrec * 2);
    out.nmask.reserve(nrec);
    out.rmeta.reserve(nrec);
    for (int r = 0; r < nrec; r++) {
        const fb_record& rc = recs[r];
        const int L2 = rc.len2;
        g.lmax = std::max(g.lmax, L2);
        const bool rev = rc.strand1 == 0;             // isReverse, figbird.cpp:1803-1812
        const bool left = rc.pos1 < gap_start;        // figbird.cpp:1822
        uint32_t cw[8] = {0, 0, 0, 0, 0, 0, 0, 0}, nm[4] = {0, 0, 0, 0};
        uint8_t* bo = &out.bytes[(size_t)r * FB_MAX_READ];
        for (int j = 0; j < L2; j++) {
            int code = rev ? revcomp_code(rc.seq2[L2 - 1 - j]) : base_code(rc.seq2[j]);
            bo[j] = (uint8_t)code;
            if (code < 4) cw[j >> 4] |= (uint32_t)code << (2 * (j & 15));
            else nm[j >> 5] |= 1u << (j & 31);
        }
        out.codes.push_back(make_uint4(cw[0], cw[1], cw[2], cw[3]));
        out.codes.push_back(make_uint4(cw[4], cw[5], cw[6], cw[7]));
        out.nmask.push_back(make_uint4(nm[0], nm[1], nm[2], nm[3]));
        FbRecMeta m;
        long tb, lo, hi;
        if (left) {          // t = u + tb
            tb = sp - rc.pos1 + L2;
            lo = k.thr_min - tb;
            hi = k.thr_max - tb;
        } else {             // t = tb - u
            tb = (long)rc.pos1 + rc.len1 - sp;
            lo = tb - k.thr_max;
            hi = tb - k.thr_min;
        }
        const long LIM = 1L << 30;
        m.tb = (int)std::max(-LIM, std::min(LIM, tb));
        m.a_lo = (int)std::max(-LIM, std::min(LIM, lo));
        m.a_hi = (int)std::max(-LIM, std::min(LIM, hi));
        m.flags = L2 | (rev ? 256 : 0) | (left ? 512 : 0);
        out.rmeta.push_back(m);
    }
}

// atoi on a token that is not NUL-terminated: blanks, an optional sign, digits (what the reference's atoi sees).
inline int tok_atoi(const char* p, const char* e) {
    while (p < e && (*p == ' ' || (*p >= '\t' && *p <= '\r'))) p++;
    bool neg = false;
    if (p < e && (*p == '+' || *p == '-')) neg = *p++ == '-';
    long v = 0;
    while (p < e && *p >= '0' && *p <= '9') v = v * 10 + (*p++ - '0');
    return (int)(neg ? -v : v);
}

// strtok(line, "\t") over [p, e): the k-th token (0-based) as [*tb, *te); runs of tabs count once, a leading run is
// skipped, the line's '\n' belongs to the last token.  false when the line has fewer tokens.
inline bool nth_token(const char* p, const char* e, int k, const char** tb, const char** te) {
    for (int t = 0;; t++) {
        while (p < e && *p == '\t') p++;
        if (p >= e) return false;
        const char* q = (const char*)memchr(p, '\t', (size_t)(e - p));
        if (!q) q = e;
        if (t == k) { *tb = p; *te = q; return true; }
        p = q;
    }
}

// placeReads' parse of one gap file (figbird.cpp:1718-1801), then pack.  The file is read in one go and cut at its
// newlines; the columns are found the way the reference's fgets + strtok("\t") finds them.  A line of 1023 bytes or
// more (which the reference's 1024-byte fgets would split) sends the whole file down the literal fgets/strtok route.
void parse_and_pack_slow(const PackConsts& k, int contig_no, long gap_start, int orig_gap, const char* contig, long contig_len,
                         const char* gap_file, PackedGap& out);

void parse_and_pack(const PackConsts& k, int contig_no, long gap_start, int orig_gap, const char* contig, long contig_len,
                    const char* gap_file, PackedGap& out) {
    FILE* fp = gap_file ? fopen(gap_file, "rb") : nullptr;
    if (!fp) return pack_fail(out, FB_EINVAL, "cannot open %s", gap_file ? gap_file : "(null)");
    std::string text;
    {
        char buf[1 << 16];
        size_t n;
        while ((n = fread(buf, 1, sizeof buf, fp)) > 0) text.append(buf, n);
    }
    fclose(fp);
    struct Line { const char *b, *e; };
    std::vector<Line> lines;
    const char* p = text.data();
    const char* end = p + text.size();
    while (p < end) {
        const char* nl = (const char*)memchr(p, '\n', (size_t)(end - p));
        const char* e = nl ? nl + 1 : end;
        if (e - p >= 1023 || memchr(p, 0, (size_t)(e - p))) return parse_and_pack_slow(k, contig_no, gap_start, orig_gap, contig, contig_len, gap_file, out);
        lines.push_back({p, e});
        p = e;
    }
    std::vector<fb_record> recs;
    std::vector<std::pair<const char*, int>> seqs;
    for (size_t i = 0; i < lines.size(); i++) {
        if (lines[i].b[0] == '@') continue;                       // figbird.cpp:1720
        if (i + 1 >= lines.size()) break;                         // the second fgets fails: placeReads stops
        const Line l1 = lines[i], l2 = lines[++i];
        const char *fb_, *fe, *pb, *pe, *sb, *se, *mb, *me;
        if (!nth_token(l1.b, l1.e, 1, &fb_, &fe) || !nth_token(l1.b, l1.e, 3, &pb, &pe) || !nth_token(l1.b, l1.e, 6, &sb, &se)) continue;
        if (!nth_token(l2.b, l2.e, 6, &mb, &me)) continue;
        fb_record r;
        r.strand1 = (tok_atoi(fb_, fe) & 16) >> 4;
        r.pos1 = tok_atoi(pb, pe);
        r.len1 = (int)(se - sb);
        r.seq2 = mb;
        r.len2 = (int)(me - mb);
        recs.push_back(r);
    }
    pack_gap(k, contig_no, gap_start, orig_gap, contig, contig_len, recs.data(), (int)recs.size(), out);
}

void parse_and_pack_slow(const PackConsts& k, int contig_no, long gap_start, int orig_gap, const char* contig, long contig_len,
                         const char* gap_file, PackedGap& out) {
    FILE* fp = gap_file ? fopen(gap_file, "r") : nullptr;
    if (!fp) return pack_fail(out, FB_EINVAL, "cannot open %s", gap_file ? gap_file : "(null)");
    std::vector<fb_record> recs;
    std::vector<std::string> seqs;
    char l1[1024], l2[1024];
    while (fgets(l1, sizeof l1, fp)) {
        if (l1[0] == '@') continue;
        if (!fgets(l2, sizeof l2, fp)) break;
        char *s1 = nullptr, *s2 = nullptr, *col[7];
        bool ok = (col[0] = strtok_r(l1, "\t", &s1)) != nullptr;
        for (int q = 1; q < 7 && ok; q++) ok = (col[q] = strtok_r(nullptr, "\t", &s1)) != nullptr;
        if (!ok) continue;
        char* m = strtok_r(l2, "\t", &s2);
        for (int q = 1; q < 7 && m; q++) m = strtok_r(nullptr, "\t", &s2);
        if (!m) continue;
        fb_record r;
        r.strand1 = (atoi(col[1]) & 16) >> 4;
        r.pos1 = atoi(col[3]);
        r.len1 = (int)strlen(col[6]);
        seqs.emplace_back(m);
        r.seq2 = nullptr;
        r.len2 = (int)seqs.back().size();
        recs.push_back(r);
    }
    fclose(fp);
    for (size_t i = 0; i < recs.size(); i++) recs[i].seq2 = seqs[i].c_str();
    pack_gap(k, contig_no, gap_start, orig_gap, contig, contig_len, recs.data(), (int)recs.size(), out);
}

int append_gap(fb_ctx* c, PackedGap& pk) {
    if (pk.status != FB_OK) return c->fail(pk.status, "%s", pk.err.c_str());
    pk.g.rec_off = (int)c->h_rmeta.size();
    c->h_flank.insert(c->h_flank.end(), pk.flank.begin(), pk.flank.end());
    c->h_bytes.insert(c->h_bytes.end(), pk.bytes.begin(), pk.bytes.end());
    c->h_codes.insert(c->h_codes.end(), pk.codes.begin(), pk.codes.end());
    c->h_nmask.insert(c->h_nmask.end(), pk.nmask.begin(), pk.nmask.end());
    c->h_rmeta.insert(c->h_rmeta.end(), pk.rmeta.begin(), pk.rmeta.end());
    FbGapMeta gm;
    gm.rec_off = pk.g.rec_off;
    gm.nrec = pk.g.nrec;
    gm.orig = pk.g.orig;
    gm.lmax = pk.g.lmax;
    c->h_gapmeta.push_back(gm);
    c->gaps.push_back(pk.g);
    c->uploaded = c->ran = c->scanned = false;
    return (int)c->gaps.size() - 1;
}

PackConsts consts_of(const fb_ctx* c) { return PackConsts{c->D, c->thr_min, c->thr_max, c->max_read_len}; }

// A small persistent pool of host threads: the library fans out three short loops per pass (parse + pack, unit
// planning, host finalize) and spawning a dozen threads each time costs more than the loops themselves.
class HostPool {
    std::vector<std::thread> workers;
    std::mutex m, run_m;
    std::condition_variable cv, done_cv;
    const std::function<void(int)>* fn = nullptr;
    std::atomic<int> next{0};
    int n = 0, limit = 0, pending = 0;
    unsigned long gen = 0;
    bool stop = false;
    const pid_t owner = getpid();

    void drain() {
        for (;;) {
            int i = next.fetch_add(1);
            if (i >= n) break;
            (*fn)(i);
        }
    }
    void loop(int id) {
        unsigned long seen = 0;
        std::unique_lock<std::mutex> lk(m);
        for (;;) {
            cv.wait(lk, [&] { return stop || gen != seen; });
            if (stop) return;
            seen = gen;
            if (id >= limit) continue;          // this job wants fewer threads
            lk.unlock();
            drain();
            lk.lock();
            if (--pending == 0) done_cv.notify_all();
        }
    }

public:
    explicit HostPool(int nthreads) {
        for (int t = 0; t < nthreads; t++) workers.emplace_back([this, t] { loop(t); });
    }
    ~HostPool() {
        {
            std::lock_guard<std::mutex> lk(m);
            stop = true;
        }
        cv.notify_all();
        for (auto& w : workers) w.join();
    }
    int size() const { return (int)workers.size(); }
    // false when another host thread is using the pool right now (the caller then runs the loop itself)
    bool run(int count, int threads, const std::function<void(int)>& f) {
        if (getpid() != owner) return false;      // a forked child has the object but not the threads
        std::unique_lock<std::mutex> busy(run_m, std::try_to_lock);
        if (!busy.owns_lock()) return false;
        {
            std::lock_guard<std::mutex> lk(m);
            fn = &f;
            n = count;
            next.store(0);
            limit = std::min(threads - 1, (int)workers.size());     // the caller works too
            pending = limit;
            gen++;
        }
        cv.notify_all();
        drain();
        std::unique_lock<std::mutex> lk(m);
        done_cv.wait(lk, [&] { return pending == 0; });
        fn = nullptr;
        return true;
    }
};

HostPool& host_pool() {      // one per process, created on first use and never destroyed (its threads end with the process)
    const int hw = (int)std::thread::hardware_concurrency();
    static HostPool* pool = new HostPool(std::max(1, std::min(hw > 0 ? hw : 1, 16) - 1));
    return *pool;
}

}  // namespace

void host_pool_warm() { host_pool(); }

namespace {

// Run fn(i) for i in [0, n) on up to `threads` host threads (0 = one per hardware thread, at most 16).
template <typename F>
void parallel_for(int n, int threads, F fn) {
    int hw = (int)std::thread::hardware_concurrency();
    if (threads <= 0) {
        // one process per GPU shares the host cores with its siblings (torchrun exports LOCAL_WORLD_SIZE)
        const char* lws = getenv("LOCAL_WORLD_SIZE");
        const int procs = lws ? std::max(1, atoi(lws)) : 1;
        // ... but never fewer than 4 threads per rank: the host loops are short bursts that rarely coincide across ranks
        threads = std::max(std::min(4, hw > 0 ? hw : 1), std::min((hw > 0 ? hw : 1) / procs, 16));
    }
    threads = std::min(threads, n);
    if (threads <= 1) {
        for (int i = 0; i < n; i++) fn(i);
        return;
    }
    const std::function<void(int)> f = fn;
    if (host_pool().run(n, threads, f)) return;
    for (int i = 0; i < n; i++) fn(i);      // pool in use by another context's thread: do it here
}

}  // namespace

extern "C" {

int fb_add_gap(fb_ctx* c, int contig_no, long gap_start, int orig_gap, const char* contig, long contig_len,
               const fb_record* recs, int nrec) {
    if (!c || !contig || nrec < 0 || (nrec > 0 && !recs)) return FB_EINVAL;
    if (c && c->piped) return c->fail(FB_ESTATE, "the gaps of fb_fill_gap_files can only be read or run again (fb_run_device); fb_clear_gaps starts over");
    PackedGap pk;
    pack_gap(consts_of(c), contig_no, gap_start, orig_gap, contig, contig_len, recs, nrec, pk);
    return append_gap(c, pk);
}

int fb_add_gap_file(fb_ctx* c, int contig_no, long gap_start, int orig_gap, const char* contig, long contig_len, const char* gap_file) {
    if (!c || !gap_file || !contig) return FB_EINVAL;
    if (c && c->piped) return c->fail(FB_ESTATE, "the gaps of fb_fill_gap_files can only be read or run again (fb_run_device); fb_clear_gaps starts over");
    PackedGap pk;
    parse_and_pack(consts_of(c), contig_no, gap_start, orig_gap, contig, contig_len, gap_file, pk);
    return append_gap(c, pk);
}

int fb_add_gap_files(fb_ctx* c, const fb_gap_desc* descs, int n, int* handles, int threads) {
    if (!c || n < 0 || (n > 0 && (!descs || !handles))) return FB_EINVAL;
    if (c && c->piped) return c->fail(FB_ESTATE, "the gaps of fb_fill_gap_files can only be read or run again (fb_run_device); fb_clear_gaps starts over");
    const PackConsts k = consts_of(c);
    std::vector<PackedGap> packed((size_t)n);
    const auto t0 = std::chrono::steady_clock::now();
    parallel_for(n, threads, [&](int i) {
        const fb_gap_desc& d = descs[i];
        if (!d.contig) pack_fail(packed[i], FB_EINVAL, "gap %d: null contig", i);
        else parse_and_pack(k, d.contig_no, d.gap_start, d.orig_gap, d.contig, d.contig_len, d.gap_file, packed[i]);
    });
    const auto t1 = std::chrono::steady_clock::now();
    // appended in the order given (handles are dense and deterministic): sizes first, one resize per array, then the
    // copies on the host threads
    int bad = 0;
    std::vector<size_t> rec_at((size_t)n + 1, c->h_rmeta.size());
    std::vector<int> slot((size_t)n, -1);
    int ngood = 0;
    for (int i = 0; i < n; i++) {
        rec_at[i + 1] = rec_at[i];
        if (packed[i].status != FB_OK) {
            handles[i] = c->fail(packed[i].status, "%s", packed[i].err.c_str());
            bad++;
            continue;
        }
        slot[i] = (int)c->gaps.size() + ngood++;
        handles[i] = slot[i];
        rec_at[i + 1] += packed[i].rmeta.size();
    }
    const size_t g0 = c->gaps.size(), nr_new = rec_at[n];
    const size_t D2 = 2 * (size_t)c->D;
    c->gaps.resize(g0 + ngood);
    c->h_gapmeta.resize(g0 + ngood);
    c->h_flank.resize((g0 + ngood) * D2);
    c->h_rmeta.resize(nr_new);
    c->h_nmask.resize(nr_new);
    c->h_codes.resize(nr_new * 2);
    c->h_bytes.resize(nr_new * FB_MAX_READ);
    parallel_for(n, threads, [&](int i) {
        if (slot[i] < 0) return;
        PackedGap& pk = packed[i];
        const size_t r0 = rec_at[i], nr = pk.rmeta.size();
        pk.g.rec_off = (int)r0;
        memcpy(&c->h_flank[(size_t)slot[i] * D2], pk.flank.data(), D2);
        if (nr) {
            memcpy(&c->h_bytes[r0 * FB_MAX_READ], pk.bytes.data(), nr * FB_MAX_READ);
            memcpy(&c->h_codes[r0 * 2], pk.codes.data(), nr * 2 * sizeof(uint4));
            memcpy(&c->h_nmask[r0], pk.nmask.data(), nr * sizeof(uint4));
            memcpy(&c->h_rmeta[r0], pk.rmeta.data(), nr * sizeof(FbRecMeta));
        }
        FbGapMeta gm;
        gm.rec_off = pk.g.rec_off;
        gm.nrec = pk.g.nrec;
        gm.orig = pk.g.orig;
        gm.lmax = pk.g.lmax;
        c->h_gapmeta[slot[i]] = gm;
        c->gaps[slot[i]] = pk.g;
    });
    c->uploaded = c->ran = c->scanned = false;
    c->t_parse_ms += std::chrono::duration<double, std::milli>(t1 - t0).count();                                  // FB_TRACE: parse + pack (threads)
    c->t_pack_ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t1).count();     // FB_TRACE: append (serial)
    if (!bad) return FB_OK;
    const std::string last = c->err;
    return c->fail(FB_EDOMAIN, "%d of %d gaps were refused (see handles); last: %s", bad, n, last.c_str());
}

}  // extern "C"

namespace {

// placements of record m that pass the window test for estimate gp
inline long placements(const FbRecMeta& m, int D, int gp) {
    const int L2 = m.flags & 0xff;
    long lo = std::max(m.a_lo, D - L2 + 1), hi = std::min(m.a_hi, D + gp - 1);
    return hi >= lo ? hi - lo + 1 : 0;
}

void gap_cost(fb_ctx* c, HostGap& g) {
    double cost = 0;
    for (int gp = g.gmin; gp <= g.gmax; gp++)
        for (int r = 0; r < g.nrec; r++) {
            const FbRecMeta& m = c->h_rmeta[g.rec_off + r];
            cost += (double)placements(m, c->D, gp) * (m.flags & 0xff);
        }
    g.cost = cost + 1000.0 * (g.gmax - g.gmin + 1);
}

}  // namespace

extern "C" {

int fb_partition_plan(const double* cost, int n, int world, int* owner) {
    if (n < 0 || world < 1 || (n > 0 && (!cost || !owner))) return FB_EINVAL;
    // longest-processing-time first: items by descending cost, each to the least loaded rank
    std::vector<int> order(n);
    for (int i = 0; i < n; i++) order[i] = i;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return cost[a] > cost[b]; });
    std::vector<double> load(world, 0.0);
    for (int i : order) {
        int best = 0;
        for (int r = 1; r < world; r++)
            if (load[r] < load[best]) best = r;
        load[best] += cost[i];
        owner[i] = best;
    }
    return FB_OK;
}

double fb_gap_cost(fb_ctx* c, int gap) {
    if (!c || gap < 0 || gap >= (int)c->gaps.size()) return -1.0;
    gap_cost(c, c->gaps[gap]);
    return c->gaps[gap].cost;
}

int fb_set_partition(fb_ctx* c, int rank, int world) {
    if (c && c->piped) return c->fail(FB_ESTATE, "the gaps of fb_fill_gap_files can only be read or run again (fb_run_device); fb_clear_gaps starts over");
    if (!c || world < 1 || rank < 0 || rank >= world) return FB_EINVAL;
    c->part_rank = rank;
    c->part_world = world;
    const int n = (int)c->gaps.size();
    std::vector<double> cost(n);
    std::vector<int> owner(n);
    for (int i = 0; i < n; i++) {
        gap_cost(c, c->gaps[i]);
        cost[i] = c->gaps[i].cost;
    }
    fb_partition_plan(cost.data(), n, world, owner.data());
    for (int i = 0; i < n; i++) {
        c->gaps[i].owned = owner[i] == rank;
        c->gaps[i].est_owned.clear();
    }
    c->unit_mode = false;
    c->rec_mode = false;
    for (HostGap& g : c->gaps) { g.scan_off = 0; g.scan_n = -1; }
    c->uploaded = c->ran = c->scanned = false;
    return FB_OK;
}

int fb_set_partition_units(fb_ctx* c, int rank, int world) {
    if (c && c->piped) return c->fail(FB_ESTATE, "the gaps of fb_fill_gap_files can only be read or run again (fb_run_device); fb_clear_gaps starts over");
    if (!c || world < 1 || rank < 0 || rank >= world) return FB_EINVAL;
    c->part_rank = rank;
    c->part_world = world;
    c->unit_mode = world > 1;
    c->rec_mode = false;
    for (HostGap& g : c->gaps) { g.scan_off = 0; g.scan_n = -1; }
    // every (gap, estimate) unit is an independent EM run (figbird.cpp:2468-2501): deal units, not gaps
    std::vector<double> cost;
    std::vector<std::pair<int, int>> id;
    for (size_t gi = 0; gi < c->gaps.size(); gi++) {
        HostGap& g = c->gaps[gi];
        g.owned = true;
        g.est_owned.assign(g.gmax - g.gmin + 1, 0);
        for (int gp = g.gmin; gp <= g.gmax; gp++) {
            double t = 2000.0;
            for (int r = 0; r < g.nrec; r++) {
                const FbRecMeta& m = c->h_rmeta[g.rec_off + r];
                t += (double)placements(m, c->D, gp) * (m.flags & 0xff);
            }
            cost.push_back(t);
            id.emplace_back((int)gi, gp);
        }
    }
    std::vector<int> owner(cost.size());
    fb_partition_plan(cost.data(), (int)cost.size(), world, owner.data());
    for (size_t k = 0; k < id.size(); k++)
        if (owner[k] == rank) c->gaps[id[k].first].est_owned[id[k].second - c->gaps[id[k].first].gmin] = 1;
    c->uploaded = c->ran = c->scanned = false;
    return FB_OK;
}

int fb_set_partition_records(fb_ctx* c, int rank, int world) {
    if (c && c->piped) return c->fail(FB_ESTATE, "the gaps of fb_fill_gap_files can only be read or run again (fb_run_device); fb_clear_gaps starts over");
    if (!c || world < 1 || world > FB_MAX_SHARDS || rank < 0 || rank >= world) return FB_EINVAL;
    if (c->f1) return c->fail(FB_EUNSUPPORTED, "the record-sharded scan runs f=0 only (gap and estimate partitions run both modes)");
    c->part_rank = rank;
    c->part_world = world;
    c->unit_mode = false;
    c->rec_mode = true;
    for (HostGap& g : c->gaps) {
        g.owned = true;
        g.est_owned.clear();
        // contiguous blocks in file order: the rank-ordered sum of the partial counts then differs from the
        // reference's record-by-record sum (figbird.cpp:1718, 1879) only in how the additions are grouped
        g.scan_off = (int)((long)rank * g.nrec / world);
        g.scan_n = (int)((long)(rank + 1) * g.nrec / world) - g.scan_off;
    }
    c->uploaded = c->ran = c->scanned = false;
    return FB_OK;
}

int fb_owns_gap(fb_ctx* c, int gap) {
    if (c && c->piped && c->child && gap >= (int)c->gaps.size()) return fb_owns_gap(c->child, gap - (int)c->gaps.size());
    if (!c || gap < 0 || gap >= (int)c->gaps.size()) return 0;
    if (c->rec_mode) return gap % c->part_world == c->part_rank ? 1 : 0;   // who finalizes the gap
    return c->gaps[gap].owned ? 1 : 0;
}

}  // extern "C"

namespace {

// fb_upload, step 1 — the work list: every (owned gap, estimate) unit, ordered for the persistent CTAs, and the size bins.
// mp / co: doubles of per-record maxima and bytes of consensus strings the units need.
int plan_units(fb_ctx* c, Trace& tr, long long& mp_out, long long& co_out) {
    const int D = c->D;
    // ---- units: every (owned gap, estimate); heaviest first so persistent CTAs balance (LPT).
    // The per-gap sums are independent: computed on several host threads, merged in gap order.
    struct U { int gap, gp; double cost; long liks; int hole_lo, hole_len; int key; };
    struct GapUnits { std::vector<U> us; long liks = 0, terms = 0, bytes = 0; };
    fb_stats& st = c->stats;
    memset(&st, 0, sizeof st);
    std::vector<GapUnits> per_gap(c->gaps.size());
    const bool compact = !c->rec_mode && !getenv("FB_NO_COMPACT");
    // bins by the rows a unit's tables hold (the shared-memory footprint decides occupancy)
    static const int edges[] = {96, 160, 256, 384, 544, 768, 1024, 1536, 2048, 3072, 4096, 6144, 8192, 1 << 30};
    auto bin_of = [&](int gp) {
        int b = 0;
        while (gp > edges[b]) b++;
        return b;
    };
    auto sort_key = [&](int bin, double cost) {          // ascending key = the launch order
        const uint64_t cst = std::max<uint64_t>(8, (uint64_t)cost);
        const int msb = 63 - __builtin_clzll(cst);
        const int cls = std::min(383, msb * 8 + (int)((cst >> (msb - 3)) & 7));
        return (15 - bin) * 384 + (383 - cls);
    };
    parallel_for((int)c->gaps.size(), 0, [&](int gi) {
        HostGap& g = c->gaps[gi];
        g.done = false;
        g.likelihoods = 0;
        if (!g.owned) return;
        GapUnits& out = per_gap[gi];
        long rec_bytes = 0;      // packed bases + mask + metadata in, per-record maximum out
        const int r0 = g.scan_off, r1 = g.scan_off + (g.scan_n < 0 ? g.nrec : g.scan_n);   // the records this context scans
        // Rows of the gap that no placement of any record can touch, whatever the estimate: the largest hole in the union
        // of the records' row intervals [lo, a_hi + L2 - 1] (estimate-independent: clamping to an estimate only shrinks
        // them).  Long gaps are reached an insert length deep from either flank and nowhere else.
        long hole0 = 0, hole1 = 0, reach = D;          // best internal hole [hole0, hole1), end of the last interval
        if (compact) {
            std::vector<std::pair<long, long>> iv;
            iv.reserve(g.nrec);
            for (int r = 0; r < g.nrec; r++) {
                const FbRecMeta& m = c->h_rmeta[g.rec_off + r];
                const int L2 = m.flags & 0xff;
                const long lo = std::max(m.a_lo, D - L2 + 1);
                if (m.a_hi >= lo) iv.emplace_back(lo, (long)m.a_hi + L2);      // rows [lo, a_hi + L2)
            }
            std::sort(iv.begin(), iv.end());
            for (const auto& v : iv) {
                if (v.first > reach && v.first - reach > hole1 - hole0) { hole0 = reach; hole1 = v.first; }
                reach = std::max(reach, v.second);
            }
        }
        for (int r = r0; r < r1; r++) {
            int L2 = c->h_rmeta[g.rec_off + r].flags & 0xff;
            rec_bytes += (L2 + 3) / 4 + (L2 + 7) / 8 + 8 + 8;
        }
        for (int gp = g.gmin; gp <= g.gmax; gp++) {
            if (!g.est_owned.empty() && !g.est_owned[gp - g.gmin]) continue;
            long liks = 0, terms = 0;
            for (int r = r0; r < r1; r++) {
                const FbRecMeta& m = c->h_rmeta[g.rec_off + r];
                long pl = placements(m, D, gp);
                liks += pl;
                terms += pl * (m.flags & 0xff);
            }
            long order_terms = terms;
            if (c->rec_mode) {      // the unit order fixes the layout of the exchange buffers: it must not depend on the rank
                order_terms = 0;
                for (int r = 0; r < g.nrec; r++) {
                    const FbRecMeta& m = c->h_rmeta[g.rec_off + r];
                    order_terms += placements(m, D, gp) * (m.flags & 0xff);
                }
            }
            int hole_lo = 0, hole_len = 0;
            if (compact) {
                // this estimate's candidates: the internal hole, or everything behind the last interval; both end at the
                // right flank at the latest; the start is rounded up to a multiple of 32 gap rows (M-step row blocks)
                const long gap_end = (long)D + gp;
                long a0 = hole0, a1 = std::min(hole1, gap_end), b0 = reach, b1 = gap_end;
                if (b1 - b0 > a1 - a0) { a0 = b0; a1 = b1; }
                a0 = D + ((a0 - D + 31) / 32) * 32;
                if (a1 - a0 >= 64) { hole_lo = (int)a0; hole_len = (int)(a1 - a0); }
            }
            const double cost = (double)order_terms + 2000.0;
            out.us.push_back({gi, gp, cost, liks, hole_lo, hole_len, sort_key(bin_of(gp - hole_len), cost)});
            out.liks += liks * c->num_iter;
            out.terms += terms * c->num_iter;
            out.bytes += (rec_bytes + 2L * 40 * gp) * c->num_iter;   // + counts out / profile in
        }
        g.likelihoods = out.liks;
    });
    tr.lap("upload:   per-gap planning");
    // Order: big bins first (they have the longest units), inside a bin by descending cost class (an eighth of a power
    // of two: persistent CTAs pull the heaviest units first, finer than that does not change the balance), inside a
    // class in (gap, estimate) order.  One counting sort over the units as the gaps planned them — a gap set of
    // chr14's size has several hundred thousand units.
    // Both passes run on the host threads: the gaps are cut into contiguous slices, every slice counts its units per key,
    // a serial prefix over (key, slice) gives every slice its place inside every key (slices in gap order, so the result
    // is the serial sort's), and the slices scatter.  The running offsets of the unit array (per-record maxima, consensus
    // bytes) are a prefix sum along the sorted order: per-slice sums, a serial prefix over the slices, then the fill.
    std::vector<U> us;
    const int kKeys = 16 * 384;
    {
        for (size_t gi = 0; gi < c->gaps.size(); gi++) {
            if (!c->gaps[gi].owned) continue;
            st.n_gaps++;
            st.n_records += c->gaps[gi].scan_n < 0 ? c->gaps[gi].nrec : c->gaps[gi].scan_n;
            st.likelihoods += per_gap[gi].liks;
            st.base_terms += per_gap[gi].terms;
            st.algorithmic_bytes += per_gap[gi].bytes;
        }
        const int ng = (int)c->gaps.size();
        const int S = std::max(1, std::min(16, ng / 64));            // slices of gaps
        std::vector<std::vector<size_t>> cnt((size_t)S, std::vector<size_t>(kKeys, 0));
        auto slice_lo = [&](int sl) { return (int)((long)ng * sl / S); };
        parallel_for(S, 0, [&](int sl) {
            std::vector<size_t>& h = cnt[sl];
            for (int gi = slice_lo(sl); gi < slice_lo(sl + 1); gi++)
                if (c->gaps[gi].owned)
                    for (const U& u : per_gap[gi].us) h[u.key]++;
        });
        size_t run = 0;
        for (int k = 0; k < kKeys; k++)
            for (int sl = 0; sl < S; sl++) {
                const size_t n = cnt[sl][k];
                cnt[sl][k] = run;
                run += n;
            }
        us.resize(run);
        parallel_for(S, 0, [&](int sl) {
            std::vector<size_t>& at = cnt[sl];
            for (int gi = slice_lo(sl); gi < slice_lo(sl + 1); gi++)
                if (c->gaps[gi].owned)
                    for (const U& u : per_gap[gi].us) us[at[u.key]++] = u;
        });
    }
    st.n_units = (long)us.size();
    tr.lap("upload:   sort");
    c->h_units.resize(us.size());
    c->unit_nrec.resize(us.size());
    c->bins.clear();
    long long mp = 0, co = 0;
    {
        const size_t n = us.size();
        const int S = (int)std::max<size_t>(1, std::min<size_t>(16, n / 4096));
        auto lo_of = [&](int sl) { return n * (size_t)sl / (size_t)S; };
        std::vector<long long> smp((size_t)S + 1, 0), sco((size_t)S + 1, 0);
        parallel_for(S, 0, [&](int sl) {
            long long a = 0, b = 0;
            for (size_t i = lo_of(sl); i < lo_of(sl + 1); i++) {
                a += c->gaps[us[i].gap].nrec;
                b += us[i].gp;
            }
            smp[sl + 1] = a;
            sco[sl + 1] = b;
        });
        for (int sl = 0; sl < S; sl++) {
            smp[sl + 1] += smp[sl];
            sco[sl + 1] += sco[sl];
        }
        mp = smp[S];
        co = sco[S];
        parallel_for(S, 0, [&](int sl) {
            long long a = smp[sl], b = sco[sl];
            for (size_t i = lo_of(sl); i < lo_of(sl + 1); i++) {
                FbUnit u;
                u.gap = us[i].gap;
                u.gp = us[i].gp;
                u.mp_off = a;
                u.cons_off = b;
                u.hole_lo = us[i].hole_lo;
                u.hole_len = us[i].hole_len;
                a += c->gaps[u.gap].nrec;
                b += u.gp;
                c->h_units[i] = u;
                c->unit_nrec[i] = c->gaps[u.gap].nrec;
            }
        });
        // bins = maximal runs of one size bin in the sorted order (the key's leading part); their extents by a scan of the runs
        size_t i0 = 0;
        while (i0 < n) {
            const int b = bin_of(us[i0].gp - us[i0].hole_len);
            size_t i1 = i0 + 1;
            // keys of one bin are contiguous: binary search for the first unit of another bin
            size_t lo = i0 + 1, hi = n;
            while (lo < hi) {
                const size_t mid = lo + (hi - lo) / 2;
                if (bin_of(us[mid].gp - us[mid].hole_len) == b) lo = mid + 1;
                else hi = mid;
            }
            i1 = lo;
            Bin nb;
            memset(&nb, 0, sizeof nb);
            nb.first_unit = (int)i0;
            nb.n_units = (int)(i1 - i0);
            const int P = (int)std::max<size_t>(1, std::min<size_t>(16, (i1 - i0) / 4096));
            std::vector<int> g1(P, 0), g2(P, 0), g3(P, 0);
            parallel_for(P, 0, [&](int q) {
                const size_t a = i0 + (i1 - i0) * (size_t)q / (size_t)P, e = i0 + (i1 - i0) * (size_t)(q + 1) / (size_t)P;
                int m1 = 0, m2 = 0, m3 = 0;
                for (size_t i = a; i < e; i++) {
                    m1 = std::max(m1, us[i].gp - us[i].hole_len);
                    m2 = std::max(m2, us[i].gp);
                    m3 = std::max(m3, c->gaps[us[i].gap].lmax);
                }
                g1[q] = m1; g2[q] = m2; g3[q] = m3;
            });
            for (int q = 0; q < P; q++) {
                nb.gpmax = std::max(nb.gpmax, g1[q]);
                nb.gp_full = std::max(nb.gp_full, g2[q]);
                nb.lmax = std::max(nb.lmax, g3[q]);
            }
            c->bins.push_back(nb);
            i0 = i1;
        }
    }
    tr.lap("upload:   unit array");
    mp_out = mp;
    co_out = co;
    return FB_OK;
}

// fb_upload, step 2 — launch geometry of every bin: table strides, records per round, CTA size, cluster size, grid.
int plan_bins(fb_ctx* c, int pwin, size_t& gws_need_out, Trace& tr) {
    // ---- shared-memory geometry per bin
    size_t gws_need = 0;
    long total_units = 0;
    for (const Bin& b : c->bins) total_units += b.n_units;
    // Few units (a single gap: its estimates are all the independent work there is): with at most one unit per SM anyway,
    // size the tables for ONE CTA per SM — 16 or 32 records per round (two CTAs per SM leave room for 4 with a wide insert
    // window: many short rounds, each ending in a barrier), 32 warps — and give every unit as many SMs (a thread-block
    // cluster) as keeps all units co-resident.  cfg1: 0.88 -> 0.53 ms per scan.
    const bool few_units = total_units <= c->sm_count && !getenv("FB_CLUSTER") && !getenv("FB_ONE_CTA") && !getenv("FB_NO_FEW");
    for (Bin& b : c->bins) {
        b.nr_stride = b.gpmax + 2 * b.lmax - 2;
        b.gp_stride = (b.gpmax + 1) & ~1;
        b.p_stride = std::max(1, std::min(pwin, b.gp_full + b.lmax - 1));
        const size_t two_cta = (size_t)(c->max_smem + 1024) / 2 - 1024 - 512;   // two CTAs per SM incl. 1 KB system reserve each
        const size_t one_cta = (size_t)c->max_smem - 1024;
        b.global_tables = false;
        b.chunk = 0;
        int ctas = 2;
        // tuning knob: records per pipeline round (the M-step sets up one record per lane, so at most 32)
        const int pref = getenv("FB_CHUNK") ? std::max(1, std::min(32, atoi(getenv("FB_CHUNK")))) : 16;
        bool force_one = getenv("FB_ONE_CTA") && atoi(getenv("FB_ONE_CTA")) != 0;   // experiment knob: size for one CTA per SM
        if (few_units) force_one = true;
        b.one_cta_few = few_units && force_one;
        for (int want : {pref, 8, 4}) {
            if (want < 1 || force_one) continue;
            const size_t need = fb_em_smem_bytes(b.nr_stride, b.gp_stride, b.p_stride, want, pwin, false);
            if (need > two_cta) continue;
            // the estimate above leaves 512 B for the kernel's static shared memory; ask the runtime whether two CTAs
            // of this size really are co-resident before planning a grid of 2 x SMs
            int fit = 0;
            if (fb_em_ctas_per_sm(FB_THREADS, false, c->f1, need, &fit) != cudaSuccess) { cudaGetLastError(); fit = 0; }
            if (fit >= 2) { b.chunk = want; break; }
        }
        if (!b.chunk) {
            ctas = 1;
            for (int want : {32, 16, 8, 4}) {
                if (fb_em_smem_bytes(b.nr_stride, b.gp_stride, b.p_stride, want, pwin, false) <= one_cta) { b.chunk = want; break; }
            }
        }
        if (!b.chunk) {
            b.global_tables = true;
            for (int want : {32, 16, 8, 4, 2, 1}) {
                if (fb_em_smem_bytes(b.nr_stride, b.gp_stride, b.p_stride, want, pwin, true) <= one_cta) { b.chunk = want; break; }
            }
            if (!b.chunk) return c->fail(FB_EINVAL, "insert window of %d placements does not fit shared memory", pwin);
        }
        b.smem = fb_em_smem_bytes(b.nr_stride, b.gp_stride, b.p_stride, b.chunk, pwin, b.global_tables);
        b.grid = std::min(b.n_units, c->sm_count * ctas);
        b.cluster = 1;
        b.ctas_per_sm = ctas;
        // one CTA per SM: give it 32 warps instead of 16 (tuning knob FB_THREADS1=512 keeps the small CTA)
        {
            const char* k1 = getenv("FB_THREADS1");
            const char* kg = getenv("FB_THREADS_GT");
            b.threads = ctas == 1 ? 1024 : FB_THREADS;
            if (ctas == 1 && !b.global_tables && k1 && atoi(k1) == 512) b.threads = 512;
            if (b.global_tables && kg && atoi(kg) == 512) b.threads = 512;
        }
        b.gws_off = gws_need;
        if (b.global_tables) gws_need += (size_t)b.grid * gws_stride_of(b.nr_stride, b.gp_stride);
    }
    // Few units (a single gap, the stress sweep): most SMs would idle with one CTA per unit.  Give every unit a
    // thread-block cluster instead — as many CTAs as keep all units of all bins co-resident, at most 4.
    {
        long cta_need = 0;      // SM-halves the launches would occupy with one CTA per unit
        for (const Bin& b : c->bins) cta_need += (long)b.n_units * (b.ctas_per_sm == 2 ? 1 : 2);
        int want = 1;
        const char* force = getenv("FB_CLUSTER");        // tuning / test knob: 1 disables, 2/4/8 forces
        if (force) want = std::max(1, std::min(8, atoi(force)));
        else
            while (want < 4 && cta_need * want * 2 <= 2L * c->sm_count) want *= 2;   // 8 measured slower than 4: too few tasks per CTA and round
        for (Bin& b : c->bins) {
            if (b.global_tables) continue;
            int bwant = want;
            if (b.one_cta_few) {                 // one CTA per SM: clusters of as many SMs as keep every unit resident
                bwant = 1;
                while (bwant < 4 && total_units * bwant * 2 <= c->sm_count) bwant *= 2;
            }
            if (b.ctas_per_sm == 2 || (getenv("FB_THREADS1") && atoi(getenv("FB_THREADS1")) == 512)) b.threads = FB_THREADS;
            for (int cl = bwant; cl > 1; cl >>= 1) {
                int fit = 0;
                if (fb_em_max_clusters(cl, b.threads, c->f1, b.smem, &fit) == cudaSuccess && fit > 0) {
                    b.cluster = cl;
                    b.grid = std::min(b.n_units, fit);
                    break;
                }
                cudaGetLastError();
            }
        }
    }
    if (tr.on)
        for (const Bin& b : c->bins)
            fprintf(stderr, "[fb_trace] bin: %d units, estimates <= %d, chunk %d, smem %zu, %s, cluster %d, grid %d\n", b.n_units, b.gpmax,
                    b.chunk, b.smem, b.global_tables ? "global tables" : b.ctas_per_sm == 2 ? "2 CTAs/SM" : "1 CTA/SM", b.cluster, b.grid);
    if (tr.on)
        for (const Bin& b : c->bins) fprintf(stderr, "[fb_trace]      threads %d\n", b.threads);
    tr.lap("upload: units + bins");
    gws_need_out = gws_need;
    return FB_OK;
}

void drop_graph(fb_ctx* c) {
    if (c->graph_exec) cudaGraphExecDestroy(c->graph_exec);
    c->graph_exec = nullptr;
}

// fb_upload, step 3 — streams and events per bin, device buffers, host -> device copies.
int upload_buffers(fb_ctx* c, long long mp, long long co, size_t gws_need, Trace& tr) {
    fb_stats& st = c->stats;
    while (c->bin_done.size() < c->bins.size()) {
        cudaEvent_t a, b;
        FB_CUDA(c, cudaEventCreateWithFlags(&a, cudaEventDisableTiming));
        FB_CUDA(c, cudaEventCreateWithFlags(&b, cudaEventDisableTiming));
        c->bin_done.push_back(a);
        c->bin_copied.push_back(b);
        cudaStream_t bs;
        FB_CUDA(c, cudaStreamCreateWithFlags(&bs, cudaStreamNonBlocking));
        c->bin_streams.push_back(bs);
    }
    // ---- device buffers + host->device copies
    const size_t ng = c->h_gapmeta.size(), nr = c->h_rmeta.size();
    FB_CUDA(c, c->d_gapmeta.ensure(ng));
    FB_CUDA(c, c->d_flank.ensure(c->h_flank.size()));
    FB_CUDA(c, c->d_codes.ensure(c->h_codes.size()));
    FB_CUDA(c, c->d_nmask.ensure(c->h_nmask.size()));
    FB_CUDA(c, c->d_rmeta.ensure(nr));
    FB_CUDA(c, c->d_st_mask.ensure(nr * 20));
    FB_CUDA(c, c->d_st_cls4.ensure(nr * 16));
    FB_CUDA(c, c->d_st_jl.ensure(nr * (FB_MAX_READ + 16)));
    FB_CUDA(c, c->d_st_cum.ensure(nr * 32));
    FB_CUDA(c, c->d_units.ensure(c->h_units.size()));
    FB_CUDA(c, c->d_counters.ensure(c->bins.size() + 1));
    FB_CUDA(c, c->d_maxprob.ensure((size_t)mp));
    FB_CUDA(c, c->d_cons.ensure((size_t)co));
    FB_CUDA(c, c->d_gaplist.ensure(ng));
    c->gapout_n = std::max<size_t>(ng, 1);
    FB_CUDA(c, c->d_gapout.ensure(28 * c->gapout_n));
    FB_CUDA(c, c->h_gapout.ensure(28 * c->gapout_n));
    FB_CUDA(c, c->d_filled_off.ensure(ng));
    FB_CUDA(c, c->d_fin_counts_off.ensure(ng));
    FB_CUDA(c, c->d_unit_lik.ensure(c->h_units.size()));
    FB_CUDA(c, c->h_unit_lik.ensure(c->h_units.size()));
    {
        size_t fill_cap = 0;                             // the filled sequences of all gaps, at most gmax bases each
        for (const HostGap& g : c->gaps) fill_cap += (size_t)std::max(g.gmax, 1);
        FB_CUDA(c, c->d_filled.ensure(fill_cap));
        FB_CUDA(c, c->h_filled.ensure(fill_cap));
    }
    if (gws_need) FB_CUDA(c, c->d_gws.ensure(gws_need));
    if (c->host_log || c->rec_mode) FB_CUDA(c, c->h_maxprob.ensure((size_t)mp));
    // ---- selection map and finalize geometry of the one-call pass
    c->fused_ok = !c->unit_mode && !c->rec_mode && !c->host_log;
    c->fused_dirty = false;
    c->runs_since_upload = 0;
    drop_graph(c);
    if (c->fused_ok) {
        c->h_sel_off.assign(ng + 1, 0);
        c->h_gmin.assign(ng, 0);
        for (size_t gi = 0; gi < ng; gi++) {
            const HostGap& g = c->gaps[gi];
            c->h_gmin[gi] = g.gmin;
            c->h_sel_off[gi + 1] = c->h_sel_off[gi] + (g.owned ? g.gmax - g.gmin + 1 : 0);
        }
        c->h_sel_unit.assign((size_t)c->h_sel_off[ng], -1);
        {
            const size_t n = c->h_units.size();
            const int S = (int)std::max<size_t>(1, std::min<size_t>(16, n / 8192));
            parallel_for(S, 0, [&](int sl) {
                for (size_t i = n * (size_t)sl / (size_t)S; i < n * (size_t)(sl + 1) / (size_t)S; i++) {
                    const FbUnit& u = c->h_units[i];
                    c->h_sel_unit[(size_t)c->h_sel_off[u.gap] + (size_t)(u.gp - c->gaps[u.gap].gmin)] = (int)i;
                }
            });
        }
        c->fin_gaplist.clear();
        c->fin_counts_off.assign(ng, 0);
        c->fin_filled_off.assign(ng, 0);
        c->fin_rows = 0;
        c->fin_gbmax = 1;
        c->fin_fill_total = c->fin_counts_total = 0;
        int max_nrec = 1;
        for (size_t gi = 0; gi < ng; gi++) {
            const HostGap& g = c->gaps[gi];
            if (!g.owned) continue;
            c->fin_gaplist.push_back((int)gi);
            c->fin_rows = std::max(c->fin_rows, g.gmax + 2 * g.lmax);
            c->fin_gbmax = std::max(c->fin_gbmax, g.gmax);
            max_nrec = std::max(max_nrec, g.nrec);
            c->fin_filled_off[gi] = (long long)c->fin_fill_total;
            c->fin_fill_total += (size_t)std::max(g.gmax, 1);
        }
        const int per_cta = FB_FIN_THREADS / 32, nfin = std::max<int>(1, (int)c->fin_gaplist.size());
        c->fin_slices = std::max(1, std::min((4 * c->sm_count + nfin - 1) / nfin, (max_nrec + per_cta - 1) / per_cta));
        if (c->fin_slices > 1)
            for (int gi : c->fin_gaplist) {
                c->fin_counts_off[gi] = (int)c->fin_counts_total;
                c->fin_counts_total += 5 * (size_t)c->gaps[gi].gmax;
            }
        c->fin_smem = sizeof(int) * 5 * (size_t)c->fin_gbmax + (size_t)c->fin_rows + 16;
        if (c->fin_smem > (size_t)c->max_smem - 8192) c->fused_ok = false;     // the two-call route sizes finalize by the winners
    }
    // the last iteration's M-step only matters for the estimate that wins its gap: with many units per CTA slot it pays to
    // leave it out and run the winners (one unit per gap) a second time; with few units the second run would add a whole
    // unit's latency instead
    c->two_pass = c->fused_ok && !getenv("FB_ONE_PASS") && (long)c->h_units.size() >= 16L * c->sm_count && c->h_units.size() >= 8 * ng;
    if (c->fused_ok) {
        FB_CUDA(c, c->d_bin_first.ensure(c->bins.size() + 1));
        if (c->two_pass) FB_CUDA(c, c->d_win_units.ensure(c->h_units.size()));
        FB_CUDA(c, c->d_counters.ensure(3 * c->bins.size() + 3));    // work counters of pass 1, of pass 2, winners per bin
        FB_CUDA(c, c->d_sel_off.ensure(ng + 1));
        FB_CUDA(c, c->d_sel_unit.ensure(c->h_sel_unit.size()));
        FB_CUDA(c, c->d_gmin.ensure(ng));
        FB_CUDA(c, c->d_fin_counts.ensure(c->fin_counts_total));
    }
    // ---- left-flank prefix table (FbPrefixParams): a record's chains with at least four left-flank terms in every lane
    {
        const bool want = !getenv("FB_NO_PREFIX");
        const int D = c->D;
        size_t chains = 0;
        bool fits = true;
        int max_nrec = 1;
        for (const HostGap& g : c->gaps) max_nrec = std::max(max_nrec, g.nrec);
        for (FbRecMeta& m : c->h_rmeta) {
            m.flags &= (1 << FB_PRE_SHIFT) - 1;
            if (!want || !fits) continue;
            const int L2 = m.flags & 0xff;
            const long lo = std::max<long>(m.a_lo, D - L2 + 1);
            if (m.a_hi < lo) continue;
            const long n = std::min<long>((D - lo - 3) / 32, ((long)m.a_hi - lo + 32) / 32);
            if (n <= 0) continue;
            if (chains + (size_t)n >= FB_PRE_MAX_CHAINS) { fits = false; continue; }
            m.flags |= (int)((unsigned)chains << FB_PRE_SHIFT);
            chains += (size_t)n;
        }
        if (!fits) {
            for (FbRecMeta& m : c->h_rmeta) m.flags &= (1 << FB_PRE_SHIFT) - 1;
            chains = 0;
        }
        c->pre_chains = chains;
        if (chains) FB_CUDA(c, c->d_pre.ensure(chains * 32));
        const int per_cta = 8, ngp = std::max<int>(1, (int)ng);
        c->pre_slices = std::max(1, std::min((4 * c->sm_count + ngp - 1) / ngp, (max_nrec + per_cta - 1) / per_cta));
    }
    long h2d = 0;
    const size_t stage_need = ng * sizeof(FbGapMeta) + c->h_flank.size() + (c->h_codes.size() + c->h_nmask.size()) * sizeof(uint4) +
                              nr * sizeof(FbRecMeta) + c->h_units.size() * sizeof(FbUnit) + 12 * 256 +
                              (c->fused_ok ? (c->h_sel_unit.size() + 4 * ng + 1) * sizeof(int) + ng * sizeof(long long) : 0);
    FB_CUDA(c, c->h_stage.ensure(stage_need));
    size_t stage_off = 0;
    auto up = [&](void* dst, const void* src, size_t bytes) -> cudaError_t {
        if (!bytes) return cudaSuccess;
        h2d += (long)bytes;
        unsigned char* pin = c->h_stage.p + stage_off;   // copies leave from pinned memory
        memcpy(pin, src, bytes);
        stage_off += (bytes + 255) & ~size_t(255);
        return cudaMemcpyAsync(dst, pin, bytes, cudaMemcpyHostToDevice, c->stream);
    };
    FB_CUDA(c, up(c->d_gapmeta.p, c->h_gapmeta.data(), ng * sizeof(FbGapMeta)));
    FB_CUDA(c, up(c->d_flank.p, c->h_flank.data(), c->h_flank.size()));
    FB_CUDA(c, up(c->d_codes.p, c->h_codes.data(), c->h_codes.size() * sizeof(uint4)));
    FB_CUDA(c, up(c->d_nmask.p, c->h_nmask.data(), c->h_nmask.size() * sizeof(uint4)));
    FB_CUDA(c, up(c->d_rmeta.p, c->h_rmeta.data(), nr * sizeof(FbRecMeta)));
    FB_CUDA(c, up(c->d_units.p, c->h_units.data(), c->h_units.size() * sizeof(FbUnit)));
    if (c->fused_ok) {
        FB_CUDA(c, up(c->d_sel_off.p, c->h_sel_off.data(), (ng + 1) * sizeof(int)));
        FB_CUDA(c, up(c->d_sel_unit.p, c->h_sel_unit.data(), c->h_sel_unit.size() * sizeof(int)));
        FB_CUDA(c, up(c->d_gmin.p, c->h_gmin.data(), ng * sizeof(int)));
        FB_CUDA(c, up(c->d_gaplist.p, c->fin_gaplist.data(), c->fin_gaplist.size() * sizeof(int)));
        FB_CUDA(c, up(c->d_filled_off.p, c->fin_filled_off.data(), ng * sizeof(long long)));
        FB_CUDA(c, up(c->d_fin_counts_off.p, c->fin_counts_off.data(), ng * sizeof(int)));
        std::vector<int> bin_first;
        for (const Bin& b : c->bins) bin_first.push_back(b.first_unit);
        bin_first.push_back((int)c->h_units.size());
        FB_CUDA(c, up(c->d_bin_first.p, bin_first.data(), bin_first.size() * sizeof(int)));
    }
    c->shard_counts = c->shard_maxprob = 0;
    if (c->rec_mode) {
        // the scan launches see only this rank's block of every gap; its maxima land at the block's place in the
        // all-records layout so that the ranks' buffers merge element-wise
        std::vector<FbGapMeta> gm_scan(c->h_gapmeta);
        for (size_t gi = 0; gi < ng; gi++) {
            gm_scan[gi].rec_off = c->gaps[gi].rec_off + c->gaps[gi].scan_off;
            gm_scan[gi].nrec = c->gaps[gi].scan_n;
        }
        std::vector<FbUnit> u_scan(c->h_units);
        for (FbUnit& u : u_scan) u.mp_off += c->gaps[u.gap].scan_off;
        c->shard_counts = 5 * (size_t)co;
        c->shard_maxprob = (size_t)mp;
        FB_CUDA(c, c->d_gapmeta_scan.ensure(ng));
        FB_CUDA(c, c->d_units_scan.ensure(u_scan.size()));
        FB_CUDA(c, c->d_unit_nrec.ensure(c->unit_nrec.size()));
        FB_CUDA(c, c->d_parts[0].ensure(c->shard_counts));
        FB_CUDA(c, c->d_parts[1].ensure(c->shard_counts));
        FB_CUDA(c, c->d_sumcounts.ensure(c->shard_counts));
        FB_CUDA(c, c->d_maxmerge.ensure(c->shard_maxprob));
        if (!c->d_flags.p || c->shard_poisoned) {     // allocated and zeroed once: the tokens the peers write only ever grow
            // (after a failed barrier every rank uploads again, which starts flags, error word and epoch afresh)
            FB_CUDA(c, c->d_flags.ensure(FB_MAX_SHARDS));
            FB_CUDA(c, c->d_shard_err.ensure(1));
            FB_CUDA(c, cudaMemset(c->d_flags.p, 0, sizeof(unsigned long long) * FB_MAX_SHARDS));
            FB_CUDA(c, cudaMemset(c->d_shard_err.p, 0, sizeof(int)));
            c->shard_epoch = 0;
            c->shard_poisoned = false;
        }
        FB_CUDA(c, cudaMemcpyAsync(c->d_gapmeta_scan.p, gm_scan.data(), ng * sizeof(FbGapMeta), cudaMemcpyHostToDevice, c->stream));
        FB_CUDA(c, cudaMemcpyAsync(c->d_units_scan.p, u_scan.data(), u_scan.size() * sizeof(FbUnit), cudaMemcpyHostToDevice, c->stream));
        FB_CUDA(c, cudaMemcpyAsync(c->d_unit_nrec.p, c->unit_nrec.data(), c->unit_nrec.size() * sizeof(int), cudaMemcpyHostToDevice, c->stream));
        FB_CUDA(c, cudaStreamSynchronize(c->stream));   // the vectors above are pageable and local
        h2d += (long)(ng * sizeof(FbGapMeta) + u_scan.size() * sizeof(FbUnit) + c->unit_nrec.size() * sizeof(int));
    }
    FB_CUDA(c, cudaStreamSynchronize(c->stream));
    tr.lap("upload: alloc + H2D");
    st.h2d_bytes = h2d;
    c->uploaded = true;
    c->ran = c->scanned = false;
    return FB_OK;
}

}  // namespace

extern "C" {

int fb_upload(fb_ctx* c) {
    if (!c) return FB_EINVAL;
    if (c && c->piped) return c->fail(FB_ESTATE, "the gaps of fb_fill_gap_files can only be read or run again (fb_run_device); fb_clear_gaps starts over");
    FB_CUDA(c, cudaSetDevice(c->device));
    const int pwin = c->thr_max - c->thr_min + 1;
    Trace tr;
    if (tr.on) fprintf(stderr, "[fb_trace] add_gap_files: parse + pack %.3f ms, append %.3f ms\n", c->t_parse_ms, c->t_pack_ms);
    c->t_parse_ms = c->t_pack_ms = 0;
    long long mp = 0, co = 0;
    size_t gws_need = 0;
    int rc = plan_units(c, tr, mp, co);
    if (rc == FB_OK) rc = plan_bins(c, pwin, gws_need, tr);
    if (rc == FB_OK) rc = upload_buffers(c, mp, co, gws_need, tr);
    return rc;
}

}  // extern "C"

namespace {

void fill_em_params(fb_ctx* c, FbEmParams& p) {
    memset(&p, 0, sizeof p);
    for (int k = 0; k < FB_MAX_READ; k++) {
        p.eo_tab[k] = c->h_eo[k];
        p.eo_tab[FB_MAX_READ + k] = c->h_eo[FB_MAX_READ - 1 - k];      // backwards, for reverse-complemented mates
    }
    p.ild = c->d_ild_em.p;
    memcpy(p.T, c->T, sizeof p.T);
    p.thr_min = c->thr_min;
    p.thr_max = c->thr_max;
    p.D = c->D;
    p.num_iter = c->num_iter;
    p.max_read_len = c->max_read_len;
    p.gaps = c->d_gapmeta.p;
    p.flank = c->d_flank.p;
    p.codes = c->d_codes.p;
    p.nmask = c->d_nmask.p;
    p.rmeta = c->d_rmeta.p;
    p.st_mask = c->d_st_mask.p;
    p.st_cls4 = c->d_st_cls4.p;
    p.st_jl = c->d_st_jl.p;
    p.st_cum = c->d_st_cum.p;
    p.pre = c->pre_chains ? c->d_pre.p : nullptr;
    memcpy(p.t0_tab, c->h_t0.data(), sizeof p.t0_tab);
    p.use_t0 = getenv("FB_NO_T0") ? 0 : 1;
    p.e_rec_major = getenv("FB_E_REC_MAJOR") ? 1 : 0;
    p.unit_maxprob = c->d_maxprob.p;
    p.unit_cons = c->d_cons.p;
    p.dump_unit = -1;
    p.gws = c->d_gws.p;
}

// The per-record tables of the EM kernel, rebuilt on the device at the start of every pass.
cudaError_t launch_stage(fb_ctx* c, cudaStream_t s) {
    FbStageParams sp;
    sp.codes = c->d_codes.p;
    sp.nmask = c->d_nmask.p;
    sp.rmeta = c->d_rmeta.p;
    sp.n_records = (int)c->h_rmeta.size();
    sp.st_mask = c->d_st_mask.p;
    sp.st_cls4 = c->d_st_cls4.p;
    sp.st_jl = c->d_st_jl.p;
    sp.st_cum = c->d_st_cum.p;
    cudaError_t e = fb_launch_stage(sp, s);
    if (e != cudaSuccess || !c->pre_chains) return e;
    FbPrefixParams pp;
    memcpy(pp.eo_tab, c->h_eo.data(), sizeof pp.eo_tab);
    memcpy(pp.T, c->T, sizeof pp.T);
    pp.ild = c->d_ild_em.p;
    pp.thr_min = c->thr_min;
    pp.D = c->D;
    pp.gaps = c->d_gapmeta.p;
    pp.n_gaps = (int)c->h_gapmeta.size();
    pp.flank = c->d_flank.p;
    pp.rmeta = c->d_rmeta.p;
    pp.st_cls4 = c->d_st_cls4.p;
    pp.pre = c->d_pre.p;
    return fb_launch_prefix(pp, c->pre_slices, c->f1, s);
}

cudaError_t launch_bin(const fb_ctx* c, const Bin& b, const FbEmParams& p, cudaStream_t s) {
    if (b.cluster > 1) return fb_launch_em_cluster(p, b.grid, b.cluster, b.threads, c->f1, b.smem, s);
    return fb_launch_em(p, b.grid, b.threads, b.global_tables, c->f1, b.smem, s);
}

void fill_bin_params(fb_ctx* c, size_t bi, FbEmParams& p) {
    const Bin& b = c->bins[bi];
    fill_em_params(c, p);
    p.units = c->d_units.p + b.first_unit;
    p.n_units = b.n_units;
    p.unit_lik = c->d_unit_lik.p + b.first_unit;
    p.work_counter = c->d_counters.p + bi;
    p.nr_stride = b.nr_stride;
    p.gp_stride = b.gp_stride;
    p.p_stride = b.p_stride;
    p.chunk = b.chunk;
    p.gws_stride = gws_stride_of(b.nr_stride, b.gp_stride);
    p.gws = c->d_gws.p + b.gws_off;
}

}  // namespace

extern "C" {

// EM scan of the units this context owns (figbird.cpp:2468-2478): one persistent launch per size bin; every unit ends
// with its likelihood (figbird.cpp:1951-1955) computed on the device — glibc's log restated bit for bit, the sum taken in
// record order (fb_glibc_math.h) — so 8 bytes per unit come home instead of 8 bytes per (unit, record).
// FB_HOST_LOG=1 keeps the round-1 route for comparison: per-record maxima to the host, log() and the sum by the host libm.
int fb_run_scan(fb_ctx* c) {
    if (!c) return FB_EINVAL;
    if (c && c->piped) return c->fail(FB_ESTATE, "the gaps of fb_fill_gap_files can only be read or run again (fb_run_device); fb_clear_gaps starts over");
    if (!c->uploaded) return c->fail(FB_ESTATE, "fb_run_scan before fb_upload");
    if (c->rec_mode && c->part_world > 1) return c->fail(FB_ESTATE, "record-sharded partition: use fb_shard_scan_iter / fb_shard_scan_finish");
    FB_CUDA(c, cudaSetDevice(c->device));
    fb_stats& st = c->stats;
    long d2h = 0;
    Trace tr;
    st.em_launches = st.finalize_launches = 0;
    st.finalize_likelihoods = 0;
    st.finalize_kernel_ms = 0;
    const bool host_log = c->host_log || c->rec_mode;

    FB_CUDA(c, cudaEventRecord(c->evs, c->stream));
    FB_CUDA(c, cudaMemsetAsync(c->d_counters.p, 0, sizeof(int) * (c->bins.size() + 1), c->stream));
    FB_CUDA(c, launch_stage(c, c->stream));
    st.aux_launches = 1 + (c->pre_chains ? 1 : 0);
    FB_CUDA(c, cudaEventRecord(c->ev0, c->stream));
    for (size_t bi = 0; bi < c->bins.size(); bi++) {
        const Bin& b = c->bins[bi];
        FbEmParams p;
        fill_bin_params(c, bi, p);
        if (host_log) p.unit_lik = nullptr;
        // every bin runs on its own stream: as the persistent CTAs of one bin drain, CTAs of the next bin take over the
        // freed SMs, so only the last (smallest) bin has a tail
        cudaStream_t bs = c->bin_streams[bi];
        FB_CUDA(c, cudaStreamWaitEvent(bs, c->ev0, 0));
        FB_CUDA(c, launch_bin(c, b, p, bs));
        st.em_launches++;
        // this bin's results go home on the copy stream while the other bins compute
        FB_CUDA(c, cudaEventRecord(c->bin_done[bi], bs));
        FB_CUDA(c, cudaStreamWaitEvent(c->stream, c->bin_done[bi], 0));
        FB_CUDA(c, cudaStreamWaitEvent(c->copy_stream, c->bin_done[bi], 0));
        if (host_log) {
            const FbUnit& u0 = c->h_units[b.first_unit];
            const FbUnit& u1 = c->h_units[b.first_unit + b.n_units - 1];
            const size_t lo = (size_t)u0.mp_off, hi = (size_t)u1.mp_off + c->unit_nrec[b.first_unit + b.n_units - 1];
            if (hi > lo) {
                FB_CUDA(c, cudaMemcpyAsync(c->h_maxprob.p + lo, c->d_maxprob.p + lo, (hi - lo) * sizeof(double), cudaMemcpyDeviceToHost, c->copy_stream));
                d2h += (long)((hi - lo) * sizeof(double));
            }
        } else {
            FB_CUDA(c, cudaMemcpyAsync(c->h_unit_lik.p + b.first_unit, c->d_unit_lik.p + b.first_unit, (size_t)b.n_units * sizeof(double),
                                       cudaMemcpyDeviceToHost, c->copy_stream));
            d2h += (long)((size_t)b.n_units * sizeof(double));
        }
        FB_CUDA(c, cudaEventRecord(c->bin_copied[bi], c->copy_stream));
    }
    FB_CUDA(c, cudaEventRecord(c->ev1, c->stream));
    tr.lap("run: launches enqueued");

    // Estimates scanned by another rank stay NaN until fb_merge_liks.
    for (HostGap& g : c->gaps) {
        g.done = false;
        if (!g.owned) continue;
        g.liks.assign(g.gmax - g.gmin + 1, std::nan(""));
        g.cons_off.assign(g.gmax - g.gmin + 1, -1);
    }
    for (size_t bi = 0; bi < c->bins.size(); bi++) {
        const Bin& b = c->bins[bi];
        FB_CUDA(c, cudaEventSynchronize(c->bin_copied[bi]));
        // units are independent (each writes its own slot of its gap's arrays)
        const int per = 512, nblock = (b.n_units + per - 1) / per;
        parallel_for(nblock, host_log ? 0 : 4, [&](int blk) {
            const int k1 = std::min(b.n_units, (blk + 1) * per);
            for (int k = blk * per; k < k1; k++) {
                const FbUnit& u = c->h_units[b.first_unit + k];
                HostGap& g = c->gaps[u.gap];
                double lik = 0;
                if (host_log) {
                    const double* mp = c->h_maxprob.p + u.mp_off;
                    for (int r = 0; r < g.nrec; r++) {
                        if (c->f1) lik += mp[r];                       // figbird.cpp:1951 (f=1: unconditional)
                        else if (mp[r] > 0) lik += log(mp[r]);         // figbird.cpp:1953-1955
                    }
                } else {
                    lik = c->h_unit_lik.p[b.first_unit + k];
                }
                g.liks[u.gp - g.gmin] = lik;
                g.cons_off[u.gp - g.gmin] = u.cons_off;
            }
        });
    }
    FB_CUDA(c, cudaStreamSynchronize(c->stream));
    tr.lap(host_log ? "run: EM + host log sums" : "run: EM");
    float ms = 0;
    cudaEventElapsedTime(&ms, c->ev0, c->ev1);
    st.em_kernel_ms = ms;
    st.d2h_bytes = d2h;
    c->scanned = true;
    c->ran = false;
    return FB_OK;
}

// Best estimate per gap (figbird.cpp:2479-2490) and finalize (figbird.cpp:2087-2393) for every gap whose winning
// estimate was scanned by this context.  The whole of finalize runs on the device (fb_finalize_kernel); the filled
// sequences and one discard count per gap come home.
int fb_run_finalize(fb_ctx* c) {
    if (!c) return FB_EINVAL;
    if (c && c->piped) return c->fail(FB_ESTATE, "the gaps of fb_fill_gap_files can only be read or run again (fb_run_device); fb_clear_gaps starts over");
    if (!c->scanned) return c->fail(FB_ESTATE, "fb_run_finalize before fb_run_scan");
    FB_CUDA(c, cudaSetDevice(c->device));
    fb_stats& st = c->stats;
    const int D = c->D;
    long d2h = st.d2h_bytes;
    Trace tr;
    const size_t ng = c->gaps.size();
    std::vector<long long> cons_off_of_best(ng, 0), filled_off(ng, 0);
    std::vector<int> gaplist, bestgp(ng, 0), counts_off(ng, 0);
    for (size_t gi = 0; gi < ng; gi++) {
        HostGap& g = c->gaps[gi];
        if (!g.owned) continue;
        double best = -DBL_MAX;
        int be = g.gmin;
        bool any = false;
        for (int gp = g.gmin; gp <= g.gmax; gp++) {
            const double l = g.liks[gp - g.gmin];
            if (std::isnan(l) && g.cons_off[gp - g.gmin] < 0)
                return c->fail(FB_ESTATE, "gap %zu: estimate %d was scanned by another rank and its likelihood has not been merged", gi, gp);
            if (l >= best) {   // last estimate wins ties; a NaN or -inf likelihood (f=1 only) never wins, as in the reference
                best = l;
                be = gp;
                any = true;
            }
        }
        g.best_lik = best;
        g.best_est = be;
        // f=1 can leave every estimate at -inf (two records without a placement in the window already do): then no
        // estimate ever passes the '>=' of figbird.cpp:2479, maxGapEstimate stays gapMin and bestString stays empty —
        // finalize scores against flank + gapMin non-bases + flank
        g.no_winner = !any;
        bestgp[gi] = be;
        if (g.cons_off[be - g.gmin] < 0) continue;   // the winner lives on another rank: that rank finalizes this gap
        if (c->rec_mode && (int)(gi % c->part_world) != c->part_rank) continue;   // every rank could; the gaps are dealt round-robin
        cons_off_of_best[gi] = any ? g.cons_off[be - g.gmin] : -1;
        gaplist.push_back((int)gi);
    }

    size_t fill_total = 0, counts_total = 0;
    if (!gaplist.empty()) {
        int rows = 0, gbmax = 1, max_nrec = 1;
        for (int gi : gaplist) {
            const HostGap& g = c->gaps[gi];
            rows = std::max(rows, g.best_est + 2 * g.lmax);
            gbmax = std::max(gbmax, g.best_est);
            max_nrec = std::max(max_nrec, g.nrec);
            filled_off[gi] = (long long)fill_total;
            fill_total += (size_t)g.best_est;
        }
        // records are independent in finalize: enough CTAs per gap to fill the GPU, at least one record per warp
        const int per_cta = FB_FIN_THREADS / 32;
        const int slices = std::max(1, std::min((4 * c->sm_count + (int)gaplist.size() - 1) / (int)gaplist.size(), (max_nrec + per_cta - 1) / per_cta));
        if (slices > 1)
            for (int gi : gaplist) {
                counts_off[gi] = (int)counts_total;
                counts_total += 5 * (size_t)c->gaps[gi].best_est;
            }
        const size_t smem = sizeof(int) * 5 * (size_t)gbmax + (size_t)rows + 16;
        if (smem > (size_t)c->max_smem - 8192) return c->fail(FB_EINVAL, "gap estimate %d too long for finalize", gbmax);
        FB_CUDA(c, c->d_fin_counts.ensure(counts_total));
        unsigned char* pin = c->h_stage.p;        // the small per-pass arrays leave from pinned memory (fb_upload sized it for far more)
        auto up = [&](void* dst, const void* src, size_t bytes) {
            memcpy(pin, src, bytes);
            cudaError_t e = cudaMemcpyAsync(dst, pin, bytes, cudaMemcpyHostToDevice, c->stream);
            pin += (bytes + 255) & ~size_t(255);
            return e;
        };
        c->fused_dirty = true;
        FB_CUDA(c, up(c->d_gaplist.p, gaplist.data(), gaplist.size() * sizeof(int)));
        FB_CUDA(c, up(c->go_gp(c->d_gapout.p), bestgp.data(), ng * sizeof(int)));
        FB_CUDA(c, up(c->go_off(c->d_gapout.p), cons_off_of_best.data(), ng * sizeof(long long)));
        FB_CUDA(c, up(c->d_filled_off.p, filled_off.data(), ng * sizeof(long long)));
        FB_CUDA(c, up(c->d_fin_counts_off.p, counts_off.data(), ng * sizeof(int)));
        FB_CUDA(c, cudaMemsetAsync(c->go_disc(c->d_gapout.p), 0, sizeof(int) * 2 * c->gapout_n, c->stream));   // discarded + tickets
        if (counts_total) FB_CUDA(c, cudaMemsetAsync(c->d_fin_counts.p, 0, sizeof(int) * counts_total, c->stream));
        FbFinParams f;
        memset(&f, 0, sizeof f);
        f.ild = c->d_ild.p;
        f.epos = c->d_epos.p;
        f.match = c->d_match.p;
        memcpy(f.T, c->T, sizeof f.T);
        f.thr_min = c->thr_min;
        f.thr_max = c->thr_max;
        f.D = D;
        f.gaps = c->d_gapmeta.p;
        f.flank = c->d_flank.p;
        f.codes = c->d_codes.p;
        f.nmask = c->d_nmask.p;
        f.rmeta = c->d_rmeta.p;
        f.gap_list = c->d_gaplist.p;
        f.best_gp = c->go_gp(c->d_gapout.p);
        f.best_cons_off = c->go_off(c->d_gapout.p);
        f.unit_cons = c->d_cons.p;
        f.gmax_stride = rows;
        f.gb_stride = gbmax;
        f.theta = c->theta;
        f.fin_counts = c->d_fin_counts.p;
        f.fin_counts_off = c->d_fin_counts_off.p;
        f.tickets = c->go_tick(c->d_gapout.p);
        f.discarded = c->go_disc(c->d_gapout.p);
        f.filled = c->d_filled.p;
        f.filled_off = c->d_filled_off.p;
        FB_CUDA(c, cudaEventRecord(c->ev2, c->stream));
        FB_CUDA(c, fb_launch_finalize(f, (int)gaplist.size(), slices, smem, c->stream));
        FB_CUDA(c, cudaEventRecord(c->ev3, c->stream));
        st.finalize_launches = 1;
        FB_CUDA(c, cudaMemcpyAsync(c->h_filled.p, c->d_filled.p, fill_total, cudaMemcpyDeviceToHost, c->stream));
        FB_CUDA(c, cudaMemcpyAsync(c->go_disc(c->h_gapout.p), c->go_disc(c->d_gapout.p), ng * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        d2h += (long)(fill_total + ng * sizeof(int));
        FB_CUDA(c, cudaEventRecord(c->eve, c->stream));
        FB_CUDA(c, cudaEventSynchronize(c->eve));
    } else {
        FB_CUDA(c, cudaEventRecord(c->eve, c->stream));
        FB_CUDA(c, cudaEventSynchronize(c->eve));
    }
    tr.lap("run: select + finalize kernel + results");
    // ---- hand the results to the gaps (figbird.cpp:2382, 2050-2063)
    for (int gi : gaplist) {
        HostGap& g = c->gaps[gi];
        g.total_reads = g.nrec;
        g.discarded = c->go_disc(c->h_gapout.p)[gi];
        g.filled.assign(c->h_filled.p + filled_off[gi], (size_t)g.best_est);
        g.done = true;
        for (int r = 0; r < g.nrec; r++) st.finalize_likelihoods += placements(c->h_rmeta[g.rec_off + r], D, g.best_est);
    }
    tr.lap("run: results");
    float ms = 0;
    cudaEventElapsedTime(&ms, c->evs, c->eve);
    st.step_ms = ms;
    st.finalize_kernel_ms = 0;
    if (!gaplist.empty()) {
        cudaEventElapsedTime(&ms, c->ev2, c->ev3);
        st.finalize_kernel_ms = ms;
    }
    st.d2h_bytes = d2h;
    c->ran = true;
    return FB_OK;
}

// ---- record-sharded scan (see include/figbird_cuda.h) ------------------------------------------------------------

int fb_shard_sizes(fb_ctx* c, size_t* counts_doubles, size_t* maxprob_doubles) {
    if (!c) return FB_EINVAL;
    if (!c->uploaded || !c->rec_mode) return c->fail(FB_ESTATE, "fb_shard_sizes needs fb_set_partition_records + fb_upload");
    if (counts_doubles) *counts_doubles = c->shard_counts;
    if (maxprob_doubles) *maxprob_doubles = c->shard_maxprob;
    return FB_OK;
}

int fb_shard_buffers(fb_ctx* c, double** counts_slot0, double** counts_slot1, double** maxprob, unsigned long long** flags) {
    if (!c) return FB_EINVAL;
    if (!c->uploaded || !c->rec_mode) return c->fail(FB_ESTATE, "fb_shard_buffers needs fb_set_partition_records + fb_upload");
    if (counts_slot0) *counts_slot0 = c->d_parts[0].p;
    if (counts_slot1) *counts_slot1 = c->d_parts[1].p;
    if (maxprob) *maxprob = c->d_maxprob.p;
    if (flags) *flags = c->d_flags.p;
    return FB_OK;
}

}  // extern "C"

namespace {

// Enqueue one EM iteration of a record-sharded scan (all bins); the caller decides when to synchronise.
int shard_enqueue_iter(fb_ctx* c, int it, const double* const* peer_counts, double* out_counts, bool first_of_pass) {
    fb_stats& st = c->stats;
    if (first_of_pass) {
        st.em_launches = st.finalize_launches = 0;
        st.finalize_likelihoods = 0;
        st.finalize_kernel_ms = st.em_kernel_ms = 0;
        st.d2h_bytes = 0;
        c->scanned = c->ran = false;
        // records scanned elsewhere keep 0 in this rank's maxima, so the ranks' buffers merge with max
        FB_CUDA(c, cudaMemsetAsync(c->d_maxprob.p, 0, sizeof(double) * c->shard_maxprob, c->stream));
        FB_CUDA(c, launch_stage(c, c->stream));
        st.aux_launches = 1 + (c->pre_chains ? 1 : 0);
    }
    FB_CUDA(c, cudaMemsetAsync(c->d_counters.p, 0, sizeof(int) * (c->bins.size() + 1), c->stream));
    FB_CUDA(c, cudaEventRecord(c->ev0, c->stream));
    for (size_t bi = 0; bi < c->bins.size(); bi++) {
        const Bin& b = c->bins[bi];
        FbEmParams p;
        fill_bin_params(c, bi, p);
        p.units = c->d_units_scan.p + b.first_unit;
        p.gaps = c->d_gapmeta_scan.p;
        p.num_iter = 1;                       // one iteration per launch: the partial counts are exchanged in between
        p.n_parts = it == 0 ? 0 : c->part_world;
        for (int q = 0; q < p.n_parts; q++) {
            if (!peer_counts[q]) return c->fail(FB_EINVAL, "peer_counts[%d] is null", q);
            p.parts_in[q] = peer_counts[q];
        }
        p.part_out = out_counts;
        p.abort = c->d_shard_err.p;
        p.unit_lik = nullptr;                 // the likelihoods come from the merged maxima of all ranks (shard_host_liks)
        cudaStream_t bs = c->bin_streams[bi];
        FB_CUDA(c, cudaStreamWaitEvent(bs, c->ev0, 0));
        FB_CUDA(c, launch_bin(c, b, p, bs));
        st.em_launches++;
        FB_CUDA(c, cudaEventRecord(c->bin_done[bi], bs));
        FB_CUDA(c, cudaStreamWaitEvent(c->stream, c->bin_done[bi], 0));
    }
    return FB_OK;
}

// The closing reduction (enqueue + wait) and the host likelihoods.
int shard_finish(fb_ctx* c, const double* const* peer_counts, const double* const* peer_maxprob) {
    FbShardFinParams f;
    memset(&f, 0, sizeof f);
    f.units = c->d_units.p;
    f.n_units = (int)c->h_units.size();
    f.unit_nrec = c->d_unit_nrec.p;
    f.n_parts = c->part_world;
    for (int q = 0; q < f.n_parts; q++) {
        if (!peer_counts[q] || !peer_maxprob[q]) return c->fail(FB_EINVAL, "peer buffer %d is null", q);
        f.parts_in[q] = peer_counts[q];
        f.maxprob_in[q] = peer_maxprob[q];
    }
    f.sum_counts = c->d_sumcounts.p;
    f.maxprob_out = c->d_maxmerge.p;
    f.unit_cons = c->d_cons.p;
    f.abort = c->d_shard_err.p;
    FB_CUDA(c, fb_launch_shard_finish(f, c->stream));
    c->stats.aux_launches++;
    FB_CUDA(c, cudaMemcpyAsync(c->h_maxprob.p, c->d_maxmerge.p, c->shard_maxprob * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    return FB_OK;
}

void shard_host_liks(fb_ctx* c) {
    c->stats.d2h_bytes = (long)(c->shard_maxprob * sizeof(double));
    // likelihood per unit over ALL records of the gap, in record order (figbird.cpp:1951-1955): the same on every rank
    for (HostGap& g : c->gaps) {
        g.done = false;
        g.liks.assign(g.gmax - g.gmin + 1, std::nan(""));
        g.cons_off.assign(g.gmax - g.gmin + 1, -1);
    }
    for (const FbUnit& u : c->h_units) {
        HostGap& g = c->gaps[u.gap];
        const double* mp = c->h_maxprob.p + u.mp_off;
        double lik = 0;
        for (int r = 0; r < g.nrec; r++)
            if (mp[r] > 0) lik += log(mp[r]);
        g.liks[u.gp - g.gmin] = lik;
        g.cons_off[u.gp - g.gmin] = u.cons_off;
    }
    c->scanned = true;
    c->ran = false;
}

}  // namespace

extern "C" {

int fb_shard_scan_iter(fb_ctx* c, int it, const double* const* peer_counts, double* out_counts) {
    if (!c || !out_counts) return FB_EINVAL;
    if (!c->uploaded || !c->rec_mode) return c->fail(FB_ESTATE, "fb_shard_scan_iter needs fb_set_partition_records + fb_upload");
    if (it < 0 || it >= c->num_iter || (it > 0 && !peer_counts)) return c->fail(FB_EINVAL, "iteration %d of %d", it, c->num_iter);
    FB_CUDA(c, cudaSetDevice(c->device));
    if (it == 0) FB_CUDA(c, cudaEventRecord(c->evs, c->stream));
    FB_CUDA(c, cudaEventRecord(c->ev2, c->stream));
    int rc = shard_enqueue_iter(c, it, peer_counts, out_counts, it == 0);
    if (rc != FB_OK) return rc;
    FB_CUDA(c, cudaEventRecord(c->ev1, c->stream));
    FB_CUDA(c, cudaStreamSynchronize(c->stream));
    float ms = 0;
    cudaEventElapsedTime(&ms, c->ev2, c->ev1);
    c->stats.em_kernel_ms += ms;
    return FB_OK;
}

int fb_shard_scan_finish(fb_ctx* c, const double* const* peer_counts, const double* const* peer_maxprob) {
    if (!c || !peer_counts || !peer_maxprob) return FB_EINVAL;
    if (!c->uploaded || !c->rec_mode) return c->fail(FB_ESTATE, "fb_shard_scan_finish needs fb_set_partition_records + fb_upload");
    FB_CUDA(c, cudaSetDevice(c->device));
    int rc = shard_finish(c, peer_counts, peer_maxprob);
    if (rc != FB_OK) return rc;
    FB_CUDA(c, cudaStreamSynchronize(c->stream));
    shard_host_liks(c);
    return FB_OK;
}

int fb_shard_scan_all(fb_ctx* c, const double* const* peer_counts0, const double* const* peer_counts1, const double* const* peer_maxprob,
                      unsigned long long* const* peer_flags) {
    if (!c || !peer_counts0 || !peer_counts1 || !peer_maxprob || !peer_flags) return FB_EINVAL;
    if (!c->uploaded || !c->rec_mode) return c->fail(FB_ESTATE, "fb_shard_scan_all needs fb_set_partition_records + fb_upload");
    FB_CUDA(c, cudaSetDevice(c->device));
    const int world = c->part_world, rank = c->part_rank, niter = c->num_iter;
    if (peer_counts0[rank] != c->d_parts[0].p || peer_counts1[rank] != c->d_parts[1].p || peer_flags[rank] != c->d_flags.p)
        return c->fail(FB_EINVAL, "entry %d of the peer arrays must be this context's own buffer", rank);
    FbFlagParams fp;
    memset(&fp, 0, sizeof fp);
    for (int q = 0; q < world; q++) {
        if (!peer_flags[q]) return c->fail(FB_EINVAL, "peer_flags[%d] is null", q);
        fp.peer_flags[q] = peer_flags[q];
    }
    fp.own_flags = c->d_flags.p;
    fp.rank = rank;
    fp.world = world;
    fp.error = c->d_shard_err.p;
    {
        // a wait gives up after FB_BARRIER_TIMEOUT_MS of SM clock (default 20 s: first-call module loads and an
        // imbalanced rank stay far below it)
        const char* tm = getenv("FB_BARRIER_TIMEOUT_MS");
        const double ms = tm ? std::max(1.0, atof(tm)) : 20000.0;
        int khz = 1965000;
        cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, c->device);
        fp.timeout_cycles = (long long)(ms * (double)khz);
    }
    if (c->shard_poisoned) return c->fail(FB_ESTATE, "a barrier of an earlier record-sharded scan failed: upload again on every rank");
    const unsigned long long stride = (unsigned long long)niter + 2;
    const unsigned long long epoch = ++c->shard_epoch, base = epoch * stride;
    auto wait_for = [&](unsigned long long token) {
        fp.token = token;
        return fb_launch_flag_wait(fp, c->stream);
    };
    auto signal = [&](unsigned long long token) {
        fp.token = token;
        return fb_launch_flag_signal(fp, c->stream);
    };
    FB_CUDA(c, cudaEventRecord(c->evs, c->stream));
    // nobody may still be reading what this pass is about to overwrite
    if (epoch > 1) FB_CUDA(c, wait_for((epoch - 1) * stride + niter + 1));
    const double* const* slots[2] = {peer_counts0, peer_counts1};
    double* own[2] = {c->d_parts[0].p, c->d_parts[1].p};
    for (int it = 0; it < niter; it++) {
        if (it > 0) FB_CUDA(c, wait_for(base + it));          // every rank has written its partial counts of iteration it-1
        int rc = shard_enqueue_iter(c, it, it ? slots[(it - 1) & 1] : nullptr, own[it & 1], it == 0);
        if (rc != FB_OK) return rc;
        FB_CUDA(c, signal(base + it + 1));
    }
    FB_CUDA(c, wait_for(base + niter));
    int rc = shard_finish(c, slots[(niter - 1) & 1], peer_maxprob);
    if (rc != FB_OK) return rc;
    FB_CUDA(c, signal(base + niter + 1));                      // this rank is done reading the peers' buffers
    c->stats.aux_launches += 2 * niter + 2;
    int err = 0;
    FB_CUDA(c, cudaMemcpyAsync(&err, c->d_shard_err.p, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    FB_CUDA(c, cudaEventRecord(c->ev1, c->stream));
    FB_CUDA(c, cudaStreamSynchronize(c->stream));
    if (err) {
        // the flags of this pass are poisoned on every rank that waited on this one (and the other way round): the
        // results of the pass are void everywhere; the next fb_upload starts the tokens afresh
        c->shard_poisoned = true;
        return c->fail(FB_ECUDA, "record-sharded scan: a rank did not reach a barrier in time (the pass is void on every rank)");
    }
    float ms = 0;
    cudaEventElapsedTime(&ms, c->evs, c->ev1);
    c->stats.em_kernel_ms = ms;            // the whole device-side sequence, barriers included
    shard_host_liks(c);
    return FB_OK;
}

int fb_shard_reset(fb_ctx* c) {
    if (!c) return FB_EINVAL;
    if (!c->d_flags.p) return FB_OK;
    FB_CUDA(c, cudaSetDevice(c->device));
    FB_CUDA(c, cudaDeviceSynchronize());
    FB_CUDA(c, cudaMemset(c->d_flags.p, 0, sizeof(unsigned long long) * FB_MAX_SHARDS));
    FB_CUDA(c, cudaMemset(c->d_shard_err.p, 0, sizeof(int)));
    c->shard_epoch = 0;
    c->shard_poisoned = false;
    return FB_OK;
}

int fb_shard_counts(fb_ctx* c, int gap, int estimate, double* counts) {
    if (!c || !counts || gap < 0 || gap >= (int)c->gaps.size()) return FB_EINVAL;
    if (!c->scanned || !c->rec_mode) return c->fail(FB_ESTATE, "fb_shard_counts needs a finished record-sharded scan");
    FB_CUDA(c, cudaSetDevice(c->device));
    for (const FbUnit& u : c->h_units) {
        if (u.gap != gap || u.gp != estimate) continue;
        std::vector<double> cm((size_t)5 * u.gp);
        FB_CUDA(c, cudaMemcpy(cm.data(), c->d_sumcounts.p + 5 * (size_t)u.cons_off, cm.size() * sizeof(double), cudaMemcpyDeviceToHost));
        for (int x = 0; x < u.gp; x++)
            for (int k = 0; k < 5; k++) counts[(size_t)x * 5 + k] = cm[(size_t)k * u.gp + x];
        return FB_OK;
    }
    return c->fail(FB_EINVAL, "estimate %d is not in the scan range of gap %d", estimate, gap);
}

int fb_ipc_export(fb_ctx* c, const void* dev_ptr, unsigned char handle[64]) {
    if (!c || !dev_ptr || !handle) return FB_EINVAL;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
    FB_CUDA(c, cudaSetDevice(c->device));
    cudaIpcMemHandle_t h;
    FB_CUDA(c, cudaIpcGetMemHandle(&h, const_cast<void*>(dev_ptr)));
    memcpy(handle, &h, 64);
    return FB_OK;
}

int fb_ipc_open(fb_ctx* c, const unsigned char handle[64], void** dev_ptr) {
    if (!c || !handle || !dev_ptr) return FB_EINVAL;
    FB_CUDA(c, cudaSetDevice(c->device));
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    FB_CUDA(c, cudaIpcOpenMemHandle(dev_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return FB_OK;
}

int fb_ipc_close(fb_ctx* c, void* dev_ptr) {
    if (!c || !dev_ptr) return FB_EINVAL;
    FB_CUDA(c, cudaSetDevice(c->device));
    FB_CUDA(c, cudaIpcCloseMemHandle(dev_ptr));
    return FB_OK;
}

// fillGap for every gap of this context in ONE enqueue: record tables, the EM scan of every size bin, the selection among
// the estimates, finalize, and the copies home; the host waits once and then hands the results to the gaps.
// run_fused_begin enqueues and returns; run_fused_end waits and collects (fb_fill_gap_files runs host work in between).
static int run_fused_begin(fb_ctx* c) {
    FB_CUDA(c, cudaSetDevice(c->device));
    fb_stats& st = c->stats;
    Trace tr;
    const size_t ng = c->gaps.size(), nu = c->h_units.size();
    st.em_launches = st.finalize_launches = 0;
    st.finalize_likelihoods = 0;
    if (c->fused_dirty) {
        FB_CUDA(c, cudaMemcpyAsync(c->d_gaplist.p, c->fin_gaplist.data(), c->fin_gaplist.size() * sizeof(int), cudaMemcpyHostToDevice, c->stream));
        FB_CUDA(c, cudaMemcpyAsync(c->d_filled_off.p, c->fin_filled_off.data(), ng * sizeof(long long), cudaMemcpyHostToDevice, c->stream));
        FB_CUDA(c, cudaMemcpyAsync(c->d_fin_counts_off.p, c->fin_counts_off.data(), ng * sizeof(int), cudaMemcpyHostToDevice, c->stream));
        FB_CUDA(c, c->d_fin_counts.ensure(c->fin_counts_total));
        FB_CUDA(c, cudaStreamSynchronize(c->stream));
        c->fused_dirty = false;
    }
    const size_t nb = c->bins.size();
    const bool fin = !c->fin_gaplist.empty();
    st.aux_launches = 1 + (c->pre_chains ? 1 : 0) + (fin ? 1 : 0);
    st.em_launches = (int)nb * ((c->two_pass && fin) ? 2 : 1);
    st.finalize_launches = fin ? 1 : 0;
    // The whole pass as a stream of operations.  `capturing`: the timing events become event-record nodes of a CUDA graph.
    auto enqueue = [&](bool capturing) -> int {
        auto rec = [&](cudaEvent_t ev) {
            return capturing ? cudaEventRecordWithFlags(ev, c->stream, cudaEventRecordExternal) : cudaEventRecord(ev, c->stream);
        };
        FB_CUDA(c, rec(c->evs));
        FB_CUDA(c, cudaMemsetAsync(c->d_counters.p, 0, sizeof(int) * (3 * nb + 3), c->stream));
        FB_CUDA(c, cudaMemsetAsync(c->go_disc(c->d_gapout.p), 0, sizeof(int) * 2 * c->gapout_n, c->stream));   // discarded + tickets
        if (c->fin_counts_total) FB_CUDA(c, cudaMemsetAsync(c->d_fin_counts.p, 0, sizeof(int) * c->fin_counts_total, c->stream));
        FB_CUDA(c, launch_stage(c, c->stream));
        FB_CUDA(c, rec(c->ev0));
        // pass: 0 the scan of every unit (without the last M-step when a second pass follows), 1 the winners again in full
        // FB_SERIAL_BINS=1 (profiling aid): every bin on the main stream, one after the other — ncu collects its metrics
        // for every launch that way (launches that overlap on side streams come back without counters)
        const bool serial = nb == 1 || (getenv("FB_SERIAL_BINS") && atoi(getenv("FB_SERIAL_BINS")));
        auto scan_pass = [&](int pass) -> int {
            if (!serial) FB_CUDA(c, cudaEventRecord(c->ev_fork, c->stream));      // every bin's stream forks from here ...
            for (size_t bi = 0; bi < nb; bi++) {
                FbEmParams p;
                fill_bin_params(c, bi, p);
                if (pass == 0) {
                    p.skip_last_m = c->two_pass ? 1 : 0;
                } else {
                    p.units = c->d_win_units.p + c->bins[bi].first_unit;
                    p.n_units_ptr = c->d_counters.p + 2 * nb + 2 + bi;
                    p.work_counter = c->d_counters.p + nb + 1 + bi;
                    p.unit_lik = nullptr;                        // pass 1 has written them, by the units' places in the full list
                }
                if (serial) {                                    // a single launch needs no stream of its own
                    FB_CUDA(c, launch_bin(c, c->bins[bi], p, c->stream));
                } else {
                    cudaStream_t bs = c->bin_streams[bi];
                    FB_CUDA(c, cudaStreamWaitEvent(bs, c->ev_fork, 0));
                    FB_CUDA(c, launch_bin(c, c->bins[bi], p, bs));
                    FB_CUDA(c, cudaEventRecord(c->bin_done[bi], bs));
                }
            }
            if (!serial)                                                         // ... and joins the main stream here
                for (size_t bi = 0; bi < nb; bi++) FB_CUDA(c, cudaStreamWaitEvent(c->stream, c->bin_done[bi], 0));
            return FB_OK;
        };
        int rc = scan_pass(0);
        if (rc != FB_OK) return rc;
        FbSelParams sp;
        memset(&sp, 0, sizeof sp);
        if (!c->fin_gaplist.empty()) {
            sp.n_gaps = (int)ng;
            sp.sel_off = c->d_sel_off.p;
            sp.sel_unit = c->d_sel_unit.p;
            sp.gmin = c->d_gmin.p;
            sp.unit_lik = c->d_unit_lik.p;
            sp.units = c->d_units.p;
            sp.best_gp = c->go_gp(c->d_gapout.p);
            sp.best_cons_off = c->go_off(c->d_gapout.p);
            sp.best_lik = c->go_lik(c->d_gapout.p);
            if (c->two_pass) {
                sp.win_units = c->d_win_units.p;
                sp.win_count = c->d_counters.p + 2 * nb + 2;
                sp.bin_first = c->d_bin_first.p;
                sp.n_bins = (int)nb;
            }
            FB_CUDA(c, fb_launch_select(sp, c->stream));
            if (c->two_pass) {
                rc = scan_pass(1);
                if (rc != FB_OK) return rc;
            }
        }
        FB_CUDA(c, rec(c->ev1));
        if (!c->fin_gaplist.empty()) {
            FbFinParams f;
            memset(&f, 0, sizeof f);
            f.ild = c->d_ild.p;
            f.epos = c->d_epos.p;
            f.match = c->d_match.p;
            memcpy(f.T, c->T, sizeof f.T);
            f.thr_min = c->thr_min;
            f.thr_max = c->thr_max;
            f.D = c->D;
            f.gaps = c->d_gapmeta.p;
            f.flank = c->d_flank.p;
            f.codes = c->d_codes.p;
            f.nmask = c->d_nmask.p;
            f.rmeta = c->d_rmeta.p;
            f.gap_list = c->d_gaplist.p;
            f.best_gp = c->go_gp(c->d_gapout.p);
            f.best_cons_off = c->go_off(c->d_gapout.p);
            f.unit_cons = c->d_cons.p;
            f.gmax_stride = c->fin_rows;
            f.gb_stride = c->fin_gbmax;
            f.theta = c->theta;
            f.fin_counts = c->d_fin_counts.p;
            f.fin_counts_off = c->d_fin_counts_off.p;
            f.tickets = c->go_tick(c->d_gapout.p);
            f.discarded = c->go_disc(c->d_gapout.p);
            f.filled = c->d_filled.p;
            f.filled_off = c->d_filled_off.p;
            FB_CUDA(c, rec(c->ev2));
            FB_CUDA(c, fb_launch_finalize(f, (int)c->fin_gaplist.size(), c->fin_slices, c->fin_smem, c->stream));
            FB_CUDA(c, rec(c->ev3));
            FB_CUDA(c, cudaMemcpyAsync(c->h_filled.p, c->d_filled.p, c->fin_fill_total, cudaMemcpyDeviceToHost, c->stream));
            FB_CUDA(c, cudaMemcpyAsync(c->h_gapout.p, c->d_gapout.p, 24 * c->gapout_n, cudaMemcpyDeviceToHost, c->stream));   // all but the tickets
        }
        FB_CUDA(c, cudaMemcpyAsync(c->h_unit_lik.p, c->d_unit_lik.p, nu * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        FB_CUDA(c, rec(c->eve));
        return FB_OK;
    };
    // Passes with little device time (single gaps) are launch-bound: the sequence is captured once per upload into a CUDA
    // graph and replayed (FB_NO_GRAPH=1 enqueues it operation by operation every time).
    // The first pass after an upload goes out directly (a gap set that is filled once does not pay for the capture).
    c->runs_since_upload++;
    const bool want_graph = c->runs_since_upload > 1 && !c->graph_failed && !(getenv("FB_NO_GRAPH") && atoi(getenv("FB_NO_GRAPH")));
    if (want_graph && !c->graph_exec) {
        cudaGraph_t g = nullptr;
        int rc = FB_ECUDA;
        if (cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
            rc = enqueue(true);
            cudaError_t e = cudaStreamEndCapture(c->stream, &g);
            if (rc == FB_OK && e == cudaSuccess && g && cudaGraphInstantiate(&c->graph_exec, g, 0) == cudaSuccess) rc = FB_OK;
            else rc = FB_ECUDA;
            if (g) cudaGraphDestroy(g);
        }
        if (rc != FB_OK) {
            cudaGetLastError();
            c->graph_exec = nullptr;
            c->graph_failed = true;          // this context enqueues directly from now on
        }
    }
    if (want_graph && c->graph_exec) {
        FB_CUDA(c, cudaGraphLaunch(c->graph_exec, c->stream));
    } else {
        int rc = enqueue(false);
        if (rc != FB_OK) return rc;
    }
    tr.lap("run: enqueued");
    return FB_OK;
}

static int run_fused_end(fb_ctx* c) {
    FB_CUDA(c, cudaSetDevice(c->device));
    fb_stats& st = c->stats;
    Trace tr;
    const size_t nu = c->h_units.size();
    const bool fin = !c->fin_gaplist.empty();
    const long d2h = (long)(nu * sizeof(double)) + (fin ? (long)(c->fin_fill_total + 24 * c->gapout_n) : 0);
    FB_CUDA(c, cudaStreamSynchronize(c->stream));     // (not the last event: a graph's record node has not marked it pending yet)
    tr.lap("run: device");
    // ---- results to the gaps (figbird.cpp:2382, 2050-2063, 2503)
    const int D = c->D;
    std::vector<long> fin_liks(c->fin_gaplist.size(), 0);
    parallel_for((int)c->fin_gaplist.size(), 4, [&](int k) {
        const int gi = c->fin_gaplist[k];
        HostGap& g = c->gaps[gi];
        const int ne = g.gmax - g.gmin + 1;
        g.liks.resize(ne);
        g.cons_off.resize(ne);
        for (int e = 0; e < ne; e++) {
            const int u = c->h_sel_unit[(size_t)c->h_sel_off[gi] + e];
            g.liks[e] = c->h_unit_lik.p[u];
            g.cons_off[e] = c->h_units[u].cons_off;
        }
        g.best_est = c->go_gp(c->h_gapout.p)[gi];
        g.best_lik = c->go_lik(c->h_gapout.p)[gi];
        g.no_winner = c->go_off(c->h_gapout.p)[gi] < 0;
        g.total_reads = g.nrec;
        g.discarded = c->go_disc(c->h_gapout.p)[gi];
        g.filled.assign(c->h_filled.p + c->fin_filled_off[gi], (size_t)g.best_est);
        g.done = true;
        for (int r = 0; r < g.nrec; r++) fin_liks[k] += placements(c->h_rmeta[g.rec_off + r], D, g.best_est);
    });
    for (long v : fin_liks) st.finalize_likelihoods += v;
    tr.lap("run: results");
    float ms = 0;
    cudaEventElapsedTime(&ms, c->ev0, c->ev1);
    st.em_kernel_ms = ms;
    cudaEventElapsedTime(&ms, c->evs, c->eve);
    st.step_ms = ms;
    st.finalize_kernel_ms = 0;
    if (!c->fin_gaplist.empty()) {
        cudaEventElapsedTime(&ms, c->ev2, c->ev3);
        st.finalize_kernel_ms = ms;
    }
    st.d2h_bytes = d2h;
    c->scanned = true;
    c->ran = true;
    return FB_OK;
}

static int run_fused(fb_ctx* c) {
    const int rc = run_fused_begin(c);
    return rc != FB_OK ? rc : run_fused_end(c);
}

// The numbers of the last run of both batches as one: kernel times and launches add up, step_ms spans both contexts.
static void merge_run_stats(fb_ctx* c) {
    fb_stats& a = c->stats;
    const fb_stats& b = c->child->stats;
    a.finalize_likelihoods += b.finalize_likelihoods;
    a.em_kernel_ms += b.em_kernel_ms; a.finalize_kernel_ms += b.finalize_kernel_ms;
    a.em_launches += b.em_launches; a.finalize_launches += b.finalize_launches; a.aux_launches += b.aux_launches;
    a.d2h_bytes += b.d2h_bytes;
    float ms = 0;
    if (cudaEventElapsedTime(&ms, c->child->evs, c->eve) == cudaSuccess) a.step_ms = ms;
    else cudaGetLastError();
}

// Both batches of a fb_fill_gap_files gap set, on the device at the same time (two contexts, two sets of streams).
static int run_both(fb_ctx* c) {
    fb_ctx* k = c->child;
    const bool two_call = getenv("FB_TWO_CALL") && atoi(getenv("FB_TWO_CALL"));
    const bool kf = k->fused_ok && !two_call, cf = c->fused_ok && !two_call;
    auto whole = [&](fb_ctx* x) {
        int rc = fb_run_scan(x);
        return rc != FB_OK ? rc : fb_run_finalize(x);
    };
    int rc = kf ? run_fused_begin(k) : whole(k);
    if (rc != FB_OK) return c->fail(rc, "first batch: %s", k->err.c_str());
    rc = cf ? run_fused_begin(c) : whole(c);
    if (rc == FB_OK && cf) rc = run_fused_end(c);
    const int rk = kf ? run_fused_end(k) : FB_OK;
    if (rk != FB_OK) return c->fail(rk, "first batch: %s", k->err.c_str());
    if (rc == FB_OK) merge_run_stats(c);
    return rc;
}

int fb_run_device(fb_ctx* c) {
    if (!c) return FB_EINVAL;
    if (!c->uploaded) return c->fail(FB_ESTATE, "fb_run_device before fb_upload");
    if (c->piped) {
        c->piped = false;                    // the two-call route goes through entry points that refuse a piped context
        const int rc = run_both(c);
        c->piped = true;
        return rc;
    }
    if (c->fused_ok && !(getenv("FB_TWO_CALL") && atoi(getenv("FB_TWO_CALL")))) return run_fused(c);
    int rc = fb_run_scan(c);
    if (rc != FB_OK) return rc;
    return fb_run_finalize(c);
}

int fb_scan_liks(fb_ctx* c, int gap, double* liks, int cap) {
    if (!c || gap < 0 || gap >= (int)c->gaps.size() || !liks) return FB_EINVAL;
    const HostGap& g = c->gaps[gap];
    if (!c->scanned || !g.owned) return c->fail(FB_ESTATE, "gap %d has not been scanned in this context", gap);
    if (cap < (int)g.liks.size()) return FB_EINVAL;
    memcpy(liks, g.liks.data(), g.liks.size() * sizeof(double));
    return (int)g.liks.size();
}

int fb_merge_liks(fb_ctx* c, int gap, const double* liks, int n) {
    if (!c || gap < 0 || gap >= (int)c->gaps.size() || !liks) return FB_EINVAL;
    HostGap& g = c->gaps[gap];
    if (!c->scanned || !g.owned) return c->fail(FB_ESTATE, "gap %d has not been scanned in this context", gap);
    if (n != (int)g.liks.size()) return c->fail(FB_EINVAL, "gap %d scans %zu estimates, %d given", gap, g.liks.size(), n);
    for (int i = 0; i < n; i++) {
        if (std::isnan(g.liks[i])) g.liks[i] = liks[i];
        else if (!std::isnan(liks[i]) && liks[i] != g.liks[i])
            return c->fail(FB_EINVAL, "gap %d estimate %d: merged likelihood differs from the one computed here", gap, g.gmin + i);
    }
    return FB_OK;
}

int fb_run(fb_ctx* c) {
    int rc = fb_upload(c);
    if (rc != FB_OK) return rc;
    return fb_run_device(c);
}

// fillGap for a whole gap set handed over as gap files, in one call.  The set goes to the device in two batches: the
// first (an eighth of the gaps) is parsed, packed, planned, uploaded and enqueued on a second context; while its EM scan
// runs, the host does the same for the rest, so all but the first batch's host work is hidden behind device work.
// Gaps are independent (figbird.cpp:2819-2855 fills them one after the other), so the results are those of
// fb_add_gap_files + fb_run, bit for bit.  handles[i] indexes the getters (fb_result, fb_estimate_liks, fb_best_string,
// fb_dump_profile): the second batch's gaps first, then the first batch's.
int fb_fill_gap_files(fb_ctx* c, const fb_gap_desc* descs, int n, int* handles, int threads) {
    if (!c || n < 0 || (n > 0 && (!descs || !handles))) return FB_EINVAL;
    if (c->piped || !c->gaps.empty()) return c->fail(FB_ESTATE, "fb_fill_gap_files needs a context without gaps (fb_clear_gaps)");
    const char* np = getenv("FB_NO_PIPE");
    const char* mp = getenv("FB_PIPE_MIN");
    const int pipe_min = mp ? std::max(2, atoi(mp)) : 128;
    if (n < pipe_min || (np && atoi(np)) || c->part_world > 1 || c->unit_mode || c->rec_mode) {
        const int ra = fb_add_gap_files(c, descs, n, handles, threads);
        if (ra != FB_OK && ra != FB_EDOMAIN) return ra;
        const int rr = fb_run(c);                 // refused gaps (FB_EDOMAIN) stay out, the others are filled
        return rr != FB_OK ? rr : ra;
    }
    if (!c->child) {
        fb_model m = c->m_model;
        m.error_pos = c->m_error_pos.data();
        m.in_pos = c->m_in_pos.data();
        m.del_pos = c->m_del_pos.data();
        m.ild = c->m_ild.data();
        const int rc = fb_create(&m, c->device, &c->child);
        if (rc != FB_OK) return c->fail(rc, "second context for the first batch could not be created");
    }
    fb_ctx* k = c->child;
    fb_clear_gaps(k);
    const int nA = std::max(pipe_min / 4, n / 8);
    // ---- first batch: on the device before the second is even read
    int refused = 0;
    std::string refused_msg;
    int rc = fb_add_gap_files(k, descs, nA, handles, threads);
    if (rc == FB_EDOMAIN) { refused++; refused_msg = k->err; rc = FB_OK; }
    if (rc == FB_OK) rc = fb_upload(k);
    const bool two_call = getenv("FB_TWO_CALL") && atoi(getenv("FB_TWO_CALL"));
    const bool kf = k->fused_ok && !two_call;
    if (rc == FB_OK) rc = kf ? run_fused_begin(k) : fb_run_device(k);
    if (rc != FB_OK) return c->fail(rc, "first batch: %s", k->err.c_str());
    // ---- second batch, while the first runs
    rc = fb_add_gap_files(c, descs + nA, n - nA, handles + nA, threads);
    if (rc == FB_EDOMAIN) { refused++; refused_msg = c->err; rc = FB_OK; }
    if (rc == FB_OK) rc = fb_upload(c);
    if (rc == FB_OK) rc = fb_run_device(c);
    const int rk = kf ? run_fused_end(k) : FB_OK;       // the first batch is waited for in any case
    if (rc != FB_OK) return rc;
    if (rk != FB_OK) return c->fail(rk, "first batch: %s", k->err.c_str());
    const int nB = (int)c->gaps.size();
    for (int i = 0; i < nA; i++)
        if (handles[i] >= 0) handles[i] += nB;
    c->piped = true;
    {
        // one set of numbers for the whole fill
        fb_stats& a = c->stats;
        const fb_stats& b = k->stats;
        a.n_gaps += b.n_gaps; a.n_units += b.n_units; a.n_records += b.n_records;
        a.likelihoods += b.likelihoods; a.base_terms += b.base_terms;
        a.algorithmic_bytes += b.algorithmic_bytes;
        a.h2d_bytes += b.h2d_bytes;
        merge_run_stats(c);
    }
    if (refused) return c->fail(FB_EDOMAIN, "some gaps were refused (see handles); last: %s", refused_msg.c_str());
    return FB_OK;
}

// Getters of a context filled by fb_fill_gap_files: indices past its own gaps belong to the first batch's context.
static fb_ctx* route(fb_ctx* c, int& gap) {
    if (c && c->piped && c->child && gap >= (int)c->gaps.size()) {
        gap -= (int)c->gaps.size();
        return c->child;
    }
    return c;
}

int fb_result(fb_ctx* c, int gap, fb_gap_result* out, char* filled, int cap) {
    c = route(c, gap);
    if (!c || gap < 0 || gap >= (int)c->gaps.size()) return FB_EINVAL;
    const HostGap& g = c->gaps[gap];
    if (!c->ran || !g.done) return c->fail(FB_ESTATE, "gap %d has no result in this context", gap);
    if (out) {
        out->gap_min = g.gmin;
        out->gap_max = g.gmax;
        out->best_estimate = g.best_est;
        out->best_lik = g.best_lik;
        out->total_reads = g.total_reads;
        out->discarded = g.discarded;
        out->filled_len = (int)g.filled.size();
        out->likelihoods = g.likelihoods;
    }
    if (filled) {
        if (cap < (int)g.filled.size() + 1) return c->fail(FB_EINVAL, "filled buffer too small");
        memcpy(filled, g.filled.data(), g.filled.size());
        filled[g.filled.size()] = 0;
    }
    return FB_OK;
}

int fb_estimate_liks(fb_ctx* c, int gap, double* liks, int cap) {
    c = route(c, gap);
    if (!c || gap < 0 || gap >= (int)c->gaps.size() || !liks) return FB_EINVAL;
    const HostGap& g = c->gaps[gap];
    if (!c->ran || !g.done) return c->fail(FB_ESTATE, "gap %d has no result in this context", gap);
    if (cap < (int)g.liks.size()) return FB_EINVAL;
    memcpy(liks, g.liks.data(), g.liks.size() * sizeof(double));
    return (int)g.liks.size();
}

int fb_best_string(fb_ctx* c, int gap, char* buf, int cap) {
    c = route(c, gap);
    if (!c || gap < 0 || gap >= (int)c->gaps.size() || !buf) return FB_EINVAL;
    const HostGap& g = c->gaps[gap];
    if (!c->ran || !g.done) return c->fail(FB_ESTATE, "gap %d has no result in this context", gap);
    if (cap < g.best_est + 1) return FB_EINVAL;
    if (g.no_winner) {      // bestString was never assigned (see fb_run_finalize)
        buf[0] = 0;
        return FB_OK;
    }
    FB_CUDA(c, cudaSetDevice(c->device));
    for (size_t ui = 0; ui < c->h_units.size(); ui++) {
        const FbUnit& u = c->h_units[ui];
        if (u.gap == gap && u.gp == g.best_est) {
            FB_CUDA(c, cudaMemcpy(buf, c->d_cons.p + u.cons_off, (size_t)u.gp, cudaMemcpyDeviceToHost));
            buf[u.gp] = 0;
            return FB_OK;
        }
    }
    return FB_ESTATE;
}

int fb_dump_profile(fb_ctx* c, int gap, int estimate, double* probs, double* counts) {
    c = route(c, gap);
    if (!c || gap < 0 || gap >= (int)c->gaps.size() || !probs || !counts) return FB_EINVAL;
    if (!c->uploaded) return c->fail(FB_ESTATE, "fb_dump_profile before fb_upload");
    FB_CUDA(c, cudaSetDevice(c->device));
    for (size_t bi = 0; bi < c->bins.size(); bi++) {
        const Bin& b = c->bins[bi];
        for (int k = 0; k < b.n_units; k++) {
            const FbUnit& u = c->h_units[b.first_unit + k];
            if (u.gap != gap || u.gp != estimate) continue;
            const size_t n = (size_t)(estimate + 2 * c->D) * 5;
            FB_CUDA(c, c->d_dump.ensure(2 * n));
            FbEmParams p;
            fill_em_params(c, p);
            p.units = c->d_units.p + b.first_unit + k;
            p.n_units = 1;
            p.work_counter = c->d_counters.p + c->bins.size();
            p.nr_stride = b.nr_stride;
            p.gp_stride = b.gp_stride;
            p.p_stride = b.p_stride;
            p.chunk = b.chunk;
            p.gws_stride = gws_stride_of(b.nr_stride, b.gp_stride);
            p.gws = c->d_gws.p + b.gws_off;
            p.dump_probs = c->d_dump.p;
            p.dump_counts = c->d_dump.p + n;
            p.dump_unit = 0;
            FB_CUDA(c, cudaMemsetAsync(p.work_counter, 0, sizeof(int), c->stream));
            FB_CUDA(c, launch_stage(c, c->stream));
            FB_CUDA(c, fb_launch_em(p, 1, b.threads, b.global_tables, c->f1, b.smem, c->stream));
            FB_CUDA(c, cudaMemcpyAsync(probs, c->d_dump.p, n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
            FB_CUDA(c, cudaMemcpyAsync(counts, c->d_dump.p + n, n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
            FB_CUDA(c, cudaStreamSynchronize(c->stream));
            return FB_OK;
        }
    }
    return c->fail(FB_EINVAL, "estimate %d is not in the scan range of gap %d (or the gap is not owned)", estimate, gap);
}

int fb_devmath_eval(int device, int what, const double* in, double* out, long n) {
    if (!in || !out || n < 0 || (what != 0 && what != 1)) return FB_EINVAL;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) return FB_ECUDA;
    if (cudaSetDevice(device) != cudaSuccess) return FB_ECUDA;
    double *di = nullptr, *dout = nullptr;
    int rc = FB_ECUDA;
    if (cudaMalloc((void**)&di, sizeof(double) * (size_t)std::max(n, 1L)) == cudaSuccess &&
        cudaMalloc((void**)&dout, sizeof(double) * (size_t)std::max(n, 1L)) == cudaSuccess &&
        cudaMemcpy(di, in, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice) == cudaSuccess &&
        fb_launch_devmath(di, dout, n, what, nullptr) == cudaSuccess &&
        cudaMemcpy(out, dout, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost) == cudaSuccess)
        rc = FB_OK;
    cudaFree(di);
    cudaFree(dout);
    return rc;
}

int fb_measure_smem_peak(int device, double* gbs) {
    if (!gbs) return FB_EINVAL;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) return FB_ECUDA;
    if (cudaSetDevice(device) != cudaSuccess) return FB_ECUDA;
    return fb_microbench_smem(gbs, 5) == cudaSuccess ? FB_OK : FB_ECUDA;
}

int fb_get_stats(fb_ctx* c, fb_stats* out) {
    if (!c || !out) return FB_EINVAL;
    *out = c->stats;
    return FB_OK;
}

}  // extern "C"
