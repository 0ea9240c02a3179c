// fb_kernels.cu — hand-written sm_100a kernels for the figbird EM hot path of GapProber.
//
// What the reference does per gap (figbird.cpp:2436-2510) and how it is mapped here:
//
//   for each gap-length estimate gp (independent)            -> one "unit"; persistent CTAs pull units
//     initialize(gp)                  figbird.cpp:1620-1675  -> profile rows built in shared memory
//     repeat num_iter times
//       placeReads                    figbird.cpp:1677-1965
//         E-step  (1826-1872/1889-1938): score = ILD[t] * prod_j (P*(1-e) + e*Q)
//                                                           -> warp per (record, 64 placements), lanes over
//                                                              placements (two per lane), j sequential per lane;
//                                                              chains that start in the left flank begin from
//                                                              fb_prefix_kernel's tabulated running score, first-
//                                                              iteration terms against gap rows are uniform operands
//         M-step  (1874-1881/1940-1947): counts[row][base] += score
//                                                           -> ordered gather: lane per gap row, warp per
//                                                              (32 rows, base class); adds arrive in the
//                                                              reference's (record, placement) order
//         return value (1951-1955)                          -> unit_likelihood(): glibc's log restated, record order
//       computeProbsGap/ErrorProbsGap figbird.cpp:1568-1608  -> fused row update, thread per row
//     computeSequence(0)              figbird.cpp:1984-2047  -> consensus string per unit
//   selection among the estimates     figbird.cpp:2479-2490  -> fb_select_kernel (thread per gap, last wins ties)
//   finalize(best)                    figbird.cpp:2087-2393  -> fb_finalize_kernel, all of it (scores, arg-max, cut-off as a
//                                                              threshold, integer counts, computeSequence(1))
//
// Bit-exactness (hard parts 1-2): the CPU evaluates P*(1-e)+e*Q as two rounded products
// and one rounded sum and multiplies the running score in j order; every FP64 operation below is an
// explicit __dmul_rn/__dadd_rn/__ddiv_rn so no FMA contraction can change a bit, and the M-step adds
// scores to a cell in exactly the order the CPU's scatter loop does.  Likelihood plateaus therefore
// tie exactly as they do on the CPU.  No tensor cores: the path is not a contraction.
//
// Roofs: ~6.5 algorithmic HBM bytes per likelihood (0.68 GB of real DRAM traffic per pass over the chr14-scale set
// against 32 GB algorithmic) vs ~1.6 KB of shared-memory reads and ~400 FP64 instructions per likelihood, so
// shared-memory bandwidth (then issue slots and the FP64 pipe) binds; DESIGN.md has the arithmetic,
// profiles/README.md the measurements.

#include "fb_device.h"
#include "fb_glibc_math.h"

#include <cooperative_groups.h>

namespace cg = cooperative_groups;

namespace {

__device__ __forceinline__ double dmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double dadd(double a, double b) { return __dadd_rn(a, b); }

// computeProbsGap + computeErrorProbsGap for one row (figbird.cpp:1568-1608).
// In: the row's five counts.  Out: P[0..3], Q[0..4]  (P[4] is 0 by definition).
__device__ __forceinline__ void profile_row(const double c[5], const double* __restrict__ T, double P[4], double Q[5]) {
    double total = dadd(dadd(dadd(dadd(c[0], c[1]), c[2]), c[3]), c[4]);
    double n4 = __ddiv_rn(c[4], 4.0);
#pragma unroll
    for (int b = 0; b < 4; b++) P[b] = __ddiv_rn(dadd(c[b], n4), total);   // 0/0 -> NaN is kept
    // Q[j] = sum_{k<4, k!=j} P[k]*T[k][j], k ascending, starting from 0.0
    Q[0] = dadd(dadd(dmul(P[1], T[5]), dmul(P[2], T[10])), dmul(P[3], T[15]));
    Q[1] = dadd(dadd(dmul(P[0], T[1]), dmul(P[2], T[11])), dmul(P[3], T[16]));
    Q[2] = dadd(dadd(dmul(P[0], T[2]), dmul(P[1], T[7])), dmul(P[3], T[17]));
    Q[3] = dadd(dadd(dmul(P[0], T[3]), dmul(P[1], T[8])), dmul(P[2], T[13]));
    Q[4] = dadd(dadd(dadd(dmul(P[0], T[4]), dmul(P[1], T[9])), dmul(P[2], T[14])), dmul(P[3], T[19]));
}

__device__ __forceinline__ unsigned long long warp_max_u64(unsigned long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        unsigned long long w = __shfl_xor_sync(0xffffffffu, v, o);
        v = w > v ? w : v;
    }
    return v;
}

// placeReads' return value in the last iteration (figbird.cpp:1951-1955): one warp, called once per unit.  Kept out of
// line so that its registers (the log's) do not weigh on the EM loop.
template <bool F1>
__device__ __noinline__ void unit_likelihood(const double* maxprob, int nrec, double* out) {
    const int lane = threadIdx.x & 31;
    double acc = 0.0;
    for (int base = 0; base < nrec; base += 32) {
        const int r = base + lane;
        double v = 0.0;
        bool use = false;
        if (r < nrec) {
            v = __ldcg(maxprob + r);              // written by this CTA (cluster) before the barrier the caller passed
            use = F1 ? true : v > 0.0;
            if (!F1 && use) v = fbm_log(v);
        }
        const unsigned used = __ballot_sync(0xffffffffu, use);
        for (int k = 0; k < 32; k++) {            // the additions one after the other, in record order, like the CPU's loop
            const double vk = __shfl_sync(0xffffffffu, v, k);
            if ((used >> k) & 1u) acc = dadd(acc, vk);
        }
    }
    if (lane == 0) *out = acc;
}

// ---------------------------------------------------------------------------------------------------
// Shared-memory plan of the EM kernel.  Per CTA:
//   ild              insert-length window ({1-e,e} per read position comes through the kernel-parameter constant bank)
//   pq[5][nrs+pad]   {P[row][c], Q[row][c]} per base class, class 4 = {0, Q[row][4]}          (E-step reads)
//   cnt[5][gps]      counts of the gap rows                                                  (M-step accumulates)
//   w[2]             placement scores of two in-flight record chunks: rows of p_stride slots separated by kPadW zeros
//   st[3]            staged record chunks: window, flags, class bit-masks, 4-bit packed classes, class-sorted position lists
// One unit runs as a software pipeline over chunks of records: in round r the bulk copies that stage chunk r+1 are
// in flight while the CTA's warps pull, from one dynamic queue, M-step tasks of chunk r-1 and E-step tasks of
// chunk r; everything meets at ONE barrier per round.
// ---------------------------------------------------------------------------------------------------
constexpr int kPadRows = 32;   // a lane's second placement sits 32 rows further; invalid lanes may read (not use) past the end
constexpr int kListBytes = FB_MAX_READ + 16;   // per record: class-sorted position lists, each class starting 4-aligned
constexpr int kPadW = 32;      // zeros either side of every score row: the M-step reads up to 31 placements outside a window

// Per-record tables of a staged chunk.  They depend on the record only, so fb_stage_kernel builds them once per pass in
// HBM (same per-record layout) and the EM kernel brings a chunk in with bulk asynchronous copies (TMA engine, mbarrier
// completion) while the previous chunks are being scored.
struct Stage {
    const FbRecMeta* meta;   // [chunk]      window, insert base, flags (the estimate-dependent clamps are applied by the readers)
    uint32_t* mask;   // [chunk][5][4] bit j of class c: base j of the record is c
    uint32_t* cls4;   // [chunk][16]  base classes, 4 bits each (8 per word), in scoring orientation
    uint8_t* jl;      // [chunk][kListBytes] read positions sorted by (class asc, position DESC), class segments 4-aligned
    uint8_t* cum;     // [chunk][32]  cum[c*4+w] = index in jl of the first position of class c that lies in mask word <= w
                      //              (i.e. class start + number of class-c positions in words above w); cum[20+c] = class end
    unsigned long long* rmax;   // [chunk] running maximum score of the record in this iteration (bit pattern of a positive double)
};
constexpr int kStageCopyBytes = 16 + 80 + 64 + kListBytes + 32;   // per record, brought in by the bulk copies (all but rmax)

struct EmSmem {
    double* ild;
    double2* pq;
    double* cnt;
    double* w0;              // two score buffers, w_stride doubles apart
    size_t w_stride;
    unsigned char* st0;      // three staged chunks, st_stride bytes apart
    size_t st_stride;
    int chunk;
};

// Carve one staged chunk out of its slab (same order as em_carve lays it out); every section starts 16-byte aligned.
__device__ __forceinline__ Stage stage_at(const EmSmem& s, int k) {
    unsigned char* b = s.st0 + (size_t)k * s.st_stride;
    Stage t;
    t.meta = reinterpret_cast<const FbRecMeta*>(b);
    b += sizeof(FbRecMeta) * s.chunk;
    t.mask = reinterpret_cast<uint32_t*>(b);
    t.cls4 = t.mask + 20 * s.chunk;
    t.jl = reinterpret_cast<uint8_t*>(t.cls4 + 16 * s.chunk);
    t.cum = t.jl + kListBytes * s.chunk;
    t.rmax = reinterpret_cast<unsigned long long*>(t.cum + 32 * s.chunk);
    return t;
}

__host__ __device__ inline size_t align16(size_t x) { return (x + 15) & ~size_t(15); }

__host__ __device__ inline size_t em_carve(unsigned char* base, EmSmem* s, int nr_stride, int gp_stride, int p_stride, int chunk,
                                           int pwin, bool global_tables) {
    size_t off = 0;
    auto take = [&](size_t bytes) {
        unsigned char* p = base ? base + off : nullptr;
        off = align16(off + bytes);
        return p;
    };
    unsigned char* a;
    a = take(sizeof(double) * pwin); if (s) s->ild = (double*)a;
    if (!global_tables) {
        a = take(sizeof(double2) * (5 * (size_t)nr_stride + kPadRows)); if (s) s->pq = (double2*)a;
        a = take(sizeof(double) * 5 * (size_t)gp_stride); if (s) s->cnt = (double*)a;
    }
    // a score buffer: kPadW zeros, then per record slot [p_stride scores][kPadW zeros]; neighbouring rows share a pad
    const size_t wbytes = align16(sizeof(double) * ((size_t)kPadW + (size_t)chunk * (p_stride + kPadW)));
    a = take(2 * wbytes);
    if (s) { s->w0 = (double*)a; s->w_stride = wbytes / sizeof(double); }
    // one staged chunk: meta[chunk] 16 B | mask[chunk][20] u32 | cls4[chunk][16] u32 | jl[chunk][144] u8 | cum[chunk][32] u8 |
    //                   rmax[chunk] u64
    const size_t sbytes = align16((size_t)chunk * (kStageCopyBytes + 8));
    a = take(3 * sbytes);
    if (s) { s->st0 = a; s->st_stride = sbytes; s->chunk = chunk; }
    return off;
}

// 2-bit codes + non-ACGT mask of one record -> this lane's four base classes (positions 4*lane .. 4*lane+3):
// 0-3 ACGT, 4 other, 5 past the end.  128-bit loads by lanes 0-2, redistributed with shuffles.
__device__ __forceinline__ void fetch_classes(const uint4* __restrict__ codes, const uint4* __restrict__ nmask, int rec, int L2,
                                              int lane, unsigned cls[4]) {
    uint4 v = make_uint4(0, 0, 0, 0);
    if (lane < 2) v = __ldg(&codes[(size_t)rec * 2 + lane]);
    else if (lane == 2) v = __ldg(&nmask[rec]);
    const int wsel = lane >> 2;          // code word holding bases 4*lane .. 4*lane+3
    const int src = wsel >> 2, comp = wsel & 3;
    unsigned x0 = __shfl_sync(0xffffffffu, v.x, src), x1 = __shfl_sync(0xffffffffu, v.y, src);
    unsigned x2 = __shfl_sync(0xffffffffu, v.z, src), x3 = __shfl_sync(0xffffffffu, v.w, src);
    unsigned cw = comp == 0 ? x0 : comp == 1 ? x1 : comp == 2 ? x2 : x3;
    const int msel = lane >> 3;          // mask word holding bases 4*lane ..
    unsigned m0 = __shfl_sync(0xffffffffu, v.x, 2), m1 = __shfl_sync(0xffffffffu, v.y, 2);
    unsigned m2 = __shfl_sync(0xffffffffu, v.z, 2), m3 = __shfl_sync(0xffffffffu, v.w, 2);
    unsigned mw = msel == 0 ? m0 : msel == 1 ? m1 : msel == 2 ? m2 : m3;
#pragma unroll
    for (int q = 0; q < 4; q++) {
        int j = lane * 4 + q;
        unsigned c = (cw >> (2 * (j & 15))) & 3u;
        if ((mw >> (j & 31)) & 1u) c = 4;
        if (j >= L2) c = 5;
        cls[q] = c;
    }
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// 8-byte load from a shared-memory byte address
__device__ __forceinline__ double lds_f64(uint32_t addr) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
    return v;
}

__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint32_t lds_u8(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!done);
}
// Bulk asynchronous copy global -> this CTA's shared memory; bytes and both addresses are multiples of 16.
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src),
                 "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// Byte-code form used by the finalize kernel.
__device__ __forceinline__ void load_record_bytes(const uint4* __restrict__ codes, const uint4* __restrict__ nmask, int rec, int L2,
                                                  uint8_t* cs, int lane) {
    unsigned cls[4];
    fetch_classes(codes, nmask, rec, L2, lane, cls);
    reinterpret_cast<uint32_t*>(cs)[lane] = cls[0] | (cls[1] << 8) | (cls[2] << 16) | (cls[3] << 24);
    __syncwarp();
}

// E-step of 64 placements of one record: lane owns placements ua and ua+32 (two independent product chains).
// The running score is multiplied base by base in read order (figbird.cpp:1840-1864).  The read-position index of base j
// is j, or L2-1-j for a reverse-complemented mate (figbird.cpp:1845-1852): FbEmParams::eo_tab holds the {1-e, e} table
// twice, forwards and backwards, so that both orientations walk it upwards from their own start k0 — one copy of the
// loop instead of two (the task loop's code no longer fits the 32 KB instruction-cache level otherwise).
// F1 is the reference's log-domain mode (argv[2] = 1): the running score is a SUM of glibc logs of the same terms
// (figbird.cpp:1856, 1861), started from log(ILD[t]) (figbird.cpp:1835).
template <bool F1>
__device__ __forceinline__ double score_step(double s, double term) {
    return F1 ? dadd(s, fbm_log(term)) : dmul(s, term);
}

// eo_tab index of base 0 of a mate of length L2
__device__ __forceinline__ int eo_start(bool rev, int L2) { return rev ? 2 * FB_MAX_READ - L2 : 0; }

// [jb, je): the bases to multiply in; jb is a multiple of 4 (a chain with a left-flank prefix starts past its first terms).
template <bool TWO, bool F1>
__device__ __forceinline__ void score_pair(const unsigned char* pa, const uint32_t* __restrict__ cls4, unsigned cstride, const FbEmParams& p,
                                           int k0, int jb, int je, double& sa, double& sb) {
    int k = k0 + jb;                   // eo_tab index of base j
    int j0 = jb;
    pa += jb * 16;
    // four bases: their classes are the low 16 bits of cw (4 bits each)
    auto four = [&](const uint32_t cw) {
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const unsigned char* a = pa + ((cw >> (4 * q)) & 7u) * cstride + q * 16;   // class table of this base, row of this placement
            const double2 va = *reinterpret_cast<const double2*>(a);
            const double2 e = p.eo_tab[k + q];                              // {1 - e[k], e[k]} via the constant cache
            sa = score_step<F1>(sa, dadd(dmul(va.x, e.x), dmul(e.y, va.y)));   // figbird.cpp:1855 / 1862
            if (TWO) {
                const double2 vb = *reinterpret_cast<const double2*>(a + 32 * 16);
                sb = score_step<F1>(sb, dadd(dmul(vb.x, e.x), dmul(e.y, vb.y)));
            }
        }
        pa += 64;
        k += 4;
    };
    if ((j0 & 4) && j0 + 4 <= je) {          // up to an 8-base boundary of the packed classes
        four(cls4[j0 >> 3] >> 16);
        j0 += 4;
    }
    for (; j0 + 8 <= je; j0 += 8) {          // one 32-bit word of classes per trip
        const uint32_t cw = cls4[j0 >> 3];
        four(cw);
        four(cw >> 16);
    }
    if (j0 + 4 <= je) {
        four(cls4[j0 >> 3]);
        j0 += 4;
    }
#pragma unroll 1
    for (; j0 < je; j0++) {                  // at most three bases
        const unsigned c = (cls4[j0 >> 3] >> ((j0 & 7) * 4)) & 7u;
        const unsigned char* a = pa + c * cstride;
        const double2 va = *reinterpret_cast<const double2*>(a);
        const double2 e = p.eo_tab[k];
        sa = score_step<F1>(sa, dadd(dmul(va.x, e.x), dmul(e.y, va.y)));
        if (TWO) {
            const double2 vb = *reinterpret_cast<const double2*>(a + 32 * 16);
            sb = score_step<F1>(sb, dadd(dmul(vb.x, e.x), dmul(e.y, vb.y)));
        }
        pa += 16;
        k++;
    }
}

// First iteration, bases [jb, je) (both multiples of 4) of a task whose placements are all on gap rows there: every
// lane multiplies by the same term (FbEmParams::t0_tab), no table loads.
template <bool TWO, bool F1>
__device__ __forceinline__ void score_uniform(const uint32_t* __restrict__ cls4, const FbEmParams& p, int L2, bool rev, int jb, int je,
                                              double& sa, double& sb) {
    const int ks = rev ? -1 : 1;
    int k = rev ? L2 - 1 - jb : jb;
    for (int j0 = jb; j0 < je; j0 += 4) {
        const uint32_t cw = cls4[j0 >> 3] >> ((j0 & 4) * 4);
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const double t = p.t0_tab[(cw >> (4 * q)) & 7u][k + q * ks];
            sa = F1 ? dadd(sa, t) : dmul(sa, t);
            if (TWO) sb = F1 ? dadd(sb, t) : dmul(sb, t);
        }
        k += 4 * ks;
    }
}

// Per-record running maximum as an unsigned key (atomicMax in shared memory).  f=0: maxProb starts at 0 and only scores
// above 0 win (figbird.cpp:1815, 1866), so the bit pattern of a positive double orders correctly and 0 means "none".
// f=1: maxProb starts at -DBL_MAX and log scores are negative (figbird.cpp:1814): the usual order-preserving flip of
// the sign-magnitude pattern; scores that do not beat -DBL_MAX (NaN, -inf, -DBL_MAX itself) give the key of -DBL_MAX.
template <bool F1>
__device__ __forceinline__ unsigned long long score_key(double s, bool valid) {
    if (!F1) return (valid && s > 0.0) ? (unsigned long long)__double_as_longlong(s) : 0ull;
    const double v = (valid && s > -1.7976931348623157e308) ? s : -1.7976931348623157e308;
    const unsigned long long b = (unsigned long long)__double_as_longlong(v);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
template <bool F1>
__device__ __forceinline__ double key_score(unsigned long long k) {
    if (!F1) return __longlong_as_double((long long)k);
    return __longlong_as_double((long long)((k >> 63) ? (k & 0x7fffffffffffffffull) : ~k));
}
template <bool F1>
__device__ __forceinline__ unsigned long long key_init() {
    return F1 ? score_key<true>(-1.7976931348623157e308, true) : 0ull;
}

// Per-record tables of the EM kernel, one warp per record (they do not depend on the estimate, so every unit of the
// gap reuses them): 4-bit packed classes, class bit-masks, class-sorted position lists and their prefix counts.
__global__ void __launch_bounds__(256) fb_stage_kernel(const FbStageParams p) {
    const int lane = threadIdx.x & 31;
    const int rec = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (rec >= p.n_records) return;
    const int L2 = p.rmeta[rec].flags & 0xff;
    unsigned cls[4];
    fetch_classes(p.codes, p.nmask, rec, L2, lane, cls);
    {
        // four classes of this lane -> 16 bits; two lanes make one 32-bit word (8 bases)
        const unsigned nib4 = min(cls[0], 4u) | (min(cls[1], 4u) << 4) | (min(cls[2], 4u) << 8) | (min(cls[3], 4u) << 12);
        const unsigned hi16 = __shfl_down_sync(0xffffffffu, nib4, 1);
        if ((lane & 1) == 0) p.st_cls4[(size_t)rec * 16 + (lane >> 1)] = nib4 | (hi16 << 16);
    }
    uint8_t* cum = p.st_cum + (size_t)rec * 32;
    uint8_t* jl = p.st_jl + (size_t)rec * kListBytes;
    for (int i = lane; i < kListBytes / 4; i += 32) reinterpret_cast<uint32_t*>(jl)[i] = 0u;   // the gaps between class segments
    if (lane >= 25) cum[lane] = 0;
    __syncwarp();
    const unsigned grp = 0xffu << (lane & 24);           // the 8 lanes sharing one 32-bit mask word
    int start = 0;
#pragma unroll
    for (int c = 0; c < 5; c++) {
        unsigned nib = 0;
#pragma unroll
        for (int q = 0; q < 4; q++) nib |= (cls[q] == (unsigned)c ? 1u : 0u) << q;
        const unsigned word = __reduce_or_sync(grp, nib << (4 * (lane & 7)));     // bit j of word w: base 32w+j is of class c
        if ((lane & 7) == 0) p.st_mask[(size_t)rec * 20 + c * 4 + (lane >> 3)] = word;
        const unsigned m0 = __shfl_sync(0xffffffffu, word, 0), m1 = __shfl_sync(0xffffffffu, word, 8);
        const unsigned m2 = __shfl_sync(0xffffffffu, word, 16), m3 = __shfl_sync(0xffffffffu, word, 24);
        // class-sorted, position-descending list: slot = class start + (positions of that class above j)
        const int n3 = __popc(m3), n2 = __popc(m2), n1 = __popc(m1), n0 = __popc(m0);
        if (lane == 0) {
            cum[c * 4 + 0] = (uint8_t)(start + n3 + n2 + n1);      // positions in words above w
            cum[c * 4 + 1] = (uint8_t)(start + n3 + n2);
            cum[c * 4 + 2] = (uint8_t)(start + n3);
            cum[c * 4 + 3] = (uint8_t)start;
            cum[20 + c] = (uint8_t)(start + n3 + n2 + n1 + n0);
        }
#pragma unroll
        for (int q = 0; q < 4; q++) {
            if (cls[q] == (unsigned)c) {
                const int j = lane * 4 + q, w = j >> 5;
                const unsigned mw = w == 0 ? m0 : w == 1 ? m1 : w == 2 ? m2 : m3;
                const int higher = __popc((mw >> (j & 31)) >> 1);
                jl[start + (w == 3 ? 0 : w == 2 ? n3 : w == 1 ? n3 + n2 : n3 + n2 + n1) + higher] = (uint8_t)j;
            }
        }
        start = (start + n3 + n2 + n1 + n0 + 3) & ~3;     // next class starts on a 32-bit boundary
    }
}

// FbEmParams::t0_tab — the first iteration's term of a base of class c at read position k against a gap row as
// initialize() leaves it (counts N = 1): the E-step's own operations on that row's {P,Q}, once per (c, k).
template <bool F1>
__global__ void fb_t0_kernel(const double2* __restrict__ eo, const double* __restrict__ T, double* __restrict__ out) {
    __shared__ double sT[25];
    __shared__ double2 row[5];
    const int tid = threadIdx.x;
    if (tid < 25) sT[tid] = T[tid];
    __syncthreads();
    if (tid == 0) {
        double c[5] = {0.0, 0.0, 0.0, 0.0, 1.0};
        double P[4], Q[5];
        profile_row(c, sT, P, Q);
#pragma unroll
        for (int k = 0; k < 4; k++) row[k] = make_double2(P[k], Q[k]);
        row[4] = make_double2(0.0, Q[4]);
    }
    __syncthreads();
    for (int i = tid; i < 5 * FB_MAX_READ; i += blockDim.x) {
        const double2 v = row[i / FB_MAX_READ], e = eo[i % FB_MAX_READ];
        const double term = dadd(dmul(v.x, e.x), dmul(e.y, v.y));
        out[i] = F1 ? fbm_log(term) : term;
    }
}

// Left-flank prefixes (FbPrefixParams in fb_device.h): warp per record, lanes over the 32 placements of a chain.  The
// operations are the E-step's own — start value, then per base one rounded term and one rounded product (or log and
// sum in f=1) in read order — against the one-hot rows initialize() gives the flank, so the E-step may start from the
// stored value and produce the same bits as if it had started at base 0.
template <bool F1>
__global__ void __launch_bounds__(256) fb_prefix_kernel(const __grid_constant__ FbPrefixParams p) {
    __shared__ double sT[25];
    __shared__ double2 sRow[5][5];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = blockDim.x >> 5;
    if (tid < 25) sT[tid] = p.T[tid];
    __syncthreads();
    if (tid < 5) {
        double c[5];
#pragma unroll
        for (int k = 0; k < 5; k++) c[k] = (k == tid) ? 1.0 : 0.0;
        double P[4], Q[5];
        profile_row(c, sT, P, Q);
#pragma unroll
        for (int k = 0; k < 4; k++) sRow[tid][k] = make_double2(P[k], Q[k]);
        sRow[tid][4] = make_double2(0.0, Q[4]);
    }
    __syncthreads();
    const int D = p.D;
    for (int g = blockIdx.x; g < p.n_gaps; g += gridDim.x) {
        const FbGapMeta gm = p.gaps[g];
        const uint8_t* fl = p.flank + (size_t)g * 2 * D;
        for (int r = blockIdx.y * nwarp + warp; r < gm.nrec; r += gridDim.y * nwarp) {
            const size_t rec = (size_t)gm.rec_off + r;
            const FbRecMeta rm = p.rmeta[rec];
            const int L2 = rm.flags & 0xff;
            const bool rev = (rm.flags >> 8) & 1, left = (rm.flags >> 9) & 1;
            const int lo = max(rm.a_lo, D - L2 + 1);
            if (rm.a_hi < lo) continue;
            const int npc = min((D - lo - 3) / 32, (rm.a_hi - lo + 32) / 32);
            const uint32_t* cls4 = p.st_cls4 + rec * 16;
            double* out = p.pre + ((size_t)((unsigned)rm.flags >> FB_PRE_SHIFT)) * 32 + lane;
            for (int c = 0; c < npc; c++) {
                const int u = lo + 32 * c + lane;
                const int skip = (D - (lo + 32 * c + 31)) & ~3;
                double s = 0.0;
                if (u <= rm.a_hi) {
                    s = p.ild[(left ? u + rm.tb : rm.tb - u) - p.thr_min];
                    for (int j = 0; j < skip; j++) {
                        const unsigned cl = (cls4[j >> 3] >> ((j & 7) * 4)) & 7u;
                        const double2 v = sRow[fl[u + j]][cl];
                        const double2 e = p.eo_tab[rev ? L2 - 1 - j : j];
                        s = score_step<F1>(s, dadd(dmul(v.x, e.x), dmul(e.y, v.y)));
                    }
                }
                out[(size_t)c * 32] = s;
            }
        }
    }
}

// CLU: one unit is run by a thread-block cluster instead of one CTA — for launches with fewer units than the GPU has
// CTA slots (single gaps, the stress sweep).  Nothing about the arithmetic changes, so results stay bit-exact:
//   * every CTA keeps a full copy of the P/Q table and of the two score buffers in its own shared memory;
//   * E tasks are dealt to the CTAs by record; a task writes its score row into every CTA's buffer (DSMEM stores:
//     one row of scores per ~100 rows' worth of table reads, so the remote traffic is negligible);
//   * M tasks are dealt by 32-row block: a CTA owns the counts of its row blocks and adds to them in the reference's
//     order from its local copy of the scores;
//   * the profile update of a row is done by the owner of its counts and broadcast to every CTA's table;
//   * the per-round barrier becomes a cluster barrier.
// NT: threads per CTA — 512 with two CTAs per SM where the tables allow it, 1024 with one CTA per SM for the bigger
// estimates, so that an SM always has 32 warps to hide the shared-memory (or, with GT, L2) latency.
template <bool GT, bool CLU, int NT, bool F1>
__global__ void __launch_bounds__(NT, (GT || NT > 512) ? 1 : 2) fb_em_kernel(const __grid_constant__ FbEmParams p) {
    static_assert(!(GT && CLU), "the cluster variant keeps its tables in shared memory");
    if (p.abort && *reinterpret_cast<const volatile int*>(p.abort)) return;   // a barrier of this record-sharded pass has failed
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned crank = 0, csize = 1;
    if (CLU) {
        crank = cg::this_cluster().block_rank();
        csize = cg::this_cluster().num_blocks();
    }
    auto sync_all = [&]() {
        if (CLU) cg::this_cluster().sync();
        else __syncthreads();
    };
    EmSmem s;
    const int pwin = p.thr_max - p.thr_min + 1;
    em_carve(smem_raw, &s, p.nr_stride, p.gp_stride, p.p_stride, p.chunk, pwin, GT);
    if (GT) {
        double* g = p.gws + (size_t)blockIdx.x * p.gws_stride;
        s.pq = reinterpret_cast<double2*>(g);
        s.cnt = g + 2 * ((size_t)5 * p.nr_stride + kPadRows);
    }
    __shared__ int s_unit;
    __shared__ int s_queue[2];
    __shared__ __align__(8) uint64_t s_bar[3];   // one per stage buffer: completion of the bulk copies that fill it
    __shared__ Stage s_stage[3];                 // the three stage buffers' sections: fixed for the launch, carved once
    uint32_t bar_phase = 0;                      // bit k: parity the next wait on s_bar[k] has to see (same in every thread)
    __shared__ double sT[25];
    __shared__ double2 sRow[5][5];   // {P,Q} per base class of a row whose counts are one-hot in base b (flanks, and b=4: untouched gap rows)

    const int tid = threadIdx.x, lane = tid & 31;
    const int D = p.D, CH = p.chunk, PS = p.p_stride;

    for (int i = tid; i < pwin; i += NT) s.ild[i] = p.ild[i];
    if (tid < 25) sT[tid] = p.T[tid];
    if (tid < 3) s_stage[tid] = stage_at(s, tid);
    if (tid == 0) {
        for (int k = 0; k < 3; k++) mbar_init(&s_bar[k], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (GT) {
        for (int i = tid; i < kPadRows; i += NT) s.pq[5 * (size_t)p.nr_stride + i] = make_double2(0.0, 0.0);
    }
    // score buffers start as zeros; E tasks only ever write the p_stride slots of their row, so the pads between rows
    // stay zero for the life of the kernel
    for (size_t i = tid; i < 2 * s.w_stride; i += NT) s.w0[i] = 0.0;
    __syncthreads();
    if (tid < 5) {
        // initialize() only ever produces five different rows (figbird.cpp:1657-1673): compute them once
        double c[5];
#pragma unroll
        for (int k = 0; k < 5; k++) c[k] = (k == tid) ? 1.0 : 0.0;
        double P[4], Q[5];
        profile_row(c, sT, P, Q);
#pragma unroll
        for (int k = 0; k < 4; k++) sRow[tid][k] = make_double2(P[k], Q[k]);
        sRow[tid][4] = make_double2(0.0, Q[4]);
    }

    for (;;) {
        sync_all();
        if (tid == 0) {
            if (!CLU) {
                s_unit = atomicAdd(p.work_counter, 1);
            } else if (crank == 0) {
                const int u = atomicAdd(p.work_counter, 1);
                for (unsigned k = 0; k < csize; k++) *cg::this_cluster().map_shared_rank(&s_unit, k) = u;
            }
            s_queue[0] = 0;
        }
        sync_all();
        const int ui = s_unit;
        if (ui >= (p.n_units_ptr ? min(p.n_units, *p.n_units_ptr) : p.n_units)) break;
        const FbUnit unit = p.units[ui];
        const FbGapMeta gm = p.gaps[unit.gap];
        const int gp = unit.gp, nrec = gm.nrec, lmax = gm.lmax;
        const int R0 = D - lmax + 1;                 // first profile row any placement can touch
        const int NR = gp + 2 * lmax - 2;            // rows [R0, R0+NR) cover every touched row
        const int nrs = p.nr_stride, gps = p.gp_stride;
        // Long gaps: no placement of any record touches the rows [hl, hl + hn) in the middle of the gap (reads reach at
        // most an insert length in from either flank), so the tables leave them out.  hl - D is a multiple of 32, so
        // an M-step row block never straddles the hole; every record's rows lie entirely on one side of it.
        const int hl = unit.hole_lo, hn = unit.hole_len;
        const int gpc = gp - hn, NRc = NR - hn;       // rows the tables hold
        const uint8_t* fl = p.flank + (size_t)unit.gap * 2 * D;

        // ---- initialize(gp): flanks one-hot from the scaffold, gap rows N=1 (figbird.cpp:1657-1673)
        for (int xr = tid; xr < NRc; xr += NT) {
            int x = xr + R0;
            if (x >= hl) x += hn;
            int b = 4;
            if (x < D) b = fl[x];
            else if (x >= D + gp) b = fl[D + (x - D - gp)];
            if (p.n_parts && x >= D && x < D + gp) {   // gap rows only: a flank row with a non-ACGT base keeps sRow[4]
                // record-sharded scan, iteration > 0: the counts of this gap row are the sum of every rank's partial
                // counts of the previous iteration, always added in rank order
                const size_t o = 5 * (size_t)unit.cons_off + (x - D);
                double c[5];
#pragma unroll
                for (int k = 0; k < 5; k++) c[k] = p.parts_in[0][o + (size_t)k * gp];
                for (int q = 1; q < p.n_parts; q++) {
#pragma unroll
                    for (int k = 0; k < 5; k++) c[k] = dadd(c[k], p.parts_in[q][o + (size_t)k * gp]);
                }
                double P[4], Q[5];
                profile_row(c, sT, P, Q);
#pragma unroll
                for (int k = 0; k < 4; k++) s.pq[k * nrs + xr] = make_double2(P[k], Q[k]);
                s.pq[4 * nrs + xr] = make_double2(0.0, Q[4]);
                continue;
            }
#pragma unroll
            for (int k = 0; k < 5; k++) s.pq[k * nrs + xr] = sRow[b][k];
        }
        for (int i = tid; i < gpc; i += NT) {
#pragma unroll
            for (int k = 0; k < 5; k++) s.cnt[k * gps + i] = 0.0;
        }
        __syncthreads();

        const int nchunks = (nrec + CH - 1) / CH;
        const int nrb = (gpc + 31) >> 5;              // M-step row blocks of 32 rows (of the rows the tables hold)
        const int nblk = (PS + 63) >> 6;              // E-step blocks of 64 placements per record
        const int WS = PS + kPadW;                    // score-row stride: the pad after a row is the pad before the next
        const bool keep_staged = nchunks <= 3;        // every chunk owns one of the three stage buffers for the whole unit
        int round = 0;                                // global round counter: selects the queue slot
        for (int it = 0; it < p.num_iter; it++) {
            const bool last = it == p.num_iter - 1;
            // Two-pass scan: the counts of the last iteration only serve computeSequence(0), which matters for the estimate
            // that wins the gap (figbird.cpp:2479-2490) — the first pass leaves the last M-step out, the winners are run
            // again in full once the likelihoods have picked them.
            const bool skip_m = last && p.skip_last_m;
            const bool uni0 = it == 0 && p.use_t0 && !p.n_parts;   // every gap row still holds initialize()'s counts
            // rounds of the pipeline: stage chunk r+1 | E-step chunk r | M-step chunk r-1   (r = -1 stages chunk 0)
            for (int r = ((it == 0 || !keep_staged) ? -1 : 0); r <= nchunks; r++, round++) {
                const int cl = r + 1, ce = r, cm = r - 1;
                const bool full_stage = it == 0 || !keep_staged;       // otherwise only the per-record maximum is reset
                const int n_l = cl < nchunks ? min(CH, nrec - cl * CH) : 0;
                const int n_e = (ce >= 0 && ce < nchunks) ? min(CH, nrec - ce * CH) : 0;
                const int n_m = (cm >= 0 && cm < nchunks) ? min(CH, nrec - cm * CH) : 0;
                const Stage &SL = s_stage[cl % 3], &SE = s_stage[(ce + 3) % 3], &SM = s_stage[(cm + 3) % 3];
                double* wE = s.w0 + ((ce + 2) & 1) * s.w_stride;
                const double* wM = s.w0 + ((cm + 2) & 1) * s.w_stride;
                if (last && tid < n_m) {   // per-record maximum of the last iteration (feeds figbird.cpp:1951-1955 on the host)
                    if (!CLU) p.unit_maxprob[unit.mp_off + cm * CH + tid] = key_score<F1>(SM.rmax[tid]);
                    else if ((unsigned)tid % csize == crank)   // all E tasks of a record run in one CTA: it holds the maximum
                        p.unit_maxprob[unit.mp_off + cm * CH + tid] = key_score<F1>(SM.rmax[tid]);
                }
                // a cluster deals E tasks by record and M tasks by row block; staging is done by every CTA for itself
                const int nrb_q = CLU ? (nrb + (int)csize - 1 - (int)crank) / (int)csize : nrb;
                const int ne_q = CLU ? (n_e + (int)csize - 1 - (int)crank) / (int)csize : n_e;
                if (n_l) {
                    // ---- stage chunk cl: bulk copies of the records' prepared tables (fb_stage_kernel), waited for at the end of the round
                    if (full_stage && tid == 0) {
                        uint64_t* bar = &s_bar[cl % 3];
                        const size_t rec0 = (size_t)gm.rec_off + (size_t)cl * CH;
                        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the buffer was last read through the generic proxy
                        mbar_expect_tx(bar, (uint32_t)n_l * kStageCopyBytes);
                        bulk_g2s(const_cast<FbRecMeta*>(SL.meta), p.rmeta + rec0, (uint32_t)n_l * 16u, bar);
                        bulk_g2s(SL.mask, p.st_mask + rec0 * 20, (uint32_t)n_l * 80u, bar);
                        bulk_g2s(SL.cls4, p.st_cls4 + rec0 * 16, (uint32_t)n_l * 64u, bar);
                        bulk_g2s(SL.jl, p.st_jl + rec0 * kListBytes, (uint32_t)n_l * kListBytes, bar);
                        bulk_g2s(SL.cum, p.st_cum + rec0 * 32, (uint32_t)n_l * 32u, bar);
                    }
                    if (tid < n_l) SL.rmax[tid] = key_init<F1>();   // new iteration (or new chunk), new maximum
                }
                const int tM = (n_m && !skip_m) ? nrb_q * 5 : 0, tE = ne_q * nblk;
                const int ntask = tM + tE;
                int* queue = &s_queue[round & 1];
                for (;;) {
                    int task = 0;
                    if (lane == 0) task = atomicAdd(queue, 1);
                    task = __shfl_sync(0xffffffffu, task, 0);
                    if (task >= ntask) break;
                    if (task < tM) {
                        // ---- M-step: a lane owns one gap row, a warp one (32 rows, base class) cell block.  Within a cell the CPU adds in
                        // (record asc, placement asc) order (figbird.cpp:1718, 1826, 1874): walking a record's class list from high j
                        // to low j visits its placements in ascending order.  Score rows are zero outside the record's window (pads +
                        // zero-filled tail), so a row that is not inside the window for some entry adds +0.0 — exact: the
                        // accumulators are sums of non-negative scores (or NaN), never -0.0 — and the loop needs no predicates:
                        // per entry one LDS.64 + DADD, list positions fetched four at a time.
                        int c = 0, rbq = task;                                   // class-major: five classes, counted off
                        while (rbq >= nrb_q) { rbq -= nrb_q; c++; }
                        const int rb = CLU ? rbq * (int)csize + (int)crank : rbq;
                        const int xg = rb * 32 + lane, x0 = D + rb * 32 + (D + rb * 32 >= hl ? hn : 0);   // table row, profile row of lane 0
                        const bool ok0 = xg < gpc;
                        double* cr = s.cnt + c * gps + xg;
                        double a0 = ok0 ? cr[0] : 0.0;                             // rows past the gap accumulate junk that is never stored
                        // per-record set-up, one record per lane: the exact sub-range [qs, qe) of its class list whose positions
                        // put some row of this block inside the window (two popcounts)
                        int v_lo = 0, v_qs = 0, v_qe = 0;
                        if (lane < n_m) {
                            const FbRecMeta rm = SM.meta[lane];
                            const int L2 = rm.flags & 0xff;
                            // placements i in [gapStart-L2+1, gapStart+gp) (figbird.cpp:1826) -> u in [D-L2+1, D+gp-1]
                            const int lo = max(rm.a_lo, D - L2 + 1), hi = min(rm.a_hi, D + gp - 1);
                            const int jhi = min(L2 - 1, x0 + 31 - lo), jlo = max(0, x0 - hi);
                            if (hi >= lo && jhi >= jlo) {
                                const uint32_t* mk = SM.mask + lane * 20 + c * 4;
                                const uint8_t* cum = SM.cum + lane * 32;
                                const int wh = jhi >> 5, wl = jlo >> 5;
                                v_qs = cum[c * 4 + wh] + __popc((mk[wh] >> (jhi & 31)) >> 1);
                                v_qe = (wl == 0 ? cum[20 + c] : cum[c * 4 + wl - 1]) - __popc(mk[wl] & ((1u << (jlo & 31)) - 1u));
                                v_lo = lo;
                            }
                        }
                        // shared-memory byte addresses by hand (ld.shared with a 32-bit address): per entry one byte extract, one
                        // multiply-add and the load, instead of an index rebuilt from its parts
                        const uint32_t wrow0 = smem_u32(wM + kPadW + x0 + lane), jl0 = smem_u32(SM.jl);
                        for (int rl = 0; rl < n_m; rl++) {
                            const int qs = __shfl_sync(0xffffffffu, v_qs, rl), qe = __shfl_sync(0xffffffffu, v_qe, rl);
                            if (qs >= qe) continue;
                            const int lo = __shfl_sync(0xffffffffu, v_lo, rl);
                            const uint32_t jl = jl0 + (uint32_t)(rl * kListBytes);       // this record's class-sorted position list
                            const uint32_t wr = wrow0 + (uint32_t)((rl * WS - lo) * 8);   // score of placement (row - j) of this record at wr - 8 j
                            uint32_t q = jl + (uint32_t)qs;
                            const uint32_t qend = jl + (uint32_t)qe;
                            for (; q < qend && (q & 3u); q++) a0 = dadd(a0, lds_f64(wr - 8u * lds_u8(q)));        // figbird.cpp:1879
#pragma unroll 2
                            for (; q + 4 <= qend; q += 4) {
                                const uint32_t j4 = lds_u32(q);
                                const double w0 = lds_f64(wr - 8u * __byte_perm(j4, 0, 0x4440)), w1 = lds_f64(wr - 8u * __byte_perm(j4, 0, 0x4441));
                                const double w2 = lds_f64(wr - 8u * __byte_perm(j4, 0, 0x4442)), w3 = lds_f64(wr - 8u * __byte_perm(j4, 0, 0x4443));
                                a0 = dadd(dadd(dadd(dadd(a0, w0), w1), w2), w3);
                            }
                            for (; q < qend; q++) a0 = dadd(a0, lds_f64(wr - 8u * lds_u8(q)));
                        }
                        if (ok0) cr[0] = a0;
                    } else {
                        // ---- E-step: warp per (record, 64 placements); lanes over placements.  Every slot of the record's score row is
                        // written: the score inside the window, 0.0 outside it (the M-step relies on those zeros).
                        const int te = task - tM;
                        // block-major: the first 64 placements of every record (full two-chain tasks) go out before the records'
                        // last blocks (partly filled or empty), so the tasks a round ends on are the short ones
                        // (a record has a handful of blocks: counting them off is cheaper than an integer division)
                        int blk = 0, rq = te;
                        if (p.e_rec_major) {
                            rq = te / nblk;
                            blk = te - rq * nblk;
                        } else {
                            while (rq >= ne_q) { rq -= ne_q; blk++; }
                        }
                        const int rl = CLU ? rq * (int)csize + (int)crank : rq;
                        const FbRecMeta rm = SE.meta[rl];
                        const int fl2 = rm.flags, L2 = fl2 & 0xff, tb = rm.tb;
                        // placements i in [gapStart-L2+1, gapStart+gp) (figbird.cpp:1826) -> u in [D-L2+1, D+gp-1]
                        const int lo = max(rm.a_lo, D - L2 + 1), hi = min(rm.a_hi, D + gp - 1);
                        double* wrow = wE + kPadW + rl * WS;                  // slot k holds placement lo + k
                        const int ka = 64 * blk + lane, kb = ka + 32;
                        if (lo + 64 * blk > hi) {                              // nothing of the window in this block (or no window at all)
                            if (!CLU) {
                                if (ka < PS) wrow[ka] = 0.0;
                                if (kb < PS) wrow[kb] = 0.0;
                            } else {
                                for (unsigned k = 0; k < csize; k++) {
                                    double* wr = cg::this_cluster().map_shared_rank(wrow, k);
                                    if (ka < PS) wr[ka] = 0.0;
                                    if (kb < PS) wr[kb] = 0.0;
                                }
                            }
                            continue;
                        }
                        const bool rev = (fl2 >> 8) & 1, left = (fl2 >> 9) & 1;
                        const int ua = lo + ka, ub = ua + 32;
                        const bool va_ok = ua <= hi, vb_ok = ub <= hi;
                        const int uc = va_ok ? ua : lo;                      // invalid lanes shadow a valid placement
                        const int ta = left ? uc + tb : tb - uc;
                        const int tbb = vb_ok ? (left ? ub + tb : tb - ub) : ta;
                        double sa = s.ild[ta - p.thr_min], sb = s.ild[tbb - p.thr_min];   // f=1: the table holds log(ILD[t]) (host libm)
                        const unsigned char* pa = reinterpret_cast<const unsigned char*>(s.pq + (uc - R0 - (lo >= hl ? hn : 0)));
                        const uint32_t* cls4 = SE.cls4 + rl * 16;
                        const unsigned cstride = (unsigned)nrs * 16u;
                        const bool two = lo + 64 * blk + 32 <= hi;           // any lane with a second placement?
                        const int k0 = eo_start(rev, L2);
                        // Left-flank prefixes (FbPrefixParams): every lane of chain A has at least jsA terms against left-flank
                        // rows, chain B 32 fewer; where that is 4 or more the chain starts from the tabulated running score.
                        int skA = 0, skB = 0;
                        if (p.pre) {
                            const int jsA = D - (lo + 64 * blk + 31);
                            if (jsA >= 4) {
                                const double* pr = p.pre + ((size_t)((unsigned)fl2 >> FB_PRE_SHIFT) + 2 * (size_t)blk) * 32 + lane;
                                skA = jsA & ~3;
                                sa = __ldg(pr);
                                if (two && jsA >= 36) {
                                    skB = skA - 32;
                                    sb = __ldg(pr + 32);
                                }
                            }
                        }
                        // First iteration: bases [g0, g1) meet gap rows in every lane of the task — the same initial row everywhere.
                        int g0 = L2, g1 = L2;
                        if (uni0) {
                            const int u0 = lo + 64 * blk;
                            const int a = max(skA, (D - u0 + 3) & ~3), b = min(L2, D + gp - (two ? 63 : 31) - u0) & ~3;
                            if (b > a) { g0 = a; g1 = b; }
                        }
                        if (!two || skB < skA) {
                            // one chain alone: the whole of a one-chain task, or chain B's terms before chain A's first
                            const bool headB = two;
                            double s1 = headB ? sb : sa, s2 = 0.0;
                            const unsigned char* p1 = headB ? pa + 32 * 16 : pa;
                            int jb = headB ? skB : skA, je = headB ? skA : g0;
                            const int jend = headB ? skA : L2;
#pragma unroll 1
                            for (;;) {
                                score_pair<false, F1>(p1, cls4, cstride, p, k0, jb, je, s1, s2);
                                if (je == jend) break;
                                score_uniform<false, F1>(cls4, p, L2, rev, g0, g1, s1, s2);
                                jb = g1;
                                je = L2;
                            }
                            if (headB) sb = s1;
                            else sa = s1;
                        }
                        if (two) {
                            int jb = skA, je = g0;
#pragma unroll 1
                            for (;;) {
                                score_pair<true, F1>(pa, cls4, cstride, p, k0, jb, je, sa, sb);
                                if (je == L2) break;
                                score_uniform<true, F1>(cls4, p, L2, rev, g0, g1, sa, sb);
                                jb = g1;
                                je = L2;
                            }
                        }
                        // what the M-step adds to the counts: the score itself, pow(10, score) in f=1 (figbird.cpp:1878, 1879)
                        double wa = sa, wb = sb;
                        if (F1) {
                            wa = fbm_pow10(sa);
                            wb = two ? fbm_pow10(sb) : 0.0;
                        }
                        if (!CLU) {
                            if (ka < PS) wrow[ka] = va_ok ? wa : 0.0;
                            if (kb < PS) wrow[kb] = vb_ok ? wb : 0.0;
                        } else {
                            for (unsigned k = 0; k < csize; k++) {   // the M tasks of every CTA read this row from their own copy
                                double* wr = cg::this_cluster().map_shared_rank(wrow, k);
                                if (ka < PS) wr[ka] = va_ok ? wa : 0.0;
                                if (kb < PS) wr[kb] = vb_ok ? wb : 0.0;
                            }
                        }
                        // tempProb > maxProb with maxProb starting at 0 (-DBL_MAX in f=1): NaN and scores not above it never win;
                        // the maximum is over the scores as the reference compares them (log scores in f=1)
                        // maxProb only feeds placeReads' return value, and only the last iteration's is ever used (figbird.cpp:2472-2478)
                        if (last) {
                            unsigned long long ba = score_key<F1>(sa, va_ok), bb = score_key<F1>(sb, vb_ok);
                            ba = warp_max_u64(ba > bb ? ba : bb);
                            if (lane == 0 && ba > key_init<F1>()) atomicMax(&SE.rmax[rl], ba);
                        }
                    }
                }
                if (tid == 0) s_queue[(round + 1) & 1] = 0;   // nobody touches the other slot during this round
                if (n_l && full_stage) {                      // chunk cl has landed (the copies had the whole round)
                    mbar_wait(&s_bar[cl % 3], (bar_phase >> (cl % 3)) & 1u);
                    bar_phase ^= 1u << (cl % 3);
                }
                sync_all();
            }

            // chunk 0 of a kept staging gets its per-record maxima reset here, a barrier ahead of the next iteration's E tasks
            if (keep_staged && !last && tid < min(CH, nrec)) s_stage[0].rmax[tid] = key_init<F1>();
            // ---- computeProbsGap + computeErrorProbsGap on the gap rows (flank rows never change)
            for (int xg = tid; xg < (skip_m ? 0 : gpc); xg += NT) {
                if (CLU && (unsigned)(xg >> 5) % csize != crank) continue;   // the counts of this row live in another CTA
                const int xa = xg + (D + xg >= hl ? hn : 0);                  // gap row in the reference's numbering
                double c[5];
#pragma unroll
                for (int k = 0; k < 5; k++) c[k] = s.cnt[k * gps + xg];
                double P[4], Q[5];
                profile_row(c, sT, P, Q);
                const int xr = xg + lmax - 1;
                if (!CLU) {
#pragma unroll
                    for (int k = 0; k < 4; k++) s.pq[k * nrs + xr] = make_double2(P[k], Q[k]);
                    s.pq[4 * nrs + xr] = make_double2(0.0, Q[4]);
                } else {
                    for (unsigned q = 0; q < csize; q++) {
                        double2* pq = cg::this_cluster().map_shared_rank(s.pq, q);
#pragma unroll
                        for (int k = 0; k < 4; k++) pq[k * nrs + xr] = make_double2(P[k], Q[k]);
                        pq[4 * nrs + xr] = make_double2(0.0, Q[4]);
                    }
                }
                if (p.part_out) {
#pragma unroll
                    for (int k = 0; k < 5; k++) p.part_out[5 * (size_t)unit.cons_off + (size_t)k * gp + xa] = c[k];
                }
                if (last) {
                    // computeSequence(0), figbird.cpp:1991-2004: first strict maximum above 0, else index 0
                    double mx = 0.0;
                    int idx = 0;
#pragma unroll
                    for (int k = 0; k < 5; k++)
                        if (c[k] > mx) { mx = c[k]; idx = k; }
                    p.unit_cons[unit.cons_off + xa] = "ACGTN"[idx];
                } else {
#pragma unroll
                    for (int k = 0; k < 5; k++) s.cnt[k * gps + xg] = 0.0;   // placeReads zeroes the gap rows, :1708-1714
                }
            }
            if (last && !skip_m && (!CLU || crank == 0)) {
                // rows no placement touches keep all-zero counts: computeSequence(0) gives index 0 there
                for (int i = tid; i < hn; i += NT) p.unit_cons[unit.cons_off + (hl - D) + i] = 'A';
            }
            sync_all();
        }

        // ---- likelihood of the unit (figbird.cpp:1951-1955, last iteration): sum over the records, IN RECORD ORDER, of
        // glibc's log of the record's best score (f=0, only where it is above 0) or of the best log score itself (f=1).
        // One warp: the logs in parallel, the additions one after the other like the CPU's loop.
        if (p.unit_lik && (!CLU || crank == 0) && tid < 32) {
            if (CLU) __threadfence();
            unit_likelihood<F1>(p.unit_maxprob + unit.mp_off, nrec, p.unit_lik + ui);
        }

        // ---- parity tap: the whole profile of one unit, laid out like probsGap/countsGap
        if (p.dump_probs && ui == p.dump_unit) {
            const int rows = gp + 2 * D;
            for (int x = tid; x < rows; x += NT) {
                double c[5];
                if (x >= hl && x < hl + hn) {
#pragma unroll
                    for (int k = 0; k < 5; k++) c[k] = 0.0;      // zeroed by placeReads, never touched
                } else if (x >= D && x < D + gp) {
#pragma unroll
                    for (int k = 0; k < 5; k++) c[k] = s.cnt[k * gps + (x - D - (x >= hl ? hn : 0))];
                } else {
                    const int b = x < D ? fl[x] : fl[D + (x - D - gp)];
#pragma unroll
                    for (int k = 0; k < 5; k++) c[k] = (k == b) ? 1.0 : 0.0;
                }
                double P[4], Q[5];
                profile_row(c, sT, P, Q);
#pragma unroll
                for (int k = 0; k < 4; k++) p.dump_probs[(size_t)x * 5 + k] = P[k];
                p.dump_probs[(size_t)x * 5 + 4] = 0.0;
#pragma unroll
                for (int k = 0; k < 5; k++) p.dump_counts[(size_t)x * 5 + k] = c[k];
            }
        }
    }
}

// finalize(), figbird.cpp:2087-2393, all of it: score every placement of every record against the winning consensus
// (+ flanks) with the match/mismatch model and keep the best placement, lowest position on ties (2239-2291 / 2310-2359);
// a record whose best score passes the cut-off adds +1 to the count of each of its bases that falls into the gap
// (2293-2306 / 2361-2374); computeSequence(1) turns the counts into the filled sequence (1984-2047).
//   * The cut-off test -log(maxProb) <= gapProbCutOff is a threshold on maxProb: the host finds, with the reference's own
//     libm, the smallest double theta that passes (fb_context.cpp: cutoff_threshold) and the kernel compares
//     maxProb >= theta — the same decision for every double, no log on the device.
//   * The counts are integers, so their order of accumulation is free: shared-memory atomics per CTA.
//   * The records of a gap are independent: CTA (gap, slice) takes every gridDim.y-th group of records, one record per
//     warp, lanes over placements.  With one slice per gap the CTA finishes the gap itself; with several, each CTA adds
//     its counts to the gap's global counts and the last one to arrive (a ticket per gap) writes the sequence.
__global__ void __launch_bounds__(FB_FIN_THREADS) fb_finalize_kernel(const FbFinParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int g = p.gap_list[blockIdx.x];
    const FbGapMeta gm = p.gaps[g];
    const int D = p.D, gb = p.best_gp[g], lmax = gm.lmax;
    int* scnt = reinterpret_cast<int*>(smem_raw);                  // [gb][5] counts of this CTA
    uint8_t* gcode = smem_raw + sizeof(int) * 5 * (size_t)p.gb_stride;   // [gmax_stride] codes of the flank+best+flank window
    __shared__ __align__(16) uint8_t cs_all[FB_FIN_THREADS / 32][FB_MAX_READ];
    __shared__ double sT[25];
    __shared__ int s_disc, s_last;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int R0 = D - lmax + 1, NR = gb + 2 * lmax - 2;
    const uint8_t* fl = p.flank + (size_t)g * 2 * D;
    const long long boff = p.best_cons_off[g];        // < 0: no estimate won, bestString is empty (non-bases throughout)
    const char* best = p.unit_cons + (boff < 0 ? 0 : boff);
    if (tid < 25) sT[tid] = p.T[tid];
    if (tid == 0) { s_disc = 0; s_last = 0; }
    for (int i = tid; i < 5 * gb; i += FB_FIN_THREADS) scnt[i] = 0;
    for (int xr = tid; xr < NR; xr += FB_FIN_THREADS) {
        int x = xr + R0, c;
        if (x < D) c = fl[x];
        else if (x >= D + gb) c = fl[D + (x - D - gb)];
        else {
            const char ch = boff < 0 ? 'N' : best[x - D];
            c = ch == 'A' ? 0 : ch == 'C' ? 1 : ch == 'G' ? 2 : ch == 'T' ? 3 : 4;
        }
        gcode[xr] = (uint8_t)c;
    }
    __syncthreads();

    uint8_t* cs = cs_all[warp];
    const int nwarp = FB_FIN_THREADS / 32;
    int disc = 0;
    for (int r = blockIdx.y * nwarp + warp; r < gm.nrec; r += gridDim.y * nwarp) {
        const int rec = gm.rec_off + r;
        const FbRecMeta rm = p.rmeta[rec];
        const int L2 = rm.flags & 0xff;
        const bool rev = (rm.flags >> 8) & 1, left = (rm.flags >> 9) & 1;
        __syncwarp();
        load_record_bytes(p.codes, p.nmask, rec, L2, cs, lane);
        const int lo = max(rm.a_lo, D - L2 + 1), hi = min(rm.a_hi, D + gb - 1);
        double bv = 0.0;
        int bu = 0x7fffffff;
        for (int u = lo + lane; u <= hi; u += 32) {
            const int t = left ? u + rm.tb : rm.tb - u;
            double sc = p.ild[t - p.thr_min];
            for (int j = 0; j < L2; j++) {
                const int to = cs[j], from = gcode[u + j - R0];
                const int k = rev ? L2 - 1 - j : j;
                if (from == to) sc = dmul(sc, p.match[k]);                          // figbird.cpp:2272
                else sc = dmul(sc, dmul(p.epos[k], sT[from * 5 + to]));             // figbird.cpp:2276
            }
            if (sc > bv) { bv = sc; bu = u; }   // ascending u per lane: first maximum kept
        }
        // warp arg-max, lowest placement wins ties (figbird.cpp:2280: strict '>' in ascending i)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            double ov = __shfl_xor_sync(0xffffffffu, bv, o);
            int ou = __shfl_xor_sync(0xffffffffu, bu, o);
            if (ov > bv || (ov == bv && ou < bu)) { bv = ov; bu = ou; }
        }
        const int pos = bv > 0.0 ? bu : 0;       // maxPos stays 0 when nothing scored above 0 (figbird.cpp:2235)
        if (bv >= p.theta) {                     // -log(maxProb) <= gapProbCutOff, figbird.cpp:2293
            for (int j = lane; j < L2; j += 32) {
                const int x = pos + j;
                if (x >= D && x < D + gb) atomicAdd(&scnt[(x - D) * 5 + cs[j]], 1);
            }
        } else {
            disc++;                              // the same in every lane
        }
    }
    if (lane == 0 && disc) atomicAdd(&s_disc, disc);
    __syncthreads();

    const int* cnt = scnt;
    if (gridDim.y > 1) {
        int* gc = p.fin_counts + p.fin_counts_off[g];
        for (int i = tid; i < 5 * gb; i += FB_FIN_THREADS)
            if (scnt[i]) atomicAdd(&gc[i], scnt[i]);
        if (tid == 0 && s_disc) atomicAdd(&p.discarded[g], s_disc);
        __threadfence();
        __syncthreads();
        if (tid == 0) s_last = atomicAdd(&p.tickets[g], 1) == (int)gridDim.y - 1;
        __syncthreads();
        if (!s_last) return;
        __threadfence();
        cnt = gc;
    } else if (tid == 0) {
        p.discarded[g] = s_disc;
    }
    // computeSequence(1), figbird.cpp:1989-2045: first strict maximum above 0, 'N' where no read was counted
    char* out = p.filled + p.filled_off[g];
    for (int i = tid; i < gb; i += FB_FIN_THREADS) {
        int mx = 0, idx = -1;
#pragma unroll
        for (int q = 0; q < 5; q++) {
            const int v = gridDim.y > 1 ? __ldcg(cnt + i * 5 + q) : cnt[i * 5 + q];
            if (v > mx) { mx = v; idx = q; }
        }
        out[i] = idx < 0 ? 'N' : "ACGTN"[idx];
    }
}

// Closing step of a record-sharded scan.  Thread per gap row: the five counts summed over the ranks in rank order
// (the same order the next iteration would have used), computeSequence(0) on the sum (figbird.cpp:1991-2004); thread
// per record: the owner's per-record maximum (every other rank holds 0 there).
__global__ void __launch_bounds__(256) fb_shard_finish_kernel(const FbShardFinParams p) {
    if (p.abort && *reinterpret_cast<const volatile int*>(p.abort)) return;
    for (int ui = blockIdx.x; ui < p.n_units; ui += gridDim.x) {
        const FbUnit unit = p.units[ui];
        const int gp = unit.gp;
        for (int xg = threadIdx.x; xg < gp; xg += blockDim.x) {
            const size_t o = 5 * (size_t)unit.cons_off + xg;
            double c[5];
#pragma unroll
            for (int k = 0; k < 5; k++) c[k] = p.parts_in[0][o + (size_t)k * gp];
            for (int q = 1; q < p.n_parts; q++) {
#pragma unroll
                for (int k = 0; k < 5; k++) c[k] = dadd(c[k], p.parts_in[q][o + (size_t)k * gp]);
            }
            double mx = 0.0;
            int idx = 0;
#pragma unroll
            for (int k = 0; k < 5; k++) {
                p.sum_counts[o + (size_t)k * gp] = c[k];
                if (c[k] > mx) { mx = c[k]; idx = k; }
            }
            p.unit_cons[unit.cons_off + xg] = "ACGTN"[idx];
        }
        const int nrec = p.unit_nrec[ui];
        for (int r = threadIdx.x; r < nrec; r += blockDim.x) {
            double m = p.maxprob_in[0][unit.mp_off + r];
            for (int q = 1; q < p.n_parts; q++) m = fmax(m, p.maxprob_in[q][unit.mp_off + r]);
            p.maxprob_out[unit.mp_off + r] = m;
        }
    }
}

// Device-side barrier of the record-sharded scan.  Every rank owns an array of flags in its exchange memory; the peers
// store into it over NVLink.  signal: after this rank's launches of a phase (stream order), tell every peer; wait: hold
// the stream until every rank has reached `want`.  Tokens only grow.  A wait gives up after ~10^10 cycles; it then
// raises this rank's error word AND stores FB_FLAG_POISON into its slot of every peer's array, so that no peer goes on
// to add up this rank's stale partial counts: a wait that reads the poison raises its own error word and passes it on.
// Once the error word is set the EM launches and the closing reduction that follow in the stream return at once, and
// every later signal of the pass carries the poison instead of its token.
__device__ __forceinline__ void flag_store(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

__global__ void fb_flag_signal_kernel(FbFlagParams p) {
    const int q = threadIdx.x;
    if (q >= p.world) return;
    __threadfence_system();          // the partial counts written by the preceding kernels are visible before the flag
    const bool failed = *reinterpret_cast<volatile int*>(p.error) != 0;
    flag_store(p.peer_flags[q] + p.rank, failed ? FB_FLAG_POISON : p.token);
}

__global__ void fb_flag_wait_kernel(FbFlagParams p) {
    const int q = threadIdx.x;
    if (q >= p.world) return;
    const long long t0 = clock64();
    bool failed = false;
    for (;;) {
        unsigned long long v;
        asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p.own_flags + q) : "memory");
        if (v == FB_FLAG_POISON) { failed = true; break; }
        if (v >= p.token) break;
        if (clock64() - t0 > p.timeout_cycles) { failed = true; break; }
        __nanosleep(200);
    }
    if (failed) {
        atomicExch(p.error, 1);
        for (int k = 0; k < p.world; k++)
            if (k != p.rank) flag_store(p.peer_flags[k] + p.rank, FB_FLAG_POISON);
    }
}

__global__ void __launch_bounds__(128) fb_select_kernel(const FbSelParams p) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= p.n_gaps) return;
    const int e0 = p.sel_off[g], e1 = p.sel_off[g + 1];
    if (e1 <= e0) return;
    double best = -1.7976931348623157e308;      // maxLikelihood = -DBL_MAX, figbird.cpp:2458
    int be = p.gmin[g], bu = -1;
    for (int e = e0; e < e1; e++) {
        const int u = p.sel_unit[e];
        const double l = p.unit_lik[u];
        if (l >= best) { best = l; be = p.gmin[g] + (e - e0); bu = u; }
    }
    p.best_gp[g] = be;
    p.best_lik[g] = best;
    p.best_cons_off[g] = bu < 0 ? -1 : p.units[bu].cons_off;
    if (p.win_units && bu >= 0) {
        // second pass of a two-pass scan: the winner joins the work list of its size bin (any order: units are independent)
        int b = 0;
        while (b + 1 < p.n_bins && bu >= p.bin_first[b + 1]) b++;
        p.win_units[p.bin_first[b] + atomicAdd(&p.win_count[b], 1)] = p.units[bu];
    }
}

// Test tap: the device's log / pow(10, .) applied element-wise (tests compare the bits with the host libm).
__global__ void fb_devmath_kernel(const double* __restrict__ in, double* __restrict__ out, long n, int what) {
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x)
        out[i] = what == 0 ? fbm_log(in[i]) : fbm_pow10(in[i]);
}

}  // namespace

cudaError_t fb_launch_select(const FbSelParams& p, cudaStream_t s) {
    if (p.n_gaps < 1) return cudaSuccess;
    fb_select_kernel<<<(p.n_gaps + 127) / 128, 128, 0, s>>>(p);
    return cudaGetLastError();
}

cudaError_t fb_launch_devmath(const double* in, double* out, long n, int what, cudaStream_t s) {
    fb_devmath_kernel<<<1184, 256, 0, s>>>(in, out, n, what);
    return cudaGetLastError();
}

cudaError_t fb_launch_flag_signal(const FbFlagParams& p, cudaStream_t s) {
    fb_flag_signal_kernel<<<1, 32, 0, s>>>(p);
    return cudaGetLastError();
}

cudaError_t fb_launch_flag_wait(const FbFlagParams& p, cudaStream_t s) {
    fb_flag_wait_kernel<<<1, 32, 0, s>>>(p);
    return cudaGetLastError();
}

size_t fb_em_smem_bytes(int nr_stride, int gp_stride, int p_stride, int chunk, int pwin, bool global_tables) {
    return em_carve(nullptr, nullptr, nr_stride, gp_stride, p_stride, chunk, pwin, global_tables);
}

// The EM kernel's variants: tables in shared or global memory (GT), one CTA or a thread-block cluster per unit (CLU),
// 512 or 1024 threads per CTA (NT), linear or log-domain scores (F1).
using EmKernel = void (*)(const FbEmParams);
static EmKernel em_variant(bool gt, bool clu, int threads, bool f1) {
    if (clu && threads == 1024) return f1 ? fb_em_kernel<false, true, 1024, true> : fb_em_kernel<false, true, 1024, false>;
    if (clu) return f1 ? fb_em_kernel<false, true, 512, true> : fb_em_kernel<false, true, 512, false>;
    if (threads == 1024) {
        if (gt) return f1 ? fb_em_kernel<true, false, 1024, true> : fb_em_kernel<true, false, 1024, false>;
        return f1 ? fb_em_kernel<false, false, 1024, true> : fb_em_kernel<false, false, 1024, false>;
    }
    if (gt) return f1 ? fb_em_kernel<true, false, 512, true> : fb_em_kernel<true, false, 512, false>;
    return f1 ? fb_em_kernel<false, false, 512, true> : fb_em_kernel<false, false, 512, false>;
}

cudaError_t fb_kernels_init(int* max_smem_optin, int* sm_count) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    e = cudaDeviceGetAttribute(max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    if (e != cudaSuccess) return e;
    e = cudaDeviceGetAttribute(sm_count, cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) return e;
    for (int f1 = 0; f1 < 2; f1++) {
        for (int v = 0; v < 6; v++) {
            const bool gt = v == 3 || v == 4, clu = v == 2 || v == 5;
            const int threads = (v == 1 || v == 4 || v == 5) ? 1024 : 512;
            e = cudaFuncSetAttribute(em_variant(gt, clu, threads, f1 != 0), cudaFuncAttributeMaxDynamicSharedMemorySize, *max_smem_optin - 1024);
            if (e != cudaSuccess) return e;
        }
    }
    return cudaFuncSetAttribute(fb_finalize_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, *max_smem_optin - 8192);
}

// CTAs of the plain (non-cluster) EM kernel that fit one SM with `smem` bytes of dynamic shared memory: the static
// shared memory of the kernel and the per-CTA system reserve are the runtime's to count.
cudaError_t fb_em_ctas_per_sm(int threads, bool global_tables, bool f1, size_t smem, int* n) {
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(n, em_variant(global_tables, false, threads, f1), threads == 1024 ? 1024 : FB_THREADS, smem);
}

cudaError_t fb_launch_em(const FbEmParams& p, int grid, int threads, bool global_tables, bool f1, size_t smem, cudaStream_t s) {
    em_variant(global_tables, false, threads, f1)<<<grid, threads == 1024 ? 1024 : FB_THREADS, smem, s>>>(p);
    return cudaGetLastError();
}

static void cluster_config(cudaLaunchConfig_t* cfg, cudaLaunchAttribute* attr, int grid, int cluster, int threads, size_t smem, cudaStream_t s) {
    *cfg = cudaLaunchConfig_t{};
    cfg->gridDim = dim3((unsigned)grid, 1, 1);
    cfg->blockDim = dim3(threads == 1024 ? 1024 : FB_THREADS, 1, 1);
    cfg->dynamicSmemBytes = smem;
    cfg->stream = s;
    attr->id = cudaLaunchAttributeClusterDimension;
    attr->val.clusterDim.x = (unsigned)cluster;
    attr->val.clusterDim.y = 1;
    attr->val.clusterDim.z = 1;
    cfg->attrs = attr;
    cfg->numAttrs = 1;
}

cudaError_t fb_em_max_clusters(int cluster, int threads, bool f1, size_t smem, int* n) {
    cudaLaunchConfig_t cfg;
    cudaLaunchAttribute attr;
    cluster_config(&cfg, &attr, cluster, cluster, threads, smem, nullptr);
    return cudaOccupancyMaxActiveClusters(n, em_variant(false, true, threads, f1), &cfg);
}

cudaError_t fb_launch_em_cluster(const FbEmParams& p, int n_clusters, int cluster, int threads, bool f1, size_t smem, cudaStream_t s) {
    cudaLaunchConfig_t cfg;
    cudaLaunchAttribute attr;
    cluster_config(&cfg, &attr, n_clusters * cluster, cluster, threads, smem, s);
    return cudaLaunchKernelEx(&cfg, em_variant(false, true, threads, f1), p);
}

cudaError_t fb_launch_finalize(const FbFinParams& p, int n, int slices, size_t smem, cudaStream_t s) {
    fb_finalize_kernel<<<dim3((unsigned)n, (unsigned)slices, 1), FB_FIN_THREADS, smem, s>>>(p);
    return cudaGetLastError();
}

cudaError_t fb_launch_shard_finish(const FbShardFinParams& p, cudaStream_t s) {
    if (p.n_units < 1) return cudaSuccess;
    fb_shard_finish_kernel<<<p.n_units < 4096 ? p.n_units : 4096, 256, 0, s>>>(p);
    return cudaGetLastError();
}

cudaError_t fb_launch_t0(const double2* eo, const double* T, double* out, bool f1, cudaStream_t s) {
    if (f1) fb_t0_kernel<true><<<1, 128, 0, s>>>(eo, T, out);
    else fb_t0_kernel<false><<<1, 128, 0, s>>>(eo, T, out);
    return cudaGetLastError();
}

cudaError_t fb_launch_prefix(const FbPrefixParams& p, int slices, bool f1, cudaStream_t s) {
    if (p.n_gaps < 1) return cudaSuccess;
    const dim3 grid((unsigned)(p.n_gaps < 65535 ? p.n_gaps : 65535), (unsigned)(slices < 1 ? 1 : slices), 1);   // CTA (gap, slice): every slices-th group of 8 records
    if (f1) fb_prefix_kernel<true><<<grid, 256, 0, s>>>(p);
    else fb_prefix_kernel<false><<<grid, 256, 0, s>>>(p);
    return cudaGetLastError();
}

cudaError_t fb_launch_stage(const FbStageParams& p, cudaStream_t s) {
    if (p.n_records < 1) return cudaSuccess;
    fb_stage_kernel<<<(p.n_records + 7) / 8, 256, 0, s>>>(p);
    return cudaGetLastError();
}
