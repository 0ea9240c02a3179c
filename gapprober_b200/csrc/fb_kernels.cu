// fb_kernels.cu — hand-written sm_100a kernels for the figbird EM hot path of GapProber.
//
// What the reference does per gap (figbird.cpp:2436-2510) and how it is mapped here:
//
//   for each gap-length estimate gp (independent)            -> one "unit"; persistent CTAs pull units
//     initialize(gp)                  figbird.cpp:1620-1675  -> profile rows built in shared memory
//     repeat num_iter times
//       placeReads                    figbird.cpp:1677-1965
//         E-step  (1826-1872/1889-1938): score = ILD[t] * prod_j (P*(1-e) + e*Q)
//                                                           -> warp per (record, 32 placements), lanes over
//                                                              placements, j sequential in each lane
//         M-step  (1874-1881/1940-1947): counts[row][base] += score
//                                                           -> ordered gather: lane per gap row, warp per
//                                                              (32 rows, base class); adds arrive in the
//                                                              reference's (record, placement) order
//       computeProbsGap/ErrorProbsGap figbird.cpp:1568-1608  -> fused row update, thread per row
//     computeSequence(0)              figbird.cpp:1984-2047  -> consensus string per unit
//   finalize(best)                    figbird.cpp:2087-2393  -> fb_finalize_kernel (scores + arg-max);
//                                                              log/cut-off/counting stay on the host
//
// Bit-exactness (hard parts 1-2): the CPU evaluates P*(1-e)+e*Q as two rounded products
// and one rounded sum and multiplies the running score in j order; every FP64 operation below is an
// explicit __dmul_rn/__dadd_rn/__ddiv_rn so no FMA contraction can change a bit, and the M-step adds
// scores to a cell in exactly the order the CPU's scatter loop does.  Likelihood plateaus therefore
// tie exactly as they do on the CPU.  No tensor cores: the path is not a contraction.
//
// Roofs: ~3 HBM bytes per likelihood vs ~1.6 KB of shared-memory reads and ~400 FP64 instructions, so
// shared-memory bandwidth (then the FP64 pipe) binds; DESIGN.md has the arithmetic.

#include "fb_device.h"

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

struct EmSmem {
    double2* eo;
    double* ild;
    double2* pq;
    double* cnt;
    double* w;
    int* rlo;
    int* rhi;
    int* rtb;
    int* rfl;
    unsigned long long* rmax;
    uint8_t* clsoff;
    uint8_t* codes;
    uint8_t* jl;
};

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
    a = take(sizeof(double2) * FB_MAX_READ); if (s) s->eo = (double2*)a;
    a = take(sizeof(double) * pwin); if (s) s->ild = (double*)a;
    if (!global_tables) {
        a = take(sizeof(double2) * 5 * (size_t)nr_stride); if (s) s->pq = (double2*)a;
        a = take(sizeof(double) * 5 * (size_t)gp_stride); if (s) s->cnt = (double*)a;
    }
    a = take(sizeof(double) * (size_t)chunk * p_stride); if (s) s->w = (double*)a;
    a = take(sizeof(unsigned long long) * chunk); if (s) s->rmax = (unsigned long long*)a;
    a = take(sizeof(int) * chunk); if (s) s->rlo = (int*)a;
    a = take(sizeof(int) * chunk); if (s) s->rhi = (int*)a;
    a = take(sizeof(int) * chunk); if (s) s->rtb = (int*)a;
    a = take(sizeof(int) * chunk); if (s) s->rfl = (int*)a;
    a = take((size_t)8 * chunk); if (s) s->clsoff = a;
    a = take((size_t)FB_MAX_READ * chunk); if (s) s->codes = a;
    a = take((size_t)FB_MAX_READ * chunk); if (s) s->jl = a;
    return off;
}

// Unpack one record's bases into byte codes (0-3 ACGT, 4 other, 5 past the end) and build, per base
// class, the list of read positions holding it in DESCENDING order (so that placements are visited in
// ascending order by the M-step).  One warp per record.
__device__ __forceinline__ void load_record(const uint4* __restrict__ codes, const uint4* __restrict__ nmask, int rec, int L2,
                                            uint8_t* cs, uint8_t* jl, uint8_t* clsoff, int lane) {
    // 128-bit loads: lanes 0,1 fetch the 2-bit codes, lane 2 the non-ACGT mask
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
    unsigned packed = 0;
#pragma unroll
    for (int q = 0; q < 4; q++) {
        int j = lane * 4 + q;
        unsigned c = (cw >> (2 * (j & 15))) & 3u;
        if ((mw >> (j & 31)) & 1u) c = 4;
        if (j >= L2) c = 5;
        packed |= c << (8 * q);
    }
    reinterpret_cast<uint32_t*>(cs)[lane] = packed;
    __syncwarp();
    int base = 0;
#pragma unroll 1
    for (int c = 0; c < 5; c++) {
        if (lane == 0) clsoff[c] = (uint8_t)base;
#pragma unroll
        for (int b = FB_MAX_READ / 32 - 1; b >= 0; b--) {
            int j = b * 32 + lane;
            bool has = cs[j] == c;
            unsigned m = __ballot_sync(0xffffffffu, has);
            if (has) jl[base + __popc(m >> lane) - 1] = (uint8_t)j;   // rank among set bits at or above this lane
            base += __popc(m);
        }
    }
    if (lane == 0) clsoff[5] = (uint8_t)base;
}

template <bool GT>
__global__ void __launch_bounds__(FB_THREADS, GT ? 1 : 2) fb_em_kernel(const FbEmParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    EmSmem s;
    const int pwin = p.thr_max - p.thr_min + 1;
    em_carve(smem_raw, &s, p.nr_stride, p.gp_stride, p.p_stride, p.chunk, pwin, GT);
    if (GT) {
        double* g = p.gws + (size_t)blockIdx.x * p.gws_stride;
        s.pq = reinterpret_cast<double2*>(g);
        s.cnt = g + (size_t)10 * p.nr_stride;
    }
    __shared__ int s_unit;
    __shared__ double sT[25];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NW = FB_THREADS / 32;
    const int D = p.D;

    for (int i = tid; i < p.max_read_len && i < FB_MAX_READ; i += FB_THREADS) s.eo[i] = p.eo[i];
    for (int i = tid; i < pwin; i += FB_THREADS) s.ild[i] = p.ild[i];
    if (tid < 25) sT[tid] = p.T[tid];

    for (;;) {
        __syncthreads();
        if (tid == 0) s_unit = atomicAdd(p.work_counter, 1);
        __syncthreads();
        const int ui = s_unit;
        if (ui >= p.n_units) break;
        const FbUnit unit = p.units[ui];
        const FbGapMeta gm = p.gaps[unit.gap];
        const int gp = unit.gp, nrec = gm.nrec, lmax = gm.lmax;
        const int R0 = D - lmax + 1;                 // first profile row any placement can touch
        const int NR = gp + 2 * lmax - 2;            // rows [R0, R0+NR) cover every touched row
        const int nrs = p.nr_stride, gps = p.gp_stride;
        const uint8_t* fl = p.flank + (size_t)unit.gap * 2 * D;

        // ---- initialize(gp): flanks one-hot from the scaffold, gap rows N=1 (figbird.cpp:1657-1673)
        for (int xr = tid; xr < NR; xr += FB_THREADS) {
            int x = xr + R0;
            int b = 4;
            if (x < D) b = fl[x];
            else if (x >= D + gp) b = fl[D + (x - D - gp)];
            double c[5];
#pragma unroll
            for (int k = 0; k < 5; k++) c[k] = (k == b) ? 1.0 : 0.0;
            double P[4], Q[5];
            profile_row(c, sT, P, Q);
#pragma unroll
            for (int k = 0; k < 4; k++) s.pq[k * nrs + xr] = make_double2(P[k], Q[k]);
            s.pq[4 * nrs + xr] = make_double2(0.0, Q[4]);
        }
        for (int i = tid; i < gp; i += FB_THREADS) {
#pragma unroll
            for (int k = 0; k < 5; k++) s.cnt[k * gps + i] = 0.0;
        }
        __syncthreads();

        for (int it = 0; it < p.num_iter; it++) {
            const bool last = it == p.num_iter - 1;
            for (int c0 = 0; c0 < nrec; c0 += p.chunk) {
                const int nch = min(p.chunk, nrec - c0);
                // ---- stage this round's records: unpack bases, class lists, placement window
                for (int rl = warp; rl < nch; rl += NW) {
                    const int rec = gm.rec_off + c0 + rl;
                    const FbRecMeta rm = p.rmeta[rec];
                    const int L2 = rm.flags & 0xff;
                    load_record(p.codes, p.nmask, rec, L2, s.codes + rl * FB_MAX_READ, s.jl + rl * FB_MAX_READ, s.clsoff + rl * 8, lane);
                    if (lane == 0) {
                        // placements i in [gapStart-L2+1, gapStart+gp) (figbird.cpp:1826) -> u in [D-L2+1, D+gp-1]
                        s.rlo[rl] = max(rm.a_lo, D - L2 + 1);
                        s.rhi[rl] = min(rm.a_hi, D + gp - 1);
                        s.rtb[rl] = rm.tb;
                        s.rfl[rl] = rm.flags;
                        s.rmax[rl] = 0ull;
                    }
                }
                __syncthreads();

                // ---- E-step: warp per (record, group of 32 placements); lanes over placements
                const int gmaxg = (p.p_stride + 31) >> 5;
                for (int task = warp; task < nch * gmaxg; task += NW) {
                    const int rl = task / gmaxg, g = task - rl * gmaxg;
                    const int lo = s.rlo[rl], hi = s.rhi[rl];
                    if (lo + 32 * g > hi) continue;
                    const int fl2 = s.rfl[rl], L2 = fl2 & 0xff;
                    const bool rev = (fl2 >> 8) & 1, left = (fl2 >> 9) & 1;
                    const int u = lo + 32 * g + lane;
                    const bool valid = u <= hi;
                    const int uu = valid ? u : lo;
                    const int t = left ? uu + s.rtb[rl] : s.rtb[rl] - uu;
                    double sc = s.ild[t - p.thr_min];
                    const double2* pq = s.pq + (uu - R0);
                    const uint32_t* cw = reinterpret_cast<const uint32_t*>(s.codes + rl * FB_MAX_READ);
                    const double2* ep = rev ? s.eo + (L2 - 1) : s.eo;
                    const int es = rev ? -1 : 1;
                    for (int j0 = 0; j0 < L2; j0 += 4) {
                        const uint32_t c4 = cw[j0 >> 2];
#pragma unroll
                        for (int q = 0; q < 4; q++) {
                            const int j = j0 + q;
                            if (j < L2) {
                                const int c = (c4 >> (8 * q)) & 0xff;
                                const double2 v = pq[c * nrs + j];     // {P[row][c], Q[row][c]}; class 4 holds {0, Q[row][4]}
                                const double2 e = ep[j * es];          // {1 - e[k], e[k]}
                                const double aa = dadd(dmul(v.x, e.x), dmul(e.y, v.y));   // figbird.cpp:1855 / 1862
                                sc = dmul(sc, aa);
                            }
                        }
                    }
                    if (valid) s.w[rl * p.p_stride + (u - lo)] = sc;
                    // tempProb > maxProb with maxProb starting at 0: NaN and non-positive scores never win
                    unsigned long long bits = (valid && sc > 0.0) ? (unsigned long long)__double_as_longlong(sc) : 0ull;
                    bits = warp_max_u64(bits);
                    if (lane == 0 && bits) atomicMax(&s.rmax[rl], bits);
                }
                __syncthreads();

                // ---- M-step: lane per gap row, warp per (32 rows, base class).  Within a cell the CPU adds in
                // (record asc, placement asc) order; walking a class list in descending j is ascending placement.
                const int nrg = (gp + 31) >> 5;
                for (int task = warp; task < nrg * 5; task += NW) {
                    const int c = task / nrg, rg = task - c * nrg;
                    const int xg = rg * 32 + lane;
                    const int x = D + xg, x0 = D + rg * 32;
                    const bool rowok = xg < gp;
                    double acc = rowok ? s.cnt[c * gps + xg] : 0.0;
                    for (int rl = 0; rl < nch; rl++) {
                        const int lo = s.rlo[rl], hi = s.rhi[rl];
                        const int L2 = s.rfl[rl] & 0xff;
                        if (hi < lo || x0 + 31 < lo || x0 > hi + L2 - 1) continue;
                        const uint8_t* jl = s.jl + rl * FB_MAX_READ;
                        const int qb = s.clsoff[rl * 8 + c], qe = s.clsoff[rl * 8 + c + 1];
                        const double* wr = s.w + rl * p.p_stride - lo;
                        for (int q = qb; q < qe; q++) {
                            const int j = jl[q];
                            if (j > x0 + 31 - lo) continue;   // every lane would sit left of the window
                            if (j < x0 - hi) break;           // ... or right of it, and j only decreases
                            const int u = x - j;
                            if (rowok && u >= lo && u <= hi) acc = dadd(acc, wr[u]);   // figbird.cpp:1879
                        }
                    }
                    if (rowok) s.cnt[c * gps + xg] = acc;
                }
                if (last && tid < nch)
                    p.unit_maxprob[unit.mp_off + c0 + tid] = __longlong_as_double((long long)s.rmax[tid]);
                __syncthreads();
            }

            // ---- computeProbsGap + computeErrorProbsGap on the gap rows (flank rows never change)
            for (int xg = tid; xg < gp; xg += FB_THREADS) {
                double c[5];
#pragma unroll
                for (int k = 0; k < 5; k++) c[k] = s.cnt[k * gps + xg];
                double P[4], Q[5];
                profile_row(c, sT, P, Q);
                const int xr = xg + lmax - 1;
#pragma unroll
                for (int k = 0; k < 4; k++) s.pq[k * nrs + xr] = make_double2(P[k], Q[k]);
                s.pq[4 * nrs + xr] = make_double2(0.0, Q[4]);
                if (last) {
                    // computeSequence(0), figbird.cpp:1991-2004: first strict maximum above 0, else index 0
                    double mx = 0.0;
                    int idx = 0;
#pragma unroll
                    for (int k = 0; k < 5; k++)
                        if (c[k] > mx) { mx = c[k]; idx = k; }
                    p.unit_cons[unit.cons_off + xg] = "ACGTN"[idx];
                } else {
#pragma unroll
                    for (int k = 0; k < 5; k++) s.cnt[k * gps + xg] = 0.0;   // placeReads zeroes the gap rows, :1708-1714
                }
            }
            __syncthreads();
        }

        // ---- parity tap: the whole profile of one unit, laid out like probsGap/countsGap
        if (p.dump_probs && ui == p.dump_unit) {
            const int rows = gp + 2 * D;
            for (int x = tid; x < rows; x += FB_THREADS) {
                double c[5];
                if (x >= D && x < D + gp) {
#pragma unroll
                    for (int k = 0; k < 5; k++) c[k] = s.cnt[k * gps + (x - D)];
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

// finalize(), figbird.cpp:2239-2291 / 2310-2359: score every placement of every record against the
// winning consensus (+ flanks) with the match/mismatch model and keep the best placement, lowest
// position on ties.  One CTA per gap.  The cut-off test (-log(maxProb) <= gapProbCutOff), the +1
// counting and computeSequence(1) are done by the host from (maxProb, maxPos).
__global__ void __launch_bounds__(FB_FIN_THREADS) fb_finalize_kernel(const FbFinParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint8_t* gcode = smem_raw;                                     // [gmax_stride] codes of flank+best+flank window
    __shared__ uint8_t cs[FB_MAX_READ];
    __shared__ uint8_t jl_unused[FB_MAX_READ];
    __shared__ uint8_t clsoff_unused[8];
    __shared__ double sT[25];
    __shared__ double red_v[FB_FIN_THREADS / 32];
    __shared__ int red_u[FB_FIN_THREADS / 32];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = p.gap_list[blockIdx.x];
    const FbGapMeta gm = p.gaps[g];
    const int D = p.D, gb = p.best_gp[g], lmax = gm.lmax;
    const int R0 = D - lmax + 1, NR = gb + 2 * lmax - 2;
    const uint8_t* fl = p.flank + (size_t)g * 2 * D;
    const char* best = p.unit_cons + p.best_cons_off[g];
    if (tid < 25) sT[tid] = p.T[tid];
    for (int xr = tid; xr < NR; xr += FB_FIN_THREADS) {
        int x = xr + R0, c;
        if (x < D) c = fl[x];
        else if (x >= D + gb) c = fl[D + (x - D - gb)];
        else {
            char ch = best[x - D];
            c = ch == 'A' ? 0 : ch == 'C' ? 1 : ch == 'G' ? 2 : ch == 'T' ? 3 : 4;
        }
        gcode[xr] = (uint8_t)c;
    }
    __syncthreads();

    for (int r = 0; r < gm.nrec; r++) {
        const int rec = gm.rec_off + r;
        const FbRecMeta rm = p.rmeta[rec];
        const int L2 = rm.flags & 0xff;
        const bool rev = (rm.flags >> 8) & 1, left = (rm.flags >> 9) & 1;
        if (warp == 0) load_record(p.codes, p.nmask, rec, L2, cs, jl_unused, clsoff_unused, lane);
        __syncthreads();
        const int lo = max(rm.a_lo, D - L2 + 1), hi = min(rm.a_hi, D + gb - 1);
        double bv = 0.0;
        int bu = 0x7fffffff;
        for (int u = lo + tid; u <= hi; u += FB_FIN_THREADS) {
            const int t = left ? u + rm.tb : rm.tb - u;
            double sc = p.ild[t - p.thr_min];
            for (int j = 0; j < L2; j++) {
                const int to = cs[j], from = gcode[u + j - R0];
                const int k = rev ? L2 - 1 - j : j;
                if (from == to) sc = dmul(sc, p.match[k]);                          // figbird.cpp:2272
                else sc = dmul(sc, dmul(p.epos[k], sT[from * 5 + to]));             // figbird.cpp:2276
            }
            if (sc > bv) { bv = sc; bu = u; }   // ascending u per thread: first maximum kept
        }
        // block arg-max, lowest placement wins ties (figbird.cpp:2280: strict '>' in ascending i)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            double ov = __shfl_xor_sync(0xffffffffu, bv, o);
            int ou = __shfl_xor_sync(0xffffffffu, bu, o);
            if (ov > bv || (ov == bv && ou < bu)) { bv = ov; bu = ou; }
        }
        if (lane == 0) { red_v[warp] = bv; red_u[warp] = bu; }
        __syncthreads();
        if (tid == 0) {
            double v = 0.0;
            int uu = 0x7fffffff;
            for (int w = 0; w < FB_FIN_THREADS / 32; w++)
                if (red_v[w] > v || (red_v[w] == v && red_u[w] < uu)) { v = red_v[w]; uu = red_u[w]; }
            p.fin_maxprob[rec] = v;
            p.fin_maxpos[rec] = v > 0.0 ? uu : 0;   // maxPos stays 0 when nothing scored above 0 (figbird.cpp:2235)
        }
        __syncthreads();
    }
}

}  // namespace

size_t fb_em_smem_bytes(int nr_stride, int gp_stride, int p_stride, int chunk, int pwin, bool global_tables) {
    return em_carve(nullptr, nullptr, nr_stride, gp_stride, p_stride, chunk, pwin, global_tables);
}

cudaError_t fb_kernels_init(int* max_smem_optin, int* sm_count) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    e = cudaDeviceGetAttribute(max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    if (e != cudaSuccess) return e;
    e = cudaDeviceGetAttribute(sm_count, cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(fb_em_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, *max_smem_optin - 1024);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(fb_em_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, *max_smem_optin - 1024);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(fb_finalize_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, *max_smem_optin - 8192);
}

cudaError_t fb_launch_em(const FbEmParams& p, int grid, bool global_tables, size_t smem, cudaStream_t s) {
    if (global_tables) fb_em_kernel<true><<<grid, FB_THREADS, smem, s>>>(p);
    else fb_em_kernel<false><<<grid, FB_THREADS, smem, s>>>(p);
    return cudaGetLastError();
}

cudaError_t fb_launch_finalize(const FbFinParams& p, int n, size_t smem, cudaStream_t s) {
    fb_finalize_kernel<<<n, FB_FIN_THREADS, smem, s>>>(p);
    return cudaGetLastError();
}
