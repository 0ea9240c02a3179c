// fb_glibc_math.h — log(x) and pow(10, y) with the bits of the host's glibc, as device (and host) functions.
//
// The reference calls the C library on the path: log() of every record's best score in f=0 mode (figbird.cpp:1951-1955),
// and in f=1 mode log() of every base term (figbird.cpp:1835, 1856, 1861) and pow(10, s) of every placement's log score
// (figbird.cpp:1878).  Those results feed the '>=' selection between estimates whose likelihoods often tie to the last
// bit (figbird.cpp:2479), so "close" is not enough: the functions below restate, operation by operation, what glibc 2.39
// computes on x86-64 hosts with FMA + AVX2 (the ifunc variants __log_fma / __pow_fma): the table-driven algorithms of
// sysdeps/ieee754/dbl-64/e_log.c, e_pow.c and the exp kernel inlined in it, WITH the fused multiply-adds that build
// contracted (read off its machine code, cited per line as "fma:" below), over the tables of that same libm
// (fb_glibc_tables.h, extracted by tools/extract_libm_tables.py).  tests/test_glibc_math.py compares these functions with
// the host libm bit for bit on the CPU (10^7 random and edge inputs), tests/test_gpu_glibc_math.py does the same for the
// device build.
//
// Only what the path can reach is restated: log of +0, positive (sub)normal numbers, +inf and NaN; pow with base 10 and
// any exponent.  Negative arguments of log cannot occur (scores are products of probabilities).
#pragma once
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define FBM_HD __device__ __forceinline__
#define FBM_TABLE static __device__ const
#else
#include <math.h>
#define FBM_HD static inline
#define FBM_TABLE static const
#endif

#if defined(__CUDA_ARCH__)
#define FBM_FMA(a, b, c) __fma_rn((a), (b), (c))
#define FBM_MUL(a, b) __dmul_rn((a), (b))
#define FBM_ADD(a, b) __dadd_rn((a), (b))
#define FBM_SUB(a, b) __dsub_rn((a), (b))
#define FBM_U64(x) ((uint64_t)__double_as_longlong(x))
#define FBM_F64(u) __longlong_as_double((long long)(u))
#else
// host build: compile with -ffp-contract=off so that only the fma() calls written here are fused
#define FBM_FMA(a, b, c) __builtin_fma((a), (b), (c))
#define FBM_MUL(a, b) ((a) * (b))
#define FBM_ADD(a, b) ((a) + (b))
#define FBM_SUB(a, b) ((a) - (b))
static inline uint64_t fbm_u64(double x) { uint64_t u; memcpy(&u, &x, 8); return u; }
static inline double fbm_f64(uint64_t u) { double x; memcpy(&x, &u, 8); return x; }
#define FBM_U64(x) fbm_u64(x)
#define FBM_F64(u) fbm_f64(u)
#endif

#include "fb_glibc_tables.h"

#if defined(__CUDA_ARCH__)
#define FBM_LD(p) __ldg(p)
#else
#define FBM_LD(p) (*(p))
#endif

// log(x), e_log.c (__log_fma).  x >= +0 or NaN.
FBM_HD double fbm_log(double x) {
    uint64_t ix = FBM_U64(x);
    const uint32_t top = (uint32_t)(ix >> 48);
    // |x - 1| small: dedicated polynomial (e_log.c "Handle close to 1.0 inputs separately")
    if (ix - 0x3fee000000000000ull < 0x3090000000000ull) {        // [1 - 0x1p-4, 1 + 0x1.09p-4)
        if (ix == 0x3ff0000000000000ull) return 0.0;
        const double* B = fbm_log_B;
        const double r = FBM_SUB(x, 1.0);
        const double r2 = FBM_MUL(r, r);
        const double r3 = FBM_MUL(r, r2);
        // fma: every Horner step of the three inner cubics and their nesting is fused
        const double q3 = FBM_FMA(r3, FBM_LD(B + 10), FBM_FMA(r2, FBM_LD(B + 9), FBM_FMA(r, FBM_LD(B + 8), FBM_LD(B + 7))));
        const double q2 = FBM_FMA(r2, FBM_LD(B + 6), FBM_FMA(r, FBM_LD(B + 5), FBM_LD(B + 4)));
        const double q1 = FBM_FMA(r2, FBM_LD(B + 3), FBM_FMA(r, FBM_LD(B + 2), FBM_LD(B + 1)));
        const double q = FBM_FMA(FBM_FMA(q3, r3, q2), r3, q1);
        // rhi = r + w - w with w = r * 2^27: fma: both uses of w are fused with the add / subtract around them
        const double t = FBM_FMA(r, 0x1p27, r);
        const double rhi = FBM_FMA(-0x1p27, r, t);
        const double rlo = FBM_SUB(r, rhi);
        const double s = FBM_MUL(rhi, rhi);
        const double b0 = FBM_LD(B + 0);                           // -0.5
        const double hi = FBM_FMA(s, b0, r);                       // fma: r + rhi*rhi*B[0]
        double lo = FBM_FMA(s, b0, FBM_SUB(r, hi));                // fma: (r - hi) + rhi*rhi*B[0]
        lo = FBM_FMA(FBM_MUL(b0, rlo), FBM_ADD(r, rhi), lo);       // fma: lo += B[0]*rlo * (rhi + r)
        const double y = FBM_FMA(q, r3, lo);                       // fma: y = r3*q + lo
        return FBM_ADD(hi, y);
    }
    if (top - 0x0010u >= 0x7ff0u - 0x0010u) {
        if (ix * 2 == 0) return FBM_F64(0xfff0000000000000ull);    // log(0) = -inf
        if (ix == 0x7ff0000000000000ull) return x;                 // log(inf) = inf
        if ((top & 0x8000u) || (top & 0x7ff0u) == 0x7ff0u) return FBM_F64(0x7ff8000000000000ull);   // negative or NaN
        ix = FBM_U64(FBM_MUL(x, 0x1p52));                          // subnormal: normalise
        ix -= 52ull << 52;
    }
    const uint64_t tmp = ix - 0x3fe6000000000000ull;
    const int i = (int)((tmp >> 45) & 127);
    const int k = (int)((int64_t)tmp >> 52);
    const uint64_t iz = ix - (tmp & (0xfffull << 52));
    const double invc = FBM_LD(fbm_log_tab + 2 * i), logc = FBM_LD(fbm_log_tab + 2 * i + 1);
    const double z = FBM_F64(iz);
    const double kd = (double)k;
    const double* A = fbm_log_A;
    const double r = FBM_FMA(z, invc, -1.0);
    const double w = FBM_FMA(kd, FBM_LN2HI, logc);                 // fma: kd*Ln2hi + logc
    const double hi = FBM_ADD(w, r);
    const double lo = FBM_FMA(kd, FBM_LN2LO, FBM_ADD(FBM_SUB(w, hi), r));   // fma: (w - hi + r) + kd*Ln2lo
    const double r2 = FBM_MUL(r, r);
    // fma: A1 + r*A2, A3 + r*A4, their combination over r2, lo + r2*A0, and the final r*r2*p + (...)
    const double p = FBM_FMA(FBM_FMA(r, FBM_LD(A + 4), FBM_LD(A + 3)), r2, FBM_FMA(r, FBM_LD(A + 2), FBM_LD(A + 1)));
    const double y = FBM_FMA(FBM_MUL(r, r2), p, FBM_FMA(r2, FBM_LD(A + 0), lo));
    return FBM_ADD(y, hi);
}

// log_inline of e_pow.c for a positive normal x: returns hi, *tail = lo with hi + lo ~ log(x) to ~68 bits.
FBM_HD double fbm_pow_log(double x, double* tail) {
    const uint64_t ix = FBM_U64(x);
    const uint64_t tmp = ix - 0x3fe6955500000000ull;
    const int i = (int)((tmp >> 45) & 127);
    const int k = (int)((int64_t)tmp >> 52);
    const uint64_t iz = ix - (tmp & (0xfffull << 52));
    const double z = FBM_F64(iz), kd = (double)k;
    const double invc = FBM_LD(fbm_pow_tab + 3 * i), logc = FBM_LD(fbm_pow_tab + 3 * i + 1), logctail = FBM_LD(fbm_pow_tab + 3 * i + 2);
    const double* A = fbm_pow_A;
    const double r = FBM_FMA(z, invc, -1.0);
    const double t1 = FBM_FMA(kd, FBM_LN2HI, logc);                // fma
    const double t2 = FBM_ADD(t1, r);
    const double lo1 = FBM_FMA(kd, FBM_LN2LO, logctail);           // fma
    const double lo2 = FBM_ADD(FBM_SUB(t1, t2), r);
    const double ar = FBM_MUL(FBM_LD(A + 0), r);
    const double ar2 = FBM_MUL(r, ar);
    const double ar3 = FBM_MUL(r, ar2);
    const double hi = FBM_ADD(t2, ar2);
    const double lo3 = FBM_FMA(ar, r, -ar2);
    const double lo4 = FBM_ADD(FBM_SUB(t2, hi), ar2);
    // fma: every step of A1 + r*A2 + ar2*(A3 + r*A4 + ar2*(A5 + r*A6)), and ar3*poly + (lo1+lo2+lo3+lo4)
    const double p56 = FBM_FMA(r, FBM_LD(A + 6), FBM_LD(A + 5));
    const double p34 = FBM_FMA(p56, ar2, FBM_FMA(r, FBM_LD(A + 4), FBM_LD(A + 3)));
    const double p12 = FBM_FMA(ar2, p34, FBM_FMA(r, FBM_LD(A + 2), FBM_LD(A + 1)));
    const double lo = FBM_FMA(ar3, p12, FBM_ADD(FBM_ADD(FBM_ADD(lo1, lo2), lo3), lo4));
    const double y = FBM_ADD(hi, lo);
    *tail = FBM_ADD(FBM_SUB(hi, y), lo);
    return y;
}

// exp_inline of e_pow.c (sign_bias = 0): exp(x + xtail).
FBM_HD double fbm_pow_exp(double x, double xtail) {
    uint32_t abstop = (uint32_t)(FBM_U64(x) >> 52) & 0x7ffu;
    if (abstop - 0x3c9u >= 0x3fu) {                                // |x| < 2^-54 or |x| >= 512
        if (abstop - 0x3c9u >= 0x80000000u) return FBM_ADD(1.0, x);
        if (abstop >= 0x409u)                                      // |x| >= 1024
            return (FBM_U64(x) >> 63) ? 0.0 : FBM_F64(0x7ff0000000000000ull);      // underflow / overflow
        abstop = 0;
    }
    const double kd0 = FBM_FMA(x, FBM_EXP_INVLN2N, FBM_EXP_SHIFT); // fma: z = InvLn2N*x is fused into z + Shift
    const uint64_t ki = FBM_U64(kd0);
    const double kd = FBM_SUB(kd0, FBM_EXP_SHIFT);
    double r = FBM_FMA(kd, FBM_EXP_NEGLN2LON, FBM_FMA(kd, FBM_EXP_NEGLN2HIN, x));   // fma: x + kd*NegLn2hiN + kd*NegLn2loN
    r = FBM_ADD(xtail, r);
    const unsigned idx = 2u * (unsigned)(ki & 127u);
    const double tail = FBM_F64(FBM_LD(fbm_exp_tab + idx));
    uint64_t sbits = FBM_LD(fbm_exp_tab + idx + 1) + (ki << 45);
    const double r2 = FBM_MUL(r, r);
    // fma: tmp = (tail + r) + r2*(C2 + r*C3) + r2*r2*(C4 + r*C5), each product-sum fused
    double tmp = FBM_FMA(FBM_FMA(r, FBM_EXP_C3, FBM_EXP_C2), r2, FBM_ADD(r, tail));
    tmp = FBM_FMA(FBM_FMA(r, FBM_EXP_C5, FBM_EXP_C4), FBM_MUL(r2, r2), tmp);
    if (abstop == 0) {
        // specialcase(): the result is huge or in the subnormal range
        if ((ki & 0x80000000ull) == 0) {
            sbits -= 1009ull << 52;
            const double scale = FBM_F64(sbits);
            return FBM_MUL(0x1p1009, FBM_FMA(tmp, scale, scale));
        }
        sbits += 1022ull << 52;
        const double scale = FBM_F64(sbits);
        const double st = FBM_MUL(scale, tmp);                     // not fused here: the product is reused below
        double y = FBM_ADD(scale, st);
        if ((FBM_U64(y) & 0x7fffffffffffffffull) < 0x3ff0000000000000ull) {        // fabs(y) < 1.0
            const double one = y < 0.0 ? -1.0 : 1.0;
            double lo = FBM_ADD(FBM_SUB(scale, y), st);
            const double hi = FBM_ADD(one, y);
            lo = FBM_ADD(FBM_ADD(FBM_SUB(one, hi), y), lo);
            y = FBM_SUB(FBM_ADD(lo, hi), one);
            if (y == 0.0) y = FBM_F64(sbits & 0x8000000000000000ull);
        }
        return FBM_MUL(0x1p-1022, y);
    }
    const double scale = FBM_F64(sbits);
    return FBM_FMA(tmp, scale, scale);                             // fma: scale + scale*tmp
}

// pow(10.0, y), e_pow.c (__pow_fma) with x = 10.
FBM_HD double fbm_pow10(double y) {
    const uint64_t iy = FBM_U64(y);
    const uint32_t topy = (uint32_t)(iy >> 52);
    if ((topy & 0x7ffu) - 0x3beu >= 0x80u) {                       // |y| < 2^-65, |y| >= 2^63, inf or NaN
        if (2 * iy - 1 >= 2 * 0x7ff0000000000000ull - 1) {         // zero, inf or NaN
            if (2 * iy == 0) return 1.0;
            if (2 * iy > 2 * 0x7ff0000000000000ull) return FBM_ADD(10.0, y);        // NaN
            return (iy >> 63) ? 0.0 : FBM_MUL(y, y);               // 10^-inf = 0, 10^inf = inf
        }
        if ((topy & 0x7ffu) < 0x3beu) return FBM_ADD(1.0, y);      // x > 1: 1 + y
        return (topy < 0x800u) ? FBM_F64(0x7ff0000000000000ull) : 0.0;             // overflow / underflow
    }
    double lo;
    const double hi = fbm_pow_log(10.0, &lo);
    const double ehi = FBM_MUL(y, hi);
    const double elo = FBM_FMA(y, lo, FBM_FMA(hi, y, -ehi));      // fma: y*lo + fma(y, hi, -ehi)
    return fbm_pow_exp(ehi, elo);
}
