/* Test helper (compiled by tests/test_glibc_math.py with gcc -O2 -ffp-contract=off -mfma): compares the restatement of
 * glibc's log and pow(10, .) in gapprober_b200/csrc/fb_glibc_math.h — the very source the device build compiles — with
 * the host libm, bit for bit, and evaluates the host libm over arrays for the GPU test. */
#include "../gapprober_b200/csrc/fb_glibc_math.h"

#include <math.h>
#include <stdio.h>

static uint64_t rng_state;
static uint64_t rnd(void) {
    rng_state ^= rng_state << 13;
    rng_state ^= rng_state >> 7;
    rng_state ^= rng_state << 17;
    return rng_state;
}

/* kind-th family of inputs of log(): positive finite, around 1, subnormal, [0.5, 2), probabilities down to underflow,
 * the near-1 branch exactly, small rationals, 1 +- a few ulps */
double gm_log_input(long i) {
    uint64_t u = rnd();
    switch (i & 7) {
        case 0: u &= 0x7fffffffffffffffull; if ((u >> 52) == 0x7ff) u &= ~(1ull << 62); return fbm_f64(u);
        case 1: return fbm_f64(0x3ff0000000000000ull + (int64_t)(u >> 12) - (1ll << 51));
        case 2: return fbm_f64(u & 0x000fffffffffffffull);
        case 3: return fbm_f64(0x3fe0000000000000ull + (u >> 11));
        case 4: return exp(-(double)(u % 745)) * (1.0 + (double)(rnd() >> 12) * 0x1p-52);
        case 5: return fbm_f64(0x3fee000000000000ull + (u % 0x3090000000000ull));
        case 6: return (double)(u >> 40) * 0x1p-24;
        default: return fbm_f64(0x3ff0000000000000ull + (int64_t)(u % 2048) - 1024);
    }
}

/* exponents of pow(10, .): log scores over the whole range down to underflow, the subnormal results, far beyond, any bit
 * pattern (NaN, inf), positive, hundredths, |y| around 2^-65 */
double gm_pow_input(long i) {
    uint64_t u = rnd();
    switch (i & 7) {
        case 0: return -(double)(u >> 11) * 0x1p-53 * 330.0;
        case 1: return -(double)(u >> 11) * 0x1p-53 * 30.0;
        case 2: return -300.0 - (double)(u >> 11) * 0x1p-53 * 30.0;
        case 3: return -(double)(u >> 11) * 0x1p-53 * 2000.0;
        case 4: return fbm_f64(u);
        case 5: return (double)(u >> 11) * 0x1p-53 * 310.0;
        case 6: return -(double)(u % 100000) * 0.01;
        default: return -fbm_f64((u & 0x000fffffffffffffull) | ((uint64_t)(0x3be - 5 + (u >> 60)) << 52));
    }
}

static int same(double a, double b) { return fbm_u64(a) == fbm_u64(b) || (a != a && b != b); }

/* mismatches between the restatement and the host libm over n generated inputs; the first few are printed */
long gm_check(int what, long n, uint64_t seed) {
    long bad = 0;
    rng_state = seed ? seed : 88172645463325252ull;
    for (long i = 0; i < n; i++) {
        const double x = what == 0 ? gm_log_input(i) : gm_pow_input(i);
        const double a = what == 0 ? log(x) : pow(10.0, x);
        const double b = what == 0 ? fbm_log(x) : fbm_pow10(x);
        if (!same(a, b)) {
            if (bad < 5) fprintf(stderr, "%s(%a): libm %a, restatement %a\n", what == 0 ? "log" : "pow10", x, a, b);
            bad++;
        }
    }
    return bad;
}

/* edge inputs, explicitly */
long gm_check_edges(void) {
    const double lx[] = {0.0, 0x1p-1074, 0x1p-1022, 0x1.fffffffffffffp-1023, 1.0, 0x1.fffffffffffffp-1, 0x1.0000000000001p0, 0x1.ep-1,
                         0x1.109p0, 0x1.108ffffffffffp0, 2.0, 10.0, 0x1.fffffffffffffp1023, INFINITY, NAN, 0.25, 1e-300, 1e-310};
    const double py[] = {0.0, -0.0, -INFINITY, INFINITY, NAN, -1.0, -2.0, -307.0, -308.0, -323.0, -323.3, -324.0, -400.0, -1e10, 1.0, 308.0,
                         308.3, 309.0, 1e300, -1e300, 0x1p-66, -0x1p-66, 0x1p-64, -0x1p-64, -0.5, -1e-5, -0x1.34413509f79ffp8};
    long bad = 0;
    for (unsigned i = 0; i < sizeof lx / sizeof *lx; i++) bad += !same(log(lx[i]), fbm_log(lx[i]));
    for (unsigned i = 0; i < sizeof py / sizeof *py; i++) bad += !same(pow(10.0, py[i]), fbm_pow10(py[i]));
    return bad;
}

/* the generated inputs themselves, and the host libm over an array (for the GPU test) */
void gm_inputs(int what, double* x, long n, uint64_t seed) {
    rng_state = seed ? seed : 88172645463325252ull;
    for (long i = 0; i < n; i++) x[i] = what == 0 ? gm_log_input(i) : gm_pow_input(i);
}
void gm_libm(int what, const double* x, double* y, long n) {
    for (long i = 0; i < n; i++) y[i] = what == 0 ? log(x[i]) : pow(10.0, x[i]);
}
