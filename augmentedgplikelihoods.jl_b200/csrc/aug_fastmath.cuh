// aug_fastmath.cuh — branch-free fp64 elementary functions for the map/reduce kernels.
//
// Why: ncu on the first fused CAVI kernel (profiles/r1a) showed 212 issued instructions per
// observation of which only ~82 were FP64 math; the rest were two UMOVs per polynomial coefficient,
// slow-path guards and exponent fiddling inside the libdevice exp / log1p / sqrt / div, and the
// branches stopped the compiler from interleaving the independent observations of a thread.  The
// kernel was issue-bound at 67% of the HBM roofline.  These versions
//   * read their coefficients from __constant__ memory (LDCU.128 -> uniform registers, hoistable),
//   * have no branch at all: they are only valid on a stated argument range, and the kernels test
//     the range of a whole thread-iteration once and re-evaluate it with the IEEE / libdevice
//     code (the SAFE instantiation of the same formulas) in the rare out-of-range case, so the
//     compiler can interleave the independent observations of a thread,
//   * seed reciprocals / reciprocal square roots with MUFU.RCP64H / MUFU.RSQ64H and finish with one
//     cubically convergent correction step.
// Accuracy (validated on the GPU against libdevice, tests/test_gpu_fastmath.py): <= 2 ulp for rcp,
// rsqrt, exp; <= 2e-16 absolute + 2 ulp for log.  The parity bar of the path is 1e-12 relative.
#pragma once
#include <math.h>

namespace augf {

static __constant__ double EXP_C[12] = {
    // Taylor coefficients 1/12! ... 1/1! of e^r, |r| <= ln2/2 (truncation 1.7e-16 relative)
    1.0 / 479001600.0, 1.0 / 39916800.0, 1.0 / 3628800.0, 1.0 / 362880.0, 1.0 / 40320.0, 1.0 / 5040.0,
    1.0 / 720.0,       1.0 / 120.0,      1.0 / 24.0,      1.0 / 6.0,      0.5,           1.0};
static __constant__ double LOG_C[10] = {
    // atanh series 2/(2k+1), k = 9 ... 1, for log m = 2 atanh(s), s = (m-1)/(m+1), |s| <= 0.1716
    2.0 / 19.0, 2.0 / 17.0, 2.0 / 15.0, 2.0 / 13.0, 2.0 / 11.0, 2.0 / 9.0, 2.0 / 7.0, 2.0 / 5.0, 2.0 / 3.0, 0.0};

constexpr double LN2_HI = 6.93147180369123816490e-01;   // 33 significant bits
constexpr double LN2_LO = 1.90821492927058770002e-10;
constexpr double LOG2E = 1.44269504088896338700e+00;
constexpr double MAGIC = 6755399441055744.0;             // 1.5 * 2^52: round-to-nearest-integer shifter
constexpr double SQRT2 = 1.41421356237309514547e+00;

// 1/x for normal positive x (no zero / inf / denormal handling)
__device__ __forceinline__ double rcp(double x) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double e = fma(-x, y, 1.0);
    const double p = fma(e, e, e);      // e + e^2
    return fma(y, p, y);                // y (1 + e + e^2): relative error e^3
}

// 1/sqrt(x) for normal positive x
__device__ __forceinline__ double rsqrt_(double x) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double t = x * y;
    const double e = fma(-t, y, 1.0);               // 1 - x y^2
    const double p = fma(e, 0.375, 0.5) * e;        // e/2 + 3e^2/8
    return fma(y, p, y);                            // relative error O(e^3)
}

// exp(a) for -708 <= a <= 708 (the caller guarantees the range; no guard, no branch)
__device__ __forceinline__ double exp_(double a) {
    const double t = fma(a, LOG2E, MAGIC);
    const int n = __double2loint(t);
    const double fn = t - MAGIC;
    double r = fma(fn, -LN2_HI, a);
    r = fma(fn, -LN2_LO, r);
    // e^r = 1 + r + r^2 q(r), q = sum_{k=2..12} r^(k-2)/k! split into its even and odd halves: two Horner chains
    // in r^2 that issue side by side (dependent depth 9 instead of 13; ncu showed the kernels stalled on
    // fixed-latency dependencies, not on the FP64 pipe)
    const double r2 = r * r;
    double qe = EXP_C[0];            // 1/12!
    double qo = EXP_C[1];            // 1/11!
    qe = fma(qe, r2, EXP_C[2]);      // 1/10!
    qo = fma(qo, r2, EXP_C[3]);      // 1/9!
    qe = fma(qe, r2, EXP_C[4]);      // 1/8!
    qo = fma(qo, r2, EXP_C[5]);      // 1/7!
    qe = fma(qe, r2, EXP_C[6]);      // 1/6!
    qo = fma(qo, r2, EXP_C[7]);      // 1/5!
    qe = fma(qe, r2, EXP_C[8]);      // 1/4!
    qo = fma(qo, r2, EXP_C[9]);      // 1/3!
    qe = fma(qe, r2, EXP_C[10]);     // 1/2!
    const double q = fma(r, qo, qe);
    double p = fma(r2, q, r) + 1.0;
    // p in [0.70, 1.42]; scale by 2^n through the exponent field (result stays normal for |a| <= 708)
    return __hiloint2double(__double2hiint(p) + n * 1048576, __double2loint(p));
}

// log(x) for normal positive finite x (the caller guarantees the range)
__device__ __forceinline__ double log_(double x) {
    int hi = __double2hiint(x);
    int k = (hi >> 20) - 1023;
    hi = (hi & 0x000fffff) | 0x3ff00000;
    double m = __hiloint2double(hi, __double2loint(x));   // [1, 2)
    const bool big = m > SQRT2;
    m = big ? 0.5 * m : m;                                 // [0.7071, 1.4142]
    k += big ? 1 : 0;
    const double s = (m - 1.0) * rcp(m + 1.0);
    const double s2 = s * s;
    double p = LOG_C[0];
#pragma unroll
    for (int j = 1; j < 9; ++j) p = fma(p, s2, LOG_C[j]);
    const double fk = (double)k;
    // log m = 2s + s^3 p(s^2);  log x = k ln2 + log m
    double res = fma(s * s2, p, fk * LN2_LO);
    res = fma(2.0, s, res);
    return fma(fk, LN2_HI, res);
}

// log(d) for d in [1, 2] (d = 1 + exp(-c)): no exponent extraction, no guards
__device__ __forceinline__ double log_1to2(double d) {
    const bool big = d > SQRT2;
    const double m = big ? 0.5 * d : d;
    const double s = (m - 1.0) * rcp(m + 1.0);
    const double s2 = s * s;
    double p = LOG_C[0];
#pragma unroll
    for (int j = 1; j < 9; ++j) p = fma(p, s2, LOG_C[j]);
    double res = fma(s * s2, p, big ? LN2_LO : 0.0);
    res = fma(2.0, s, res);
    return big ? res + LN2_HI : res;
}

// c = sqrt(s2) and 1/c in one go (MUFU.RSQ64H + one correction) for 1e-290 <= s2 <= 1e290
__device__ __forceinline__ void sqrt_inv(double s2, double& c, double& inv_c) {
    inv_c = rsqrt_(s2);
    c = s2 * inv_c;
}

// range in which every function above is valid for a second moment / squared residual
__device__ __forceinline__ bool in_range(double s2) { return s2 >= 1e-290 && s2 <= 1e290; }

}  // namespace augf
