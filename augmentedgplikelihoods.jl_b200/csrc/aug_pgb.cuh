// aug_pgb.cuh — PG(b, c) for general b: the building blocks of the warp-compacted sampler pgb_kernel (aug_gibbs.cu).
//
// Law sampled: rand(PolyaGamma(b, c)), SpecialDistributions/polyagamma.jl:112-164.  Three exact or certified pieces:
//
//  (1) integer part, b <= PGB_BX: a sum of Devroye PG(1, c) draws — the reference's own method (draw_sum :129-134,
//      sample_pg1 :225-257), exact.  Arithmetic as in aug_pg.cuh (normalised series, squeezes, r(z) table).
//
//  (2) fractional part e = b - floor(b) in (0, 1), b <= PGB_BX: EXACT rejection sampler (the reference truncates an
//      infinite Gamma convolution at 200 terms here, rand_gamma_sum :157-164, biased low by e/(2 pi^2 200)).
//      With J*(e, z) = 4 PG(e, 2z), the density of J*(e, 0) is the alternating series the reference's logpdf uses
//      (:37-53)   f(x) = 2^e/Gamma(e) sum_n (-1)^n Gamma(n+e)/n! (2n+e)/sqrt(2 pi x^3) exp(-(2n+e)^2/(2x));
//      its n = 0 term a_0(x) = 2^e * Levy(e^2)(x), tilted by exp(-z^2 x/2), is exp(-e z) 2^e times the
//      InverseGaussian(mean e/z, shape e^2) density, and
//          R(x) = f(x)/a_0(x) = sum_n (-1)^n c_n q^{n(n+e)},  q = exp(-2/x),  c_n = Gamma(n+e)(2n+e)/(Gamma(e+1) n!)
//      lies in [0, 1] and decreases from 1 to 0 (for e = 1 it is Jacobi's product prod (1 - q^{2m})^3; checked to
//      100 digits on a grid of (e, x) in tests/test_pg_laws_cpu.py).  So: X ~ IG(e/z, e^2), accept with probability
//      R(X); acceptance rate (1 + exp(-2z))^{-e} >= 2^{-e} > 1/2.  The terms t_n = c_n q^{n(n+e)} are unimodal in n
//      (the ratio t_{n+1}/t_n decreases in n), so once they decrease the partial sums bracket R and the test stops
//      as soon as U is outside the bracket.  R(x) < 2^-63 for x > 48: rejected outright (U lives on a 2^-53 grid).
//
//  (3) b > PGB_BX: the Gamma convolution PG(b, c) = (1/2pi^2) sum_k G_k/d_k, d_k = (k-1/2)^2 + (c/2pi)^2,
//      G_k ~ Gamma(b, 1) (the representation behind rand_gamma_sum) with KT explicit terms and the whole tail
//      sum_{k>KT} replaced by loc + theta * Gamma(shape) matching its first THREE cumulants b T1, b T2, 2 b T3
//      (T_j = sum_{k>KT} d_k^-j):  theta = T3/T2, shape = b T2^3/T3^2, loc = b (T1 - T2^2/T3) >= 0 (Cauchy-Schwarz).
//      KT = 2 for |c| <= 3 and grows like 0.4 |c| beyond, which keeps the Kolmogorov distance to the exact law below
//      3e-6 for every b > 4 (3e-7 for |c| <= 2.5) — computed from the two characteristic functions by Fourier
//      inversion in tests/test_pg_laws_cpu.py, not estimated from samples.  Cost is independent of b, where the
//      reference's exact summation costs b Devroye draws (b ~ 20 for the NegBin / Poisson workloads).
#pragma once
#include "aug_math.cuh"
#include "aug_pg.cuh"

namespace augb {

using augp::PI;
using augp::T;

#ifndef PGB_BX
#define PGB_BX 4.0   // b <= PGB_BX: exact pieces (1) + (2); b > PGB_BX: convolution (3)
#endif
#define PGB_KT_MAX 32

// Philox block counter of a step: tag (3 bits) | sub-draw / term index (7 bits) | round (8 bits) | attempt (14 bits)
// tags: 0 Devroye round, 1 truncated-IG attempt, 2 series refinement, 3 Gamma attempt, 4 fractional attempt (block 1),
//       5 fractional attempt (block 2), 6 Poisson stream of the fresh step, 7 sequential fall-back stream
__device__ __forceinline__ uint32_t ctr(uint32_t tag, uint32_t sub, uint32_t round, uint32_t attempt) {
    return (tag << 29) | (sub << 22) | (round << 14) | attempt;
}
#define PGB_MAXROUND 255u
#define PGB_MAXATT 16000u

struct Key {
    uint32_t k0, k1, c3;
};

// tail sums T_j(w) = sum_{k>=3} ((k-1/2)^2 + w)^-j as power series in w (radius 6.25; used for w <= 0.41, i.e.
// |c| <= 4, where 14 / 11 / 11 terms give 1e-17 / 1e-12 / 1e-11 relative): coefficients
// (-1)^m binom(j+m-1, m) zeta(2(j+m), 5/2)
static __constant__ double TAIL1[14] = {
    -7.2063434910111825475e-12, 4.5043155479529543806e-11, -2.815627611113582269e-10, 1.760295677618269014e-9,
    -1.1008345466854143143e-8,  6.8882254155130349311e-8,  -4.3150565296917940235e-7, 2.7092759794669097368e-6,
    -0.000017089167198843548335, 0.00010882584206868630821, -0.00070738816518316098251, 0.0048214098213931957046,
    -0.037317641469542008543,   0.49035775610023486497};
static __constant__ double TAIL2[11] = {
    3.0971903722249404959e-9,  -1.760295677618269014e-8,  9.9075109201687288283e-8,  -5.5105803324104279449e-7,
    3.0205395707842558165e-6,  -0.000016255655876801458421, 0.000085445835994217741673, -0.00043530336827474523286,
    0.0021221644955494829475,  -0.0096428196427863914092, 0.037317641469542008543};
static __constant__ double TAIL3[11] = {
    2.9728482616489498912e-9,  -1.548595186112470248e-8,  7.9213305492822105631e-8,  -3.9630043680674915313e-7,
    1.9287031163436497807e-6,  -9.0616187123527674494e-6, 0.000040639139692003646051, -0.00017089167198843548335,
    0.00065295505241211784929, -0.0021221644955494829475, 0.0048214098213931957046};

// For KT = 2 (|c| <= 3) the tail parameters themselves as polynomials in w (degree 8, 2e-16 relative on w <= 0.41):
// theta(w) = T3/T2 and kappa(w) = T2^3/T3^2 (shape = b kappa); loc = b (T1 - theta kappa) is formed from the SAME
// theta and kappa that are sampled with, so the mean b T1 of the tail is exact whatever their rounding.
static __constant__ double THETA2[9] = {
    5.28127610569258015e-8, -4.31455767197214569e-7, 2.80415742473835293e-6, -0.000017721113520866235,
    0.000110869988004362895, -0.000682087731559447308, 0.00408207521761980359, -0.023482722238011115,
    0.129199210655591495};
static __constant__ double KAPPA2[9] = {
    3.01471237519757539e-9, -4.83814298092573481e-8, 6.34800672867450131e-7, -7.06935670734417957e-6,
    0.0000628052840369596483, -0.000341645633489901802, -0.00256658071170994153, 0.234991786141717408,
    2.23560188539702316};

struct Conv {
    double w;            // (c / 2pi)^2
    double loc, theta, shape;   // tail = loc + theta * Gamma(shape)
    int kt;              // explicit terms
};

__device__ __forceinline__ int conv_kt(double absc) {
    if (absc <= 3.0) return 2;
    const int k = (int)ceil(0.4 * absc) + 1;
    return k > PGB_KT_MAX ? PGB_KT_MAX : k;
}

// tail parameters for |c| > 3 (KT > 2 explicit terms): out of line — it is rare and heavy (exp, divisions)
static __device__ __noinline__ Conv conv_setup_general(double b, double c) {
    Conv s;
    const double x = 0.5 * fabs(c);
    const double xp = x * (1.0 / PI);
    s.w = xp * xp;
    s.kt = conv_kt(fabs(c));
    double t1, t2, t3;
    if (x <= 2.0) {
        double p = TAIL1[0], q = TAIL2[0], r = TAIL3[0];
#pragma unroll 1
        for (int m = 1; m < 14; ++m) p = fma(p, s.w, TAIL1[m]);
#pragma unroll 1
        for (int m = 1; m < 11; ++m) { q = fma(q, s.w, TAIL2[m]); r = fma(r, s.w, TAIL3[m]); }
        t1 = p; t2 = q; t3 = r;
    } else {
        // totals over all k >= 1 in closed form (x >= 2: no cancellation), minus the terms k = 1, 2
        const double e = exp(-2.0 * x);
        const double th = (1.0 - e) / (1.0 + e);
        const double se = 1.0 - th * th;                       // sech^2 x
        const double x2 = x * x, x3 = x2 * x;
        const double h = th - x * se;
        const double P2 = PI * PI;
        const double s1 = 0.5 * P2 * th / x;
        const double s2 = 0.25 * P2 * P2 * h / x3;
        const double s3 = (P2 * P2 * P2 / 16.0) * (3.0 * h / (x3 * x2) - 2.0 * se * th / x3);
        const double i1 = 1.0 / (0.25 + s.w), i2 = 1.0 / (2.25 + s.w);
        t1 = s1 - i1 - i2;
        t2 = s2 - i1 * i1 - i2 * i2;
        t3 = s3 - i1 * i1 * i1 - i2 * i2 * i2;
    }
#pragma unroll 1
    for (int k = 3; k <= s.kt; ++k) {                          // |c| > 3 only
        const double km = (double)k - 0.5;
        const double id = 1.0 / fma(km, km, s.w);
        t1 -= id;
        t2 -= id * id;
        t3 -= id * id * id;
    }
    s.theta = t3 / t2;
    const double ka = t2 * t2 * t2 / (t3 * t3);
    s.shape = b * ka;
    s.loc = b * fma(-s.theta, ka, t1);
    return s;
}

// tail parameters for b, c  (b > 0)
__device__ __forceinline__ Conv conv_setup(double b, double c) {
    Conv s;
    const double x = 0.5 * fabs(c);
    const double xp = x * (1.0 / PI);
    s.w = xp * xp;
    s.kt = conv_kt(fabs(c));
    if (s.kt != 2) return conv_setup_general(b, c);            // |c| > 3: rare, out of line
    {                                                          // |c| <= 3: straight-line, no division
        double p = TAIL1[0], th = THETA2[0], ka = KAPPA2[0];
#pragma unroll
        for (int m = 1; m < 14; ++m) p = fma(p, s.w, TAIL1[m]);
#pragma unroll
        for (int m = 1; m < 9; ++m) { th = fma(th, s.w, THETA2[m]); ka = fma(ka, s.w, KAPPA2[m]); }
        s.theta = th;
        s.shape = b * ka;
        s.loc = b * fma(-th, ka, p);
        return s;
    }
}

using augr::rand_unit_vector;
using augr::sqrt_pos;

// Marsaglia-Tsang accept test for a N(0,1) variate x and a uniform u; returns Gamma(shape) or < 0
__device__ __forceinline__ double mt_test(double x, double u, double d, double ci) {
    double v = fma(ci, x, 1.0);
    if (v <= 1e-90) return -1.0;
    v = v * v * v;
    const double x2 = x * x;
    if (u < fma(-0.0331 * x2, x2, 1.0)) return d * v;
    if (augr::mt_squeeze2(x, u, d, ci)) return d * v;               // rigorous second squeeze: no logarithm (aug_rng.cuh)
    if (augf::log_(u) < fma(0.5, x2, d * (1.0 - v + augf::log_(v)))) return d * v;
    return -1.0;
}
// TWO independent attempts for Gamma(shape >= 1, 1) out of ONE Philox block: both normals of a Box-Muller pair
// (w0 radius, w1 angle — 32 bits each: the pair lives on a 2^64-point lattice and is cut at 6.6 sigma, a 3e-11
// perturbation used only inside the convolution, whose certified distance to the exact law is 3e-6) and w2, w3 as
// the two accept uniforms.
__device__ __forceinline__ void gamma_pair_attempt(const uint32_t (&w)[4], double shape, double& v1, double& v2) {
    const double d = shape - (1.0 / 3.0);
    const double ci = augf::rsqrt_(9.0 * d);
    const double r = sqrt_pos(-2.0 * augf::log_(augr::u32_mid(w[0])));
    double cs, sn;
    rand_unit_vector(w[1], cs, sn);
    v1 = mt_test(r * cs, augr::u32_mid(w[2]), d, ci);
    v2 = mt_test(r * sn, augr::u32_mid(w[3]), d, ci);
}

// one Marsaglia-Tsang attempt for Gamma(shape >= 1, 1) out of one Philox block; < 0: rejected
// (w0, w1) radius of a Box-Muller normal, w2 its angle, w3 the accept uniform
__device__ __forceinline__ double gamma_attempt(const uint32_t (&w)[4], double shape) {
    const double d = shape - (1.0 / 3.0);
    const double ci = augf::rsqrt_(9.0 * d);
    const double rad2 = -2.0 * augf::log_(augr::u53_open0(w[0], w[1]));
    double cs, sn;
    rand_unit_vector(w[2], cs, sn);
    return mt_test(sqrt_pos(rad2) * cs, augr::u32_mid(w[3]), d, ci);
}

// ---- (2) fractional part: one proposal X ~ IG(mean e/z, shape e^2) (J* scale) out of one Philox block
// (w0, w1) radius, w2 angle of a normal N; w3 picks the root (Michael-Schucany-Haas).  Stable for z -> 0 (Levy).
__device__ __forceinline__ double frac_propose(const uint32_t (&w)[4], double e, double z) {
    const double rad2 = -2.0 * augf::log_(augr::u53_open0(w[0], w[1]));
    double cs, sn;
    rand_unit_vector(w[2], cs, sn);
    const double y = fmax(rad2 * cs * cs, 1e-280);                  // N^2
    const double h = y * augf::rcp(2.0 * e);
    const double x1 = e * augf::rcp(z + h + sqrt_pos(h * (2.0 * z + h)));   // smaller root; = e^2/y at z = 0
    // P(x1) = mu/(mu + x1) = e/(e + z x1); other root mu^2/x1 = e^2/(z^2 x1)
    if (augr::u32_mid(w[3]) * (e + z * x1) <= e) return x1;
    const double ez = e / z;
    return ez * ez * augf::rcp(fmax(x1, 1e-280));
}
// The same proposal AND the uniform of the acceptance test out of ONE block.  N^2 = rad2 cos^2 uses 53 + 30 bits of
// (w0, w1, w2) — the two sign bits of the random unit vector do not matter for a square — which leaves 11 + 2 spare bits;
// with w3 they make a 45-bit uniform u.  u picks the root (u <= P(x1)), and GIVEN the root the rescaled u — u / P(x1) or
// (u - P(x1)) / (1 - P(x1)) — is again uniform and independent of the proposal: it is the acceptance uniform.  The total
// probability of a decision that differs from infinitely fine uniforms is below 2^-45 per attempt.
__device__ __forceinline__ double frac_propose1(const uint32_t (&w)[4], double e, double z, double& uacc) {
    const double rad2 = -2.0 * augf::log_(augr::u53_open0(w[0], w[1]));
    double cs, sn;
    rand_unit_vector(w[2], cs, sn);
    const double y = fmax(rad2 * cs * cs, 1e-280);                  // N^2
    const double h = y * augf::rcp(2.0 * e);
    const double x1 = e * augf::rcp(z + h + sqrt_pos(h * (2.0 * z + h)));   // smaller root; = e^2/y at z = 0
    const uint64_t bits = ((uint64_t)w[3] << 13) | ((uint64_t)(w[0] & 0x7ffu) << 2) | (uint64_t)(w[2] >> 30);
    const double u = fma((double)bits, 0x1.0p-45, 0x1.0p-46);       // (0, 1) on a 2^-45 grid
    const double p1 = e * augf::rcp(e + z * x1);                    // P(x1) = mu/(mu + x1) in (0, 1]
    // both roots and both rescaled uniforms in straight-line code, then selects: the other root is taken by ~20 % of the
    // lanes, i.e. by some lane of nearly every step (z = 0: p1 = 1, the first root always; the inf below is never selected)
    const bool first = u <= p1;
    const double ez = e * augf::rcp(fmax(z, 1e-300));
    const double x2 = ez * ez * augf::rcp(fmax(x1, 1e-280));        // mu^2 / x1
    uacc = first ? u * augf::rcp(p1) : (u - p1) * augf::rcp(fmax(1.0 - p1, 1e-300));
    return first ? x1 : x2;
}
// accept X with probability R(x) = sum_n (-1)^n c_n q^{n(n+e)}; u in (0, 1]
__device__ __forceinline__ bool frac_accept(double x, double e, double u) {
    if (!(x <= 48.0)) return false;                                  // R < 2^-63 (also catches inf / nan)
    const double ix = augf::rcp(fmax(x, 1e-290));
    const double q = augf::exp_(fmax(-2.0 * ix, -700.0));
    double step = q * augf::exp_(fmax(-2.0 * e * ix, -700.0));       // q^{2n+1+e} for n = 0
    const double q2 = q * q;
    // n = 1: t_1 = (2 + e) q^{1+e}; R >= 1 - t_1 whenever t_1 <= 1 (then all terms decrease)
    double cn = 2.0 + e, P = step;
    double t = cn * P;
    double S = 1.0 - t;
    if (t <= 1.0 && u <= S) return true;
    double tprev = t;
    step *= q2;
#pragma unroll 1
    for (int n = 1; n < 400; ++n) {
        const double dn = (double)n;
        cn *= ((dn + e) * (2.0 * dn + 2.0 + e)) * augf::rcp((dn + 1.0) * (2.0 * dn + e));
        P *= step;
        step *= q2;
        t = cn * P;
        const bool dec = t <= tprev;
        tprev = t;
        if (n & 1) {             // term n + 1 even: add -> S is an upper bound once the terms decrease
            S += t;
            if (dec && u > S) return false;
        } else {
            S -= t;
            if (dec && u <= S) return true;
        }
        if (dec && t < 1e-18) break;
    }
    return u <= S;
}

// ---- (1) Devroye pieces with this file's counter layout (same arithmetic as aug_pg.cuh: pg1_accept)
static __device__ __noinline__ bool dev_accept_series(double x, uint32_t uacc, uint32_t k0, uint32_t k1, uint32_t c3, uint32_t e_lo,
                                               uint32_t e_hi, uint32_t sub, uint32_t round) {
    double u = augr::u32_mid(uacc);
    const double q = x > T ? -0.5 * PI * PI * x : -2.0 / x;
    uint32_t w[4];
    augr::philox4x32_10(k0, k1, e_lo, e_hi, ctr(2u, sub, round, 0u), c3, w);
    u += ((double)w[0] - 2147483648.0) * 0x1.0p-64;
    double sum = 1.0;
    for (int n = 1;; ++n) {
        const double rho = (double)(2 * n + 1) * exp(q * (double)(n * (n + 1)));
        if (n & 1) {
            sum -= rho;
            if (u <= sum) return true;
        } else {
            sum += rho;
            if (u > sum) return false;
        }
    }
}
__device__ __forceinline__ bool dev_accept(double x, uint32_t uacc, const Key& k, uint32_t e_lo, uint32_t e_hi,
                                           uint32_t sub, uint32_t round) {
    const int xh = __double2hiint(x);
    const bool mid = xh >= 0x3fe00000 && xh < 0x3fe99999;      // [0.5, 0.8)
    const bool wide = xh >= 0x3fd99999 && xh < 0x3ff00000;     // [0.39999, 1.0)
    const uint32_t thr = mid ? 4269197491u : (wide ? 4289813334u : 4294108302u);
    if (uacc <= thr) return true;                              // squeeze: >= 99.4% of the proposals
    return dev_accept_series(x, uacc, k.k0, k.k1, k.c3, e_lo, e_hi, sub, round);
}

// rand(Poisson(lam)) on the element's own stream (tag 6): chop-down inversion in line for lam < 40 (a lone lane's 30 extra search steps cost a quarter of the out-of-line call), PTRS (Hörmann 1993)
// out of line above (rare for the rates of poisson.jl:26-28 / heteroscedasticgaussian.jl:28-32 and heavy: lgamma, logs)
static __device__ __noinline__ int64_t poisson_ptrs(uint64_t seed, uint64_t offset, uint64_t gi, double lam) {
    augr::Philox g;
    g.init(seed, offset, gi, 193u);
    return augr::poisson_rand(g, lam);
}
__device__ __forceinline__ int64_t poisson_draw(const augr::PhiloxKeys& rk, const Key& key, uint32_t e_lo, uint32_t e_hi,
                                                uint64_t seed, uint64_t offset, uint64_t gi, double lam) {
    if (!(lam > 0.0)) return 0;
    if (lam >= 40.0) return poisson_ptrs(seed, offset, gi, lam);
    // chop-down inversion from 0 with ONE 53-bit uniform (block tag 6); the search multiplies by a table of 1/k.
    // k reaches 200 only through the 1e-16 round-off tail of the cumulative sum: restart with the block's other half.
    uint32_t w[4];
    AUG_PHILOX_RK(rk, e_lo, e_hi, ctr(6u, 0u, 0u, 0u), key.c3, w);
    const double p0 = augf::exp_(-lam);
    for (int half = 0;; ++half) {
        double u = half == 0 ? augr::u53_open0(w[0], w[1]) : augr::u53_open0(w[2], w[3]);
        double p = p0;
        int k = 0;
        while (u > p && k < 200) {
            u -= p;
            ++k;
            p *= k < 64 ? lam * augr::INV_K[k] : lam / (double)k;
        }
        if (k < 200 || half == 1) return (int64_t)(k < 200 ? k : 0);
    }
}
// logistic(x) with the LogExpFunctions saturation (augm::logistic) on the straight-line exp / reciprocal
__device__ __forceinline__ double logistic_fast(double x) {
    if (x < augm::LOGISTIC_LO) return 0.0;
    if (x > augm::LOGISTIC_HI) return 1.0;
    const double e = augf::exp_(-fmin(fabs(x), 700.0));
    const double r = augf::rcp(1.0 + e);
    return x >= 0.0 ? r : e * r;
}

// ---- the whole draw on ONE sequential stream: the three pieces in one place.  Used by the fall-backs (counters exhausted,
// probability < 1e-70 per draw), by the rare b >= 2 elements of the categorical sampler, and by the one-thread-one-draw
// kernels (AUGCUDA_NO_COMPACT=1, shards of 2^32 elements and more): every route samples the same exact / certified law.
static __device__ __noinline__ double pg_draw_stream(augr::Philox& g, double b, bool b_is_int, double c, const double* tab) {
    if (!(b > 0.0)) return 0.0;                                      // Dirac at 0  polyagamma.jl:122-124
    if (b_is_int) b = rint(b);
    if (b > PGB_BX) {
        const Conv s = conv_setup(b, c);
        double acc = s.loc;
        for (int k = 1; k <= s.kt + 1; ++k) {
            const double km = (double)k - 0.5;
            const double sh = k <= s.kt ? b : s.shape;
            const double wt = k <= s.kt ? 1.0 / fma(km, km, s.w) : s.theta;
            double v;
            do {
                uint32_t w[4] = {g.next32(), g.next32(), g.next32(), g.next32()};
                v = gamma_attempt(w, sh);
            } while (v < 0.0);
            acc = fma(v, wt, acc);
        }
        return acc * (0.5 / (PI * PI));
    }
    const double fl = floor(b);
    double e = b - fl;
    if (e < 1e-250) e = 0.0;
    const double z = 0.5 * fabs(c);
    double acc = 0.0;
    if (e > 0.0) {
        for (;;) {
            uint32_t w[4] = {g.next32(), g.next32(), g.next32(), g.next32()};
            const double x = frac_propose(w, e, z);
            if (frac_accept(x, e, g.u01_open0())) { acc = 0.25 * x; break; }
        }
    }
    const augp::PG1 s = augp::pg1_setup(c, tab);
    for (int k = 0; k < (int)fl; ++k) acc += augp::pg1_draw(g, s);
    return acc;
}
__device__ __forceinline__ double pgb_sequential(uint64_t seed, uint64_t offset, uint64_t gi, double b, bool b_is_int,
                                                 double c, const double* tab) {
    augr::Philox g;
    g.init(seed, offset, gi, 224u);                                  // tag 7
    return pg_draw_stream(g, b, b_is_int, c, tab);
}

}  // namespace augb
