// aug_pg.cuh — Pólya-Gamma sampler (device).
//
// Law sampled: PG(b, c) as in SpecialDistributions/polyagamma.jl:121-257 of the reference
// (Devroye's exact J*(1, z) sampler, Polson-Scott-Windle Alg. 1; integer b by summation).
// What differs from the reference's arithmetic, without changing the law:
//   * the alternating-series test is normalised by a_0(x): rho_n = a_n/a_0 = (2n+1) exp(q n(n+1))
//     with q = -pi^2 x/2 (x > t) or q = -2/x (x <= t), so neither a_0 nor log(x) is evaluated;
//   * a squeeze (rho_1 <= 0.0058 on both sides of t = 0.64) accepts 99.4% of proposals with no exp;
//   * the proposal mass r(z) comes from a host-built piecewise polynomial table instead of two
//     logcdf + three exp per observation, and is 0 to double precision for z >= 20;
//   * accept/reject decisions use 32-bit uniforms (refined to 64 bits in the 0.6% of cases that reach the
//     alternating series) and squeezes 1 - a <= exp(-a) <= 1 - a + a^2/2, values use 53-bit uniforms;
//   * the non-integer remainder of b is drawn as a KT-term Gamma convolution plus a Gamma matched
//     to the exact mean and variance of the infinite tail, instead of the reference's biased
//     200-term truncation (polyagamma.jl:157-164) — see DESIGN.md "PG(b) for real b".
#pragma once
#include "aug_common.cuh"
#include "aug_rng.cuh"

namespace augp {

constexpr double T = 0.64;                           // PG_T  polyagamma.jl:3
constexpr double PI = 3.14159265358979323846;
constexpr double PI2_8 = PI * PI / 8.0;              // polyagamma.jl:4
constexpr double R0 = 0.5776972428360435;            // r(z = 0)  polyagamma.jl:231
constexpr double INV_SQRT_T = 1.25;                  // 1/sqrt(0.64)
constexpr double INV_SQRT2 = 0.70710678118654752440;
constexpr int KT = 12;                               // explicit terms of the Gamma convolution

struct PG1 {
    double z, K, invK, r;
};

// r(z) = p/(p+q), the mass of the exponential part of the proposal (mass_texpon, polyagamma.jl:179-192).
// It only depends on z, so it is read from the degree-7 piecewise polynomial table the host builds once per
// context from the log-domain formula (aug_ctx.cu: build_pg_table; |error| < 1e-15); r < 1e-40 beyond z = 20.
__device__ __forceinline__ PG1 pg1_setup(double c, const double* __restrict__ tab) {
    PG1 s;
    s.z = 0.5 * fabs(c);
    s.K = fma(0.5 * s.z, s.z, PI2_8);
    s.invK = s.K < 1e290 ? augf::rcp(s.K) : 0.0;
    if (s.z >= AUG_PGTAB_N * AUG_PGTAB_H) {
        s.r = 0.0;
    } else {
        const double u = s.z * (1.0 / AUG_PGTAB_H);
        const int k = (int)u;
        const double x = fma(2.0, u - (double)k, -1.0);
        const double2* row = reinterpret_cast<const double2*>(tab + k * AUG_PGTAB_DEG);
        const double2 a = __ldg(row), b = __ldg(row + 1), cc = __ldg(row + 2), d = __ldg(row + 3);
        double p = fma(a.x, x, a.y);
        p = fma(p, x, b.x);
        p = fma(p, x, b.y);
        p = fma(p, x, cc.x);
        p = fma(p, x, cc.y);
        p = fma(p, x, d.x);
        s.r = fma(p, x, d.y);
    }
    return s;
}

// one proposal from the truncated inverse Gaussian IG(1/z, 1) on (0, t]  (polyagamma.jl:195-221)
// returns x > 0 when the proposal stage accepted, a negative value to ask for another attempt
__device__ __forceinline__ double trunc_ig_attempt(augr::Philox& g, double z) {
    if (z < 1.0 / T) {  // mu = 1/z > t: x = t/(1 + tE)^2 with E from exp(-E) exp(-t E^2/2), then alpha = exp(-z^2 x/2)
        const double E = g.expo();
        const double a1 = 0.5 * T * E * E;     // accept E iff E'>= t E^2/2  <=>  U' <= exp(-a1)
        const double u1 = g.u01_32();
        if (u1 > 1.0 - a1) {                   // not decided by the squeeze exp(-a) >= 1 - a
            if (u1 > fma(0.5 * a1, a1, 1.0 - a1) || u1 > exp(-a1)) return -1.0;
        }
        const double d = fma(T, E, 1.0);
        const double x = T * augf::rcp(d * d);
        const double a2 = 0.5 * z * z * x;
        const double u2 = g.u01_32();
        if (u2 > 1.0 - a2) {
            if (u2 > fma(0.5 * a2, a2, 1.0 - a2) || u2 > exp(-a2)) return -1.0;
        }
        return x;
    }
    const double mu = 1.0 / z;
    const double n = g.normal();
    const double muy = mu * n * n;
    double x = mu + 0.5 * mu * muy - 0.5 * mu * sqrt(fma(muy, muy, 4.0 * muy));
    if (g.u01_32() * (mu + x) > mu) x = mu * mu / x;
    return x > T ? -1.0 : x;
}

// one draw of PG(1, 2z) = J*(1, z)/4   (sample_pg1, polyagamma.jl:225-257)
__device__ __forceinline__ double pg1_draw(augr::Philox& g, const PG1& s) {
    for (;;) {
        double x, q;
        if (g.u01_32() < s.r) {
            x = fma(g.expo(), s.invK, T);          // truncated exponential on (t, inf)
            q = -0.5 * PI * PI * x;
        } else {
            do { x = trunc_ig_attempt(g, s.z); } while (x < 0.0);
            q = -2.0 / x;
        }
        double u = g.u01_32();
        if (u <= 0.994) return 0.25 * x;           // squeeze: rho_1 <= 0.0058 on both sides of t
        u += ((double)g.next32() - 2147483648.0) * 0x1.0p-64;   // refine the uniform to 64 bits
        double sum = 1.0;
        for (int n = 1;; ++n) {
            const double rho = (double)(2 * n + 1) * exp(q * (double)(n * (n + 1)));
            if (n & 1) {
                sum -= rho;
                if (u <= sum) return 0.25 * x;
            } else {
                sum += rho;
                if (u > sum) break;
            }
        }
    }
}

// Non-integer remainder e in (0,1): PG(e, c) = (1/2pi^2) sum_k G_k / ((k-1/2)^2 + c^2/4pi^2), G_k ~ Gamma(e,1)
__device__ __forceinline__ double pg_frac(augr::Philox& g, double e, double c) {
    const double w = (c * (0.5 / PI)) * (c * (0.5 / PI));
    double acc = 0.0;
    for (int k = 1; k <= KT; ++k) {
        const double d = fma((double)k - 0.5, (double)k - 0.5, w);
        acc += augr::gamma_rand(g, e) / d;
    }
    // tail k > KT: mean e*T1, variance e*T2 with T1 = sum 1/d_k, T2 = sum 1/d_k^2 (midpoint integrals)
    const double kt = (double)KT;
    const double s = w / (kt * kt);
    double T1, T2;
    if (s < 0.01) {
        T1 = (1.0 / kt) * (1.0 - s * (1.0 / 3.0 - s * (1.0 / 5.0 - s / 7.0)));
        T2 = (1.0 / (kt * kt * kt)) * (1.0 / 3.0 - s * (2.0 / 5.0 - s * (3.0 / 7.0 - s * (4.0 / 9.0))));
    } else {
        const double sw = sqrt(w);
        T1 = atan(sw / kt) / sw;
        T2 = (T1 - kt / (kt * kt + w)) / (2.0 * w);
    }
    const double shape = e * T1 * T1 / T2;
    const double scale = T2 / T1;
    acc += scale * augr::gamma_rand(g, shape);
    return acc * (0.5 / (PI * PI));
}

// rand(PolyaGamma(b, c))  polyagamma.jl:121-154
__device__ __forceinline__ double pg_draw(augr::Philox& g, double b, bool b_is_int, double c,
                                          const double* __restrict__ tab) {
    if (!(b > 0.0)) return 0.0;                    // b == 0 -> Dirac at 0 (:122-124)
    const double fl = b_is_int ? rint(b) : floor(b);
    double acc = 0.0;
    if (fl >= 1.0) {
        const PG1 s = pg1_setup(c, tab);
        const long long nb = (long long)fl;
        for (long long k = 0; k < nb; ++k) acc += pg1_draw(g, s);
    }
    if (!b_is_int) {
        const double e = b - fl;
        if (e > 0.0) acc += pg_frac(g, e, c);
    }
    return acc;
}

}  // namespace augp
