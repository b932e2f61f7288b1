// aug_pg.cuh — Pólya-Gamma sampler (device).
//
// Law sampled: PG(b, c) as in SpecialDistributions/polyagamma.jl:121-257 of the reference
// (Devroye's exact J*(1, z) sampler, Polson-Scott-Windle Alg. 1; integer b by summation).
// What differs from the reference's arithmetic, without changing the law:
//   * the alternating-series test is normalised by a_0(x): rho_n = a_n/a_0 = (2n+1) exp(q n(n+1))
//     with q = -pi^2 x/2 (x > t) or q = -2/x (x <= t), so neither a_0 nor log(x) is evaluated;
//   * a squeeze (rho_1 <= 0.0058 on both sides of t = 0.64) accepts 99.4% of proposals with no exp;
//   * the proposal mass r(z) is evaluated directly (2 erfc + 2 exp) instead of through logcdf,
//     and is 0 to double precision for z >= 20;
//   * the non-integer remainder of b is drawn as a KT-term Gamma convolution plus a Gamma matched
//     to the exact mean and variance of the infinite tail, instead of the reference's biased
//     200-term truncation (polyagamma.jl:157-164) — see DESIGN.md "PG(b) for real b".
#pragma once
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

// mass_texpon(z, K) polyagamma.jl:179-192: r = p/(p+q)
__device__ __forceinline__ PG1 pg1_setup(double c) {
    PG1 s;
    s.z = 0.5 * fabs(c);
    s.K = fma(0.5 * s.z, s.z, PI2_8);
    s.invK = 1.0 / s.K;
    if (s.z == 0.0) {
        s.r = R0;
    } else if (s.z >= 20.0) {
        s.r = 0.0;
    } else {
        const double b = INV_SQRT_T * (T * s.z - 1.0);
        const double a = -INV_SQRT_T * (T * s.z + 1.0);
        const double Kt = s.K * T;
        // Phi(x) = erfc(-x/sqrt2)/2
        const double qb = exp(Kt - s.z) * 0.5 * erfc(-b * INV_SQRT2);
        const double qa = exp(Kt + s.z) * 0.5 * erfc(-a * INV_SQRT2);
        const double qdivp = (4.0 / PI) * s.K * (qb + qa);
        s.r = 1.0 / (1.0 + qdivp);
    }
    return s;
}

// rand_truncated_inverse_gaussian(z) on (0, t]  polyagamma.jl:195-221
__device__ __forceinline__ double trunc_ig(augr::Philox& g, double z) {
    if (z < 1.0 / T) {  // mu = 1/z > t
        const double hz2 = 0.5 * z * z;
        for (;;) {
            double E, E2;
            do {
                E = g.expo();
                E2 = g.expo();
            } while (E * E > 2.0 * E2 / T);
            const double d = fma(T, E, 1.0);
            const double x = T / (d * d);
            const double a = hz2 * x;          // alpha = exp(-a)
            const double u = g.u01();
            if (u <= 1.0 - a || u <= exp(-a)) return x;
        }
    }
    const double mu = 1.0 / z;
    double x;
    do {
        const double n = g.normal();
        const double muy = mu * n * n;
        x = mu + 0.5 * mu * muy - 0.5 * mu * sqrt(fma(muy, muy, 4.0 * muy));
        if (g.u01() * (mu + x) > mu) x = mu * mu / x;
    } while (x > T);
    return x;
}

// one draw of PG(1, 2z) = J*(1, z)/4   (sample_pg1, polyagamma.jl:225-257)
__device__ __forceinline__ double pg1_draw(augr::Philox& g, const PG1& s) {
    for (;;) {
        double x, q;
        if (g.u01() < s.r) {
            x = fma(g.expo(), s.invK, T);          // truncated exponential on (t, inf)
            q = -0.5 * PI * PI * x;
        } else {
            x = trunc_ig(g, s.z);
            q = -2.0 / x;
        }
        const double u = g.u01();
        if (u <= 0.994) return 0.25 * x;           // squeeze
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
__device__ __forceinline__ double pg_draw(augr::Philox& g, double b, bool b_is_int, double c) {
    if (!(b > 0.0)) return 0.0;                    // b == 0 -> Dirac at 0 (:122-124)
    const double fl = b_is_int ? rint(b) : floor(b);
    double acc = 0.0;
    if (fl >= 1.0) {
        const PG1 s = pg1_setup(c);
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
