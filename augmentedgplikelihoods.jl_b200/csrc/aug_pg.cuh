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
//   * b != 1 (sums of draws, the fractional remainder, the certified Gamma convolution) lives in aug_pgb.cuh.
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

struct PG1 {
    double z, K, invK, r;
};

// r(z) = p/(p+q), the mass of the exponential part of the proposal (mass_texpon, polyagamma.jl:179-192).
// It only depends on z, so it is read from the degree-7 piecewise polynomial table the host builds once per
// context from the log-domain formula (aug_ctx.cu: build_pg_table; |error| < 1e-15); r < 1e-40 beyond z = 20.
template <bool TAB_SHARED = false>
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
        double2 a, b, cc, d;
        if (TAB_SHARED) { a = row[0]; b = row[1]; cc = row[2]; d = row[3]; }
        else { a = __ldg(row); b = __ldg(row + 1); cc = __ldg(row + 2); d = __ldg(row + 3); }
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

// The same from a COEFFICIENT-major copy of the table in shared memory: tabT[j * AUG_PGTAB_N + k].  Lanes of a warp
// hold different intervals k (z = |f|/2 spreads over ~12 of them): with the row-major layout their 16-byte loads hit
// the same banks (rows are 64 bytes apart: ncu counted 1.1e8 bank conflicts per 1e8 draws in pg1_compact_kernel),
// here lane addresses differ by 8 bytes per unit of k and the eight LDS.64 are conflict-free.
__device__ __forceinline__ PG1 pg1_setup_cm(double c, const double* __restrict__ tabT) {
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
        const double* col = tabT + k;
        double p = fma(col[0], x, col[AUG_PGTAB_N]);
#pragma unroll
        for (int j = 2; j < AUG_PGTAB_DEG; ++j) p = fma(p, x, col[j * AUG_PGTAB_N]);
        s.r = p;
    }
    return s;
}
// cooperative load of the table into shared memory, transposed to coefficient-major
__device__ __forceinline__ void pg1_load_table_cm(double* __restrict__ tabT, const double* __restrict__ tab, int nthreads) {
    for (int t = threadIdx.x; t < AUG_PGTAB_N * AUG_PGTAB_DEG; t += nthreads) {
        const int k = t / AUG_PGTAB_DEG, j = t - k * AUG_PGTAB_DEG;
        tabT[j * AUG_PGTAB_N + k] = __ldg(tab + t);
    }
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

// ------------------------------------------------------------------ PG(1, c), warp-compacted (aug_gibbs.cu: pg1_compact_kernel)
// The per-thread loop above diverges: ncu on aux_sample_kernel<BERNOULLI> counted 9.6 active lanes per issued
// instruction (the cheap exponential proposal, 58% of draws, waits for the truncated-inverse-Gaussian rejection loop
// of its neighbours).  The compacted kernel instead runs every draw as a sequence of uniform STEPS and keeps the
// unfinished ones in a per-warp shared-memory queue, so that each step is executed by (up to) 32 lanes that all need
// it.  A step owns one Philox block keyed by (element, round, attempt): the law of the output and its dependence on
// (seed, offset, global element index) only are unchanged, the order in which a warp works is irrelevant.
//   block A (tag 0; round):          w0 branch selector, (w1,w2) the Exp(1) of the exponential proposal, w3 the
//                                    uniform of the final alternating-series test;
//   block B (tag 1; round, attempt): one truncated-IG proposal attempt: (w0,w1) Exp(1) / radius, w2, w3 decisions;
//   block C (tag 2; round):          32 more bits for the series uniform in the 0.6% of rounds the squeeze leaves open.
__device__ __forceinline__ uint32_t pg1_ctr(uint32_t tag, uint32_t round, uint32_t attempt) {
    return (tag << 28) | (round << 14) | attempt;
}

// the alternating series itself (0.6 % of the proposals), out of line: in line it costs its callers ~10 registers
static __device__ __noinline__ bool pg1_accept_series(double x, uint32_t uacc, uint32_t k0, uint32_t k1, uint32_t e_lo, uint32_t e_hi, uint32_t c3,
                       uint32_t round) {
    double u = augr::u32_mid(uacc);
    const double q = x > T ? -0.5 * PI * PI * x : -2.0 / x;
    uint32_t w[4];
    augr::philox4x32_10(k0, k1, e_lo, e_hi, pg1_ctr(2u, round, 0u), c3, w);
    u += ((double)w[0] - 2147483648.0) * 0x1.0p-64;   // refine the uniform to 64 bits
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

// final accept/reject of a proposal x: alternating series of rho_n = a_n/a_0 = (2n+1) exp(q n(n+1)) with
// q = -pi^2 x/2 for x > t, -2/x for x <= t   (polyagamma.jl:243-255, 167-177)
__device__ __forceinline__ bool pg1_accept(double x, uint32_t uacc, uint32_t k0, uint32_t k1, uint32_t e_lo,
                                           uint32_t e_hi, uint32_t c3, uint32_t round) {
    // squeeze on rho_1(x) = 3 exp(-pi^2 x) (x > t) resp. 3 exp(-4/x) (x <= t), which peaks at x = t (0.0058):
    //   x in [0.5, 0.8]: rho_1 <= 0.0058 -> accept if u <= 0.994;   x in [0.4, 1.0]: rho_1 <= 0.00111 -> u <= 0.9988;
    //   elsewhere: rho_1 <= 1.6e-4 -> u <= 0.9998.   (u = (uacc + 0.5) 2^-32; compares are on the high word of x > 0)
    const int xh = __double2hiint(x);
    const bool mid = xh >= 0x3fe00000 && xh < 0x3fe99999;      // [0.5, 0.8)
    const bool wide = xh >= 0x3fd99999 && xh < 0x3ff00000;     // [0.39999, 1.0)
    const uint32_t thr = mid ? 4269197491u : (wide ? 4289813334u : 4294108302u);
    if (uacc <= thr) return true;
    return pg1_accept_series(x, uacc, k0, k1, e_lo, e_hi, c3, round);
}

// one proposal attempt from the truncated inverse Gaussian on (0, t] out of one Philox block; < 0: rejected
// mu = 1/z > t: x = t/(1 + tE)^2 with E ~ Exp(1) kept with probability exp(-t E^2/2), then x kept with probability
// exp(-z^2 x/2) (polyagamma.jl:195-213).  Both rejections lead to the same retry, so ONE uniform decides both:
// keep iff U <= exp(-(t E^2/2 + z^2 x/2)) — the same law with one exp and one decision.
__device__ __forceinline__ double trunc_ig_small_z(double E, double z, double& a) {
    const double d = fma(T, E, 1.0);
    const double x = T * augf::rcp(d * d);
    a = fma(0.5 * T * E, E, 0.5 * z * z * x);
    return x;
}
__device__ __forceinline__ double trunc_ig_attempt_w(const uint32_t (&w)[4], double z) {
    if (z < 1.0 / T) {
        const double E = -augf::log_(augr::u53_open0(w[0], w[1]));
        double a;
        const double x = trunc_ig_small_z(E, z, a);
        return augr::u32_mid(w[2]) > augf::exp_(-fmin(a, 700.0)) ? -1.0 : x;
    }
    const double mu = 1.0 / z;
    const double rad2 = -2.0 * augf::log_(augr::u53_open0(w[0], w[1]));
    const double cs = cospi(2.0 * augr::u32_mid(w[2]));
    const double muy = mu * rad2 * cs * cs;                      // mu * N(0,1)^2
    double x = mu + 0.5 * mu * muy - 0.5 * mu * sqrt(fma(muy, muy, 4.0 * muy));
    if (augr::u32_mid(w[3]) * (mu + x) > mu) x = mu * mu / x;
    return x > T ? -1.0 : x;
}

}  // namespace augp
