// aug_rng.cuh — counter-based Philox4x32-10 streams and the elementary samplers of the Gibbs path.
//
// Replaces the `rng::AbstractRNG` of aux_sample!/init_aux_variables (generic.jl:1-20,32-34) and the
// Distributions.jl / Random samplers they call (rand, randexp, randn, Gamma, Poisson,
// InverseGaussian).  A stream is keyed by (seed, verb offset) and positioned by the GLOBAL element
// index, so draws do not depend on how the observation axis is sharded over GPUs or on which
// thread processes an element.  Only the laws have to match the reference, not the bit streams.
#pragma once
#include <stdint.h>

#include "aug_fastmath.cuh"

namespace augr {

// One Philox4x32-10 block: 128 random bits as a pure function of (key, counter).
__device__ __forceinline__ void philox4x32_10(uint32_t k0, uint32_t k1, uint32_t c0, uint32_t c1, uint32_t c2,
                                              uint32_t c3, uint32_t (&w)[4]) {
    uint32_t x0 = c0, x1 = c1, x2 = c2, x3 = c3, a = k0, b = k1;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, x0), lo0 = 0xD2511F53u * x0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, x2), lo1 = 0xCD9E8D57u * x2;
        const uint32_t y0 = hi1 ^ x1 ^ a, y1 = lo1, y2 = hi0 ^ x3 ^ b, y3 = lo0;
        x0 = y0; x1 = y1; x2 = y2; x3 = y3;
        a += 0x9E3779B9u;
        b += 0xBB67AE85u;
    }
    w[0] = x0; w[1] = x1; w[2] = x2; w[3] = x3;
}

// the same block with the ten round keys precomputed by the host (rk[2r] = k0 + r*0x9E3779B9, rk[2r+1] = k1 + r*0xBB67AE85):
// the key is launch-uniform, so the per-round key additions (20 of ~75 instructions of a block) become constant-bank operands
struct PhiloxKeys {
    uint32_t rk[20];
};
static inline void philox_round_keys(uint32_t k0, uint32_t k1, PhiloxKeys* out) {
    for (int r = 0; r < 10; ++r) {
        out->rk[2 * r] = k0 + (uint32_t)r * 0x9E3779B9u;
        out->rk[2 * r + 1] = k1 + (uint32_t)r * 0xBB67AE85u;
    }
}
#define AUG_PHILOX_RK(KEYS, C0, C1, C2, C3, W)                                              \
    do {                                                                                    \
        uint32_t x0_ = (C0), x1_ = (C1), x2_ = (C2), x3_ = (C3);                            \
        _Pragma("unroll") for (int r_ = 0; r_ < 10; ++r_) {                                 \
            const uint32_t hi0_ = __umulhi(0xD2511F53u, x0_), lo0_ = 0xD2511F53u * x0_;     \
            const uint32_t hi1_ = __umulhi(0xCD9E8D57u, x2_), lo1_ = 0xCD9E8D57u * x2_;     \
            const uint32_t y0_ = hi1_ ^ x1_ ^ (KEYS).rk[2 * r_], y2_ = hi0_ ^ x3_ ^ (KEYS).rk[2 * r_ + 1]; \
            x0_ = y0_; x1_ = lo1_; x2_ = y2_; x3_ = lo0_;                                   \
        }                                                                                   \
        (W)[0] = x0_; (W)[1] = x1_; (W)[2] = x2_; (W)[3] = x3_;                             \
    } while (0)

// uniforms from raw words: 53-bit (0,1] for values that are returned, 32-bit mid-point (0,1) for decisions
__device__ __forceinline__ double u53_open0(uint32_t lo, uint32_t hi) {
    return (double)(((((uint64_t)hi << 32) | lo) >> 11) + 1ull) * 0x1.0p-53;
}
__device__ __forceinline__ double u32_mid(uint32_t w) { return fma((double)w, 0x1.0p-32, 0x1.0p-33); }

struct Philox {
    uint32_t k0, k1;          // key   = seed
    uint32_t c0, c1, c2, c3;  // counter = (element lo, element hi, block counter, verb offset)
    uint32_t buf[4];
    int have;                 // 32-bit words left in buf
    double spare;             // second Box-Muller normal
    bool has_spare;

    __device__ __forceinline__ void init(uint64_t seed, uint64_t offset, uint64_t element, uint32_t lane_tag = 0) {
        k0 = (uint32_t)seed;
        k1 = (uint32_t)(seed >> 32) ^ (uint32_t)(offset >> 32);
        c0 = (uint32_t)element;
        c1 = (uint32_t)(element >> 32);
        c2 = lane_tag << 24;  // up to 2^24 blocks (6.7e7 uniforms) per element and tag
        c3 = (uint32_t)offset;
        have = 0;
        has_spare = false;
    }
    __device__ __forceinline__ void refill() {
        philox4x32_10(k0, k1, c0, c1, c2, c3, buf);
        c2++;
        have = 4;
    }
    // words are consumed from buf[3] down to buf[0] (selects, not a dynamically indexed array)
    __device__ __forceinline__ uint32_t next32() {
        if (have < 1) refill();
        const uint32_t w = have == 4 ? buf[3] : (have == 3 ? buf[2] : (have == 2 ? buf[1] : buf[0]));
        have -= 1;
        return w;
    }
    __device__ __forceinline__ uint64_t next64() {
        const uint32_t lo = next32();
        const uint32_t hi = next32();
        return ((uint64_t)hi << 32) | lo;
    }
    // U in [0,1): 53 random bits (rand(rng))
    __device__ __forceinline__ double u01() { return (double)(next64() >> 11) * 0x1.0p-53; }
    // U in (0,1]
    __device__ __forceinline__ double u01_open0() { return (double)((next64() >> 11) + 1ull) * 0x1.0p-53; }
    // U in (0,1) on a 2^-32 grid: for accept/reject DECISIONS only (never for a returned value)
    __device__ __forceinline__ double u01_32() { return ((double)next32() + 0.5) * 0x1.0p-32; }
    // randexp(rng): u in (0,1] is a normal double, so the straight-line log is valid
    __device__ __forceinline__ double expo() { return -augf::log_(u01_open0()); }
    // randn(rng): Box-Muller; the sine branch is kept for the next call
    __device__ __forceinline__ double normal() {
        if (has_spare) {
            has_spare = false;
            return spare;
        }
        const double u = u01_open0();
        const double v = u01();
        const double r = sqrt(-2.0 * augf::log_(u));
        double sn, cs;
        sincospi(2.0 * v, &sn, &cs);
        spare = r * sn;
        has_spare = true;
        return r * cs;
    }
};

// sqrt of a non-negative normal-range number without the IEEE slow path
__device__ __forceinline__ double sqrt_pos(double v) {
    v = fmax(v, 1e-290);
    return v * augf::rsqrt_(v);
}

// (cos, sin) of a UNIFORMLY RANDOM angle out of 32 random bits: 29 bits pick a in (0, pi/4), Taylor polynomials give
// (cos a, sin a) to 1e-16, and the top 3 bits apply a random symmetry of the octagon (swap, two sign flips) — the point
// is uniform on the circle, which is all Box-Muller needs (no argument reduction: ~25 instructions against ~40 for cospi)
__device__ __forceinline__ void rand_unit_vector(uint32_t w, double& cs, double& sn) {
    const double a = fma((double)(w & 0x1fffffffu), 0x1.0p-29 * (3.14159265358979323846 / 4.0), 0x1.0p-30 * (3.14159265358979323846 / 4.0));
    const double a2 = a * a;
    double ps = -1.0 / 1307674368000.0;                       // sin a = a (1 - a^2/3! + ... - a^14/15!)
    ps = fma(ps, a2, 1.0 / 6227020800.0);
    ps = fma(ps, a2, -1.0 / 39916800.0);
    ps = fma(ps, a2, 1.0 / 362880.0);
    ps = fma(ps, a2, -1.0 / 5040.0);
    ps = fma(ps, a2, 1.0 / 120.0);
    ps = fma(ps, a2, -1.0 / 6.0);
    const double sa = fma(a * a2, ps, a);
    double pc = 1.0 / 20922789888000.0;                       // cos a = 1 - a^2/2! + ... + a^16/16!
    pc = fma(pc, a2, -1.0 / 87178291200.0);
    pc = fma(pc, a2, 1.0 / 479001600.0);
    pc = fma(pc, a2, -1.0 / 3628800.0);
    pc = fma(pc, a2, 1.0 / 40320.0);
    pc = fma(pc, a2, -1.0 / 720.0);
    pc = fma(pc, a2, 1.0 / 24.0);
    pc = fma(pc, a2, -0.5);
    const double ca = fma(pc, a2, 1.0);
    const bool sw = (w >> 29) & 1u;
    double x = sw ? sa : ca, y = sw ? ca : sa;
    cs = (w >> 30) & 1u ? -x : x;
    sn = (w >> 31) ? -y : y;
}

// Second squeeze of the Marsaglia-Tsang test  log u < x^2/2 + d (1 - v + log v),  v = (1 + t)^3, t = c x, 9 d c^2 = 1.
// The first squeeze u < 1 - 0.0331 x^4 holds for every d >= 2/3 and is loose for large d (it leaves 10 % of the attempts
// undecided where the true rejection rate at d = 20 is 0.1 %), and the exact test costs two logarithms that a warp
// executes for a handful of lanes.  With log(1+t) = t - t^2/2 + ... - t^6/6 + R7, |R7| <= (2/7)|t|^7 for |t| <= 1/2, the
// x^2 terms cancel (9 d c^2 = 1) and   rhs >= d t^4 (-3/4 + 3t/5 - t^2/2) - (6/7) d |t|^7;   with w = 1 - u,
// log u <= -w - w^2/2.  Returns true when these rigorous bounds already prove acceptance (the 1e-12 covers the rounding
// of 9 d c^2 and of the polynomial); false = undecided, run the exact test.
__device__ __forceinline__ bool mt_squeeze2(double x, double u, double d, double c) {
    const double t = c * x, t2 = t * t, t4 = t2 * t2;
    const double w = 1.0 - u;
    const double lhs = fma(-0.5 * w, w, -w);
    const double rhs = d * fma(t4, fma(t, fma(t, -0.5, 0.6), -0.75), (-6.0 / 7.0) * t4 * t2 * fabs(t)) - 1e-12;
    return fabs(t) <= 0.5 && lhs < rhs;
}

// rand(Gamma(shape, 1)) — Marsaglia & Tsang (2000); shape < 1 through the U^(1/shape) boost
__device__ __forceinline__ double gamma_rand(Philox& g, double shape) {
    double boost = 1.0;
    if (shape < 1.0) {
        const double la = augf::log_(g.u01_open0()) / shape;   // U^(1/shape)
        boost = la > -700.0 ? augf::exp_(la) : 0.0;
        shape += 1.0;
    }
    const double d = shape - 1.0 / 3.0;
    const double c = augf::rsqrt_(9.0 * d);
    for (;;) {
        double x, v;
        do {
            x = g.normal();
            v = fma(c, x, 1.0);
        } while (v <= 0.0);
        v = v * v * v;
        const double u = g.u01_32();
        const double x2 = x * x;
        if (u < 1.0 - 0.0331 * x2 * x2) return boost * d * v;
        if (d >= 3.0 && mt_squeeze2(x, u, d, c)) return boost * d * v;   // pays off for larger shapes only (StudentT: d = 5/3)
        if (augf::log_(u) < 0.5 * x2 + d * (1.0 - v + augf::log_(v))) return boost * d * v;
    }
}

// 1/k, k = 0..63 (entry 0 unused): the chop-down search below multiplies instead of dividing
static __constant__ double INV_K[64] = {
    0.0, 1.0, 1.0 / 2, 1.0 / 3, 1.0 / 4, 1.0 / 5, 1.0 / 6, 1.0 / 7, 1.0 / 8, 1.0 / 9, 1.0 / 10, 1.0 / 11, 1.0 / 12, 1.0 / 13,
    1.0 / 14, 1.0 / 15, 1.0 / 16, 1.0 / 17, 1.0 / 18, 1.0 / 19, 1.0 / 20, 1.0 / 21, 1.0 / 22, 1.0 / 23, 1.0 / 24, 1.0 / 25,
    1.0 / 26, 1.0 / 27, 1.0 / 28, 1.0 / 29, 1.0 / 30, 1.0 / 31, 1.0 / 32, 1.0 / 33, 1.0 / 34, 1.0 / 35, 1.0 / 36, 1.0 / 37,
    1.0 / 38, 1.0 / 39, 1.0 / 40, 1.0 / 41, 1.0 / 42, 1.0 / 43, 1.0 / 44, 1.0 / 45, 1.0 / 46, 1.0 / 47, 1.0 / 48, 1.0 / 49,
    1.0 / 50, 1.0 / 51, 1.0 / 52, 1.0 / 53, 1.0 / 54, 1.0 / 55, 1.0 / 56, 1.0 / 57, 1.0 / 58, 1.0 / 59, 1.0 / 60, 1.0 / 61,
    1.0 / 62, 1.0 / 63};

// rand(Poisson(lambda)): sequential inversion for small rates, PTRS (Hörmann 1993) otherwise
__device__ __forceinline__ int64_t poisson_rand(Philox& g, double lam) {
    if (!(lam > 0.0)) return 0;
    if (lam < 12.0) {
        // inversion by chop-down search from 0 (exact in law; restart guards the 1e-16 round-off tail)
        for (;;) {
            double u = g.u01();
            double p = augf::exp_(-lam);
            int k = 0;
            while (u > p && k < 200) {
                u -= p;
                ++k;
                p *= k < 64 ? lam * INV_K[k] : lam / (double)k;
            }
            if (k < 200) return (int64_t)k;
        }
    }
    const double slam = sqrt(lam), loglam = log(lam);
    const double b = 0.931 + 2.53 * slam;
    const double a = -0.059 + 0.02483 * b;
    const double inv_alpha = 1.1239 + 1.1328 / (b - 3.4);
    const double vr = 0.9277 - 3.6224 / (b - 2.0);
    for (;;) {
        const double u = g.u01() - 0.5;
        const double v = g.u01_open0();
        const double us = 0.5 - fabs(u);
        const double kf = floor((2.0 * a / us + b) * u + lam + 0.43);
        if (us >= 0.07 && v <= vr) return (int64_t)kf;
        if (kf < 0.0 || (us < 0.013 && v > us)) continue;
        if (log(v) + log(inv_alpha) - log(a / (us * us) + b) <= -lam + kf * loglam - lgamma(kf + 1.0))
            return (int64_t)kf;
    }
}

// rand(InverseGaussian(mu, lambda)) — Michael, Schucany & Haas (1976)
__device__ __forceinline__ double invgauss_rand(Philox& g, double mu, double lam) {
    const double z = g.normal();
    const double w = mu * z * z;
    const double x1 = mu + mu / (2.0 * lam) * (w - sqrt(w * (4.0 * lam + w)));
    const double u = g.u01();
    return u * (mu + x1) >= mu ? mu * mu / x1 : x1;
}

}  // namespace augr
