// aug_math.cuh — per-observation closed forms (device).  Each function cites the reference
// formula it evaluates; the arithmetic is re-derived for the FP64 pipe of sm_100a (one shared
// exp(-c), no pow, series near c = 0) and is NOT a transcription of the Julia code.
#pragma once
#include <math.h>

#include "aug_fastmath.cuh"

namespace augm {

constexpr double LN2 = 0.69314718055994530942;
constexpr double LOGISTIC_LO = -744.4400719213812;  // LogExpFunctions._logistic_bounds(Float64)
constexpr double LOGISTIC_HI = 36.7368005696771;
constexpr double PI = 3.14159265358979323846;

// Everything the Pólya-Gamma moments need from the tilt c >= 0:
//   h   = tanh(c/2)/(2c)           -> mean(PG(b,c)) = b*h        (polyagamma.jl:25-31; 1/4 at c == 0)
//   lch = logcosh(c/2)             -> KL(PG(b,c)||PG(b,0)) = b*lch - c^2*b*h/2   (polyagamma.jl:99-110)
//   e   = exp(-c), inv1pe = 1/(1+e) (re-used by approx_expected_logistic)
struct PGTerms {
    double h, lch, e, inv1pe;
    double l1pe;   // log(1 + e) (only when NEED_LCH)
};

// SAFE = false: straight-line code on top of aug_fastmath.cuh, valid for 0 <= c <= 1e290 (exp(-c) is
// evaluated at min(c, 708): beyond that e < 1e-307 changes nothing in double precision); inv_c = 1/c is
// only read when c >= 1/16 and the c < 1/16 series is a select.
// SAFE = true: the same formulas with IEEE sqrt / div and the libdevice exp / log1p for any input.
// SERIES = false drops the c < 1/16 branch: for callers that route such c to the SAFE instantiation themselves.
template <bool NEED_LCH, bool SAFE, bool SERIES = true>
__device__ __forceinline__ PGTerms pg_terms_ic(double c, double inv_c) {
    PGTerms t;
    double e, inv;
    if (SAFE) {
        e = exp(-c);
        inv = 1.0 / (1.0 + e);
        t.h = (1.0 - e) * inv / (2.0 * c);
        t.l1pe = NEED_LCH ? log1p(e) : 0.0;
        t.lch = NEED_LCH ? fma(0.5, c, t.l1pe - LN2) : 0.0;
    } else {
        e = augf::exp_(-fmin(c, 708.0));
        const double d = 1.0 + e;
        inv = augf::rcp(d);
        t.h = (1.0 - e) * inv * (0.5 * inv_c);
        t.l1pe = NEED_LCH ? augf::log_1to2(d) : 0.0;
        t.lch = NEED_LCH ? fma(0.5, c, t.l1pe - LN2) : 0.0;
    }
    t.e = e;
    t.inv1pe = inv;
    if (SERIES && c < 0.0625) {
        // series in x = c/2 <= 1/32: truncation error < 1e-17 relative (also covers c == 0 -> 1/4)
        const double x2 = 0.25 * c * c;
        double p = fma(x2, 62.0 / 2835.0, -17.0 / 315.0);
        p = fma(x2, p, 2.0 / 15.0);
        p = fma(x2, p, -1.0 / 3.0);
        p = fma(x2, p, 1.0);          // tanh(x)/x
        t.h = 0.25 * p;
        if (NEED_LCH) {
            double q = fma(x2, -17.0 / 2520.0, 1.0 / 45.0);
            q = fma(x2, q, -1.0 / 12.0);
            q = fma(x2, q, 0.5);
            t.lch = x2 * q;           // logcosh(x)
        }
    }
    return t;
}

template <bool NEED_LCH, bool SAFE = true>
__device__ __forceinline__ PGTerms pg_terms(double c) {
    const double ic = SAFE ? 0.0 : augf::rcp(fmin(fmax(c, 0.0625), 1e290));
    return pg_terms_ic<NEED_LCH, SAFE>(c, ic);
}

// approx_expected_logistic(mu, c) = exp(mu/2) sech(c/2)/2, saturating on mu alone (utils.jl:11-14).
// sech(c/2)/2 = exp(-c/2)/(1+exp(-c)): one extra exp on top of pg_terms.
// SAFE = false needs |mu - c| <= 1416.
template <bool SAFE>
__device__ __forceinline__ double approx_expected_logistic(double mu, double c, const PGTerms& t) {
    const double v = (SAFE ? exp(0.5 * (mu - c)) : augf::exp_(0.5 * (mu - c))) * t.inv1pe;
    return mu < LOGISTIC_LO ? 0.0 : (mu > LOGISTIC_HI ? 1.0 : v);
}

// kldivergence(Poisson(q), Poisson(p)) (Distributions.jl): q == 0 ? p : p - q + q (log q - log p).
// SAFE = false: q log q is dropped below 1e-290 (it is < 1e-287 there).
// KL(Poisson(q) || Poisson(p)) when log(q/p) is already known (fused CAVI: q = p * sigma~, so
// log(q/p) = log sigma~ = (mu - c)/2 - log(1+e) comes for free); q == 0 -> p.
__device__ __forceinline__ double kl_poisson_lr(double q, double p, double log_q_over_p) {
    return q == 0.0 ? p : p - q + q * log_q_over_p;
}

template <bool SAFE>
__device__ __forceinline__ double kl_poisson(double q, double p, double logp) {
    if (SAFE) {
        if (q == 0.0) return p;
        return p - q + q * (log(q) - logp);
    }
    const double lq = augf::log_(fmax(q, 1e-290));
    return q < 1e-290 ? p - q : p - q + q * (lq - logp);
}

// logistic(x) with the LogExpFunctions saturation
__device__ __forceinline__ double logistic(double x) {
    if (x < LOGISTIC_LO) return 0.0;
    if (x > LOGISTIC_HI) return 1.0;
    const double e = exp(-fabs(x));
    const double r = 1.0 / (1.0 + e);
    return x >= 0.0 ? r : e * r;
}

// log y! / negbin log-constant with a table for small integer y (device lgamma beyond it)
__device__ __forceinline__ double lfact(double y, const double* __restrict__ table) {
    if (y < 0.0) return __longlong_as_double(0x7ff8000000000000ll);   // logfactorial(y < 0): DomainError in the reference
    if (table != nullptr && y < (double)AUG_TABLE_N) return __ldg(&table[(int)y]);
    return lgamma(y + 1.0);
}

}  // namespace augm
