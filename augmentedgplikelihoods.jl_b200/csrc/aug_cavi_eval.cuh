// aug_cavi_eval.cuh — per-observation closed forms of the CAVI update (aux_posterior! + expected potential /
// precision + ELBO terms) shared by the streaming kernels (aug_cavi.cu) and the sparse-GP sweep (aug_sparse.cu).
// Reference lines are cited next to each formula.
#pragma once
#include "aug_common.cuh"
#include "aug_math.cuh"

namespace {

template <int KIND> struct YT { typedef double T; };
template <> struct YT<AUG_BERNOULLI> { typedef uint8_t T; };
template <> struct YT<AUG_NEGBIN> { typedef int64_t T; };
template <> struct YT<AUG_POISSON> { typedef int64_t T; };

struct Obs {
    double y, ys;              // observation; state copy of y used for the PG shape (φ.y)
    double m, v, mg, vg;       // marginal moments of q(f) (and q(g))
    double s0, s1, s2;         // state (in when FROM_STATE, out otherwise)
    double b0, g0, b1, g1;     // E[β], E[γ] per latent
    double elt, kl;            // per-observation ELBO terms
};

// c = sqrt(second moment) (fused) or the stored c (FROM_STATE), and the PG terms of that tilt
template <bool FROM_STATE, bool ELBO, bool SAFE>
__device__ __forceinline__ augm::PGTerms terms(double s2, double& c) {
    if (FROM_STATE) return augm::pg_terms<ELBO, SAFE>(c);
    double ic = 0.0;
    if (SAFE) c = sqrt(s2);
    else augf::sqrt_inv(s2, c, ic);
    return augm::pg_terms_ic<ELBO, SAFE>(c, ic);
}

// Can the straight-line (SAFE = false) instantiation be used for this observation?  False for zero /
// denormal / huge / non-finite second moments, |m| beyond the exp range, integer y beyond the
// constant table — inputs the IEEE + libdevice instantiation handles instead.
template <int KIND, bool FROM_STATE, bool ELBO>
__device__ __forceinline__ bool fast_ok(const Obs& o) {
    bool ok = true;
    if (KIND == AUG_BERNOULLI || KIND == AUG_NEGBIN || KIND == AUG_POISSON) {
        if (FROM_STATE) ok = o.s0 >= 0.0 && o.s0 <= 700.0;
        if (!FROM_STATE || ELBO) {
            const double s2 = fma(o.m, o.m, o.v);
            ok = ok && augf::in_range(s2) && fabs(o.m) <= 700.0 && o.v <= 4e5;
            // approx_expected_logistic evaluates exp_((-m-c)/2): s2 <= 2.4e5 keeps |m|, c <= 490
            if (KIND == AUG_POISSON) ok = ok && s2 <= 2.4e5;
        }
        if (ELBO && KIND != AUG_BERNOULLI) ok = ok && o.y >= 0.0 && o.y < (double)AUG_TABLE_N;   // y < 0: DomainError in the reference -> NaN on the SAFE path
        if (KIND == AUG_POISSON && FROM_STATE) ok = ok && o.s1 >= 0.0 && o.s1 <= 1e290;
    } else if (KIND == AUG_LAPLACE || KIND == AUG_STUDENTT) {
        const double d = o.m - o.y;
        if (!FROM_STATE || ELBO) ok = augf::in_range(fma(d, d, o.v));
        if (FROM_STATE) ok = ok && o.s0 >= 1e-290 && o.s0 <= 1e290;
    } else if (KIND == AUG_HETERO) {
        const double d = o.m - o.y;
        if (FROM_STATE) ok = o.s0 >= 0.0 && o.s0 <= 700.0 && o.s1 >= 0.0 && o.s1 <= 1e290;
        if (!FROM_STATE || ELBO) ok = ok && augf::in_range(fma(o.mg, o.mg, o.vg)) && fma(o.mg, o.mg, o.vg) <= 2.4e5 &&
                                      augf::in_range(fma(d, d, o.v));
        ok = ok && fabs(o.mg) <= 490.0;
    }
    return ok;
}

// ------------------------------------------------------------------ per-kind closed forms
template <int KIND, bool FROM_STATE, bool ELBO, bool SAFE>
__device__ __forceinline__ void eval(const LikConst& L, Obs& o) {
    using namespace augm;
    o.elt = 0.0;
    o.kl = 0.0;
    o.b1 = o.g1 = 0.0;
    if (KIND == AUG_BERNOULLI) {
        const double s2m = fma(o.m, o.m, o.v);                  // second_moment utils.jl:1-3
        const PGTerms t = terms<FROM_STATE, ELBO, SAFE>(s2m, o.s0);   // c = sqrt(.)  bernoulli.jl:23
        const double sg = o.y > 0.5 ? 0.5 : -0.5;               // sign(y - 0.5)/2  :28
        o.b0 = sg;
        o.g0 = t.h;                                             // mean(PG(1,c))    :44
        if (ELBO) {
            o.elt = -LN2 + fma(sg, o.m, -0.5 * s2m * t.h);      // :62-64
            o.kl = fma(-0.5 * o.s0 * o.s0, t.h, t.lch);         // KL(PG(1,c)||PG(1,0))
        }
    } else if (KIND == AUG_NEGBIN) {
        const double s2m = fma(o.m, o.m, o.v);
        const PGTerms t = terms<FROM_STATE, ELBO, SAFE>(s2m, o.s0);   // negativebinomial.jl:29-31
        const double r = L.p0;
        const double b = o.ys + r;
        const double th = b * t.h;                              // mean(PG(y+r,c)) :48
        o.b0 = 0.5 * (o.y - r);                                 // :36
        o.g0 = th;
        if (ELBO) {
            double lc;                                          // negbin_logconst :51-52
            if (!SAFE) lc = __ldg(&L.table[(int)fmin(fmax(o.y, 0.0), (double)(AUG_TABLE_N - 1))]);
            else if (o.y < 0.0) lc = __longlong_as_double(0x7ff8000000000000ll);   // loggamma / binomial of a negative count
            else if (o.y < (double)AUG_TABLE_N) lc = __ldg(&L.table[(int)o.y]);
            else lc = lgamma(o.y + r) - lgamma(o.y + 1.0) - L.c0;
            o.elt = lc - (o.y + r) * LN2 + 0.5 * fma(o.m, o.y - r, -s2m * th);   // :62-64
            o.kl = fma(b, t.lch, -0.5 * o.s0 * o.s0 * th);
        }
    } else if (KIND == AUG_POISSON) {
        const double s2m = fma(o.m, o.m, o.v);
        const PGTerms t = terms<FROM_STATE, ELBO, SAFE>(s2m, o.s0);   // poisson.jl:35
        if (!FROM_STATE) o.s1 = L.p0 * approx_expected_logistic<SAFE>(-o.m, o.s0, t);   // :37
        const double lam = o.s1;
        const double b = o.ys + lam;
        const double th = b * t.h;                              // polyagammapoisson.jl:35-41
        o.b0 = 0.5 * (o.y - lam);                               // poisson.jl:59
        o.g0 = th;
        if (ELBO) {
            double lf;                                          // logfactorial(y)
            if (!SAFE) lf = __ldg(&L.table[(int)fmin(fmax(o.y, 0.0), (double)(AUG_TABLE_N - 1))]);
            else lf = lfact(o.y, L.table);
            o.elt = -(o.y + lam) * LN2 + 0.5 * fma(o.y - lam, o.m, -s2m * th) + o.y * L.c0 - lf;   // :81-83
            // KL(Po(λ̂) || Po(λ)): fused, log(λ̂/λ) = log σ̃ = (−m − c)/2 − log(1+e) needs no extra log
            double klp;
            if (FROM_STATE) klp = kl_poisson<SAFE>(lam, L.p0, L.c0);
            else {
                const double mneg = -o.m;
                const double lst = mneg > LOGISTIC_HI ? 0.0 : fma(0.5, mneg - o.s0, -t.l1pe);
                klp = kl_poisson_lr(lam, L.p0, lst);
            }
            o.kl = fma(b, t.lch, -0.5 * o.s0 * o.s0 * th) + klp;
        }
    } else if (KIND == AUG_LAPLACE) {
        const double d = o.m - o.y;
        const double s2y = fma(d, d, o.v);                      // second_moment(q, y) utils.jl:5-7
        if (!FROM_STATE) {                                      // 1/(2β sqrt(.))  laplace.jl:48-50
            if (SAFE) {
                o.s0 = 1.0 / (2.0 * L.p0 * sqrt(s2y));
            } else {
                double c_, ic_;
                augf::sqrt_inv(s2y, c_, ic_);
                o.s0 = L.c0 * ic_;
            }
        }
        o.b0 = 2.0 * o.s0 * o.y;                                // :63
        o.g0 = 2.0 * o.s0;                                      // :67
        if (ELBO) {
            o.elt = L.c2 - s2y * o.s0;                          // :84-87
            o.kl = L.c3 + (SAFE ? L.c1 / o.s0 : L.c1 * augf::rcp(o.s0));   // :98-104
        }
    } else if (KIND == AUG_STUDENTT) {
        const double d = o.m - o.y;
        const double s2y = fma(d, d, o.v);
        if (!FROM_STATE) o.s0 = 0.5 * (L.c0 + s2y);             // studentt.jl:54-56
        const double ib = SAFE ? 1.0 / o.s0 : augf::rcp(o.s0);
        const double th = L.c1 * ib;                            // mean(Gamma(α, 1/β)) :69,73
        o.b0 = th * o.y;
        o.g0 = th;
        if (ELBO) {
            // logpdf(Normal(y, θ^-1/2), m) - vθ/2 = -log(2π)/2 + log(θ)/2 - θ((m-y)² + v)/2   :80-83
            // KL(Gamma(α, 1/β) || Gamma(ν/2, 2σ²/ν)), Distributions.jl closed form
            const double r = ib / L.c4;
            const double lth = SAFE ? log(th) : augf::log_(th);
            const double lr = SAFE ? log(r) : lth - L.p1;       // log r = log θ − log(α θq)  (L.p1 holds it)
            o.elt = L.c5 + 0.5 * lth - 0.5 * th * s2y;
            o.kl = L.c2 - L.c3 * lr + L.c1 * r;
        }
    } else if (KIND == AUG_HETERO) {
        const double d = o.m - o.y;
        const double s2f = fma(d, d, o.v);
        const double s2g = fma(o.mg, o.mg, o.vg);
        if (!FROM_STATE) o.s2 = 0.5 * s2f;                      // ψ  heteroscedasticgaussian.jl:42
        const PGTerms t = terms<FROM_STATE, ELBO, SAFE>(s2g, o.s0);   // c  :43
        const double st = approx_expected_logistic<SAFE>(-o.mg, o.s0, t);
        if (!FROM_STATE) o.s1 = L.p0 * st * o.s2;               // λ  :44
        const double lam = o.s1;
        const double lsg = L.p0 * (1.0 - st);                   // :102
        const double b = 0.5 + lam;
        const double th = b * t.h;
        o.b0 = 0.5 * o.y * lsg;
        o.b1 = 0.5 * (0.5 - lam);
        o.g0 = lsg;
        o.g1 = th;
        if (ELBO) {                                             // :129-145
            o.elt = L.c0 - b * LN2 + 0.5 * fma(0.5 - lam, o.mg, -s2g * th);
            const double pl = 0.5 * L.p0 * s2f;
            double klp;   // KL(Po(λ̂) || Po(λψ)) with λ̂ = λψσ̃: log(λ̂/(λψ)) = log σ̃, no log needed when fused
            if (FROM_STATE) klp = kl_poisson<SAFE>(lam, pl, SAFE ? log(pl) : augf::log_(pl));
            else {
                const double mneg = -o.mg;
                const double lst = mneg > LOGISTIC_HI ? 0.0 : fma(0.5, mneg - o.s0, -t.l1pe);
                klp = kl_poisson_lr(lam, pl, lst);
            }
            o.kl = fma(b, t.lch, -0.5 * o.s0 * o.s0 * th) + klp;
        }
    }
}

// s2 holds ψ (double) for HETERO and a copy of y (int64) for NEGBIN / POISSON
template <int KIND> struct S2T { typedef double T; };
template <> struct S2T<AUG_NEGBIN> { typedef int64_t T; };
template <> struct S2T<AUG_POISSON> { typedef int64_t T; };

}  // namespace
