// aug_loglik.cu — sampled log-likelihood terms (logtilt, logdensity of the auxiliary prior,
// aug_loglik) and the element-wise SpecialDistributions primitives exported for callers/tests.
//
// Reference behaviour replaced (paths relative to /root/reference/src):
//   generic.jl:40-50 (logtilt sum, aug_loglik), likelihoods/*.jl logtilt + aux_prior methods,
//   heteroscedasticgaussian.jl:117-127, SpecialDistributions/polyagamma.jl:25-31 (mean),
//   :37-91 (logpdf: 101 paired terms of the alternating series; log-domain variant for x < 1e-2),
//   :99-110 (KL), polyagammapoisson.jl:29-33, polyagammanegativemultinomial.jl:33-39,
//   negativemultinomial.jl:47-52, utils.jl:11-14.
#include "aug_common.cuh"
#include "aug_math.cuh"

namespace {

constexpr double LOG2PI = 1.83787706640934548356;

__device__ __forceinline__ double log1mexp_dev(double x) {   // LogExpFunctions.log1mexp
    return x < -augm::LN2 ? log1p(-exp(x)) : log(-expm1(x));
}

// logpdf(PolyaGamma(b, c), x)  polyagamma.jl:37-91.  The loops stop once a term is below 2^-70 of the running sum (or
// underflows to an exact zero): the terms then fall super-exponentially, each later one is far below half an ulp of the
// sum, so the reference's 101-pair sum is unchanged bit for bit — 3 pairs instead of 10 at x = 1/4.
__device__ double pg_logpdf_dev(double b, double c, double x) {
    if (b == 0.0) return x == 0.0 ? 0.0 : -INFINITY;
    const double hc = 0.5 * fabs(c);
    const double lch = hc < 0.03125 ? augm::pg_terms<true>(fabs(c)).lch : hc + log1p(exp(-2.0 * hc)) - augm::LN2;
    const double ext = b * lch - 0.5 * c * c * x + (b - 1.0) * augm::LN2 - 0.5 * (LOG2PI + 3.0 * log(x));
    const double bm1 = b - 1.0;
    const double inv8x = -1.0 / (8.0 * x), inv2x = -1.0 / (2.0 * x);
    if (x < 1e-2) {   // calc_log_series :55-73 with an online logsumexp
        double lprod = 0.0, mx = -INFINITY, acc = 0.0;
        for (int n = 0; n <= 200; n += 2) {
            const double Rn = 2.0 * n + b;
            const double log_c_nb = log(n + b) - log(n + 1.0) + log(2.0 / Rn + 1.0);
            const double v = lprod + log(Rn) + Rn * Rn * inv8x + log1mexp_dev(log_c_nb + (Rn + 1.0) * inv2x);
            if (v > mx) { acc = acc * exp(mx - v) + 1.0; mx = v; }
            else acc += exp(v - mx);
            if (v < mx - 50.0) break;                      // exp(-50) < 2^-70
            lprod += log(1.0 + bm1 / (n + 1.0)) + log(1.0 + bm1 / (n + 2.0));
        }
        return ext + mx + log(acc);
    }
    double prod = 1.0, sum = 0.0;   // calc_series :75-91
    for (int n = 0; n <= 200; n += 2) {
        const double Rn = 2.0 * n + b;
        const double ea = Rn * Rn * inv8x;
        if (ea < -746.0) break;
        const double c_nb = ((n + b) / (n + 1.0)) * (2.0 / Rn + 1.0);
        const double inner = 1.0 - c_nb * exp((Rn + 1.0) * inv2x);
        const double term = prod * Rn * exp(ea) * inner;
        sum += term;
        if (ea < -48.6 && fabs(term) < fabs(sum) * 0x1.0p-70) break;   // (inner < 0 for x >> b: compare magnitudes)
        prod *= (1.0 + bm1 / (n + 1.0)) * (1.0 + bm1 / (n + 2.0));
    }
    return ext + log(fmax(sum, 2.2250738585072014e-308));
}

__device__ __forceinline__ double poislogpdf(double lam, double loglam, double x) {   // StatsFuns.poislogpdf
    return (x == 0.0 ? 0.0 : x * loglam) - lam - lgamma(x + 1.0);
}

struct LLArgs {
    int64_t n;
    const void* y;
    const double* f;
    const double* g;
    const double* omega;
    const int64_t* nvar;
    int with_prior;
    double k0, k1, k2, k3;   // prior constants per kind
    double* partials;
    unsigned int* counter;
    double* scalars;
    AugXchDev* xch;          // fused multi-GPU mode: all-reduce over peer memory in the finaliser
    LikConst L;
};

template <int KIND>
__global__ void __launch_bounds__(AUG_BLOCK) loglik_kernel(const LLArgs a) {
    using namespace augm;
    const int64_t nth = (int64_t)gridDim.x * blockDim.x;
    double acc[2] = {0.0, 0.0};
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += nth) {
        const double w = a.omega[i], f = a.f[i];
        double lt = 0.0, lp = 0.0;
        if (KIND == AUG_BERNOULLI) {                          // bernoulli.jl:47-49,57
            const double sg = reinterpret_cast<const uint8_t*>(a.y)[i] ? 0.5 : -0.5;
            lt = -LN2 + fma(sg, f, -0.5 * f * f * w);
            if (a.with_prior) lp = pg_logpdf_dev(1.0, 0.0, w);
        } else if (KIND == AUG_NEGBIN) {                      // negativebinomial.jl:54-57,73
            const double y = (double)reinterpret_cast<const int64_t*>(a.y)[i], r = a.L.p0;
            double lc;
            if (y < 0.0) lc = __longlong_as_double(0x7ff8000000000000ll);
            else if (y < (double)AUG_TABLE_N) lc = __ldg(&a.L.table[(int)y]);
            else lc = lgamma(y + r) - lgamma(y + 1.0) - a.L.c0;
            lt = lc - (y + r) * LN2 + 0.5 * (f * (y - r) - f * f * w);
            if (a.with_prior) lp = pg_logpdf_dev(r + y, 0.0, w);
        } else if (KIND == AUG_POISSON) {                     // poisson.jl:62-65,74; polyagammapoisson.jl:29-33
            const double y = (double)reinterpret_cast<const int64_t*>(a.y)[i], nn = (double)a.nvar[i];
            lt = y * a.L.c0 - (y + nn) * LN2 - lfact(y, a.L.table) + 0.5 * ((y - nn) * f - f * f * w);
            if (a.with_prior) lp = pg_logpdf_dev(y + nn, 0.0, w) + poislogpdf(a.L.p0, a.L.c0, nn);
        } else if (KIND == AUG_LAPLACE) {                     // laplace.jl:79-81,96
            const double d = reinterpret_cast<const double*>(a.y)[i] - f;
            lt = a.L.c2 - d * d * w;
            // logpdf(InverseGamma(1/2, λ), ω) = log(λ)/2 − lgamma(1/2) − (3/2) log ω − λ/ω
            if (a.with_prior) lp = a.k0 - 1.5 * log(w) - a.L.c1 / w;
        } else if (KIND == AUG_STUDENTT) {                    // studentt.jl:76-78,91
            const double d = reinterpret_cast<const double*>(a.y)[i] - f;
            const double lw = log(w);
            lt = -0.5 * (d * d * w + LOG2PI) + 0.5 * lw;      // logpdf(Normal(f, ω^-1/2), y)
            // logpdf(Gamma(k = ν/2, θ = 2σ²/ν), ω) = −lgamma(k) − k log θ + (k−1) log ω − ω/θ
            if (a.with_prior) lp = a.k0 + (a.L.c3 - 1.0) * lw - w / a.L.c4;
        } else if (KIND == AUG_HETERO) {                      // heteroscedasticgaussian.jl:117-127
            const double gg = a.g[i], nn = (double)a.nvar[i];
            lt = -(0.5 + nn) * LN2 + 0.5 * ((0.5 - nn) * gg - gg * gg * w);
            if (a.with_prior) {
                const double d = reinterpret_cast<const double*>(a.y)[i] - f;
                const double pl = 0.5 * a.L.p0 * d * d;
                lp = pg_logpdf_dev(0.5 + nn, 0.0, w) + poislogpdf(pl, log(pl), nn);
            }
        }
        acc[0] += lt;
        acc[1] += lp;
    }
    double out[2];
    if (block_reduce_and_finalize<2>(acc, a.partials, a.counter, out)) {
        if (a.xch) xch_allreduce<2>(a.xch, out);
        a.scalars[AUG_S_LOGTILT] = out[0];
        a.scalars[AUG_S_LOGPRIOR] = out[1];
        a.scalars[AUG_S_AUGLL] = out[0] + out[1];             // generic.jl:48-50
        scal_zero_except(a.scalars, 0x38u);
    }
}

// CAT: one warp per row.  categorical.jl:138-163; pgnm.jl:33-39 (all classes — the reference sums
// only 1:2); negativemultinomial.jl:47-52 with x₀ = 1.
struct CatLLArgs {
    int64_t n;
    int nl;
    const uint8_t* y;
    const double* f;
    const double* omega;
    const int64_t* nvar;
    int with_prior;
    int quirks;              // AUG_LIK_FAITHFUL_QUIRKS: PG log-densities of classes 1:2 only (pgnm.jl:35)
    double log_prior_p, log_p0;
    double* partials;
    unsigned int* counter;
    double* scalars;
    AugXchDev* xch;          // fused multi-GPU mode: all-reduce over peer memory in the finaliser
};

__global__ void __launch_bounds__(AUG_BLOCK) cat_loglik_kernel(const CatLLArgs a) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    double acc[2] = {0.0, 0.0};
    for (int64_t i = warp; i < a.n; i += nwarps) {
        const int64_t base = i * a.nl;
        double lt = 0.0, lp = 0.0, sn = 0.0;
        for (int j = lane; j < a.nl; j += 32) {
            const double y = (double)a.y[base + j], nn = (double)a.nvar[base + j];
            const double f = a.f[base + j], w = a.omega[base + j];
            lt += -(y + nn) * augm::LN2 + 0.5 * ((y - nn) * f - f * f * w);
            if (a.with_prior) {
                if (!a.quirks || j < 2) lp += pg_logpdf_dev(y + nn, 0.0, w);
                lp += (nn == 0.0 ? 0.0 : nn * a.log_prior_p) - lgamma(nn + 1.0);
                sn += nn;
            }
        }
        acc[0] += lt;
        if (a.with_prior) {
            sn = warp_sum(sn);
            if (lane == 0) lp += lgamma(1.0 + sn) + a.log_p0;
            acc[1] += lp;
        }
    }
    double out[2];
    if (block_reduce_and_finalize<2>(acc, a.partials, a.counter, out)) {
        if (a.xch) xch_allreduce<2>(a.xch, out);
        a.scalars[AUG_S_LOGTILT] = out[0];
        a.scalars[AUG_S_LOGPRIOR] = out[1];
        a.scalars[AUG_S_AUGLL] = out[0] + out[1];
        scal_zero_except(a.scalars, 0x38u);
    }
}

// ---- element-wise primitives
__global__ void __launch_bounds__(AUG_BLOCK) ew_kernel(int op, int64_t n, const double* p, const double* q,
                                                        const double* r, double s0, double s1, double* out) {
    const int64_t nth = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += nth) {
        if (op == 0) {                // mean(PolyaGamma(b,c)) polyagamma.jl:25-31
            out[i] = p[i] * augm::pg_terms<false>(fabs(q[i])).h;
        } else if (op == 1) {         // KL(PG(b,c)||PG(b,0)) polyagamma.jl:99-110
            const double c = fabs(q[i]);
            const augm::PGTerms t = augm::pg_terms<true>(c);
            out[i] = p[i] * fma(-0.5 * c * c, t.h, t.lch);
        } else if (op == 2) {         // logpdf(PolyaGamma(s0, s1), x)
            out[i] = pg_logpdf_dev(s0, s1, p[i]);
        } else if (op == 3) {         // approx_expected_logistic(mu, c) utils.jl:11-14
            const double c = fabs(q[i]);
            out[i] = augm::approx_expected_logistic<true>(p[i], c, augm::pg_terms<false>(c));
        } else if (op == 4) {         // second_moment(q[, y]) utils.jl:1-7
            const double d = r ? p[i] - r[i] : p[i];
            out[i] = fma(d, d, q[i]);
        } else {                      // op >= 10: the straight-line functions of aug_fastmath.cuh (accuracy tests)
            const double x = p[i];
            double c_, ic_;
            switch (op) {
                case 10: out[i] = augf::rcp(x); break;
                case 11: out[i] = augf::rsqrt_(x); break;
                case 12: out[i] = augf::exp_(x); break;
                case 13: out[i] = augf::log_(x); break;
                case 14: out[i] = augf::log_1to2(x); break;
                default: augf::sqrt_inv(x, c_, ic_); out[i] = c_; break;
            }
        }
    }
}

template <typename K, typename A>
int32_t launch_red(aug_ctx* ctx, K kernel, const A& a, int64_t items) {
    const int grid = aug_grid_for(ctx, (const void*)kernel, items, AUG_BLOCK);
    kernel<<<grid, AUG_BLOCK, 0, ctx->stream>>>(a);
    ctx->launches++;
    return (int32_t)cudaGetLastError();
}

int32_t ew_launch(aug_ctx* c, int op, int64_t n, const double* p, const double* q, double s0, double s1,
                  double* out, const double* r = nullptr) {
    if (!c) return AUG_ERR_NOT_INIT;
    if (n < 0 || !p || !out) return AUG_ERR_BAD_ARG;
    AUG_CUDA(cudaSetDevice(c->device));
    if (n == 0) return AUG_OK;
    const int grid = aug_grid_for(c, (const void*)ew_kernel, n, AUG_BLOCK);
    ew_kernel<<<grid, AUG_BLOCK, 0, c->stream>>>(op, n, p, q, r, s0, s1, out);
    c->launches++;
    return (int32_t)cudaGetLastError();
}

}  // namespace

extern "C" {

int32_t aug_sampled_loglik_terms(aug_ctx* c, const aug_lik* lik, int64_t n, const void* y, const double* f,
                                 int64_t ld, const double* omega, const int64_t* nvar, int32_t with_prior,
                                 double* scalars) {
    if (!c) return AUG_ERR_NOT_INIT;
    if (!lik || n < 0 || !y || !f || !omega || !scalars) return AUG_ERR_BAD_ARG;
    AUG_CUDA(cudaSetDevice(c->device));
    if (aug_xch_for(c)) {                    // exchanges happen in call order: complete a pending split-phase one first
        int32_t rf = aug_xch_flush(c);
        if (rf) return rf;
    }
    if (n == 0) {
        AUG_CUDA(cudaMemsetAsync(scalars, 0, AUG_NSCALARS * sizeof(double), c->stream));
        if (aug_xch_for(c)) return aug_xch_zero_contribution(c, scalars, AUG_S_LOGTILT, 2);
        return AUG_OK;
    }
    if (lik->kind == AUG_CAT || lik->kind == AUG_CAT_BIJ) {
        if (!nvar) return AUG_ERR_BAD_ARG;
        // aux_prior of the non-bijective link is NegativeMultinomial(1, fill(1/nl, nl)) (categorical.jl:158-163,
        // "we just pretend this will work"): sum(p) = 1 fails the constructor check (negativemultinomial.jl:18)
        if (with_prior && lik->kind == AUG_CAT) return AUG_ERR_PRECONDITION;
        LikConst L;
        int32_t rc = aug_lik_const(c, lik, &L, false, false);
        if (rc) return rc;
        CatLLArgs a{};
        a.n = n;
        a.nl = lik->nlatent;
        a.y = (const uint8_t*)y;
        a.f = f;
        a.omega = omega;
        a.nvar = nvar;
        a.with_prior = with_prior;
        a.quirks = L.quirks;
        if (with_prior && L.quirks && lik->nlatent < 2) return AUG_ERR_PRECONDITION;   // BoundsError in the reference
        a.log_prior_p = L.c2;
        a.log_p0 = L.c3;
        a.partials = c->partials;
        a.counter = c->counter;
        a.scalars = scalars;
    a.xch = aug_xch_for(c);
        return launch_red(c, cat_loglik_kernel, a, n * 32);
    }
    LLArgs a{};
    a.n = n;
    a.y = y;
    a.f = f;
    a.omega = omega;
    a.nvar = nvar;
    a.with_prior = with_prior;
    a.partials = c->partials;
    a.counter = c->counter;
    a.scalars = scalars;
    a.xch = aug_xch_for(c);
    int32_t rc = aug_lik_const(c, lik, &a.L, true, false);
    if (rc) return rc;
    if ((lik->kind == AUG_POISSON || lik->kind == AUG_HETERO) && !nvar) return AUG_ERR_BAD_ARG;
    switch (lik->kind) {
        case AUG_BERNOULLI: return launch_red(c, loglik_kernel<AUG_BERNOULLI>, a, n);
        case AUG_NEGBIN: return launch_red(c, loglik_kernel<AUG_NEGBIN>, a, n);
        case AUG_POISSON: return launch_red(c, loglik_kernel<AUG_POISSON>, a, n);
        case AUG_LAPLACE:
            a.k0 = 0.5 * log(a.L.c1) - lgamma(0.5);
            return launch_red(c, loglik_kernel<AUG_LAPLACE>, a, n);
        case AUG_STUDENTT:
            a.k0 = -lgamma(a.L.c3) - a.L.c3 * log(a.L.c4);
            return launch_red(c, loglik_kernel<AUG_STUDENTT>, a, n);
        case AUG_HETERO:
            if (ld < n) return AUG_ERR_BAD_ARG;
            a.g = f + ld;
            return launch_red(c, loglik_kernel<AUG_HETERO>, a, n);
        default: return AUG_ERR_BAD_KIND;
    }
}

int32_t aug_pg_mean(aug_ctx* c, int64_t n, const double* b, const double* cc, double* out) {
    if (!cc) return AUG_ERR_BAD_ARG;
    return ew_launch(c, 0, n, b, cc, 0, 0, out);
}
int32_t aug_pg_kl(aug_ctx* c, int64_t n, const double* b, const double* cc, double* out) {
    if (!cc) return AUG_ERR_BAD_ARG;
    return ew_launch(c, 1, n, b, cc, 0, 0, out);
}
int32_t aug_pg_logpdf(aug_ctx* c, int64_t n, double b, double cc, const double* x, double* out) {
    return ew_launch(c, 2, n, x, nullptr, b, cc, out);
}
int32_t aug_approx_expected_logistic(aug_ctx* c, int64_t n, const double* mu, const double* cc, double* out) {
    if (!cc) return AUG_ERR_BAD_ARG;
    return ew_launch(c, 3, n, mu, cc, 0, 0, out);
}

/* diagnostics: evaluate one of the aug_fastmath.cuh functions element-wise (0 rcp, 1 rsqrt, 2 exp, 3 log,
 * 4 log on [1,2], 5 sqrt) on arguments inside its stated range */
int32_t aug_fastmath_eval(aug_ctx* c, int32_t fn, int64_t n, const double* x, double* out) {
    if (fn < 0 || fn > 5) return AUG_ERR_BAD_ARG;
    return ew_launch(c, 10 + fn, n, x, nullptr, 0, 0, out);
}

int32_t aug_second_moment(aug_ctx* c, int64_t n, const double* mu, const double* var, const double* y,
                          double* out) {
    if (!var) return AUG_ERR_BAD_ARG;
    return ew_launch(c, 4, n, mu, var, 0, 0, out, y);
}

}  // extern "C"
