// aug_next.cu — the callers on either side of the hot path:
//   * the sums behind the likelihood-parameter update of the heteroscedastic model, which reuse the per-observation
//     quantities of the CAVI / Gibbs step (ψ, σ̃_g resp. σ(g), (y − f)²):
//       opt_lik, examples/heteroscedasticgaussian/script.jl:41-51              (variational)
//       Gamma full conditional of λ, docs/src/likelihoods/heteroscedasticgaussian.md:80-84   (Gibbs)
//   * the prediction-side links: logisticsoftmax / LogisticSoftMaxLink(f) (likelihoods/categorical.jl:1-4, 32-35,
//     with the BijectiveSimplexLink's appended zero latent, :12-14) and approx_expected_logisticsoftmax
//     (utils.jl:17-22).
// Both reductions are HBM-bound map/reduce passes (40 resp. 24 B per observation) on the same block-tree /
// last-block / peer-mailbox plumbing as the CAVI kernels; the links are one-warp-per-row kernels.
#include <math.h>
#include <stdlib.h>

#include "aug_common.cuh"
#include "aug_math.cuh"

namespace {

struct LamArgs {
    int64_t n;
    const double* y;
    const double* a0;   // variational: mu_f      sampled: f
    const double* a1;   // variational: var_f     sampled: g
    const double* a2;   // variational: mu_g
    const double* a3;   // variational: var_g
    double* partials;
    unsigned int* counter;
    double* out;
    AugXchDev* xch;
};

// ψ (1 − σ̃_g) with ψ = ((μ_f − y)² + σ²_f)/2, c = sqrt(μ_g² + σ²_g), σ̃_g = approx_expected_logistic(−μ_g, c)
__device__ __forceinline__ double lam_term_vi(double y, double mf, double vf, double mg, double vg) {
    const double d = mf - y;
    const double psi = 0.5 * fma(d, d, vf);                 // second_moment(qf − y)/2   utils.jl:5-7
    const double s2 = fma(mg, mg, vg);                      // second_moment(qg)         utils.jl:1-3
    const unsigned hi = (unsigned)__double2hiint(s2);
    double sg;
    if (hi - 0x3f700000u < 0x410d4c00u - 0x3f700000u) {     // 2^-8 <= s2 < 2.4e5: straight-line evaluation
        double c, ic;
        augf::sqrt_inv(s2, c, ic);
        const double e = augf::exp_(-c);
        const double w = augf::exp_(0.5 * (-mg - c)) * augf::rcp(1.0 + e);   // exp(−μ_g/2) sech(c/2)/2
        sg = -mg > augm::LOGISTIC_HI ? 1.0 : w;
    } else {
        const double c = sqrt(s2);
        sg = augm::approx_expected_logistic<true>(-mg, c, augm::pg_terms<false>(c));
    }
    return psi * (1.0 - sg);
}

// σ(g)/2 · (y − f)²
__device__ __forceinline__ double lam_term_gibbs(double y, double f, double g) {
    const double d = y - f;
    return 0.5 * augm::logistic(g) * d * d;
}

template <bool SAMPLED, bool VEC>
__global__ void __launch_bounds__(AUG_BLOCK) hetero_lambda_kernel(const LamArgs a) {
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t nth = (int64_t)gridDim.x * blockDim.x;
    double acc[1] = {0.0};
    if (VEC) {
        const int64_t np = a.n >> 1;
        for (int64_t p = tid; p < np; p += nth) {
            const double2 y = ld_stream2(a.y + 2 * p), u = ld_stream2(a.a0 + 2 * p), v = ld_stream2(a.a1 + 2 * p);
            if (SAMPLED) {
                acc[0] += lam_term_gibbs(y.x, u.x, v.x) + lam_term_gibbs(y.y, u.y, v.y);
            } else {
                const double2 mg = ld_stream2(a.a2 + 2 * p), vg = ld_stream2(a.a3 + 2 * p);
                acc[0] += lam_term_vi(y.x, u.x, v.x, mg.x, vg.x) + lam_term_vi(y.y, u.y, v.y, mg.y, vg.y);
            }
        }
        if ((a.n & 1) && tid == 0) {
            const int64_t i = a.n - 1;
            acc[0] += SAMPLED ? lam_term_gibbs(a.y[i], a.a0[i], a.a1[i])
                              : lam_term_vi(a.y[i], a.a0[i], a.a1[i], a.a2[i], a.a3[i]);
        }
    } else {
        for (int64_t i = tid; i < a.n; i += nth)
            acc[0] += SAMPLED ? lam_term_gibbs(a.y[i], a.a0[i], a.a1[i])
                              : lam_term_vi(a.y[i], a.a0[i], a.a1[i], a.a2[i], a.a3[i]);
    }
    double out[1];
    if (block_reduce_and_finalize<1>(acc, a.partials, a.counter, out)) {
        if (a.xch) xch_allreduce<1>(a.xch, out);
        a.out[0] = out[0];
    }
}

__global__ void xch_one_kernel(AugXchDev* x, double* out) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    double v[1] = {0.0};
    xch_allreduce<1>(x, v);
    out[0] = v[0];
}

template <bool SAMPLED>
int32_t launch_lambda(aug_ctx* c, LamArgs& a) {
    const bool vec = aug_aligned16(a.y) && aug_aligned16(a.a0) && aug_aligned16(a.a1) && aug_aligned16(a.a2) &&
                     aug_aligned16(a.a3);
    const void* k = vec ? (const void*)hetero_lambda_kernel<SAMPLED, true> : (const void*)hetero_lambda_kernel<SAMPLED, false>;
    const int grid = aug_grid_for(c, k, vec ? (a.n + 1) / 2 : a.n, AUG_BLOCK * 4);
    if (vec) hetero_lambda_kernel<SAMPLED, true><<<grid, AUG_BLOCK, 0, c->stream>>>(a);
    else hetero_lambda_kernel<SAMPLED, false><<<grid, AUG_BLOCK, 0, c->stream>>>(a);
    c->launches++;
    return (int32_t)cudaGetLastError();
}

// ---------------------------------------------------------------- links: one warp per row
struct LinkArgs {
    int64_t n;
    int nl, K;              // latents per row, classes per output row
    const double* f;        // [n][nl]  (logisticsoftmax: f;  approx_expected: mu)
    const double* c;        // [n][nl]  approx_expected only
    const double* theta;    // device, K entries exp(logθ_j)
    double* out;            // [n][K] resp. [n][nl]
};

// EXPECTED = false: out[i][j] = θ_j σ(f_ij) / Σ_k θ_k σ(f_ik), the appended latent of the bijective link being 0
// EXPECTED = true:  out[i][j] = θ_j σ̃_ij / (θ_K σ(0) + Σ_k θ_k σ̃_ik),  σ̃ = approx_expected_logistic(μ, c)
template <bool EXPECTED>
__global__ void __launch_bounds__(AUG_BLOCK) link_kernel(const LinkArgs a) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int nl = a.nl, K = a.K;
    for (int64_t i = warp; i < a.n; i += nwarps) {
        const double* fr = a.f + i * nl;
        auto term = [&](int j) {
            if (EXPECTED) {
                const double cc = fabs(a.c[i * nl + j]);
                return __ldg(a.theta + j) * augm::approx_expected_logistic<true>(fr[j], cc, augm::pg_terms<false>(cc));
            }
            return __ldg(a.theta + j) * augm::logistic(j < nl ? fr[j] : 0.0);
        };
        const int terms = EXPECTED ? nl : K;
        double s = 0.0;
        for (int j = lane; j < terms; j += 32) s += term(j);
        s = warp_sum(s);
        if (EXPECTED) s += __ldg(a.theta + nl) * 0.5;       // θ[end] * logistic(0)
        double* o = a.out + i * (EXPECTED ? nl : K);
        for (int j = lane; j < terms; j += 32) o[j] = term(j) / s;
    }
}

int32_t upload_theta(aug_ctx* c, const aug_lik* lik, int K, const double** dtheta) {
    if (c->dtheta_cap < K) {
        if (c->dtheta) cudaFree(c->dtheta);
        c->dtheta = nullptr;
        c->dtheta_cap = 0;
        AUG_CUDA(cudaMalloc(&c->dtheta, sizeof(double) * K));
        c->dtheta_cap = K;
    }
    double* h = (double*)malloc(sizeof(double) * K);
    if (!h) return AUG_ERR_BAD_ARG;
    for (int j = 0; j < K; ++j) h[j] = exp(lik->logtheta ? lik->logtheta[j] : 0.0);
    cudaError_t e = cudaMemcpyAsync(c->dtheta, h, sizeof(double) * K, cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    free(h);
    if (e != cudaSuccess) return (int32_t)e;
    *dtheta = c->dtheta;
    return AUG_OK;
}

}  // namespace

extern "C" {

int32_t aug_hetero_lambda_stats(aug_ctx* c, int64_t n, const double* y, const double* mu, const double* var,
                                int64_t ld, double* out) {
    if (!c) return AUG_ERR_NOT_INIT;
    if (n < 0 || !out || (n > 0 && (!y || !mu || !var || ld < n))) return AUG_ERR_BAD_ARG;
    AUG_CUDA(cudaSetDevice(c->device));
    if (n == 0) {
        AUG_CUDA(cudaMemsetAsync(out, 0, sizeof(double), c->stream));
        if (aug_xch_for(c)) {
            xch_one_kernel<<<1, 32, 0, c->stream>>>(c->xch, out);
            c->launches++;
            return (int32_t)cudaGetLastError();
        }
        return AUG_OK;
    }
    LamArgs a{};
    a.n = n;
    a.y = y;
    a.a0 = mu;
    a.a1 = var;
    a.a2 = mu + ld;
    a.a3 = var + ld;
    a.partials = c->partials;
    a.counter = c->counter;
    a.out = out;
    a.xch = aug_xch_for(c);
    if (a.xch) { int32_t rf = aug_xch_flush(c); if (rf) return rf; }
    return launch_lambda<false>(c, a);
}

int32_t aug_hetero_lambda_stats_sampled(aug_ctx* c, int64_t n, const double* y, const double* f, int64_t ld,
                                        double* out) {
    if (!c) return AUG_ERR_NOT_INIT;
    if (n < 0 || !out || (n > 0 && (!y || !f || ld < n))) return AUG_ERR_BAD_ARG;
    AUG_CUDA(cudaSetDevice(c->device));
    if (n == 0) {
        AUG_CUDA(cudaMemsetAsync(out, 0, sizeof(double), c->stream));
        if (aug_xch_for(c)) {
            xch_one_kernel<<<1, 32, 0, c->stream>>>(c->xch, out);
            c->launches++;
            return (int32_t)cudaGetLastError();
        }
        return AUG_OK;
    }
    LamArgs a{};
    a.n = n;
    a.y = y;
    a.a0 = f;
    a.a1 = f + ld;
    a.partials = c->partials;
    a.counter = c->counter;
    a.out = out;
    a.xch = aug_xch_for(c);
    if (a.xch) { int32_t rf = aug_xch_flush(c); if (rf) return rf; }
    return launch_lambda<true>(c, a);
}

int32_t aug_logisticsoftmax(aug_ctx* c, const aug_lik* lik, int64_t n, const double* f, double* out) {
    if (!c) return AUG_ERR_NOT_INIT;
    if (!lik || n < 0 || (n > 0 && (!f || !out))) return AUG_ERR_BAD_ARG;
    if (lik->kind != AUG_CAT && lik->kind != AUG_CAT_BIJ) return AUG_ERR_BAD_KIND;
    if (lik->nlatent < 1) return AUG_ERR_BAD_ARG;
    AUG_CUDA(cudaSetDevice(c->device));
    if (n == 0) return AUG_OK;
    LinkArgs a{};
    a.n = n;
    a.nl = lik->nlatent;
    a.K = lik->kind == AUG_CAT_BIJ ? a.nl + 1 : a.nl;
    a.f = f;
    a.out = out;
    int32_t rc = upload_theta(c, lik, a.K, &a.theta);
    if (rc) return rc;
    const int grid = aug_grid_for(c, (const void*)link_kernel<false>, n * 32, AUG_BLOCK);
    link_kernel<false><<<grid, AUG_BLOCK, 0, c->stream>>>(a);
    c->launches++;
    return (int32_t)cudaGetLastError();
}

int32_t aug_approx_expected_logisticsoftmax(aug_ctx* c, const aug_lik* lik, int64_t n, const double* mu,
                                            const double* cc, double* out) {
    if (!c) return AUG_ERR_NOT_INIT;
    if (!lik || n < 0 || (n > 0 && (!mu || !cc || !out))) return AUG_ERR_BAD_ARG;
    if (lik->kind != AUG_CAT_BIJ) return AUG_ERR_BAD_KIND;   // utils.jl:17-22 takes θ with one more entry than μ
    AUG_CUDA(cudaSetDevice(c->device));
    if (n == 0) return AUG_OK;
    LinkArgs a{};
    a.n = n;
    a.nl = lik->nlatent;
    a.K = a.nl + 1;
    a.f = mu;
    a.c = cc;
    a.out = out;
    int32_t rc = upload_theta(c, lik, a.K, &a.theta);
    if (rc) return rc;
    const int grid = aug_grid_for(c, (const void*)link_kernel<true>, n * 32, AUG_BLOCK);
    link_kernel<true><<<grid, AUG_BLOCK, 0, c->stream>>>(a);
    c->launches++;
    return (int32_t)cudaGetLastError();
}

}  // extern "C"
