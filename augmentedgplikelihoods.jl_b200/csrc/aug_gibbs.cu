// aug_gibbs.cu — Gibbs side: aux_sample!, init_aux_variables, the raw PG sampler, and the sampled
// (non-expected) potential / precision map for the scalar-latent likelihoods and HETERO.
//
// Reference behaviour replaced (paths relative to /root/reference/src):
//   generic.jl:1-20,32-34 (aux_sample!, aux_sample, init_aux_variables),
//   likelihoods/*.jl aux_full_conditional + auglik_potential/precision,
//   SpecialDistributions/polyagamma.jl:112-257, polyagammapoisson.jl:23-27.
// One thread owns one observation and one Philox stream positioned by the GLOBAL observation
// index, so the draws are identical for any sharding of the observation axis.
#include "aug_common.cuh"
#include "aug_math.cuh"
#include "aug_pg.cuh"

namespace {

struct GibbsArgs {
    int64_t n, i0;
    uint64_t seed, offset;
    const void* y;
    const double* f;
    const double* g;   // HETERO second latent
    double* omega;
    int64_t* nvar;
    LikConst L;
};

template <int KIND>
__global__ void __launch_bounds__(AUG_BLOCK) aux_sample_kernel(const GibbsArgs a) {
    const int64_t nth = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += nth) {
        augr::Philox g;
        g.init(a.seed, a.offset, (uint64_t)(a.i0 + i));
        const double f = ld_stream1(a.f + i);
        if (KIND == AUG_BERNOULLI) {                              // PG(1, |f|)  bernoulli.jl:13-15
            st_stream1(a.omega + i, augp::pg_draw(g, 1.0, true, f, a.L.pgtab));
        } else if (KIND == AUG_NEGBIN) {                          // PG(y + r, |f|)  negativebinomial.jl:20-22
            const double y = (double)__ldg(reinterpret_cast<const int64_t*>(a.y) + i);
            st_stream1(a.omega + i, augp::pg_draw(g, y + a.L.p0, a.L.r_is_int != 0, f, a.L.pgtab));
        } else if (KIND == AUG_POISSON) {                         // poisson.jl:26-28, polyagammapoisson.jl:23-27
            const int64_t y = __ldg(reinterpret_cast<const int64_t*>(a.y) + i);
            const int64_t nn = augr::poisson_rand(g, a.L.p0 * augm::logistic(-f));
            a.nvar[i] = nn;
            st_stream1(a.omega + i, augp::pg_draw(g, (double)(nn + y), true, f, a.L.pgtab));
        } else if (KIND == AUG_LAPLACE) {                         // IG(1/(2β|y-f|), 2λ)  laplace.jl:40-42
            const double y = ld_stream1(reinterpret_cast<const double*>(a.y) + i);
            const double mu = a.L.c0 / fabs(y - f);
            st_stream1(a.omega + i, augr::invgauss_rand(g, mu, 2.0 * a.L.c1));
        } else if (KIND == AUG_STUDENTT) {                        // Gamma(α, 2/(ν/σ² + (y-f)²))  studentt.jl:46-48
            const double y = ld_stream1(reinterpret_cast<const double*>(a.y) + i);
            const double d = y - f;
            st_stream1(a.omega + i, augr::gamma_rand(g, a.L.c1) * 2.0 / fma(d, d, a.L.c0));
        } else if (KIND == AUG_HETERO) {                          // heteroscedasticgaussian.jl:28-32
            const double y = ld_stream1(reinterpret_cast<const double*>(a.y) + i);
            const double gg = ld_stream1(a.g + i);
            const double d = f - y;
            const double rate = a.L.p0 * augm::logistic(-gg) * d * d * 0.5;
            const int64_t nn = augr::poisson_rand(g, rate);
            a.nvar[i] = nn;
            st_stream1(a.omega + i, augp::pg_draw(g, (double)nn + 0.5, false, gg, a.L.pgtab));
        }
    }
}

// init_aux_variables: PG(1,0) [+ Poisson(1)] / InverseGamma(1,1) / Gamma(1,1)
// bernoulli.jl:3-5, negativebinomial.jl:10-12, poisson.jl:14-18, laplace.jl:29-31, studentt.jl:35-37,
// heteroscedasticgaussian.jl:16-20, categorical.jl:52-57 (m = nl * n elements)
__global__ void __launch_bounds__(AUG_BLOCK) init_aux_kernel(int kind, int64_t m, int64_t e0, uint64_t seed,
                                                              uint64_t offset, double* omega, int64_t* nvar, const double* pgtab) {
    const int64_t nth = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += nth) {
        augr::Philox g;
        g.init(seed, offset, (uint64_t)(e0 + i));
        if (kind == AUG_LAPLACE) {
            omega[i] = 1.0 / augr::gamma_rand(g, 1.0);
        } else if (kind == AUG_STUDENTT) {
            omega[i] = augr::gamma_rand(g, 1.0);
        } else {
            omega[i] = augp::pg_draw(g, 1.0, true, 0.0, pgtab);
            if (nvar) nvar[i] = augr::poisson_rand(g, 1.0);
        }
    }
}

__global__ void __launch_bounds__(AUG_BLOCK) pg_rand_kernel(int64_t n, int64_t i0, uint64_t seed, uint64_t offset,
                                                             const double* b, const double* c, double bs, double cs,
                                                             int b_is_int, double* out, const double* pgtab) {
    const int64_t nth = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += nth) {
        augr::Philox g;
        g.init(seed, offset, (uint64_t)(i0 + i));
        const double bi = b ? b[i] : bs;
        const double ci = c ? c[i] : cs;
        out[i] = augp::pg_draw(g, bi, b_is_int != 0, ci, pgtab);
    }
}

// auglik_potential_and_precision (sampled): bernoulli.jl:27-33, negativebinomial.jl:35-41,
// poisson.jl:41-47, laplace.jl:54-60, studentt.jl:60-66, heteroscedasticgaussian.jl:48-66
struct PotArgs {
    int64_t n;
    const void* y;
    const double* g;  // HETERO
    const double* omega;
    const int64_t* nvar;
    double* beta;
    double* gamma;
    double* beta_g;
    double* gamma_g;
    LikConst L;
};

template <int KIND>
__global__ void __launch_bounds__(AUG_BLOCK) potential_kernel(const PotArgs a) {
    const int64_t nth = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += nth) {
        const double w = a.omega[i];
        double b0, g0 = w;
        if (KIND == AUG_BERNOULLI) {
            b0 = reinterpret_cast<const uint8_t*>(a.y)[i] ? 0.5 : -0.5;
        } else if (KIND == AUG_NEGBIN) {
            b0 = 0.5 * ((double)reinterpret_cast<const int64_t*>(a.y)[i] - a.L.p0);
        } else if (KIND == AUG_POISSON) {
            b0 = 0.5 * (double)(reinterpret_cast<const int64_t*>(a.y)[i] - a.nvar[i]);
        } else if (KIND == AUG_LAPLACE) {
            b0 = 2.0 * w * reinterpret_cast<const double*>(a.y)[i];
            g0 = 2.0 * w;
        } else if (KIND == AUG_STUDENTT) {
            b0 = reinterpret_cast<const double*>(a.y)[i] * w;
        } else {  // HETERO
            const double il = a.L.p0 * augm::logistic(a.g[i]);     // inv(invlink(g)) = λ σ(g)
            b0 = reinterpret_cast<const double*>(a.y)[i] * il;
            g0 = il;
            if (a.beta_g) a.beta_g[i] = 0.5 * (0.5 - (double)a.nvar[i]);
            if (a.gamma_g) a.gamma_g[i] = w;
        }
        if (a.beta) a.beta[i] = b0;
        if (a.gamma) a.gamma[i] = g0;
    }
}

template <typename K, typename A>
int32_t launch_map(aug_ctx* ctx, K kernel, const A& a, int64_t n) {
    const int grid = aug_grid_for(ctx, (const void*)kernel, n, AUG_BLOCK);
    kernel<<<grid, AUG_BLOCK, 0, ctx->stream>>>(a);
    ctx->launches++;
    return (int32_t)cudaGetLastError();
}

bool is_cat(int k) { return k == AUG_CAT || k == AUG_CAT_BIJ; }

}  // namespace

int32_t aug_cat_sample(aug_ctx* ctx, const aug_lik* lik, int64_t n, int64_t i0, const void* y, const double* f,
                       double* omega, int64_t* nvar, uint64_t offset);
int32_t aug_cat_potential(aug_ctx* ctx, const aug_lik* lik, int64_t n, const void* y, const double* omega,
                          const int64_t* nvar, double* beta, double* gamma, int64_t ldo);

// device-pointer implementation with an explicit RNG tick (shared with the host-buffer pipeline,
// which must use ONE tick for all chunks of a call)
int32_t aug_aux_sample_dev(aug_ctx* c, const aug_lik* lik, int64_t n, int64_t i0, const void* y, const double* f,
                           int64_t ld, double* omega, int64_t* nvar, uint64_t off) {
    if (n == 0) return AUG_OK;
    if (is_cat(lik->kind)) return aug_cat_sample(c, lik, n, i0, y, f, omega, nvar, off);
    GibbsArgs a{};
    a.n = n;
    a.i0 = i0;
    a.seed = c->seed;
    a.offset = off;
    a.y = y;
    a.f = f;
    a.omega = omega;
    a.nvar = nvar;
    int32_t rc = aug_lik_const(c, lik, &a.L, false, false);
    if (rc) return rc;
    const bool needs_y = lik->kind != AUG_BERNOULLI;
    const bool needs_n = lik->kind == AUG_POISSON || lik->kind == AUG_HETERO;
    if ((needs_y && !y) || (needs_n && !nvar)) return AUG_ERR_BAD_ARG;
    if (lik->kind == AUG_HETERO) {
        if (ld < n) return AUG_ERR_BAD_ARG;
        a.g = f + ld;
    }
    switch (lik->kind) {
        case AUG_BERNOULLI: return launch_map(c, aux_sample_kernel<AUG_BERNOULLI>, a, n);
        case AUG_NEGBIN: return launch_map(c, aux_sample_kernel<AUG_NEGBIN>, a, n);
        case AUG_POISSON: return launch_map(c, aux_sample_kernel<AUG_POISSON>, a, n);
        case AUG_LAPLACE: return launch_map(c, aux_sample_kernel<AUG_LAPLACE>, a, n);
        case AUG_STUDENTT: return launch_map(c, aux_sample_kernel<AUG_STUDENTT>, a, n);
        case AUG_HETERO: return launch_map(c, aux_sample_kernel<AUG_HETERO>, a, n);
        default: return AUG_ERR_BAD_KIND;
    }
}

extern "C" {

int32_t aug_aux_sample(aug_ctx* c, const aug_lik* lik, int64_t n, int64_t i0, const void* y, const double* f,
                       int64_t ld, double* omega, int64_t* nvar) {
    if (!c) return AUG_ERR_NOT_INIT;
    if (!lik || n < 0 || !f || !omega) return AUG_ERR_BAD_ARG;
    if (lik->kind < 0 || lik->kind >= AUG_NKINDS) return AUG_ERR_BAD_KIND;
    AUG_CUDA(cudaSetDevice(c->device));
    const uint64_t off = c->offset++;
    return aug_aux_sample_dev(c, lik, n, i0, y, f, ld, omega, nvar, off);
}

int32_t aug_init_aux_variables(aug_ctx* c, const aug_lik* lik, int64_t n, int64_t i0, double* omega,
                               int64_t* nvar) {
    if (!c) return AUG_ERR_NOT_INIT;
    if (!lik || n < 0 || !omega) return AUG_ERR_BAD_ARG;
    if (lik->kind < 0 || lik->kind >= AUG_NKINDS) return AUG_ERR_BAD_KIND;
    AUG_CUDA(cudaSetDevice(c->device));
    const uint64_t off = c->offset++;
    if (n == 0) return AUG_OK;
    const bool needs_n = lik->kind == AUG_POISSON || lik->kind == AUG_HETERO || is_cat(lik->kind);
    if (needs_n && !nvar) return AUG_ERR_BAD_ARG;
    const int64_t per = is_cat(lik->kind) ? lik->nlatent : 1;
    const int64_t m = n * per;
    const int grid = aug_grid_for(c, (const void*)init_aux_kernel, m, AUG_BLOCK);
    init_aux_kernel<<<grid, AUG_BLOCK, 0, c->stream>>>(lik->kind, m, i0 * per, c->seed, off, omega,
                                                        needs_n ? nvar : nullptr, c->pgtab);
    c->launches++;
    return (int32_t)cudaGetLastError();
}

static int32_t pg_rand_common(aug_ctx* c, int64_t n, int64_t i0, const double* b, const double* cc, double bs,
                              double cs, int32_t b_is_int, double* out) {
    if (!c) return AUG_ERR_NOT_INIT;
    if (n < 0 || !out) return AUG_ERR_BAD_ARG;
    AUG_CUDA(cudaSetDevice(c->device));
    const uint64_t off = c->offset++;
    if (n == 0) return AUG_OK;
    const int grid = aug_grid_for(c, (const void*)pg_rand_kernel, n, AUG_BLOCK);
    pg_rand_kernel<<<grid, AUG_BLOCK, 0, c->stream>>>(n, i0, c->seed, off, b, cc, bs, cs, b_is_int, out, c->pgtab);
    c->launches++;
    return (int32_t)cudaGetLastError();
}

int32_t aug_pg_rand(aug_ctx* c, int64_t n, int64_t i0, const double* b, const double* cc, int32_t b_is_int,
                    double* out) {
    if (!b || !cc) return AUG_ERR_BAD_ARG;
    return pg_rand_common(c, n, i0, b, cc, 0.0, 0.0, b_is_int, out);
}

int32_t aug_pg_rand_bc(aug_ctx* c, int64_t n, int64_t i0, double b, double cc, int32_t b_is_int, double* out) {
    return pg_rand_common(c, n, i0, nullptr, nullptr, b, cc, b_is_int, out);
}

int32_t aug_potential_precision(aug_ctx* c, const aug_lik* lik, int64_t n, const void* y, const double* f,
                                int64_t ld, const double* omega, const int64_t* nvar, double* beta,
                                double* gamma, int64_t ldo) {
    if (!c) return AUG_ERR_NOT_INIT;
    if (!lik || n < 0 || !y || !omega) return AUG_ERR_BAD_ARG;
    AUG_CUDA(cudaSetDevice(c->device));
    if (n == 0) return AUG_OK;
    if (is_cat(lik->kind)) return aug_cat_potential(c, lik, n, y, omega, nvar, beta, gamma, ldo);
    PotArgs a{};
    a.n = n;
    a.y = y;
    a.omega = omega;
    a.nvar = nvar;
    a.beta = beta;
    a.gamma = gamma;
    int32_t rc = aug_lik_const(c, lik, &a.L, false, false);
    if (rc) return rc;
    if ((lik->kind == AUG_POISSON || lik->kind == AUG_HETERO) && !nvar) return AUG_ERR_BAD_ARG;
    if (lik->kind == AUG_HETERO) {
        if (!f || ld < n || ldo < n) return AUG_ERR_BAD_ARG;
        a.g = f + ld;
        a.beta_g = beta ? beta + ldo : nullptr;
        a.gamma_g = gamma ? gamma + ldo : nullptr;
    }
    switch (lik->kind) {
        case AUG_BERNOULLI: return launch_map(c, potential_kernel<AUG_BERNOULLI>, a, n);
        case AUG_NEGBIN: return launch_map(c, potential_kernel<AUG_NEGBIN>, a, n);
        case AUG_POISSON: return launch_map(c, potential_kernel<AUG_POISSON>, a, n);
        case AUG_LAPLACE: return launch_map(c, potential_kernel<AUG_LAPLACE>, a, n);
        case AUG_STUDENTT: return launch_map(c, potential_kernel<AUG_STUDENTT>, a, n);
        case AUG_HETERO: return launch_map(c, potential_kernel<AUG_HETERO>, a, n);
        default: return AUG_ERR_BAD_KIND;
    }
}

}  // extern "C"
