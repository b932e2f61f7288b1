/*
 * augcuda.h — C ABI of libaugcuda.so
 *
 * B200-native (sm_100a) implementation of the per-observation augmentation hot
 * path of AugmentedGPLikelihoods.jl (CAVI update, Gibbs auxiliary draw, ELBO
 * terms).  The reference has no FFI: its extension mechanism is Julia multiple
 * dispatch on the verbs exported at src/AugmentedGPLikelihoods.jl:18-30.  Each
 * entry point below is what a `ccall` from a Julia method of that verb,
 * specialised on a device-array type, binds to (see INTEGRATION.md and
 * augmentedgplikelihoods.jl_b200/julia/AugCUDA.jl).  All reference citations are relative to
 * /root/reference.
 *
 * Conventions
 *  - every function returns int32: 0 = OK, <0 = aug_status, >0 = cudaError_t,
 *    >=1000 = ncclResult_t + 1000.  No exception crosses the boundary.
 *  - all data pointers are DEVICE pointers owned by the caller unless the
 *    name ends in `_host`; the library owns only ctx-internal scratch.
 *  - all verbs are asynchronous on the ctx stream; scalars are written to
 *    device memory (read them after aug_ctx_sync or with aug_memcpy_d2h).
 *  - n is the number of observations held by THIS rank (shard), i0 the global
 *    index of its first observation (RNG counters use the global index, so
 *    draws are invariant under re-sharding).
 *  - dtype of y per likelihood kind: BERNOULLI uint8 (Julia Bool),
 *    NEGBIN / POISSON int64 (Julia Int), LAPLACE / STUDENTT / HETERO double,
 *    CAT / CAT_BIJ uint8 one-hot [n][nl] (class index fastest, the flat view of
 *    the reference's ArrayOfSimilarArrays, src/likelihoods/categorical.jl:63).
 *  - multi-latent layouts: HETERO latent-major, f at ptr, g at ptr + ld
 *    (src/likelihoods/heteroscedasticgaussian.jl:38: qfg = (qf, qg));
 *    CAT observation-major [n][nl] (src/likelihoods/categorical.jl:84).
 *    Outputs beta/gamma are always latent-major [nlatent][ldo]
 *    (class-major for CAT: src/likelihoods/categorical.jl:112-136, utils.jl:24).
 */
#ifndef AUGCUDA_H
#define AUGCUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AUGCUDA_VERSION 100 /* 0.1.0 */

typedef struct aug_ctx aug_ctx;

/* Likelihood kinds = the dispatch tags of src/likelihoods/ (one file per kind) */
enum aug_kind {
    AUG_BERNOULLI = 0, /* BernoulliLikelihood{<:LogisticLink}           bernoulli.jl */
    AUG_NEGBIN    = 1, /* NegativeBinomialLikelihood{<:NBParamFailure}  negativebinomial.jl */
    AUG_POISSON   = 2, /* PoissonLikelihood{<:ScaledLogistic}           poisson.jl */
    AUG_LAPLACE   = 3, /* LaplaceLikelihood                             laplace.jl */
    AUG_STUDENTT  = 4, /* StudentTLikelihood                            studentt.jl */
    AUG_HETERO    = 5, /* HeteroscedasticGaussianLikelihood{<:InvScaledLogistic} */
    AUG_CAT_BIJ   = 6, /* CategoricalLikelihood{<:BijectiveSimplexLink{<:LogisticSoftMaxLink}} */
    AUG_CAT       = 7, /* CategoricalLikelihood{<:LogisticSoftMaxLink} */
    AUG_NKINDS    = 8
};

enum aug_status {
    AUG_OK               = 0,
    AUG_ERR_BAD_KIND     = -1,
    AUG_ERR_BAD_ARG      = -2,
    AUG_ERR_PRECONDITION = -3, /* mirrors the reference's error(...) calls:
                                  non-bijective KL (categorical.jl:165-170),
                                  sum(p) >= 1 (negativemultinomial.jl:18-22) */
    AUG_ERR_NOT_INIT     = -4,
    AUG_ERR_NO_NCCL      = -5,
    AUG_ERR_DEVICE_FLAG  = -6, /* a kernel raised the device-side error flag */
    AUG_ERR_NO_CUBLAS    = -7  /* m > 128 needs libcublas.so.12 (dlopen), which could not be loaded */
};

/*
 * Likelihood descriptor (POD, passed by pointer).  Parameters are the struct
 * fields of the reference likelihood / link types:
 *   NEGBIN   p[0] = failures r   (negativebinomial.jl:16); r_is_int selects the
 *            Int methods (negbin_logconst :52, integer-b PG sampler polyagamma.jl:129)
 *   POISSON  p[0] = lambda       (ScaledLogistic.λ, poisson.jl:1-3)
 *   LAPLACE  p[0] = beta         (laplace.jl:13-15)
 *   STUDENTT p[0] = nu, p[1] = sigma (studentt.jl:14-21)
 *   HETERO   p[0] = lambda       (InvScaledLogistic.λ, heteroscedasticgaussian.jl:1-3)
 *   CAT*     logtheta = HOST pointer to K doubles (LogisticSoftMaxLink.logθ,
 *            categorical.jl:6-10); NULL means zeros.  K = nlatent (CAT) or
 *            nlatent + 1 (CAT_BIJ).
 */
/* aug_lik.flags */
#define AUG_LIK_FAITHFUL_QUIRKS 1 /* reproduce what the reference's CODE returns where it differs from the intended
                                     mathematics (default 0 = intended):
                                     - logdensity_def(::PolyaGammaNegativeMultinomial) sums the PG log-densities of
                                       the first TWO classes only (`sum(1:length(x))` with x a 2-field NamedTuple,
                                       SpecialDistributions/polyagammanegativemultinomial.jl:33-39; a BoundsError ->
                                       AUG_ERR_PRECONDITION when nlatent < 2);
                                     - kldivergence(::NegativeMultinomial, ...) is NaN when a variational p_j is
                                       exactly 0 (`p * (log(p) - log(q))` = 0 * -Inf,
                                       SpecialDistributions/negativemultinomial.jl:80) instead of the limit 0. */
typedef struct aug_lik {
    int32_t kind;
    int32_t nlatent;
    int32_t r_is_int;
    int32_t flags;
    double  p[4];
    const double* logtheta;
} aug_lik;

/* Slots of the device scalar block written by the reducing verbs.  A verb owns the WHOLE block of its call: it
 * writes its own slots and zeroes the others, so the caller never has to clear it. */
enum aug_scalar_slot {
    AUG_S_EXPECTED_LOGTILT = 0, /* api.jl:219-223 and per-likelihood methods */
    AUG_S_KL               = 1, /* generic.jl:56-62, laplace.jl:98-104 */
    AUG_S_EXPECTED_AUGLL   = 2, /* generic.jl:52-54 (the sum, "+"); hetero :129-145 */
    AUG_S_LOGTILT          = 3, /* generic.jl:40-46 */
    AUG_S_LOGPRIOR         = 4, /* logdensity_def(aux_prior(lik,y), Ω), generic.jl:49 */
    AUG_S_AUGLL            = 5, /* generic.jl:48-50; hetero :117-127 */
    AUG_S_FLAGS            = 6, /* count of rows that tripped a precondition */
    AUG_S_RESERVED         = 7,
    AUG_NSCALARS           = 8
};

/* ---- library / context ------------------------------------------------- */
int32_t     aug_version(void);
const char* aug_strerror(int32_t rc);

/* stream: a cudaStream_t to run on, or NULL to let the ctx create its own. */
int32_t aug_ctx_create(aug_ctx** out, int32_t device, void* stream);
int32_t aug_ctx_destroy(aug_ctx* ctx);
/* Counter-based RNG state: replaces the `rng::AbstractRNG` argument of
 * aux_sample!/init_aux_variables (generic.jl:1-20,32-34).  Every sampling verb
 * consumes one `offset` tick, so a (seed, offset) pair makes a run resumable. */
int32_t aug_ctx_seed(aug_ctx* ctx, uint64_t seed, uint64_t offset);
int32_t aug_ctx_get_offset(aug_ctx* ctx, uint64_t* offset);
int32_t aug_ctx_sync(aug_ctx* ctx);
int32_t aug_ctx_stream(aug_ctx* ctx, void** stream);
int32_t aug_ctx_sm_count(aug_ctx* ctx, int32_t* n);
/* number of kernels launched by this ctx since creation (bench "gpu_launches") */
int32_t aug_ctx_launch_count(aug_ctx* ctx, uint64_t* n);
/* Synchronises, returns and clears the device-side error flag word: bit 0 = a
 * NegativeMultinomial row had sum(p) >= 1 (the reference throws ArgumentError at
 * negativemultinomial.jl:17-22; an asynchronous kernel can only flag it). */
int32_t aug_ctx_error_flag(aug_ctx* ctx, uint32_t* flag);

/* memory helpers for callers without their own device allocator */
int32_t aug_malloc(aug_ctx* ctx, void** dev, size_t bytes);
int32_t aug_free(aug_ctx* ctx, void* dev);
int32_t aug_host_alloc(void** host, size_t bytes); /* pinned */
int32_t aug_host_free(void* host);
int32_t aug_memcpy_h2d(aug_ctx* ctx, void* dev, const void* host, size_t bytes);
int32_t aug_memcpy_d2h(aug_ctx* ctx, void* host, const void* dev, size_t bytes);

/* ---- variational (CAVI) side ------------------------------------------- */
/*
 * State arrays of qΩ (the struct-of-arrays `only(qΩ.inds)`):
 *   kind       s0        s1            s2 (optional copy, may be NULL)
 *   BERNOULLI  c         -             -              bernoulli.jl:7-11
 *   NEGBIN     c         -             y (int64)      negativebinomial.jl:14-18
 *   POISSON    c         λ             y (int64)      poisson.jl:20-24
 *   LAPLACE    μ         -             -              laplace.jl:33-38
 *   STUDENTT   β         -             -              studentt.jl:39-44
 *   HETERO     c         λ             ψ (required)   heteroscedasticgaussian.jl:22-26
 *   CAT*       c [n][nl] p [n][nl]     y (uint8)      categorical.jl:59-70
 * When s2 is NULL for NEGBIN/POISSON/CAT the y copy of the reference
 * (`φ.y .= y`) is not materialised and the verbs read the y argument.
 */

/* init_aux_posterior(T, lik, n): zero-filled state.  (a3) */
int32_t aug_init_aux_posterior(aug_ctx* ctx, const aug_lik* lik, int64_t n,
                               void* s0, void* s1, void* s2);

/* aux_posterior!(qΩ, lik, y, qf): bernoulli.jl:17-25, negativebinomial.jl:24-33,
 * poisson.jl:30-39, laplace.jl:44-52, studentt.jl:50-58,
 * heteroscedasticgaussian.jl:34-46, categorical.jl:80-110.  (a5)
 * mu/var: means and variances of the marginals of q(f). */
int32_t aug_aux_posterior(aug_ctx* ctx, const aug_lik* lik, int64_t n,
                          const void* y, const double* mu, const double* var, int64_t ld,
                          void* s0, void* s1, void* s2);

/* expected_auglik_potential_and_precision(lik, qΩ, y[, qf]) from an existing
 * state: bernoulli.jl:35-45, negativebinomial.jl:43-49, poisson.jl:49-60,
 * laplace.jl:62-68, studentt.jl:68-74, heteroscedasticgaussian.jl:68-104,
 * categorical.jl:121-136.  (a7)  mu is only read for HETERO (means of q(g)).
 * Either of beta/gamma may be NULL (expected_auglik_potential / _precision). */
int32_t aug_expected_potential_precision(aug_ctx* ctx, const aug_lik* lik, int64_t n,
                                         const void* y, const double* mu, int64_t ld,
                                         const void* s0, const void* s1, const void* s2,
                                         double* beta, double* gamma, int64_t ldo);

/* Fused single-pass CAVI step = aux_posterior! + expected_auglik_potential_and_precision
 * (+ expected_logtilt, aux_kldivergence, expected_aug_loglik partial sums when
 * scalars != NULL, written to scalars[AUG_S_EXPECTED_LOGTILT / _KL / _EXPECTED_AUGLL]).
 * One read of y, mu, var; one write of state, beta, gamma.  This is the call
 * pattern of examples/bernoulli/script.jl:29-39.  Any of s0..s2, beta, gamma
 * may be NULL to skip that output. */
int32_t aug_cavi_step(aug_ctx* ctx, const aug_lik* lik, int64_t n,
                      const void* y, const double* mu, const double* var, int64_t ld,
                      void* s0, void* s1, void* s2,
                      double* beta, double* gamma, int64_t ldo,
                      double* scalars);

/* expected_logtilt / aux_kldivergence / expected_aug_loglik from an existing
 * state (a9, a11, a13): api.jl:219-223, generic.jl:52-62 and the per-likelihood
 * methods.  Writes scalars[0..2].  CAT (non-bijective) returns
 * AUG_ERR_PRECONDITION like categorical.jl:165-170. */
int32_t aug_expected_elbo_terms(aug_ctx* ctx, const aug_lik* lik, int64_t n,
                                const void* y, const double* mu, const double* var, int64_t ld,
                                const void* s0, const void* s1, const void* s2,
                                double* scalars);

/* ---- sampling (Gibbs) side --------------------------------------------- */
/* Ω is the TupleVector of samples: omega (double) and, for POISSON / HETERO /
 * CAT*, nvar (int64) — fields ω and n (bernoulli.jl:3-5, poisson.jl:14-18).
 * CAT*: both are [n][nl]. */

/* init_aux_variables(rng, lik, n): bernoulli.jl:3-5 etc.  (a4) */
int32_t aug_init_aux_variables(aug_ctx* ctx, const aug_lik* lik, int64_t n, int64_t i0,
                               double* omega, int64_t* nvar);

/* aux_sample!(rng, Ω, lik, y, f): generic.jl:5-12 with the full conditionals of
 * bernoulli.jl:13-15, negativebinomial.jl:20-22, poisson.jl:26-28,
 * laplace.jl:40-42, studentt.jl:46-48, heteroscedasticgaussian.jl:28-32,
 * categorical.jl:72-78.  (a14-a19) */
int32_t aug_aux_sample(aug_ctx* ctx, const aug_lik* lik, int64_t n, int64_t i0,
                       const void* y, const double* f, int64_t ld,
                       double* omega, int64_t* nvar);

/* auglik_potential_and_precision(lik, Ω, y[, f]) (a20).  f only read for HETERO. */
int32_t aug_potential_precision(aug_ctx* ctx, const aug_lik* lik, int64_t n,
                                const void* y, const double* f, int64_t ld,
                                const double* omega, const int64_t* nvar,
                                double* beta, double* gamma, int64_t ldo);

/* logtilt (a21) and, when with_prior != 0, also logdensity(aux_prior, Ω) and
 * aug_loglik (a22-a24: the PG(b,0) log-density series, polyagamma.jl:37-91).
 * Writes scalars[AUG_S_LOGTILT / _LOGPRIOR / _AUGLL]. */
int32_t aug_sampled_loglik_terms(aug_ctx* ctx, const aug_lik* lik, int64_t n,
                                 const void* y, const double* f, int64_t ld,
                                 const double* omega, const int64_t* nvar,
                                 int32_t with_prior, double* scalars);

/* ---- SpecialDistributions primitives (element-wise, for tests and callers) */
/* rand(PolyaGamma(b_i, c_i)) (polyagamma.jl:121-164).  b_is_int selects the
 * Int method (draw_sum :129-134) — b is then rounded to the nearest integer. */
int32_t aug_pg_rand(aug_ctx* ctx, int64_t n, int64_t i0,
                    const double* b, const double* c, int32_t b_is_int, double* out);
/* same with scalar (b, c) broadcast over n draws */
int32_t aug_pg_rand_bc(aug_ctx* ctx, int64_t n, int64_t i0,
                       double b, double c, int32_t b_is_int, double* out);
/* mean(PolyaGamma(b,c)) polyagamma.jl:25-31 */
int32_t aug_pg_mean(aug_ctx* ctx, int64_t n, const double* b, const double* c, double* out);
/* kldivergence(PolyaGamma(b,c), PolyaGamma(b,0)) polyagamma.jl:99-110 */
int32_t aug_pg_kl(aug_ctx* ctx, int64_t n, const double* b, const double* c, double* out);
/* logpdf(PolyaGamma(b,c), x) polyagamma.jl:37-91 (scalar b, c) */
int32_t aug_pg_logpdf(aug_ctx* ctx, int64_t n, double b, double c, const double* x, double* out);
/* approx_expected_logistic(mu, c) utils.jl:11-14 */
int32_t aug_approx_expected_logistic(aug_ctx* ctx, int64_t n, const double* mu, const double* c,
                                     double* out);

/* second_moment(q) = mu^2 + var, or (mu - y)^2 + var when y != NULL — utils.jl:1-7 */
int32_t aug_second_moment(aug_ctx* ctx, int64_t n, const double* mu, const double* var, const double* y,
                          double* out);

/* diagnostics: evaluates one of the straight-line fp64 functions the kernels use (csrc/aug_fastmath.cuh)
 * element-wise, for accuracy tests.  fn: 0 rcp, 1 rsqrt, 2 exp (|x| <= 708), 3 log (normal x > 0),
 * 4 log on [1,2], 5 sqrt (1e-290 <= x <= 1e290). */
int32_t aug_fastmath_eval(aug_ctx* ctx, int32_t fn, int64_t n, const double* x, double* out);

/* ---- callers on either side of the path -------- */
/* opt_lik of examples/heteroscedasticgaussian/script.jl:41-51: out[0] = dot(ψ, 1 .- σ̃g) with
 * ψ = second_moment.(qf .- y)/2, c = sqrt.(second_moment.(qg)), σ̃g = approx_expected_logistic.(-mean.(qg), c);
 * the caller forms λ = max(N / (2 out[0]), λ_old).  mu/var latent-major [2][ld] like aug_cavi_step (HETERO).
 * One 40 B/obs map-reduce pass; in fused multi-GPU mode the sum is over all ranks. */
int32_t aug_hetero_lambda_stats(aug_ctx* ctx, int64_t n, const double* y, const double* mu, const double* var,
                                int64_t ld, double* out /* device, 1 double */);
/* Gibbs counterpart, docs/src/likelihoods/heteroscedasticgaussian.md:80-84: out[0] = Σᵢ σ(gᵢ)/2 (yᵢ − fᵢ)², the
 * rate increment of the Gamma full conditional of λ.  f latent-major [2][ld] (f then g). */
int32_t aug_hetero_lambda_stats_sampled(aug_ctx* ctx, int64_t n, const double* y, const double* f, int64_t ld,
                                        double* out /* device, 1 double */);
/* (l::LogisticSoftMaxLink)(f) / logisticsoftmax(x): likelihoods/categorical.jl:1-4, 32-35; for AUG_CAT_BIJ the
 * BijectiveSimplexLink appends a zero latent (categorical.jl:12-14).  f [n][nl] class fastest;
 * out [n][K], K = nl (AUG_CAT) or nl + 1 (AUG_CAT_BIJ). */
int32_t aug_logisticsoftmax(aug_ctx* ctx, const aug_lik* lik, int64_t n, const double* f, double* out);
/* approx_expected_logisticsoftmax(μ, c, θ) utils.jl:17-22 (AUG_CAT_BIJ only: θ has nl + 1 entries).
 * mu, c, out: [n][nl]. */
int32_t aug_approx_expected_logisticsoftmax(aug_ctx* ctx, const aug_lik* lik, int64_t n, const double* mu,
                                            const double* c, double* out);

/* ---- the sparse-GP steps either side of the path ----------------------------------
 * The reference's user loop (examples/bernoulli/script.jl:29-39; sparse form docs/src/index.md:154-163) is
 *   qf = marginals(post_u(x));  aux_posterior!(qΩ, lik, y, qf);
 *   S = inv(K_Z⁻¹ + κ·Diagonal(γ)·κᵀ);  m = S·(κ·β + K_Z⁻¹ μ₀(Z)),        κ = K_Z⁻¹ K_{Z,X}  (M×N)
 * `kappa` is that Julia matrix as stored (column-major M×N) = [n][m], inducing index fastest.  m <= 128 runs the
 * fused tensor-core kernels below; m > 128 runs the same steps as a chunked composition of cuBLAS DGEMM / DGEMV calls
 * (libcublas.so.12 bound with dlopen, like NCCL) around the library's own kernels — same results, more passes over kappa.
 * B = K_Z − S (M×M, symmetric; read as given, the quadratic form does not depend on its storage order).
 * Outputs: Pr = [m*m + m] doubles, Pr[i*m + j] = P0[i*m+j] + Σ_t γ_t κ_it κ_jt (exactly symmetric), Pr[m*m + i] =
 * r0[i] + Σ_t β_t κ_it;  P0 / r0 may be NULL (zeros).  Sums over observations run in a fixed order for a given
 * device (bit-reproducible).  Multi-GPU: shard kappa / y by observations, then either pass P0 / r0 on rank 0 only and
 * aug_allreduce_scalars(ctx, Pr, m*m + m) (and the scalar block) afterwards, or — in fused mode (aug_comm_set_fused,
 * m <= 128) — do nothing: the finalise launch of the verb pushes its sums into every rank's mailbox over NVLink, publishes
 * the epoch flag, adds the ranks' sums in rank order and only then adds P0 / r0 (pass the same P0 / r0 on every rank)
 * and writes Pr and the scalars: kernel and collective are one launch, every rank ends with identical bits.
 * FP64 tensor-core kernels (DMMA m8n8k4): 3·m² flops per observation for the fused sweep (2·m² producer + m²
 * consumer, the symmetric half), bound by the FP64 pipe for m >= 32 and by HBM (8·m bytes per observation) below. */

/* row 2 — marginals(post_u(x)), examples/bernoulli/script.jl:32-33 (SVGP posterior with q(u) = N(mvec, S), zero
 * prior mean): mu_t = κ_tᵀ·mvec, var_t = kdiag_t − κ_tᵀ·B·κ_t. */
int32_t aug_sparse_marginals(aug_ctx* ctx, int64_t n, int32_t m, const double* kappa, const double* mvec,
                             const double* B, const double* kdiag, double* mu, double* var);
/* the same with an element stride on the outputs: mu[t*stride], var[t*stride].  stride = nlatent writes latent j of the
 * class-fastest [n][nl] marginals of the Categorical likelihood (categorical.jl:63,84) in place, one call per latent GP
 * (pass mu + j, var + j). */
int32_t aug_sparse_marginals_strided(aug_ctx* ctx, int64_t n, int32_t m, const double* kappa, const double* mvec,
                                     const double* B, const double* kdiag, double* mu, double* var, int64_t stride);
/* row 1 — docs/src/index.md:156-160: P = P0 + κ·Diagonal(gamma)·κᵀ, rhs = r0 + κ·beta. */
int32_t aug_sparse_precision_potential(aug_ctx* ctx, int64_t n, int32_t m, const double* kappa,
                                       const double* gamma, const double* beta, const double* P0,
                                       const double* r0, double* Pr);
/* rows 2 + path + 1 in ONE pass over kappa (scalar-latent likelihoods): marginals → aux_posterior! →
 * expected_auglik_potential_and_precision (+ expected_logtilt / aux_kldivergence sums) → P, rhs.  mu / var /
 * s0 / s1 / s2 / beta / gamma are optional per-observation outputs (NULL: not materialised); scalars as in
 * aug_cavi_step or NULL. */
int32_t aug_sparse_cavi_sweep(aug_ctx* ctx, const aug_lik* lik, int64_t n, int32_t m, const void* y,
                              const double* kappa, const double* mvec, const double* B, const double* kdiag,
                              double* mu, double* var, void* s0, void* s1, void* s2, double* beta,
                              double* gamma, const double* P0, const double* r0, double* Pr, double* scalars);
/* dense (non-sparse) form, examples/bernoulli/script.jl:35-36: P = Kinv + Diagonal(gamma) (n×n, P may alias
 * Kinv), rhs = r0 + beta (r0 = K \ mean(fz) or NULL). */
int32_t aug_dense_precision_potential(aug_ctx* ctx, int64_t n, const double* Kinv, const double* gamma,
                                      const double* beta, const double* r0, double* P, double* rhs);

/* ---- multi-GPU: shard over observations, all-reduce only the scalars ---- */
int32_t aug_comm_get_unique_id(char uid[128]);
int32_t aug_comm_init(aug_ctx* ctx, int32_t nranks, int32_t rank, const char uid[128]);
int32_t aug_comm_destroy(aug_ctx* ctx);
/* in-place sum over ranks of `count` device doubles, on the ctx stream */
int32_t aug_allreduce_scalars(aug_ctx* ctx, double* dev, int32_t count);

/* Peer-memory mailbox over NVLink / NVSwitch: the all-reduce of the scalar block FUSED into the reducing kernel.
 * The block is 64 bytes, so an NCCL call is pure launch + protocol latency; with the mailbox attached the
 * finalising thread of aug_cavi_step / aug_expected_elbo_terms / aug_sampled_loglik_terms pushes its sums into
 * every rank's mailbox with peer stores (each aligned 8-byte word carries 32 data bits and a 32-bit epoch flag, so there is
 * no fence on the publishing side and no ordering between words to rely on), gathers the other ranks' sums from its own
 * mailbox and adds them in rank order (bit-identical on all ranks) before it writes `scalars` — kernel and
 * collective are ONE launch.  Set-up: every rank calls aug_comm_p2p_export, the caller all-gathers the 64-byte
 * cudaIpc handles (torch.distributed / MPI / Julia Distributed), every rank calls aug_comm_p2p_attach, and after a
 * barrier aug_comm_set_fused(ctx, 1).  From then on the scalar-producing verbs are COLLECTIVE: every rank must call
 * the same sequence of them (an empty shard, n = 0, still takes part).  A peer that does not arrive within 5 s
 * (AUGCUDA_XCH_TIMEOUT_MS) makes the kernel write NaN and raise bit 1 of aug_ctx_error_flag instead of hanging.
 * The reference has no counterpart (it is single-process); the sums are those of api.jl:219-223,
 * generic.jl:40-62 over the union of the shards. */
int32_t aug_comm_p2p_export(aug_ctx* ctx, char handle[64], void** local_ptr);
int32_t aug_comm_p2p_attach(aug_ctx* ctx, int32_t nranks, int32_t rank, const char* handles /* nranks x 64 B */);
/* several ctxs inside one process: raw mailbox pointers (local_ptr of each rank's export) and their devices */
int32_t aug_comm_p2p_attach_ptrs(aug_ctx* ctx, int32_t nranks, int32_t rank, void* const* ptrs,
                                 const int32_t* devices);
int32_t aug_comm_p2p_detach(aug_ctx* ctx);
int32_t aug_comm_set_fused(aug_ctx* ctx, int32_t on);
int32_t aug_comm_get_fused(aug_ctx* ctx, int32_t* on);
/* Split-phase exchange (fused mode only).  With on != 0 the finaliser of aug_cavi_step / aug_expected_elbo_terms only
 * PUBLISHES its sums to the peers' mailboxes and leaves this rank's LOCAL sums in the scalar block; the gather — wait
 * for every rank's epoch flag, add the sums in rank order, complete the block — runs later on the same stream: in one
 * extra CTA of the next aug_aux_sample launch (Bernoulli / NegBin / Poisson / heteroscedastic kernels; the CTA is
 * scheduled as that kernel drains, by when the peers have long published) or in a 1-thread kernel at the next
 * aug_comm_flush, aug_ctx_sync, aug_ctx_error_flag or scalar-producing verb.  This takes the per-call barrier between
 * the ranks out of the reducing kernel: a rank may run up to one sampling kernel ahead of its slowest peer (the two
 * epochs of mailbox slots are exactly enough for that), so rank-to-rank skew is absorbed instead of added to every
 * step.  Read the scalars only after one of the completing calls. */
int32_t aug_comm_set_deferred(aug_ctx* ctx, int32_t on);
int32_t aug_comm_flush(aug_ctx* ctx);
/* in-place sum over ranks of count <= 7 device doubles through the mailbox (one 32-thread kernel, no NCCL) */
int32_t aug_allreduce_scalars_p2p(aug_ctx* ctx, double* dev, int32_t count);

/* ---- host-buffer plugin calls (end-to-end path) ------------------------------
 * The reference's verbs take plain host `Vector`s (src/generic.jl:1-88); these are the entry points a method
 * specialised on `Vector` arguments binds to.  Same arguments as the device verbs above, but EVERY data pointer
 * is a HOST buffer (pinned for full speed; pageable works) and scalars come back in a host block of
 * AUG_NSCALARS doubles.  The library stages chunks of the observation axis (2^22 elements) through device
 * memory, overlapping H2D, kernel and D2H on internal streams, and returns when the outputs are on the host.
 * The arithmetic is done by the same kernels as the device verbs: arrays are bit-identical to theirs, samples are
 * bit-identical for the same (seed, offset, i0), scalars are the per-chunk sums added in chunk order.
 * Optional outputs (NULL = not computed / not copied back): s0, s1, s2, beta, gamma, scalars_host — e.g. a
 * Bernoulli caller that already holds beta = sign(y - 1/2)/2 (bernoulli.jl:28) passes beta = NULL and saves
 * 8 B/observation of D2H.  In fused multi-GPU mode these calls stay rank-local (no in-kernel exchange). */

/* init_aux_posterior(T, lik, n): zero-fills host arrays (bernoulli.jl:7-11 ... categorical.jl:59-70) */
int32_t aug_init_aux_posterior_host(aug_ctx* ctx, const aug_lik* lik, int64_t n, void* s0, void* s1, void* s2);
/* aux_posterior!(qΩ, lik, y, qf) — see aug_aux_posterior */
int32_t aug_aux_posterior_host(aug_ctx* ctx, const aug_lik* lik, int64_t n,
                               const void* y, const double* mu, const double* var, int64_t ld,
                               void* s0, void* s1, void* s2);
/* expected_auglik_potential_and_precision(lik, qΩ, y[, qf]) — see aug_expected_potential_precision */
int32_t aug_expected_potential_precision_host(aug_ctx* ctx, const aug_lik* lik, int64_t n,
                                              const void* y, const double* mu, int64_t ld,
                                              const void* s0, const void* s1, const void* s2,
                                              double* beta, double* gamma, int64_t ldo);
/* fused CAVI step — see aug_cavi_step (examples/bernoulli/script.jl:29-39) */
int32_t aug_cavi_step_host(aug_ctx* ctx, const aug_lik* lik, int64_t n,
                           const void* y, const double* mu, const double* var, int64_t ld,
                           void* s0, void* s1, void* s2,
                           double* beta, double* gamma, int64_t ldo,
                           double* scalars_host);
/* expected_logtilt / aux_kldivergence / expected_aug_loglik — see aug_expected_elbo_terms */
int32_t aug_expected_elbo_terms_host(aug_ctx* ctx, const aug_lik* lik, int64_t n,
                                     const void* y, const double* mu, const double* var, int64_t ld,
                                     const void* s0, const void* s1, const void* s2,
                                     double* scalars_host);
/* init_aux_variables(rng, lik, n) — see aug_init_aux_variables */
int32_t aug_init_aux_variables_host(aug_ctx* ctx, const aug_lik* lik, int64_t n, int64_t i0,
                                    double* omega, int64_t* nvar);
/* aux_sample!(rng, Ω, lik, y, f) — see aug_aux_sample (generic.jl:1-12) */
int32_t aug_aux_sample_host(aug_ctx* ctx, const aug_lik* lik, int64_t n, int64_t i0,
                            const void* y, const double* f, int64_t ld,
                            double* omega, int64_t* nvar);
/* auglik_potential_and_precision(lik, Ω, y[, f]) — see aug_potential_precision */
int32_t aug_potential_precision_host(aug_ctx* ctx, const aug_lik* lik, int64_t n,
                                     const void* y, const double* f, int64_t ld,
                                     const double* omega, const int64_t* nvar,
                                     double* beta, double* gamma, int64_t ldo);
/* logtilt / logdensity(aux_prior) / aug_loglik — see aug_sampled_loglik_terms */
int32_t aug_sampled_loglik_terms_host(aug_ctx* ctx, const aug_lik* lik, int64_t n,
                                      const void* y, const double* f, int64_t ld,
                                      const double* omega, const int64_t* nvar,
                                      int32_t with_prior, double* scalars_host);

#ifdef __cplusplus
}
#endif
#endif /* AUGCUDA_H */
