// aug_cat.cu — Categorical (logistic-softmax) likelihood: K-class row kernels.
//
// Layouts (as the reference stores them): y / mu / var / f and the state c, p and the samples ω, n
// are observation-major [n][nl] (class index fastest — the flat view of the ArrayOfSimilarArrays,
// categorical.jl:63,84); β, γ are returned class-major [nl][ldo] (categorical.jl:112-136,
// utils.jl:24).  A CTA owns a tile of R consecutive rows, which is ONE contiguous span of R*nl
// elements, so every global access in the element phases is fully coalesced whatever nl is.
//   phase 1 (element-major): load, c = sqrt(m²+v), σ̃, p; write state; stage p, h (and the
//            p-weighted ELBO factors) in shared memory with an odd row stride;
//   phase 2 (one warp per row): Σ_j p_ij -> 1/p0 (NegativeMultinomial mean, negativemultinomial.jl:54)
//            and the row-wise ELBO sums;
//   phase 3 (class-major): n̄ = p/p0, β = (y − n̄)/2, γ = (y + n̄) h, written transposed and
//            coalesced along the observation axis.
// Reference behaviour replaced: likelihoods/categorical.jl:59-180,
// SpecialDistributions/polyagammanegativemultinomial.jl:27-65, negativemultinomial.jl:35-82.
#include "aug_common.cuh"
#include "aug_math.cuh"
#include "aug_pg.cuh"

namespace {

struct CatArgs {
    int64_t n;
    int nl, nlp, R;           // classes, padded (odd) smem row stride, rows per tile
    const uint8_t* y;
    const double* mu;
    const double* var;
    double* s0;               // c  [n][nl]
    double* s1;               // p  [n][nl]
    uint8_t* s2;              // y copy
    const double* rs0;
    const double* rs1;
    const uint8_t* rs2;
    double* beta;             // [nl][ldo]
    double* gamma;
    int64_t ldo;
    double* partials;
    unsigned int* counter;
    double* scalars;
    unsigned int* dflag;
    LikConst L;
};

// dynamic shared memory carve-up: P, H [, X1, X2, X3] (R*nlp doubles each), rinv[R], rows[R][3], Y bytes
template <bool FROM_STATE, bool ELBO>
__global__ void __launch_bounds__(AUG_BLOCK) cat_cavi_kernel(const CatArgs a) {
    using namespace augm;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int nl = a.nl, nlp = a.nlp, R = a.R;
    const int tile_sz = R * nlp;
    double* P = reinterpret_cast<double*>(smem_raw);
    double* H = P + tile_sz;
    double* X1 = ELBO ? H + tile_sz : nullptr;      // p (−ln2 − m/2 − s2 h/2)
    double* X2 = ELBO ? X1 + tile_sz : nullptr;     // p (lch − c² h/2)
    double* rinv = (ELBO ? X2 + tile_sz : H + tile_sz);
    uint8_t* Y = reinterpret_cast<uint8_t*>(rinv + R);

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t ntiles = (a.n + R - 1) / R;
    double acc[3] = {0.0, 0.0, 0.0};                // elt, kl, flags
    const double inv_denom = 1.0 / a.L.c0;

    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t row0 = tile * R;
        const int rows = (int)min((int64_t)R, a.n - row0);
        const int E = rows * nl;
        const int64_t base = row0 * nl;
        // ---- phase 1
        for (int e = threadIdx.x; e < E; e += AUG_BLOCK) {
            const int i = e / nl, j = e - i * nl;
            const int se = i * nlp + j;
            const double yv = (double)__ldg(a.y + base + e);
            double m = 0.0, v = 0.0, c, p;
            if (!FROM_STATE || ELBO) {
                m = ld_stream1(a.mu + base + e);
                v = ld_stream1(a.var + base + e);
            }
            const double s2m = fma(m, m, v);
            // straight-line fast math when the element is in range, IEEE / libdevice otherwise
            PGTerms t;
            if (FROM_STATE) {
                c = ld_stream1(a.rs0 + base + e);
                p = ld_stream1(a.rs1 + base + e);
                if (c >= 0.0 && c <= 700.0) t = pg_terms<ELBO, false>(c);
                else t = pg_terms<ELBO, true>(c);
            } else if (augf::in_range(s2m) && fabs(m) <= 700.0 && v <= 4e5) {
                double ic;
                augf::sqrt_inv(s2m, c, ic);                           // categorical.jl:88,105
                t = pg_terms_ic<ELBO, false>(c, ic);
                p = approx_expected_logistic<false>(-m, c, t) * inv_denom;   // :90-92, :107
            } else {
                c = sqrt(s2m);
                t = pg_terms_ic<ELBO, true>(c, 0.0);
                p = approx_expected_logistic<true>(-m, c, t) * inv_denom;
            }
            double ys = yv;
            if (FROM_STATE && a.rs2) ys = (double)__ldg(a.rs2 + base + e);
            if (!FROM_STATE) {
                if (a.s0) st_stream1(a.s0 + base + e, c);
                if (a.s1) st_stream1(a.s1 + base + e, p);
                if (a.s2) a.s2[base + e] = (uint8_t)yv;               // φᵢ.y .= y[i]  :89,106
            }
            P[se] = p;
            H[se] = t.h;
            Y[i * nl + j] = (uint8_t)((yv != 0.0) | ((ys != 0.0) << 1));
            if (ELBO) {
                const double hb = 0.5 * s2m * t.h;
                const double d = fma(-0.5 * c * c, t.h, t.lch);
                X1[se] = p * (-LN2 - 0.5 * m - hb);
                X2[se] = p * d;
                // row-independent parts: y_j (−ln2 + m/2 − s2 h/2) and y_j (lch − c² h/2)
                acc[0] += yv * (-LN2 + 0.5 * m - hb);
                acc[1] += ys * d;
                if (ys != yv) {                       // state copy of y differs from the argument (rare)
                    acc[0] += (ys - yv) * (-hb);
                }
            }
        }
        __syncthreads();
        // ---- phase 2: one warp per row
        for (int i = warp; i < rows; i += AUG_BLOCK / 32) {
            double sp = 0.0, sx1 = 0.0, sx2 = 0.0, sx3 = 0.0;
            for (int j = lane; j < nl; j += 32) {
                const double p = P[i * nlp + j];
                sp += p;
                if (ELBO) {
                    sx1 += X1[i * nlp + j];
                    sx2 += X2[i * nlp + j];
                    if (p > 0.0) sx3 += p * ((p >= 1e-290 ? augf::log_(p) : log(p)) - a.L.c2);       // negativemultinomial.jl:79-81
                }
            }
            sp = warp_sum(sp);
            if (ELBO) { sx1 = warp_sum(sx1); sx2 = warp_sum(sx2); sx3 = warp_sum(sx3); }
            if (lane == 0) {
                const double p0 = 1.0 - sp;                           // _p₀ negativemultinomial.jl:27
                const double ri = 1.0 / p0;
                rinv[i] = ri;
                if (!(sp < 1.0)) acc[2] += 1.0;                       // ctor precondition :18-22
                if (ELBO) {
                    acc[0] += ri * sx1;
                    // KL(NM(1,q)||NM(1,p)) = log p0q − log p0p + (1/p0q) Σ q_j (log q_j − log p_j)  :72-82
                    acc[1] += ri * sx2 + (log(p0) - a.L.c3 + ri * sx3);
                }
            }
        }
        __syncthreads();
        // ---- phase 3: class-major, coalesced along the observation axis
        if (a.beta || a.gamma) {
            const int cols = nl;
            for (int t = threadIdx.x; t < rows * cols; t += AUG_BLOCK) {
                const int j = t / rows, i = t - j * rows;
                const int se = i * nlp + j;
                const uint8_t yb = Y[i * nl + j];
                const double yv = (double)(yb & 1), ys = (double)((yb >> 1) & 1);
                const double nbar = P[se] * rinv[i];                  // mean(NM(1,p)) :54
                const int64_t o = (int64_t)j * a.ldo + row0 + i;
                if (a.beta) st_stream1(a.beta + o, 0.5 * (yv - nbar));            // categorical.jl:124,135
                if (a.gamma) st_stream1(a.gamma + o, (ys + nbar) * H[se]);        // :128; pgnm.jl:41-54
            }
        }
        __syncthreads();
    }
    if (ELBO) {
        double out[3];
        if (block_reduce_and_finalize<3>(acc, a.partials, a.counter, out)) {
            a.scalars[AUG_S_EXPECTED_LOGTILT] = out[0];
            a.scalars[AUG_S_KL] = out[1];
            a.scalars[AUG_S_EXPECTED_AUGLL] = out[0] + out[1];
            a.scalars[AUG_S_FLAGS] = out[2];
            if (out[2] > 0.0) atomicOr(a.dflag, 1u);
        }
    } else if (acc[2] > 0.0) {
        atomicOr(a.dflag, 1u);
    }
}

// ------------------------------------------------------------------ Gibbs: aux_sample! for CAT
struct CatSampleArgs {
    int64_t n, i0;
    int nl, R;
    uint64_t seed, offset;
    const uint8_t* y;
    const double* f;
    double* omega;
    int64_t* nvar;
    unsigned int* dflag;
    LikConst L;
};

// categorical.jl:72-78 (p_j = θ_j σ(f_j)/Σθ), negativemultinomial.jl:35-45 (Gamma-Poisson mixture),
// polyagammanegativemultinomial.jl:27-31 (ω_j ~ PG(y_j + n_j, |f_j|))
__global__ void __launch_bounds__(AUG_BLOCK) cat_sample_kernel(const CatSampleArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int nl = a.nl, R = a.R;
    double* P = reinterpret_cast<double*>(smem_raw);   // [R][nl]
    double* rscale = P + R * nl;                       // [R]  τ/(1−p0)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t ntiles = (a.n + R - 1) / R;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t row0 = tile * R;
        const int rows = (int)min((int64_t)R, a.n - row0);
        const int E = rows * nl;
        const int64_t base = row0 * nl;
        for (int e = threadIdx.x; e < E; e += AUG_BLOCK) {
            const int j = e % nl;
            P[e] = __ldg(a.L.theta + j) * augm::logistic(ld_stream1(a.f + base + e));
        }
        __syncthreads();
        for (int i = warp; i < rows; i += AUG_BLOCK / 32) {
            double sp = 0.0;
            for (int j = lane; j < nl; j += 32) sp += P[i * nl + j];
            sp = warp_sum(sp);
            if (lane == 0) {
                const double p0 = 1.0 - sp;
                if (!(sp < 1.0)) atomicOr(a.dflag, 1u);
                augr::Philox g;
                g.init(a.seed, a.offset, (uint64_t)(a.i0 + row0 + i), 1u);   // row stream (tag 1)
                const double tau = g.expo() * (1.0 / p0 - 1.0);              // Gamma(1, 1/p0 − 1)
                rscale[i] = tau / (1.0 - p0);
            }
        }
        __syncthreads();
        for (int e = threadIdx.x; e < E; e += AUG_BLOCK) {
            const int i = e / nl;
            augr::Philox g;
            g.init(a.seed, a.offset, (uint64_t)((a.i0 + row0) * nl + e), 2u);  // element stream (tag 2)
            const int64_t nn = augr::poisson_rand(g, P[e] * rscale[i]);
            const int64_t yv = (int64_t)__ldg(a.y + base + e);
            a.nvar[base + e] = nn;
            st_stream1(a.omega + base + e, augp::pg_draw(g, (double)(nn + yv), true, a.f[base + e], a.L.pgtab));
        }
        __syncthreads();
    }
}

// auglik_potential / auglik_precision (sampled), transposed: categorical.jl:112-119
struct CatPotArgs {
    int64_t n;
    int nl, nlp, R;
    const uint8_t* y;
    const double* omega;
    const int64_t* nvar;
    double* beta;
    double* gamma;
    int64_t ldo;
};

__global__ void __launch_bounds__(AUG_BLOCK) cat_potential_kernel(const CatPotArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int nl = a.nl, nlp = a.nlp, R = a.R;
    double* B = reinterpret_cast<double*>(smem_raw);
    double* G = B + R * nlp;
    const int64_t ntiles = (a.n + R - 1) / R;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t row0 = tile * R;
        const int rows = (int)min((int64_t)R, a.n - row0);
        const int E = rows * nl;
        const int64_t base = row0 * nl;
        for (int e = threadIdx.x; e < E; e += AUG_BLOCK) {
            const int i = e / nl, j = e - i * nl;
            B[i * nlp + j] = 0.5 * ((double)__ldg(a.y + base + e) - (double)__ldg(a.nvar + base + e));
            G[i * nlp + j] = ld_stream1(a.omega + base + e);
        }
        __syncthreads();
        for (int t = threadIdx.x; t < rows * nl; t += AUG_BLOCK) {
            const int j = t / rows, i = t - j * rows;
            const int64_t o = (int64_t)j * a.ldo + row0 + i;
            if (a.beta) st_stream1(a.beta + o, B[i * nlp + j]);
            if (a.gamma) st_stream1(a.gamma + o, G[i * nlp + j]);
        }
        __syncthreads();
    }
}

// rows per tile: ~2048 elements, a multiple of 16 when possible (128-byte segments in the
// transposed stores), shrunk until the tile fits in shared memory
int pick_rows(int nl, size_t bytes_per_elem, size_t fixed_per_row, size_t budget) {
    int R = 2048 / nl;
    if (R >= 16) R &= ~15;
    else if (R >= 2) R &= ~1;
    else R = 2;
    if (R > 256) R = 256;
    const int nlp = nl | 1;
    while (R > 2 && (size_t)R * nlp * bytes_per_elem + (size_t)R * fixed_per_row > budget) R = (R / 2) & ~1;
    if (R < 2) R = 2;
    return R;
}

}  // namespace

int32_t aug_cat_dispatch(aug_ctx* ctx, const aug_lik* lik, int64_t n, const void* y, const double* mu,
                         const double* var, void* s0, void* s1, void* s2, const void* rs0, const void* rs1,
                         const void* rs2, double* beta, double* gamma, int64_t ldo, double* scalars,
                         bool from_state) {
    if (n < 0 || !y) return AUG_ERR_BAD_ARG;
    const bool elbo = scalars != nullptr;
    if (elbo && lik->kind == AUG_CAT) return AUG_ERR_PRECONDITION;   // categorical.jl:165-170
    if (n == 0) {
        if (scalars) AUG_CUDA(cudaMemsetAsync(scalars, 0, AUG_NSCALARS * sizeof(double), ctx->stream));
        return AUG_OK;
    }
    if ((!from_state || elbo) && (!mu || !var)) return AUG_ERR_BAD_ARG;
    if (from_state && (!rs0 || !rs1)) return AUG_ERR_BAD_ARG;
    if ((beta || gamma) && ldo < n) return AUG_ERR_BAD_ARG;
    CatArgs a{};
    a.n = n;
    a.nl = lik->nlatent;
    a.nlp = a.nl | 1;
    a.y = (const uint8_t*)y;
    a.mu = mu;
    a.var = var;
    a.s0 = (double*)s0;
    a.s1 = (double*)s1;
    a.s2 = (uint8_t*)s2;
    a.rs0 = (const double*)rs0;
    a.rs1 = (const double*)rs1;
    a.rs2 = (const uint8_t*)rs2;
    a.beta = beta;
    a.gamma = gamma;
    a.ldo = ldo;
    a.partials = ctx->partials;
    a.counter = ctx->counter;
    a.scalars = scalars;
    a.dflag = ctx->dflag;
    int32_t rc = aug_lik_const(ctx, lik, &a.L, false, false);
    if (rc) return rc;
    const size_t per_elem = (elbo ? 4 : 2) * sizeof(double);
    const size_t budget = 96 * 1024;
    a.R = pick_rows(a.nl, per_elem, sizeof(double) + a.nl, budget);
    const size_t smem = (size_t)a.R * a.nlp * per_elem + (size_t)a.R * sizeof(double) + (size_t)a.R * a.nl + 16;
    if (smem > 200 * 1024) return AUG_ERR_BAD_ARG;   // nl too large for a 2-row tile
    const void* k;
    if (from_state) k = elbo ? (const void*)cat_cavi_kernel<true, true> : (const void*)cat_cavi_kernel<true, false>;
    else k = elbo ? (const void*)cat_cavi_kernel<false, true> : (const void*)cat_cavi_kernel<false, false>;
    AUG_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 1;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k, AUG_BLOCK, smem) != cudaSuccess || occ < 1) occ = 1;
    const int64_t ntiles = (n + a.R - 1) / a.R;
    int64_t grid = (int64_t)ctx->sms * occ;
    if (grid > AUG_MAX_GRID) grid = AUG_MAX_GRID;
    if (grid > ntiles) grid = ntiles;
    void* args[] = {(void*)&a};
    AUG_CUDA(cudaLaunchKernel(k, dim3((unsigned)grid), dim3(AUG_BLOCK), args, smem, ctx->stream));
    ctx->launches++;
    return AUG_OK;
}

int32_t aug_cat_sample(aug_ctx* ctx, const aug_lik* lik, int64_t n, int64_t i0, const void* y, const double* f,
                       double* omega, int64_t* nvar, uint64_t offset) {
    if (!y || !f || !omega || !nvar) return AUG_ERR_BAD_ARG;
    CatSampleArgs a{};
    a.n = n;
    a.i0 = i0;
    a.nl = lik->nlatent;
    a.seed = ctx->seed;
    a.offset = offset;
    a.y = (const uint8_t*)y;
    a.f = f;
    a.omega = omega;
    a.nvar = nvar;
    a.dflag = ctx->dflag;
    int32_t rc = aug_lik_const(ctx, lik, &a.L, false, true);
    if (rc) return rc;
    a.R = pick_rows(a.nl, sizeof(double), sizeof(double), 64 * 1024);
    const size_t smem = (size_t)a.R * a.nl * sizeof(double) + (size_t)a.R * sizeof(double) + 16;
    if (smem > 200 * 1024) return AUG_ERR_BAD_ARG;
    AUG_CUDA(cudaFuncSetAttribute(cat_sample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 1;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, cat_sample_kernel, AUG_BLOCK, smem) != cudaSuccess ||
        occ < 1)
        occ = 1;
    const int64_t ntiles = (n + a.R - 1) / a.R;
    int64_t grid = (int64_t)ctx->sms * occ;
    if (grid > ntiles) grid = ntiles;
    cat_sample_kernel<<<(unsigned)grid, AUG_BLOCK, smem, ctx->stream>>>(a);
    ctx->launches++;
    return (int32_t)cudaGetLastError();
}

int32_t aug_cat_potential(aug_ctx* ctx, const aug_lik* lik, int64_t n, const void* y, const double* omega,
                          const int64_t* nvar, double* beta, double* gamma, int64_t ldo) {
    if (!nvar || ldo < n) return AUG_ERR_BAD_ARG;
    CatPotArgs a{};
    a.n = n;
    a.nl = lik->nlatent;
    a.nlp = a.nl | 1;
    a.y = (const uint8_t*)y;
    a.omega = omega;
    a.nvar = nvar;
    a.beta = beta;
    a.gamma = gamma;
    a.ldo = ldo;
    a.R = pick_rows(a.nl, 2 * sizeof(double), 0, 64 * 1024);
    const size_t smem = (size_t)a.R * a.nlp * 2 * sizeof(double) + 16;
    if (smem > 200 * 1024) return AUG_ERR_BAD_ARG;
    AUG_CUDA(cudaFuncSetAttribute(cat_potential_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 1;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, cat_potential_kernel, AUG_BLOCK, smem) !=
            cudaSuccess || occ < 1)
        occ = 1;
    const int64_t ntiles = (n + a.R - 1) / a.R;
    int64_t grid = (int64_t)ctx->sms * occ;
    if (grid > ntiles) grid = ntiles;
    cat_potential_kernel<<<(unsigned)grid, AUG_BLOCK, smem, ctx->stream>>>(a);
    ctx->launches++;
    return (int32_t)cudaGetLastError();
}
