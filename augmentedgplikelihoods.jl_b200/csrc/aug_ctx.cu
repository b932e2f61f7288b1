// aug_ctx.cu — context, likelihood constants, memory helpers, NCCL plumbing and the C-ABI entry
// points that route each verb to its kernels.
#include <dlfcn.h>
#include <math.h>
#include <stdio.h>
#include <string.h>

#include <map>
#include <mutex>

#include "aug_common.cuh"

void aug_cublas_destroy(aug_ctx* c);   // aug_sparse.cu: releases the cuBLAS handle of the m > 128 composition

// kernels implemented in the other translation units
int32_t aug_cavi_dispatch(aug_ctx* ctx, const aug_lik* lik, int64_t n, const void* y, const double* mu,
                          const double* var, int64_t ld, void* s0, void* s1, void* s2, const void* rs0,
                          const void* rs1, const void* rs2, double* beta, double* gamma, int64_t ldo,
                          double* scalars, bool from_state);
int32_t aug_cat_dispatch(aug_ctx* ctx, const aug_lik* lik, int64_t n, const void* y, const double* mu,
                         const double* var, void* s0, void* s1, void* s2, const void* rs0, const void* rs1,
                         const void* rs2, double* beta, double* gamma, int64_t ldo, double* scalars,
                         bool from_state);

namespace {

const double LOGTWO = 0.69314718055994530942;
const double PI_ = 3.14159265358979323846;

// digamma for the Student-t KL constant (host, once per call)
double digamma_host(double x) {
    double r = 0.0;
    while (x < 12.0) { r -= 1.0 / x; x += 1.0; }
    const double f = 1.0 / (x * x);
    const double t = f * (-1.0 / 12 + f * (1.0 / 120 + f * (-1.0 / 252 + f * (1.0 / 240 + f * (-1.0 / 132 +
                     f * (691.0 / 32760 + f * (-1.0 / 12)))))));
    return r + log(x) - 0.5 / x + t;
}

bool is_cat(int kind) { return kind == AUG_CAT || kind == AUG_CAT_BIJ; }

// r(z) = mass of the truncated-exponential part of Devroye's PG(1,z) proposal (mass_texpon,
// SpecialDistributions/polyagamma.jl:179-192), evaluated on the host in the log domain.
double pg_tail_mass_host(double z) {
    const double t = 0.64, K = PI_ * PI_ / 8 + z * z / 2;
    const double b = sqrt(1 / t) * (t * z - 1), a = -sqrt(1 / t) * (t * z + 1);
    auto logcdf = [](double x) {   // log Phi(x)
        const double y = -x * 0.70710678118654752440;
        if (y > 5.0) {             // log(erfc(y)/2) through the scaled complementary error function
            const double w = 1.0 / (y * y);
            const double asym = 1 + w * (-0.5 + w * (0.75 + w * (-1.875 + w * (6.5625 + w * (-29.53125 + w * 162.421875)))));
            return y < 25.0 ? log(erfc(y) / 2) : -y * y + log(asym / (y * sqrt(PI_)) / 2);
        }
        return log(erfc(y) / 2);
    };
    const double x0 = log(K) + K * t;
    const double xb = x0 - z + logcdf(b), xa = x0 + z + logcdf(a);
    const double qdivp = 4 / PI_ * (exp(xb) + exp(xa));
    return 1 / (1 + qdivp);
}

// Degree-7 interpolant of r on each [k h, (k+1) h) at the Chebyshev nodes, stored as monomial coefficients in
// s = 2 (z - k h)/h - 1; interpolation error < 1e-15.
void build_pg_table(double* tab) {
    const int D = AUG_PGTAB_DEG;
    for (int k = 0; k < AUG_PGTAB_N; ++k) {
        const double a = k * AUG_PGTAB_H;
        double fx[AUG_PGTAB_DEG], cheb[AUG_PGTAB_DEG];
        for (int j = 0; j < D; ++j) {
            const double xj = cos(PI_ * (j + 0.5) / D);
            fx[j] = pg_tail_mass_host(a + 0.5 * AUG_PGTAB_H * (xj + 1));
        }
        for (int m = 0; m < D; ++m) {
            double acc = 0;
            for (int j = 0; j < D; ++j) acc += fx[j] * cos(PI_ * m * (j + 0.5) / D);
            cheb[m] = acc * (m == 0 ? 1.0 : 2.0) / D;
        }
        // Chebyshev -> monomial: T_0 = 1, T_1 = s, T_{m+1} = 2 s T_m - T_{m-1}
        double mono[AUG_PGTAB_DEG] = {0}, Tm1[AUG_PGTAB_DEG] = {0}, Tm[AUG_PGTAB_DEG] = {0}, Tn[AUG_PGTAB_DEG];
        Tm1[0] = 1;
        Tm[1] = 1;
        for (int i = 0; i < D; ++i) mono[i] += cheb[0] * Tm1[i] + cheb[1] * Tm[i];
        for (int m = 2; m < D; ++m) {
            for (int i = 0; i < D; ++i) Tn[i] = (i > 0 ? 2 * Tm[i - 1] : 0) - Tm1[i];
            for (int i = 0; i < D; ++i) { mono[i] += cheb[m] * Tn[i]; Tm1[i] = Tm[i]; Tm[i] = Tn[i]; }
        }
        for (int i = 0; i < D; ++i) tab[k * D + i] = mono[D - 1 - i];   // highest degree first (Horner order)
    }
}

int32_t check_lik(const aug_lik* lik) {
    if (lik == nullptr) return AUG_ERR_BAD_ARG;
    if (lik->kind < 0 || lik->kind >= AUG_NKINDS) return AUG_ERR_BAD_KIND;
    if (is_cat(lik->kind) && lik->nlatent < 1) return AUG_ERR_BAD_ARG;
    if (lik->kind == AUG_HETERO && lik->nlatent != 2) return AUG_ERR_BAD_ARG;
    return AUG_OK;
}

std::mutex g_occ_mutex;
std::map<const void*, int> g_occ;

}  // namespace

// CTAs for a grid-stride kernel: enough to cover the work, at most one resident wave
// (SM count x occupancy), so every CTA is co-resident and the last-block ticket is cheap.
int aug_grid_for(aug_ctx* ctx, const void* kernel, int64_t work_items, int items_per_block) {
    int occ = 0;
    {
        std::lock_guard<std::mutex> lk(g_occ_mutex);
        auto it = g_occ.find(kernel);
        if (it != g_occ.end()) occ = it->second;
    }
    if (occ == 0) {
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, AUG_BLOCK, 0) != cudaSuccess || occ < 1)
            occ = 2;
        std::lock_guard<std::mutex> lk(g_occ_mutex);
        g_occ[kernel] = occ;
    }
    int64_t want = (work_items + items_per_block - 1) / items_per_block;
    int64_t cap = (int64_t)ctx->sms * occ;
    if (cap > AUG_MAX_GRID) cap = AUG_MAX_GRID;
    if (want > cap) want = cap;
    if (want < 1) want = 1;
    return (int)want;
}

// Host-side constants of one likelihood (a handful of libm calls per ABI call, never per observation).
int32_t aug_lik_const(aug_ctx* ctx, const aug_lik* lik, LikConst* L, bool need_table, bool need_theta) {
    int32_t rc = check_lik(lik);
    if (rc) return rc;
    memset(L, 0, sizeof(*L));
    L->kind = lik->kind;
    L->nl = lik->nlatent;
    L->r_is_int = lik->r_is_int;
    L->quirks = (lik->flags & AUG_LIK_FAITHFUL_QUIRKS) ? 1 : 0;
    L->p0 = lik->p[0];
    L->p1 = lik->p[1];
    L->pgtab = ctx->pgtab;
    double tab_param = 0.0;
    bool want_table = false;
    switch (lik->kind) {
        case AUG_BERNOULLI: break;
        case AUG_NEGBIN:
            if (!(lik->p[0] > 0.0)) return AUG_ERR_BAD_ARG;
            L->c0 = lgamma(lik->p[0]);
            want_table = need_table;
            tab_param = lik->p[0];
            break;
        case AUG_POISSON:
            if (!(lik->p[0] > 0.0)) return AUG_ERR_BAD_ARG;
            L->c0 = log(lik->p[0]);
            want_table = need_table;
            break;
        case AUG_LAPLACE: {
            const double beta = lik->p[0];
            if (!(beta > 0.0)) return AUG_ERR_BAD_ARG;
            const double lam = 1.0 / ((2 * beta) * (2 * beta));            // laplace_λ laplace.jl:25
            L->c0 = 1.0 / (2 * beta);
            L->c1 = lam;
            L->c2 = lgamma(0.5) - log(sqrt(PI_)) - log(2 * beta);           // laplace.jl:84
            L->c3 = log(2 * lam) / 2 - log(2 * PI_) / 2 - log(lam) / 2 + lgamma(0.5);  // laplace.jl:102
            break;
        }
        case AUG_STUDENTT: {
            const double nu = lik->p[0], sig = lik->p[1];
            if (!(nu > 0.0) || !(sig > 0.0)) return AUG_ERR_BAD_ARG;
            const double alpha = (nu + 1) / 2, aq = nu / 2, tq = sig * sig / aq;   // studentt.jl:27,91
            L->c0 = nu / (sig * sig);
            L->c1 = alpha;
            L->c2 = (alpha - aq) * digamma_host(alpha) - lgamma(alpha) + lgamma(aq) - alpha;
            L->c3 = aq;
            L->c4 = tq;
            L->c5 = -0.5 * log(2 * PI_);
            L->p1 = log(alpha * tq);   // log(θ/r): lets the kernel reuse log θ for the KL's log r
            break;
        }
        case AUG_HETERO:
            if (!(lik->p[0] > 0.0)) return AUG_ERR_BAD_ARG;
            L->c0 = 0.5 * (log(lik->p[0]) + log(2.0 / PI_));                // hetero :133
            break;
        case AUG_CAT:
        case AUG_CAT_BIJ: {
            const int nl = lik->nlatent;
            const bool bij = lik->kind == AUG_CAT_BIJ;
            const int K = bij ? nl + 1 : nl;
            L->bij = bij;
            auto lt = [&](int k) { return lik->logtheta ? lik->logtheta[k] : 0.0; };
            const int m = bij ? K - 1 : K;
            double mx = -INFINITY;
            for (int k = 0; k < m; ++k) mx = fmax(mx, lt(k));
            double acc = 0.0;
            for (int k = 0; k < m; ++k) acc += exp(lt(k) - mx);
            const double se = exp(mx + log(acc));                            // exp(logsumexp) categorical.jl:16-19
            double D = 0.0, sum_theta, prior_p;
            if (bij) {
                D = exp(lt(K - 1)) * 0.5;                                    // _get_const :12-14
                sum_theta = D + se;                                          // :18-20
                L->c0 = D + nl;                                              // :92
                prior_p = 1.0 / sum_theta;                                   // :155
            } else {
                sum_theta = se;
                L->c0 = nl;                                                  // :107
                prior_p = 1.0 / nl;                                          // :161
            }
            double spp = 0.0;
            for (int j = 0; j < nl; ++j) spp += prior_p;
            L->c1 = prior_p;
            L->c2 = log(prior_p);
            L->c3 = log(1.0 - spp);                                          // log p0 of the prior NM
            L->c4 = sum_theta;
            if (need_theta) {
                double* h = (double*)malloc(sizeof(double) * nl);
                if (!h) return (int32_t)cudaErrorMemoryAllocation;
                for (int j = 0; j < nl; ++j) h[j] = exp(lt(j)) / sum_theta;  // _scale_σf / _sum_θ :72-78
                // cached by value, like the integer-y table below: the chunks of a host-buffer call and the iterations of
                // a Gibbs loop pass the same logθ, and only a change of it costs an upload + stream synchronisation
                if (ctx->htheta && ctx->htheta_n == nl && memcmp(ctx->htheta, h, sizeof(double) * nl) == 0) {
                    free(h);
                } else {
                    ctx->htheta_n = 0;
                    free(ctx->htheta);
                    ctx->htheta = nullptr;
                    if (ctx->dtheta_cap < nl) {
                        if (ctx->dtheta) cudaFree(ctx->dtheta);
                        ctx->dtheta = nullptr;
                        ctx->dtheta_cap = 0;
                        cudaError_t em = cudaMalloc(&ctx->dtheta, sizeof(double) * nl);
                        if (em != cudaSuccess) { free(h); return (int32_t)em; }
                        ctx->dtheta_cap = nl;
                    }
                    cudaError_t e = cudaMemcpyAsync(ctx->dtheta, h, sizeof(double) * nl, cudaMemcpyHostToDevice,
                                                    ctx->stream);
                    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
                    if (e != cudaSuccess) { free(h); return (int32_t)e; }
                    ctx->htheta = h;
                    ctx->htheta_n = nl;
                }
                L->theta = ctx->dtheta;
            }
            break;
        }
        default: return AUG_ERR_BAD_KIND;
    }
    if (want_table) {
        if (ctx->table_kind != lik->kind || ctx->table_param != tab_param ||
            ctx->table_r_is_int != lik->r_is_int) {
            double h[AUG_TABLE_N];
            for (int y = 0; y < AUG_TABLE_N; ++y) {
                if (lik->kind == AUG_POISSON) {
                    h[y] = lgamma(y + 1.0);                                  // logfactorial(y) poisson.jl:83
                } else {
                    const double r = lik->p[0];                              // negbin_logconst :51-52
                    if (!lik->r_is_int) h[y] = lgamma(y + r) - lgamma(y + 1.0) - lgamma(r);
                    else {
                        const double n = y + r - 1, k = y;
                        h[y] = -log1p(n) - (lgamma(n - k + 1) + lgamma(k + 1) - lgamma(n + 2));
                    }
                }
            }
            cudaError_t e = cudaMemcpyAsync(ctx->table, h, sizeof(h), cudaMemcpyHostToDevice, ctx->stream);
            if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
            if (e != cudaSuccess) return (int32_t)e;
            ctx->table_kind = lik->kind;
            ctx->table_param = tab_param;
            ctx->table_r_is_int = lik->r_is_int;
        }
        L->table = ctx->table;
    }
    return AUG_OK;
}

// ---------------------------------------------------------------------------- NCCL via dlopen
// libaugcuda.so does not link libnccl: it binds to whichever libnccl.so.2 the process already
// holds (e.g. the one torch loaded) or loads the system one, so there is a single NCCL per process.
namespace {
typedef struct { char internal[128]; } nccl_uid;
typedef int (*fn_getuid)(nccl_uid*);
typedef int (*fn_initrank)(void**, int, nccl_uid, int);
typedef int (*fn_allreduce)(const void*, void*, size_t, int, int, void*, cudaStream_t);
typedef int (*fn_destroy)(void*);
struct NcclApi {
    void* lib = nullptr;
    fn_getuid getuid = nullptr;
    fn_initrank initrank = nullptr;
    fn_allreduce allreduce = nullptr;
    fn_destroy destroy = nullptr;
} g_nccl;

int32_t nccl_load() {
    if (g_nccl.lib) return AUG_OK;
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return AUG_ERR_NO_NCCL;
    g_nccl.getuid = (fn_getuid)dlsym(h, "ncclGetUniqueId");
    g_nccl.initrank = (fn_initrank)dlsym(h, "ncclCommInitRank");
    g_nccl.allreduce = (fn_allreduce)dlsym(h, "ncclAllReduce");
    g_nccl.destroy = (fn_destroy)dlsym(h, "ncclCommDestroy");
    if (!g_nccl.getuid || !g_nccl.initrank || !g_nccl.allreduce || !g_nccl.destroy) return AUG_ERR_NO_NCCL;
    g_nccl.lib = h;
    return AUG_OK;
}
const int NCCL_FLOAT64 = 8;  // ncclDouble
const int NCCL_SUM = 0;      // ncclSum
}  // namespace

void aug_pipe_destroy(aug_ctx* ctx);  // aug_host.cu

namespace {
__global__ void xch_only_kernel(AugXchDev* x, double* vals, int count, int first, int mode) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    double v[7];
#pragma unroll
    for (int k = 0; k < 7; ++k) v[k] = (mode == 0 && k < count) ? vals[k] : 0.0;
    xch_allreduce<7>(x, v);
    if (mode == 0) {                       // plain all-reduce of vals[0..count)
        for (int k = 0; k < count; ++k) vals[k] = v[k];
    } else {                               // zero contribution of an empty shard to a verb's slots
        vals[first] = v[0];
        vals[first + 1] = v[1];
        vals[first + 2] = v[0] + v[1];
        if (count == 3) vals[AUG_S_FLAGS] = v[2];
    }
}
__global__ void xch_finish_kernel(AugXchDev* x) {
    if (threadIdx.x == 0 && blockIdx.x == 0) xch_finish_pending(x);
}
}  // namespace

int32_t aug_xch_flush(aug_ctx* c) {
    if (!c->pending || !c->xch) return AUG_OK;
    c->pending = 0;
    xch_finish_kernel<<<1, 32, 0, c->stream>>>(c->xch);
    c->launches++;
    return (int32_t)cudaGetLastError();
}

extern "C" {
static void p2p_close(aug_ctx* c);


int32_t aug_version(void) { return AUGCUDA_VERSION; }

const char* aug_strerror(int32_t rc) {
    switch (rc) {
        case AUG_OK: return "ok";
        case AUG_ERR_BAD_KIND: return "unknown likelihood kind";
        case AUG_ERR_BAD_ARG: return "bad argument (null pointer, negative size, leading dimension or parameter)";
        case AUG_ERR_PRECONDITION:
            return "precondition of the reference violated (non-bijective logistic-softmax KL, or sum(p) >= 1)";
        case AUG_ERR_NOT_INIT: return "context or communicator not initialised";
        case AUG_ERR_NO_NCCL: return "libnccl.so.2 could not be loaded";
        case AUG_ERR_DEVICE_FLAG: return "a kernel raised the device-side error flag";
        case AUG_ERR_NO_CUBLAS: return "libcublas.so.12 could not be loaded (needed for m > 128 inducing points)";
        default: break;
    }
    if (rc >= 2000) return "cuBLAS error (cublasStatus_t = rc - 2000)";
    if (rc >= 1000) return "NCCL error (ncclResult_t = rc - 1000)";
    if (rc > 0) return cudaGetErrorString((cudaError_t)rc);
    return "unknown error";
}

int32_t aug_ctx_create(aug_ctx** out, int32_t device, void* stream) {
    if (out == nullptr) return AUG_ERR_BAD_ARG;
    *out = nullptr;
    int ndev = 0;
    AUG_CUDA(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) return AUG_ERR_BAD_ARG;
    AUG_CUDA(cudaSetDevice(device));
    aug_ctx* c = new aug_ctx();
    memset(c, 0, sizeof(*c));
    c->device = device;
    c->table_kind = -1;
    if (stream) {
        c->stream = (cudaStream_t)stream;
        c->own_stream = false;
    } else {
        cudaError_t e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
        if (e != cudaSuccess) { delete c; return (int32_t)e; }
        c->own_stream = true;
    }
    cudaDeviceProp prop;
    cudaError_t e = cudaGetDeviceProperties(&prop, device);
    if (e == cudaSuccess) {
        c->sms = prop.multiProcessorCount;
        c->smem_per_sm = (int)prop.sharedMemPerMultiprocessor;     // 233472 on sm_100
        c->smem_optin = (int)prop.sharedMemPerBlockOptin;          // 232448
        e = cudaMalloc(&c->partials, sizeof(double) * AUG_MAX_GRID * AUG_NRED);
    }
    if (e == cudaSuccess) e = cudaMalloc(&c->counter, sizeof(unsigned int) * 4);
    if (e == cudaSuccess) e = cudaMemset(c->counter, 0, sizeof(unsigned int) * 4);
    if (e == cudaSuccess) e = cudaMalloc(&c->dscalars, sizeof(double) * AUG_NSCALARS);
    if (e == cudaSuccess) e = cudaMalloc(&c->dflag, sizeof(unsigned int));
    if (e == cudaSuccess) e = cudaMemset(c->dflag, 0, sizeof(unsigned int));
    if (e == cudaSuccess) e = cudaMalloc(&c->table, sizeof(double) * AUG_TABLE_N);
    if (e == cudaSuccess) e = cudaMalloc(&c->pgtab, sizeof(double) * AUG_PGTAB_N * AUG_PGTAB_DEG);
    if (e == cudaSuccess) {
        static double htab[AUG_PGTAB_N * AUG_PGTAB_DEG];
        static bool built = false;
        if (!built) { build_pg_table(htab); built = true; }
        e = cudaMemcpy(c->pgtab, htab, sizeof(htab), cudaMemcpyHostToDevice);
    }
    if (e != cudaSuccess) { aug_ctx_destroy(c); return (int32_t)e; }
    c->seed = 0x243F6A8885A308D3ull;
    c->offset = 0;
    *out = c;
    return AUG_OK;
}

int32_t aug_ctx_destroy(aug_ctx* c) {
    if (c == nullptr) return AUG_OK;
    cudaSetDevice(c->device);
    aug_pipe_destroy(c);
    if (c->nccl_comm && g_nccl.destroy) g_nccl.destroy(c->nccl_comm);
    p2p_close(c);
    if (c->mailbox) cudaFree(c->mailbox);
    if (c->partials) cudaFree(c->partials);
    if (c->counter) cudaFree(c->counter);
    if (c->dscalars) cudaFree(c->dscalars);
    if (c->dflag) cudaFree(c->dflag);
    if (c->table) cudaFree(c->table);
    if (c->pgtab) cudaFree(c->pgtab);
    if (c->dtheta) cudaFree(c->dtheta);
    free(c->htheta);
    if (c->sparse_scratch) cudaFree(c->sparse_scratch);
    aug_cublas_destroy(c);
    if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
    delete c;
    return AUG_OK;
}

int32_t aug_ctx_seed(aug_ctx* c, uint64_t seed, uint64_t offset) {
    if (!c) return AUG_ERR_NOT_INIT;
    c->seed = seed;
    c->offset = offset;
    return AUG_OK;
}
int32_t aug_ctx_get_offset(aug_ctx* c, uint64_t* offset) {
    if (!c || !offset) return AUG_ERR_NOT_INIT;
    *offset = c->offset;
    return AUG_OK;
}
int32_t aug_ctx_sync(aug_ctx* c) {
    if (!c) return AUG_ERR_NOT_INIT;
    AUG_CUDA(cudaSetDevice(c->device));
    { int32_t rc = aug_xch_flush(c); if (rc) return rc; }     // a pending split-phase exchange completes here
    AUG_CUDA(cudaStreamSynchronize(c->stream));
    return AUG_OK;
}
int32_t aug_ctx_stream(aug_ctx* c, void** stream) {
    if (!c || !stream) return AUG_ERR_NOT_INIT;
    *stream = (void*)c->stream;
    return AUG_OK;
}
int32_t aug_ctx_sm_count(aug_ctx* c, int32_t* n) {
    if (!c || !n) return AUG_ERR_NOT_INIT;
    *n = c->sms;
    return AUG_OK;
}
int32_t aug_ctx_launch_count(aug_ctx* c, uint64_t* n) {
    if (!c || !n) return AUG_ERR_NOT_INIT;
    *n = c->launches;
    return AUG_OK;
}
int32_t aug_ctx_error_flag(aug_ctx* c, uint32_t* flag) {
    if (!c || !flag) return AUG_ERR_NOT_INIT;
    AUG_CUDA(cudaSetDevice(c->device));
    { int32_t rc = aug_xch_flush(c); if (rc) return rc; }
    AUG_CUDA(cudaMemcpyAsync(flag, c->dflag, sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
    AUG_CUDA(cudaMemsetAsync(c->dflag, 0, sizeof(uint32_t), c->stream));
    AUG_CUDA(cudaStreamSynchronize(c->stream));
    return AUG_OK;
}

int32_t aug_malloc(aug_ctx* c, void** dev, size_t bytes) {
    if (!c || !dev) return AUG_ERR_NOT_INIT;
    AUG_CUDA(cudaSetDevice(c->device));
    AUG_CUDA(cudaMalloc(dev, bytes ? bytes : 1));
    return AUG_OK;
}
int32_t aug_free(aug_ctx* c, void* dev) {
    if (!c) return AUG_ERR_NOT_INIT;
    AUG_CUDA(cudaSetDevice(c->device));
    AUG_CUDA(cudaFree(dev));
    return AUG_OK;
}
int32_t aug_host_alloc(void** host, size_t bytes) {
    if (!host) return AUG_ERR_BAD_ARG;
    AUG_CUDA(cudaHostAlloc(host, bytes ? bytes : 1, cudaHostAllocDefault));
    return AUG_OK;
}
int32_t aug_host_free(void* host) {
    AUG_CUDA(cudaFreeHost(host));
    return AUG_OK;
}
int32_t aug_memcpy_h2d(aug_ctx* c, void* dev, const void* host, size_t bytes) {
    if (!c) return AUG_ERR_NOT_INIT;
    AUG_CUDA(cudaSetDevice(c->device));
    AUG_CUDA(cudaMemcpyAsync(dev, host, bytes, cudaMemcpyHostToDevice, c->stream));
    return AUG_OK;
}
int32_t aug_memcpy_d2h(aug_ctx* c, void* host, const void* dev, size_t bytes) {
    if (!c) return AUG_ERR_NOT_INIT;
    AUG_CUDA(cudaSetDevice(c->device));
    AUG_CUDA(cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, c->stream));
    AUG_CUDA(cudaStreamSynchronize(c->stream));
    return AUG_OK;
}

// ---------------------------------------------------------------------------- variational verbs
int32_t aug_init_aux_posterior(aug_ctx* c, const aug_lik* lik, int64_t n, void* s0, void* s1, void* s2) {
    if (!c) return AUG_ERR_NOT_INIT;
    int32_t rc = check_lik(lik);
    if (rc) return rc;
    if (n < 0) return AUG_ERR_BAD_ARG;
    AUG_CUDA(cudaSetDevice(c->device));
    const size_t m = is_cat(lik->kind) ? (size_t)n * lik->nlatent : (size_t)n;
    if (s0) AUG_CUDA(cudaMemsetAsync(s0, 0, m * 8, c->stream));
    if (s1) AUG_CUDA(cudaMemsetAsync(s1, 0, m * 8, c->stream));
    if (s2) AUG_CUDA(cudaMemsetAsync(s2, 0, m * (is_cat(lik->kind) ? 1 : 8), c->stream));
    return AUG_OK;
}

int32_t aug_cavi_step(aug_ctx* c, const aug_lik* lik, int64_t n, const void* y, const double* mu,
                      const double* var, int64_t ld, void* s0, void* s1, void* s2, double* beta, double* gamma,
                      int64_t ldo, double* scalars) {
    if (!c) return AUG_ERR_NOT_INIT;
    int32_t rc = check_lik(lik);
    if (rc) return rc;
    AUG_CUDA(cudaSetDevice(c->device));
    if (is_cat(lik->kind))
        return aug_cat_dispatch(c, lik, n, y, mu, var, s0, s1, s2, nullptr, nullptr, nullptr, beta, gamma, ldo,
                                scalars, false);
    return aug_cavi_dispatch(c, lik, n, y, mu, var, ld, s0, s1, s2, nullptr, nullptr, nullptr, beta, gamma, ldo,
                             scalars, false);
}

int32_t aug_aux_posterior(aug_ctx* c, const aug_lik* lik, int64_t n, const void* y, const double* mu,
                          const double* var, int64_t ld, void* s0, void* s1, void* s2) {
    return aug_cavi_step(c, lik, n, y, mu, var, ld, s0, s1, s2, nullptr, nullptr, 0, nullptr);
}

int32_t aug_expected_potential_precision(aug_ctx* c, const aug_lik* lik, int64_t n, const void* y,
                                         const double* mu, int64_t ld, const void* s0, const void* s1,
                                         const void* s2, double* beta, double* gamma, int64_t ldo) {
    if (!c) return AUG_ERR_NOT_INIT;
    int32_t rc = check_lik(lik);
    if (rc) return rc;
    AUG_CUDA(cudaSetDevice(c->device));
    if (is_cat(lik->kind))
        return aug_cat_dispatch(c, lik, n, y, nullptr, nullptr, nullptr, nullptr, nullptr, s0, s1, s2, beta,
                                gamma, ldo, nullptr, true);
    return aug_cavi_dispatch(c, lik, n, y, mu, nullptr, ld, nullptr, nullptr, nullptr, s0, s1, s2, beta, gamma,
                             ldo, nullptr, true);
}

int32_t aug_expected_elbo_terms(aug_ctx* c, const aug_lik* lik, int64_t n, const void* y, const double* mu,
                                const double* var, int64_t ld, const void* s0, const void* s1, const void* s2,
                                double* scalars) {
    if (!c) return AUG_ERR_NOT_INIT;
    if (!scalars) return AUG_ERR_BAD_ARG;
    int32_t rc = check_lik(lik);
    if (rc) return rc;
    AUG_CUDA(cudaSetDevice(c->device));
    if (is_cat(lik->kind))
        return aug_cat_dispatch(c, lik, n, y, mu, var, nullptr, nullptr, nullptr, s0, s1, s2, nullptr, nullptr, 0,
                                scalars, true);
    return aug_cavi_dispatch(c, lik, n, y, mu, var, ld, nullptr, nullptr, nullptr, s0, s1, s2, nullptr, nullptr, 0,
                             scalars, true);
}

// ---------------------------------------------------------------------------- collective
int32_t aug_comm_get_unique_id(char uid[128]) {
    int32_t rc = nccl_load();
    if (rc) return rc;
    nccl_uid u;
    int r = g_nccl.getuid(&u);
    if (r) return 1000 + r;
    memcpy(uid, u.internal, 128);
    return AUG_OK;
}
int32_t aug_comm_init(aug_ctx* c, int32_t nranks, int32_t rank, const char uid[128]) {
    if (!c) return AUG_ERR_NOT_INIT;
    if (nranks < 1 || rank < 0 || rank >= nranks || !uid) return AUG_ERR_BAD_ARG;
    int32_t rc = nccl_load();
    if (rc) return rc;
    AUG_CUDA(cudaSetDevice(c->device));
    if (c->nccl_comm) { g_nccl.destroy(c->nccl_comm); c->nccl_comm = nullptr; }
    nccl_uid u;
    memcpy(u.internal, uid, 128);
    int r = g_nccl.initrank(&c->nccl_comm, nranks, u, rank);
    if (r) { c->nccl_comm = nullptr; return 1000 + r; }
    c->nranks = nranks;
    c->rank = rank;
    return AUG_OK;
}
int32_t aug_comm_destroy(aug_ctx* c) {
    if (!c) return AUG_ERR_NOT_INIT;
    if (c->nccl_comm && g_nccl.destroy) g_nccl.destroy(c->nccl_comm);
    c->nccl_comm = nullptr;
    c->nranks = 0;
    return AUG_OK;
}
// ---------------------------------------------------------------------------- peer-memory mailbox (NVLink / NVSwitch)
// The scalar block is 64 bytes: an NCCL all-reduce of it is pure launch + protocol latency.  With the mailbox attached
// (one cudaMalloc'ed kilobyte per rank, exported by cudaIpc or shared as a raw pointer inside one process), the
// finalising thread of the reducing kernel itself pushes the sums to all peers and gathers theirs (xch_allreduce,
// aug_common.cuh): kernel + collective in ONE launch, no host round trip, deterministic rank-order sum.
int32_t aug_comm_p2p_export(aug_ctx* c, char handle[64], void** local_ptr) {
    if (!c) return AUG_ERR_NOT_INIT;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
    AUG_CUDA(cudaSetDevice(c->device));
    if (!c->mailbox) AUG_CUDA(cudaMalloc(&c->mailbox, sizeof(unsigned long long) * AUG_XCH_TOTAL_WORDS));
    // an export opens a SESSION: every rank exports before any rank attaches and the device epoch restarts at 0 in
    // p2p_finish_attach, so flags a previous session left behind (epoch 1, 2, ...) must not survive into this one
    AUG_CUDA(cudaStreamSynchronize(c->stream));
    AUG_CUDA(cudaMemset(c->mailbox, 0, sizeof(unsigned long long) * AUG_XCH_TOTAL_WORDS));
    if (handle) {
        cudaIpcMemHandle_t h;
        AUG_CUDA(cudaIpcGetMemHandle(&h, c->mailbox));
        memcpy(handle, &h, 64);
    }
    if (local_ptr) *local_ptr = c->mailbox;
    return AUG_OK;
}

static void p2p_close(aug_ctx* c) {
    aug_xch_flush(c);
    c->deferred = 0;
    cudaStreamSynchronize(c->stream);   // no exchanging kernel may still be polling / pushing when the maps go away
    for (int r = 0; r < AUG_MAX_RANKS; ++r) {
        if (c->peer_ipc[r] && c->peer_box[r]) cudaIpcCloseMemHandle(c->peer_box[r]);
        c->peer_box[r] = nullptr;
        c->peer_ipc[r] = false;
    }
    if (c->xch) cudaFree(c->xch);
    c->xch = nullptr;
    c->xch_ranks = 0;
    c->fused = 0;
}

static int32_t p2p_finish_attach(aug_ctx* c, int32_t nranks, int32_t rank) {
    AugXchDev h;
    memset(&h, 0, sizeof(h));
    for (int r = 0; r < nranks; ++r) h.box[r] = c->peer_box[r];
    h.nranks = nranks;
    h.rank = rank;
    h.epoch = 0;
    h.err = c->dflag;
    h.timeout_ns = 5000000000ull;      // 5 s: a rank that never launches the matching verb raises flag bit 1
    const char* e = getenv("AUGCUDA_XCH_TIMEOUT_MS");
    if (e && atof(e) > 0) h.timeout_ns = (unsigned long long)(atof(e) * 1e6);
    cudaError_t ce = cudaMalloc(&c->xch, sizeof(AugXchDev));
    if (ce == cudaSuccess) ce = cudaMemcpy(c->xch, &h, sizeof(h), cudaMemcpyHostToDevice);
    if (ce != cudaSuccess) { p2p_close(c); return (int32_t)ce; }
    c->xch_ranks = nranks;
    c->xch_rank = rank;
    c->pending = 0;
    c->deferred = 0;
    return AUG_OK;
}

// one process per GPU: handles = nranks x 64 bytes (cudaIpcMemHandle_t of every rank's mailbox, own slot ignored).
// Every rank must have exported before any rank attaches, and all ranks must attach before the first fused verb
// (the caller's rendezvous — torch.distributed all_gather + barrier — provides both).
int32_t aug_comm_p2p_attach(aug_ctx* c, int32_t nranks, int32_t rank, const char* handles) {
    if (!c || !c->mailbox) return AUG_ERR_NOT_INIT;
    if (nranks < 1 || nranks > AUG_MAX_RANKS || rank < 0 || rank >= nranks || !handles) return AUG_ERR_BAD_ARG;
    AUG_CUDA(cudaSetDevice(c->device));
    p2p_close(c);
    for (int r = 0; r < nranks; ++r) {
        if (r == rank) { c->peer_box[r] = c->mailbox; continue; }
        cudaIpcMemHandle_t h;
        memcpy(&h, handles + 64 * (size_t)r, 64);
        void* p = nullptr;
        cudaError_t ce = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
        if (ce != cudaSuccess) { p2p_close(c); return (int32_t)ce; }
        c->peer_box[r] = (unsigned long long*)p;
        c->peer_ipc[r] = true;
    }
    return p2p_finish_attach(c, nranks, rank);
}

// several ctxs inside ONE process: ptrs[r] = the local_ptr aug_comm_p2p_export returned for rank r, devices[r] its device
int32_t aug_comm_p2p_attach_ptrs(aug_ctx* c, int32_t nranks, int32_t rank, void* const* ptrs, const int32_t* devices) {
    if (!c || !c->mailbox) return AUG_ERR_NOT_INIT;
    if (nranks < 1 || nranks > AUG_MAX_RANKS || rank < 0 || rank >= nranks || !ptrs || !devices) return AUG_ERR_BAD_ARG;
    AUG_CUDA(cudaSetDevice(c->device));
    p2p_close(c);
    for (int r = 0; r < nranks; ++r) {
        if (r != rank && devices[r] != c->device) {
            int can = 0;
            AUG_CUDA(cudaDeviceCanAccessPeer(&can, c->device, devices[r]));
            if (!can) return AUG_ERR_PRECONDITION;
            cudaError_t ce = cudaDeviceEnablePeerAccess(devices[r], 0);
            if (ce != cudaSuccess && ce != cudaErrorPeerAccessAlreadyEnabled) return (int32_t)ce;
            (void)cudaGetLastError();
        }
        c->peer_box[r] = (unsigned long long*)ptrs[r];
    }
    return p2p_finish_attach(c, nranks, rank);
}

int32_t aug_comm_p2p_detach(aug_ctx* c) {
    if (!c) return AUG_ERR_NOT_INIT;
    AUG_CUDA(cudaSetDevice(c->device));
    AUG_CUDA(cudaStreamSynchronize(c->stream));
    p2p_close(c);
    return AUG_OK;
}

// on != 0: every scalar-producing verb (aug_cavi_step, aug_expected_elbo_terms, aug_sampled_loglik_terms) returns
// the sums over ALL ranks, exchanged inside its own reducing kernel.  Collective semantics: every rank must call
// the same sequence of such verbs (an empty shard, n = 0, still takes part).
int32_t aug_comm_set_fused(aug_ctx* c, int32_t on) {
    if (!c) return AUG_ERR_NOT_INIT;
    if (on && !c->xch) return AUG_ERR_NOT_INIT;
    if (!on) {
        AUG_CUDA(cudaSetDevice(c->device));
        int32_t rc = aug_xch_flush(c);
        if (rc) return rc;
    }
    c->fused = on ? 1 : 0;
    return AUG_OK;
}
// Split-phase exchange: with on != 0 (fused mode only) aug_cavi_step / aug_expected_elbo_terms only PUBLISH their sums
// from the reducing kernel; the gather runs in an extra CTA of the next aug_aux_sample launch on the ctx (scheduled
// when that kernel drains, by when the peers have long published), or in a 1-thread kernel at the next
// aug_comm_flush / aug_ctx_sync / other scalar-producing verb.  Until then the block holds this rank's LOCAL sums.
// The per-step barrier between the ranks disappears from the reducing kernel: a rank may run up to one sampling kernel
// ahead of its slowest peer (two epochs of mailbox slots are exactly enough for that).
int32_t aug_comm_set_deferred(aug_ctx* c, int32_t on) {
    if (!c) return AUG_ERR_NOT_INIT;
    if (on && !c->xch) return AUG_ERR_NOT_INIT;
    AUG_CUDA(cudaSetDevice(c->device));
    int32_t rc = aug_xch_flush(c);
    if (rc) return rc;
    c->deferred = on ? 1 : 0;
    return AUG_OK;
}
int32_t aug_comm_flush(aug_ctx* c) {
    if (!c) return AUG_ERR_NOT_INIT;
    AUG_CUDA(cudaSetDevice(c->device));
    return aug_xch_flush(c);
}
int32_t aug_comm_get_fused(aug_ctx* c, int32_t* on) {
    if (!c || !on) return AUG_ERR_NOT_INIT;
    *on = c->fused;
    return AUG_OK;
}


// in-place sum over ranks of count <= 7 device doubles through the mailbox, on the ctx stream (no NCCL involved)
int32_t aug_allreduce_scalars_p2p(aug_ctx* c, double* dev, int32_t count) {
    if (!c || !c->xch) return AUG_ERR_NOT_INIT;
    if (!dev || count < 1 || count > 7) return AUG_ERR_BAD_ARG;
    AUG_CUDA(cudaSetDevice(c->device));
    { int32_t rc = aug_xch_flush(c); if (rc) return rc; }
    xch_only_kernel<<<1, 32, 0, c->stream>>>(c->xch, dev, count, 0, 0);
    c->launches++;
    return (int32_t)cudaGetLastError();
}

int32_t aug_allreduce_scalars(aug_ctx* c, double* dev, int32_t count) {
    if (!c || !c->nccl_comm) return AUG_ERR_NOT_INIT;
    if (!dev || count < 1) return AUG_ERR_BAD_ARG;
    AUG_CUDA(cudaSetDevice(c->device));
    int r = g_nccl.allreduce(dev, dev, (size_t)count, NCCL_FLOAT64, NCCL_SUM, c->nccl_comm, c->stream);
    if (r) return 1000 + r;
    return AUG_OK;
}

}  // extern "C"

int32_t aug_xch_zero_contribution(aug_ctx* c, double* scalars, int first_slot, int nslots) {
    { int32_t rc = aug_xch_flush(c); if (rc) return rc; }
    xch_only_kernel<<<1, 32, 0, c->stream>>>(c->xch, scalars, nslots, first_slot, 1);
    c->launches++;
    return (int32_t)cudaGetLastError();
}
