// aug_host.cu — host-buffer ("plugin") entry points: every verb of the path with HOST pointers.
//
// These are what a method of the reference's generic functions specialised on plain `Vector` arguments binds to
// (src/generic.jl:1-88 dispatches on host vectors for every verb).  The observation axis is cut into chunks that
// are staged through three device slots; H2D copies, the kernel and D2H copies of consecutive chunks overlap on
// three streams (copy-in, the ctx stream, copy-out) ordered by events.  Per-chunk scalar blocks land in pinned
// host memory and are added in chunk order, so the result does not depend on timing.  The arithmetic is done by
// the same device kernels as the device-pointer verbs: arrays are bit-identical to those, sampled values are
// bit-identical for the same (seed, offset, i0) because the RNG is keyed by the global element index.
#include <stdlib.h>
#include <string.h>

#include <functional>

#include "aug_common.cuh"

int32_t aug_cavi_dispatch(aug_ctx* ctx, const aug_lik* lik, int64_t n, const void* y, const double* mu,
                          const double* var, int64_t ld, void* s0, void* s1, void* s2, const void* rs0,
                          const void* rs1, const void* rs2, double* beta, double* gamma, int64_t ldo,
                          double* scalars, bool from_state);
int32_t aug_cat_dispatch(aug_ctx* ctx, const aug_lik* lik, int64_t n, const void* y, const double* mu,
                         const double* var, void* s0, void* s1, void* s2, const void* rs0, const void* rs1,
                         const void* rs2, double* beta, double* gamma, int64_t ldo, double* scalars,
                         bool from_state);
int32_t aug_aux_sample_dev(aug_ctx* c, const aug_lik* lik, int64_t n, int64_t i0, const void* y, const double* f,
                           int64_t ld, double* omega, int64_t* nvar, uint64_t offset);
int32_t aug_init_aux_variables_dev(aug_ctx* c, const aug_lik* lik, int64_t n, int64_t i0, double* omega,
                                   int64_t* nvar, uint64_t offset);

#define PIPE_SLOTS 3

struct aug_pipe {
    cudaStream_t s_in, s_out;
    cudaEvent_t ev_in[PIPE_SLOTS], ev_k[PIPE_SLOTS], ev_out[PIPE_SLOTS];
    unsigned char* dbuf[PIPE_SLOTS];
    size_t dbuf_bytes;
    double* hscal;       // pinned [max_chunks][AUG_NSCALARS]
    size_t hscal_chunks;
    double* dscal;       // device [PIPE_SLOTS][AUG_NSCALARS]
};

void aug_pipe_destroy(aug_ctx* ctx) {
    aug_pipe* p = ctx->pipe;
    if (!p) return;
    for (int s = 0; s < PIPE_SLOTS; ++s) {
        if (p->dbuf[s]) cudaFree(p->dbuf[s]);
        cudaEventDestroy(p->ev_in[s]);
        cudaEventDestroy(p->ev_k[s]);
        cudaEventDestroy(p->ev_out[s]);
    }
    if (p->hscal) cudaFreeHost(p->hscal);
    if (p->dscal) cudaFree(p->dscal);
    cudaStreamDestroy(p->s_in);
    cudaStreamDestroy(p->s_out);
    delete p;
    ctx->pipe = nullptr;
}

namespace {

int32_t pipe_get(aug_ctx* ctx, size_t slot_bytes, size_t chunks) {
    if (!ctx->pipe) {
        aug_pipe* p = new aug_pipe();
        memset(p, 0, sizeof(*p));
        AUG_CUDA(cudaStreamCreateWithFlags(&p->s_in, cudaStreamNonBlocking));
        AUG_CUDA(cudaStreamCreateWithFlags(&p->s_out, cudaStreamNonBlocking));
        for (int s = 0; s < PIPE_SLOTS; ++s) {
            AUG_CUDA(cudaEventCreateWithFlags(&p->ev_in[s], cudaEventDisableTiming));
            AUG_CUDA(cudaEventCreateWithFlags(&p->ev_k[s], cudaEventDisableTiming));
            AUG_CUDA(cudaEventCreateWithFlags(&p->ev_out[s], cudaEventDisableTiming));
        }
        AUG_CUDA(cudaMalloc(&p->dscal, sizeof(double) * PIPE_SLOTS * AUG_NSCALARS));
        ctx->pipe = p;
    }
    aug_pipe* p = ctx->pipe;
    if (p->dbuf_bytes < slot_bytes) {
        // a previous call may have been abandoned on an error with copies still in flight
        AUG_CUDA(cudaStreamSynchronize(p->s_in));
        AUG_CUDA(cudaStreamSynchronize(ctx->stream));
        AUG_CUDA(cudaStreamSynchronize(p->s_out));
        for (int s = 0; s < PIPE_SLOTS; ++s) {
            if (p->dbuf[s]) cudaFree(p->dbuf[s]);
            p->dbuf[s] = nullptr;
        }
        p->dbuf_bytes = 0;
        for (int s = 0; s < PIPE_SLOTS; ++s) AUG_CUDA(cudaMalloc(&p->dbuf[s], slot_bytes));
        p->dbuf_bytes = slot_bytes;
    }
    if (p->hscal_chunks < chunks) {
        if (p->hscal) cudaFreeHost(p->hscal);
        p->hscal = nullptr;
        p->hscal_chunks = 0;
        AUG_CUDA(cudaHostAlloc(&p->hscal, sizeof(double) * AUG_NSCALARS * chunks, cudaHostAllocDefault));
        p->hscal_chunks = chunks;
    }
    return AUG_OK;
}

inline size_t up256(size_t x) { return (x + 255) & ~(size_t)255; }
// elements per staged chunk (AUGCUDA_HOST_CHUNK_LOG2 = 10..26 overrides the default 2^22: tuning runs, and the
// chunk-boundary parity tests, which use small chunks to cross many boundaries at small n)
inline int64_t host_chunk_elems() {
    const char* e = getenv("AUGCUDA_HOST_CHUNK_LOG2");
    const int v = e ? atoi(e) : 22;
    return (int64_t)1 << ((v >= 10 && v <= 26) ? v : 22);
}
inline bool is_cat(int k) { return k == AUG_CAT || k == AUG_CAT_BIJ; }
inline size_t y_size(int kind) { return (kind == AUG_BERNOULLI || is_cat(kind)) ? 1 : 8; }

// One array of a host verb.  Host layout: `planes` planes of n*per elements, `ld` elements apart (latent-major
// HETERO inputs and all beta / gamma outputs), or a single obs-major plane (per = nlatent for the Categorical
// [n][nl] arrays).  Device layout inside a slot: planes `chunk*per` elements apart.
struct Arr {
    const void* in = nullptr;   // host source: copied H2D before the kernel
    void* out = nullptr;        // host destination: copied D2H after the kernel
    size_t esz = 8;
    int64_t per = 1;
    int planes = 1;
    int copy_plane0 = 0, copy_planes = -1;   // sub-range of planes that is actually copied (default: all)
    int64_t ld = 0;
    bool scratch = false;       // device space without a host side (a kernel needs the buffer)
    size_t off = 0;
    bool used() const { return in || out || scratch; }
    unsigned char* dev(unsigned char* base) const { return used() ? base + off : nullptr; }
};

struct FusedOff {   // per-chunk scalars are summed on the host: the in-kernel peer exchange (fused multi-GPU mode) stays off
    aug_ctx* c;
    int was;
    explicit FusedOff(aug_ctx* cc) : c(cc), was(cc->fused) { cc->fused = 0; }
    ~FusedOff() { c->fused = was; }
};

typedef std::function<int32_t(int64_t rows, int64_t r0, int64_t chunk, unsigned char* d, double* dscal)> Launch;

// The chunk pipeline shared by all host verbs.  `per` = elements per observation of the widest obs-major array
// (sizes the chunk); scal_host != nullptr: per-chunk device scalar blocks are brought back and added in chunk order.
int32_t run_pipeline(aug_ctx* c, int64_t n, int64_t per, Arr* arrs, int narr, double* scal_host, const Launch& launch) {
    if (scal_host) memset(scal_host, 0, sizeof(double) * AUG_NSCALARS);
    if (n == 0) return AUG_OK;
    int64_t chunk = host_chunk_elems() / per;
    if (chunk < 2) chunk = 2;
    chunk &= ~(int64_t)1;          // even: keeps every plane of a slot 16-byte aligned for the 128-bit kernels
    if (chunk > n) chunk = n;
    const int64_t nchunks = (n + chunk - 1) / chunk;
    size_t off = 0;
    for (int i = 0; i < narr; ++i) {
        Arr& a = arrs[i];
        if (!a.used()) continue;
        if (a.copy_planes < 0) a.copy_planes = a.planes;
        a.off = off;
        off += up256((size_t)chunk * a.per * a.esz * a.planes);
    }
    int32_t rc = pipe_get(c, off, scal_host ? (size_t)nchunks : 1);
    if (rc) return rc;
    aug_pipe* p = c->pipe;
    for (int64_t k = 0; k < nchunks; ++k) {
        const int s = (int)(k % PIPE_SLOTS);
        const int64_t r0 = k * chunk;
        const int64_t rows = (n - r0 < chunk) ? n - r0 : chunk;
        unsigned char* d = p->dbuf[s];
        if (k >= PIPE_SLOTS) AUG_CUDA(cudaStreamWaitEvent(p->s_in, p->ev_out[s], 0));
        for (int i = 0; i < narr; ++i) {
            const Arr& a = arrs[i];
            if (!a.in) continue;
            const size_t w = (size_t)rows * a.per * a.esz;
            const unsigned char* src = (const unsigned char*)a.in + (size_t)r0 * a.per * a.esz;
            if (a.planes == 1) {
                AUG_CUDA(cudaMemcpyAsync(d + a.off, src, w, cudaMemcpyHostToDevice, p->s_in));
            } else {
                const size_t dp = (size_t)chunk * a.per * a.esz, hp = (size_t)a.ld * a.esz;
                AUG_CUDA(cudaMemcpy2DAsync(d + a.off + dp * a.copy_plane0, dp, src + hp * a.copy_plane0, hp, w,
                                           a.copy_planes, cudaMemcpyHostToDevice, p->s_in));
            }
        }
        AUG_CUDA(cudaEventRecord(p->ev_in[s], p->s_in));
        AUG_CUDA(cudaStreamWaitEvent(c->stream, p->ev_in[s], 0));
        if (k >= PIPE_SLOTS) AUG_CUDA(cudaStreamWaitEvent(c->stream, p->ev_out[s], 0));
        double* dsc = scal_host ? p->dscal + (size_t)s * AUG_NSCALARS : nullptr;
        rc = launch(rows, r0, chunk, d, dsc);
        if (rc) return rc;
        AUG_CUDA(cudaEventRecord(p->ev_k[s], c->stream));
        AUG_CUDA(cudaStreamWaitEvent(p->s_out, p->ev_k[s], 0));
        for (int i = 0; i < narr; ++i) {
            const Arr& a = arrs[i];
            if (!a.out) continue;
            const size_t w = (size_t)rows * a.per * a.esz;
            unsigned char* dst = (unsigned char*)a.out + (size_t)r0 * a.per * a.esz;
            if (a.planes == 1) {
                AUG_CUDA(cudaMemcpyAsync(dst, d + a.off, w, cudaMemcpyDeviceToHost, p->s_out));
            } else {
                const size_t dp = (size_t)chunk * a.per * a.esz, hp = (size_t)a.ld * a.esz;
                AUG_CUDA(cudaMemcpy2DAsync(dst + hp * a.copy_plane0, hp, d + a.off + dp * a.copy_plane0, dp, w,
                                           a.copy_planes, cudaMemcpyDeviceToHost, p->s_out));
            }
        }
        if (dsc) AUG_CUDA(cudaMemcpyAsync(p->hscal + (size_t)k * AUG_NSCALARS, dsc, sizeof(double) * AUG_NSCALARS,
                                          cudaMemcpyDeviceToHost, p->s_out));
        AUG_CUDA(cudaEventRecord(p->ev_out[s], p->s_out));
    }
    AUG_CUDA(cudaStreamSynchronize(p->s_out));
    AUG_CUDA(cudaStreamSynchronize(c->stream));
    if (scal_host) {
        for (int64_t k = 0; k < nchunks; ++k) {
            const double* h = p->hscal + (size_t)k * AUG_NSCALARS;
            for (int j = 0; j < AUG_NSCALARS; ++j) scal_host[j] += h[j];
        }
        // the two derived slots are re-formed from the totals (generic.jl:52-54 "+", :48-50)
        scal_host[AUG_S_EXPECTED_AUGLL] = scal_host[AUG_S_EXPECTED_LOGTILT] + scal_host[AUG_S_KL];
        scal_host[AUG_S_AUGLL] = scal_host[AUG_S_LOGTILT] + scal_host[AUG_S_LOGPRIOR];
    }
    return AUG_OK;
}

struct Shape {
    int kind;
    bool cat, het;
    int64_t per;        // elements per observation of the obs-major arrays (nl for the Categorical likelihood)
    int nlat_in;        // latent-major planes of mu / var / f (2 for HETERO)
    int nlat_out;       // planes of beta / gamma
    size_t ysz;
    bool has_s1;
    size_t s2sz;        // 0: the kind has no third state array
    bool needs_n;       // the sample carries an integer field n
};

int32_t shape_of(const aug_lik* lik, Shape* sh) {
    if (!lik) return AUG_ERR_BAD_ARG;
    if (lik->kind < 0 || lik->kind >= AUG_NKINDS) return AUG_ERR_BAD_KIND;
    const int kind = lik->kind;
    sh->kind = kind;
    sh->cat = is_cat(kind);
    sh->het = kind == AUG_HETERO;
    if (sh->cat && lik->nlatent < 1) return AUG_ERR_BAD_ARG;
    sh->per = sh->cat ? lik->nlatent : 1;
    sh->nlat_in = sh->het ? 2 : 1;
    sh->nlat_out = sh->cat ? lik->nlatent : (sh->het ? 2 : 1);
    sh->ysz = y_size(kind);
    sh->has_s1 = kind == AUG_POISSON || sh->het || sh->cat;
    sh->s2sz = (kind == AUG_NEGBIN || kind == AUG_POISSON || sh->het) ? 8 : (sh->cat ? 1 : 0);
    sh->needs_n = kind == AUG_POISSON || sh->het || sh->cat;
    return AUG_OK;
}

// arrays of the verbs, in a fixed order
enum { A_Y, A_MU, A_VAR, A_S0, A_S1, A_S2, A_B, A_G, A_W, A_N, A_COUNT };

void set_latent_in(Arr& a, const Shape& sh, const void* host, int64_t ld) {   // mu / var / f
    a.in = host;
    a.per = sh.per;
    a.planes = sh.nlat_in;
    a.ld = ld;
}
void set_bg(Arr& a, const Shape& sh, void* host, int64_t n, int64_t ldo) {     // beta / gamma (latent-major)
    a.out = host;
    a.per = 1;
    a.planes = sh.nlat_out;
    a.ld = sh.nlat_out > 1 ? ldo : n;
}

// CAVI-side verbs share one body: `from_state` selects aug_expected_* (state is an input) or aug_cavi_step /
// aug_aux_posterior (state is an output)
int32_t cavi_host(aug_ctx* c, const aug_lik* lik, int64_t n, const void* y, const double* mu, const double* var,
                  int64_t ld, void* s0, void* s1, void* s2, double* beta, double* gamma, int64_t ldo,
                  double* scalars_host, bool from_state) {
    if (!c) return AUG_ERR_NOT_INIT;
    Shape sh;
    int32_t rc = shape_of(lik, &sh);
    if (rc) return rc;
    if (n < 0 || !y) return AUG_ERR_BAD_ARG;
    const bool need_moments = !from_state || scalars_host != nullptr;
    if (need_moments && (!mu || !var)) return AUG_ERR_BAD_ARG;
    if (from_state && (!s0 || (sh.has_s1 && !sh.cat && !s1) || (sh.het && !s2))) return AUG_ERR_BAD_ARG;
    if (from_state && sh.cat && !s1) return AUG_ERR_BAD_ARG;
    if (sh.het && (!mu || ld < n)) return AUG_ERR_BAD_ARG;
    if ((sh.het || sh.cat) && (beta || gamma) && ldo < n) return AUG_ERR_BAD_ARG;
    if (scalars_host && sh.kind == AUG_CAT) return AUG_ERR_PRECONDITION;   // categorical.jl:165-170
    AUG_CUDA(cudaSetDevice(c->device));
    FusedOff fused_off(c);
    Arr A[A_COUNT];
    A[A_Y].in = y; A[A_Y].esz = sh.ysz; A[A_Y].per = sh.per;
    if (mu && (need_moments || sh.het)) {
        set_latent_in(A[A_MU], sh, mu, ld);
        if (!need_moments) { A[A_MU].copy_plane0 = 1; A[A_MU].copy_planes = 1; }   // only E[g] is read (hetero :68-104)
    }
    if (var && need_moments) set_latent_in(A[A_VAR], sh, var, ld);
    A[A_S0].per = A[A_S1].per = A[A_S2].per = sh.per;
    A[A_S2].esz = sh.s2sz ? sh.s2sz : 8;
    if (from_state) {
        A[A_S0].in = s0;
        if (sh.has_s1) A[A_S1].in = s1;
        if (sh.s2sz) A[A_S2].in = s2;          // NEGBIN / POISSON / CAT: optional y copy (NULL: the verbs read y)
    } else {
        A[A_S0].out = s0;
        if (sh.has_s1) A[A_S1].out = s1;
        if (sh.s2sz) A[A_S2].out = s2;
        if (sh.het && !s2) A[A_S2].scratch = true;   // the kernel always materialises ψ for HETERO
    }
    if (beta) set_bg(A[A_B], sh, beta, n, ldo);
    if (gamma) set_bg(A[A_G], sh, gamma, n, ldo);
    Launch launch = [&](int64_t rows, int64_t, int64_t chunk, unsigned char* d, double* dsc) -> int32_t {
        const void* dy = A[A_Y].dev(d);
        const double* dmu = (const double*)A[A_MU].dev(d);
        const double* dvar = (const double*)A[A_VAR].dev(d);
        void *w0 = nullptr, *w1 = nullptr, *w2 = nullptr;
        const void *r0 = nullptr, *r1 = nullptr, *r2 = nullptr;
        if (from_state) { r0 = A[A_S0].dev(d); r1 = A[A_S1].dev(d); r2 = A[A_S2].dev(d); }
        else { w0 = A[A_S0].dev(d); w1 = A[A_S1].dev(d); w2 = A[A_S2].dev(d); }
        double* db = (double*)A[A_B].dev(d);
        double* dg = (double*)A[A_G].dev(d);
        if (sh.cat)
            return aug_cat_dispatch(c, lik, rows, dy, dmu, dvar, w0, w1, w2, r0, r1, r2, db, dg, chunk, dsc, from_state);
        return aug_cavi_dispatch(c, lik, rows, dy, dmu, dvar, chunk, w0, w1, w2, r0, r1, r2, db, dg, chunk, dsc,
                                 from_state);
    };
    return run_pipeline(c, n, sh.per, A, A_COUNT, scalars_host, launch);
}

}  // namespace

extern "C" {

int32_t aug_potential_precision(aug_ctx* c, const aug_lik* lik, int64_t n, const void* y, const double* f,
                                int64_t ld, const double* omega, const int64_t* nvar, double* beta,
                                double* gamma, int64_t ldo);
int32_t aug_sampled_loglik_terms(aug_ctx* c, const aug_lik* lik, int64_t n, const void* y, const double* f,
                                 int64_t ld, const double* omega, const int64_t* nvar, int32_t with_prior,
                                 double* scalars);

// init_aux_posterior(T, lik, n) into host arrays: a zero fill (bernoulli.jl:7-11 ... categorical.jl:59-70) — there
// is no arithmetic to put on the device
int32_t aug_init_aux_posterior_host(aug_ctx* c, const aug_lik* lik, int64_t n, void* s0, void* s1, void* s2) {
    if (!c) return AUG_ERR_NOT_INIT;
    Shape sh;
    int32_t rc = shape_of(lik, &sh);
    if (rc) return rc;
    if (n < 0) return AUG_ERR_BAD_ARG;
    const size_t m = (size_t)n * sh.per;
    if (s0) memset(s0, 0, m * 8);
    if (s1 && sh.has_s1) memset(s1, 0, m * 8);
    if (s2 && sh.s2sz) memset(s2, 0, m * sh.s2sz);
    return AUG_OK;
}

int32_t aug_cavi_step_host(aug_ctx* c, const aug_lik* lik, int64_t n, const void* y, const double* mu,
                           const double* var, int64_t ld, void* s0, void* s1, void* s2, double* beta,
                           double* gamma, int64_t ldo, double* scalars_host) {
    return cavi_host(c, lik, n, y, mu, var, ld, s0, s1, s2, beta, gamma, ldo, scalars_host, false);
}

int32_t aug_aux_posterior_host(aug_ctx* c, const aug_lik* lik, int64_t n, const void* y, const double* mu,
                               const double* var, int64_t ld, void* s0, void* s1, void* s2) {
    return cavi_host(c, lik, n, y, mu, var, ld, s0, s1, s2, nullptr, nullptr, 0, nullptr, false);
}

int32_t aug_expected_potential_precision_host(aug_ctx* c, const aug_lik* lik, int64_t n, const void* y,
                                              const double* mu, int64_t ld, const void* s0, const void* s1,
                                              const void* s2, double* beta, double* gamma, int64_t ldo) {
    return cavi_host(c, lik, n, y, mu, nullptr, ld, (void*)s0, (void*)s1, (void*)s2, beta, gamma, ldo, nullptr, true);
}

int32_t aug_expected_elbo_terms_host(aug_ctx* c, const aug_lik* lik, int64_t n, const void* y, const double* mu,
                                     const double* var, int64_t ld, const void* s0, const void* s1,
                                     const void* s2, double* scalars_host) {
    if (!scalars_host) return AUG_ERR_BAD_ARG;
    return cavi_host(c, lik, n, y, mu, var, ld, (void*)s0, (void*)s1, (void*)s2, nullptr, nullptr, 0, scalars_host,
                     true);
}

int32_t aug_aux_sample_host(aug_ctx* c, const aug_lik* lik, int64_t n, int64_t i0, const void* y, const double* f,
                            int64_t ld, double* omega, int64_t* nvar) {
    if (!c) return AUG_ERR_NOT_INIT;
    Shape sh;
    int32_t rc = shape_of(lik, &sh);
    if (rc) return rc;
    if (n < 0 || !f || !omega) return AUG_ERR_BAD_ARG;
    const bool needs_y = sh.kind != AUG_BERNOULLI;
    if ((needs_y && !y) || (sh.needs_n && !nvar) || (sh.het && ld < n)) return AUG_ERR_BAD_ARG;
    AUG_CUDA(cudaSetDevice(c->device));
    const uint64_t offset = c->offset++;   // one RNG tick for the whole call, whatever the chunking
    Arr A[A_COUNT];
    if (needs_y) { A[A_Y].in = y; A[A_Y].esz = sh.ysz; A[A_Y].per = sh.per; }
    set_latent_in(A[A_MU], sh, f, ld);
    A[A_W].out = omega; A[A_W].per = sh.per;
    if (sh.needs_n) { A[A_N].out = nvar; A[A_N].per = sh.per; }
    Launch launch = [&](int64_t rows, int64_t r0, int64_t chunk, unsigned char* d, double*) -> int32_t {
        return aug_aux_sample_dev(c, lik, rows, i0 + r0, A[A_Y].dev(d), (const double*)A[A_MU].dev(d), chunk,
                                  (double*)A[A_W].dev(d), (int64_t*)A[A_N].dev(d), offset);
    };
    return run_pipeline(c, n, sh.per, A, A_COUNT, nullptr, launch);
}

int32_t aug_init_aux_variables_host(aug_ctx* c, const aug_lik* lik, int64_t n, int64_t i0, double* omega,
                                    int64_t* nvar) {
    if (!c) return AUG_ERR_NOT_INIT;
    Shape sh;
    int32_t rc = shape_of(lik, &sh);
    if (rc) return rc;
    if (n < 0 || !omega || (sh.needs_n && !nvar)) return AUG_ERR_BAD_ARG;
    AUG_CUDA(cudaSetDevice(c->device));
    const uint64_t offset = c->offset++;
    Arr A[A_COUNT];
    A[A_W].out = omega; A[A_W].per = sh.per;
    if (sh.needs_n) { A[A_N].out = nvar; A[A_N].per = sh.per; }
    Launch launch = [&](int64_t rows, int64_t r0, int64_t, unsigned char* d, double*) -> int32_t {
        return aug_init_aux_variables_dev(c, lik, rows, i0 + r0, (double*)A[A_W].dev(d), (int64_t*)A[A_N].dev(d),
                                          offset);
    };
    return run_pipeline(c, n, sh.per, A, A_COUNT, nullptr, launch);
}

int32_t aug_potential_precision_host(aug_ctx* c, const aug_lik* lik, int64_t n, const void* y, const double* f,
                                     int64_t ld, const double* omega, const int64_t* nvar, double* beta,
                                     double* gamma, int64_t ldo) {
    if (!c) return AUG_ERR_NOT_INIT;
    Shape sh;
    int32_t rc = shape_of(lik, &sh);
    if (rc) return rc;
    if (n < 0 || !y || !omega || (sh.needs_n && !nvar)) return AUG_ERR_BAD_ARG;
    if (sh.het && (!f || ld < n)) return AUG_ERR_BAD_ARG;
    if ((sh.het || sh.cat) && (beta || gamma) && ldo < n) return AUG_ERR_BAD_ARG;
    AUG_CUDA(cudaSetDevice(c->device));
    Arr A[A_COUNT];
    A[A_Y].in = y; A[A_Y].esz = sh.ysz; A[A_Y].per = sh.per;
    if (sh.het) {                       // only g is read: inv(invlink(g)) (heteroscedasticgaussian.jl:48-66)
        set_latent_in(A[A_MU], sh, f, ld);
        A[A_MU].copy_plane0 = 1;
        A[A_MU].copy_planes = 1;
    }
    A[A_W].in = omega; A[A_W].per = sh.per;
    if (sh.needs_n) { A[A_N].in = nvar; A[A_N].per = sh.per; }
    if (beta) set_bg(A[A_B], sh, beta, n, ldo);
    if (gamma) set_bg(A[A_G], sh, gamma, n, ldo);
    Launch launch = [&](int64_t rows, int64_t, int64_t chunk, unsigned char* d, double*) -> int32_t {
        return aug_potential_precision(c, lik, rows, A[A_Y].dev(d), (const double*)A[A_MU].dev(d), chunk,
                                       (const double*)A[A_W].dev(d), (const int64_t*)A[A_N].dev(d),
                                       (double*)A[A_B].dev(d), (double*)A[A_G].dev(d), chunk);
    };
    return run_pipeline(c, n, sh.per, A, A_COUNT, nullptr, launch);
}

int32_t aug_sampled_loglik_terms_host(aug_ctx* c, const aug_lik* lik, int64_t n, const void* y, const double* f,
                                      int64_t ld, const double* omega, const int64_t* nvar, int32_t with_prior,
                                      double* scalars_host) {
    if (!c) return AUG_ERR_NOT_INIT;
    Shape sh;
    int32_t rc = shape_of(lik, &sh);
    if (rc) return rc;
    if (n < 0 || !y || !f || !omega || !scalars_host || (sh.needs_n && !nvar)) return AUG_ERR_BAD_ARG;
    if (sh.het && ld < n) return AUG_ERR_BAD_ARG;
    if (with_prior && sh.kind == AUG_CAT) return AUG_ERR_PRECONDITION;   // categorical.jl:158-163
    AUG_CUDA(cudaSetDevice(c->device));
    FusedOff fused_off(c);
    Arr A[A_COUNT];
    A[A_Y].in = y; A[A_Y].esz = sh.ysz; A[A_Y].per = sh.per;
    set_latent_in(A[A_MU], sh, f, ld);
    A[A_W].in = omega; A[A_W].per = sh.per;
    if (sh.needs_n) { A[A_N].in = nvar; A[A_N].per = sh.per; }
    Launch launch = [&](int64_t rows, int64_t, int64_t chunk, unsigned char* d, double* dsc) -> int32_t {
        return aug_sampled_loglik_terms(c, lik, rows, A[A_Y].dev(d), (const double*)A[A_MU].dev(d), chunk,
                                        (const double*)A[A_W].dev(d), (const int64_t*)A[A_N].dev(d), with_prior, dsc);
    };
    return run_pipeline(c, n, sh.per, A, A_COUNT, scalars_host, launch);
}

}  // extern "C"
