// aug_host.cu — host-buffer ("plugin") entry points: the same verbs with HOST pointers.
//
// The observation axis is cut into chunks that are staged through three device slots; H2D copies,
// the kernel and D2H copies of consecutive chunks overlap on three streams (copy-in, the ctx
// stream, copy-out) ordered by events.  Per-chunk scalars land in pinned host memory and are
// added in chunk order, so the result does not depend on timing.
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "aug_common.cuh"

int32_t aug_cavi_dispatch(aug_ctx* ctx, const aug_lik* lik, int64_t n, const void* y, const double* mu,
                          const double* var, int64_t ld, void* s0, void* s1, void* s2, const void* rs0,
                          const void* rs1, const void* rs2, double* beta, double* gamma, int64_t ldo,
                          double* scalars, bool from_state);
int32_t aug_cat_dispatch(aug_ctx* ctx, const aug_lik* lik, int64_t n, const void* y, const double* mu,
                         const double* var, void* s0, void* s1, void* s2, const void* rs0, const void* rs1,
                         const void* rs2, double* beta, double* gamma, int64_t ldo, double* scalars,
                         bool from_state);

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
        for (int s = 0; s < PIPE_SLOTS; ++s) {
            if (p->dbuf[s]) cudaFree(p->dbuf[s]);
            p->dbuf[s] = nullptr;
            AUG_CUDA(cudaMalloc(&p->dbuf[s], slot_bytes));
        }
        p->dbuf_bytes = slot_bytes;
    }
    if (p->hscal_chunks < chunks) {
        if (p->hscal) cudaFreeHost(p->hscal);
        p->hscal = nullptr;
        AUG_CUDA(cudaHostAlloc(&p->hscal, sizeof(double) * AUG_NSCALARS * chunks, cudaHostAllocDefault));
        p->hscal_chunks = chunks;
    }
    return AUG_OK;
}

inline size_t up256(size_t x) { return (x + 255) & ~(size_t)255; }
// elements per staged chunk (AUGCUDA_HOST_CHUNK_LOG2 = 18..26 overrides the default 2^22 for tuning runs)
inline int64_t host_chunk_elems() {
    static int lg = -1;
    if (lg < 0) {
        const char* e = getenv("AUGCUDA_HOST_CHUNK_LOG2");
        int v = e ? atoi(e) : 22;
        lg = (v >= 18 && v <= 26) ? v : 22;
    }
    return (int64_t)1 << lg;
}
inline bool is_cat(int k) { return k == AUG_CAT || k == AUG_CAT_BIJ; }
inline size_t y_size(int kind) {
    return (kind == AUG_BERNOULLI || is_cat(kind)) ? 1 : 8;
}

}  // namespace

int32_t aug_aux_sample_dev(aug_ctx* c, const aug_lik* lik, int64_t n, int64_t i0, const void* y, const double* f,
                           int64_t ld, double* omega, int64_t* nvar, uint64_t offset);

extern "C" {

int32_t aug_cavi_step_host(aug_ctx* c, const aug_lik* lik, int64_t n, const void* y, const double* mu,
                           const double* var, int64_t ld, void* s0, void* s1, void* s2, double* beta,
                           double* gamma, int64_t ldo, double* scalars_host) {
    if (!c) return AUG_ERR_NOT_INIT;
    if (!lik || n < 0 || !y || !mu || !var) return AUG_ERR_BAD_ARG;
    if (lik->kind < 0 || lik->kind >= AUG_NKINDS) return AUG_ERR_BAD_KIND;
    AUG_CUDA(cudaSetDevice(c->device));
    const int kind = lik->kind;
    const bool cat = is_cat(kind), het = kind == AUG_HETERO;
    if (scalars_host && kind == AUG_CAT) return AUG_ERR_PRECONDITION;
    const int64_t per = cat ? lik->nlatent : 1;       // elements per observation in the obs-major arrays
    const int nlat_in = het ? 2 : 1;                  // latent-major planes of mu / var
    const int nlat_out = cat ? lik->nlatent : (het ? 2 : 1);
    const size_t ysz = y_size(kind);
    const bool has_s1 = kind == AUG_POISSON || het || cat;
    const size_t s2sz = (kind == AUG_NEGBIN || kind == AUG_POISSON || het) ? 8 : (cat ? 1 : 0);
    if (het && (ld < n || ((beta || gamma) && ldo < n))) return AUG_ERR_BAD_ARG;
    if (cat && (beta || gamma) && ldo < n) return AUG_ERR_BAD_ARG;
    if (scalars_host) memset(scalars_host, 0, sizeof(double) * AUG_NSCALARS);
    if (n == 0) return AUG_OK;
    // per-chunk scalars are summed on the host: the in-kernel peer exchange (fused multi-GPU mode) stays off here
    struct FusedOff {
        aug_ctx* c;
        int was;
        explicit FusedOff(aug_ctx* cc) : c(cc), was(cc->fused) { cc->fused = 0; }
        ~FusedOff() { c->fused = was; }
    } fused_off(c);

    int64_t chunk = host_chunk_elems() / per;
    if (chunk < 2) chunk = 2;
    chunk &= ~(int64_t)1;
    if (chunk > n) chunk = n;
    const int64_t nchunks = (n + chunk - 1) / chunk;
    // device slot layout
    size_t off = 0;
    const size_t o_y = off;   off += up256((size_t)chunk * per * ysz);
    const size_t o_mu = off;  off += up256((size_t)chunk * per * 8 * nlat_in);
    const size_t o_var = off; off += up256((size_t)chunk * per * 8 * nlat_in);
    const size_t o_s0 = off;  off += up256((size_t)chunk * per * 8);
    const size_t o_s1 = off;  off += has_s1 ? up256((size_t)chunk * per * 8) : 0;
    const size_t o_s2 = off;  off += s2sz ? up256((size_t)chunk * per * s2sz) : 0;
    const size_t o_b = off;   off += up256((size_t)chunk * 8 * nlat_out);
    const size_t o_g = off;   off += up256((size_t)chunk * 8 * nlat_out);
    int32_t rc = pipe_get(c, off, (size_t)nchunks);
    if (rc) return rc;
    aug_pipe* p = c->pipe;

    for (int64_t k = 0; k < nchunks; ++k) {
        const int s = (int)(k % PIPE_SLOTS);
        const int64_t r0 = k * chunk;
        const int64_t rows = (n - r0 < chunk) ? n - r0 : chunk;
        unsigned char* d = p->dbuf[s];
        if (k >= PIPE_SLOTS) AUG_CUDA(cudaStreamWaitEvent(p->s_in, p->ev_out[s], 0));
        // H2D
        AUG_CUDA(cudaMemcpyAsync(d + o_y, (const unsigned char*)y + (size_t)r0 * per * ysz,
                                 (size_t)rows * per * ysz, cudaMemcpyHostToDevice, p->s_in));
        if (het) {
            AUG_CUDA(cudaMemcpy2DAsync(d + o_mu, (size_t)chunk * 8, mu + r0, (size_t)ld * 8, (size_t)rows * 8, 2,
                                       cudaMemcpyHostToDevice, p->s_in));
            AUG_CUDA(cudaMemcpy2DAsync(d + o_var, (size_t)chunk * 8, var + r0, (size_t)ld * 8, (size_t)rows * 8, 2,
                                       cudaMemcpyHostToDevice, p->s_in));
        } else {
            AUG_CUDA(cudaMemcpyAsync(d + o_mu, mu + r0 * per, (size_t)rows * per * 8, cudaMemcpyHostToDevice,
                                     p->s_in));
            AUG_CUDA(cudaMemcpyAsync(d + o_var, var + r0 * per, (size_t)rows * per * 8, cudaMemcpyHostToDevice,
                                     p->s_in));
        }
        AUG_CUDA(cudaEventRecord(p->ev_in[s], p->s_in));
        // kernel on the ctx stream
        AUG_CUDA(cudaStreamWaitEvent(c->stream, p->ev_in[s], 0));
        if (k >= PIPE_SLOTS) AUG_CUDA(cudaStreamWaitEvent(c->stream, p->ev_out[s], 0));
        double* dsc = scalars_host ? p->dscal + (size_t)s * AUG_NSCALARS : nullptr;
        void* ds0 = s0 ? d + o_s0 : nullptr;
        void* ds1 = (s1 && has_s1) ? d + o_s1 : nullptr;
        void* ds2 = ((s2 && s2sz) || het) ? d + o_s2 : nullptr;
        double* db = beta ? (double*)(d + o_b) : nullptr;
        double* dg = gamma ? (double*)(d + o_g) : nullptr;
        if (cat)
            rc = aug_cat_dispatch(c, lik, rows, d + o_y, (const double*)(d + o_mu), (const double*)(d + o_var), ds0,
                                  ds1, ds2, nullptr, nullptr, nullptr, db, dg, chunk, dsc, false);
        else
            rc = aug_cavi_dispatch(c, lik, rows, d + o_y, (const double*)(d + o_mu), (const double*)(d + o_var),
                                   chunk, ds0, ds1, ds2, nullptr, nullptr, nullptr, db, dg, chunk, dsc, false);
        if (rc) return rc;
        AUG_CUDA(cudaEventRecord(p->ev_k[s], c->stream));
        // D2H
        AUG_CUDA(cudaStreamWaitEvent(p->s_out, p->ev_k[s], 0));
        if (s0) AUG_CUDA(cudaMemcpyAsync((double*)s0 + r0 * per, d + o_s0, (size_t)rows * per * 8,
                                         cudaMemcpyDeviceToHost, p->s_out));
        if (s1 && has_s1) AUG_CUDA(cudaMemcpyAsync((double*)s1 + r0 * per, d + o_s1, (size_t)rows * per * 8,
                                                   cudaMemcpyDeviceToHost, p->s_out));
        if (s2 && s2sz) AUG_CUDA(cudaMemcpyAsync((unsigned char*)s2 + (size_t)r0 * per * s2sz, d + o_s2,
                                                 (size_t)rows * per * s2sz, cudaMemcpyDeviceToHost, p->s_out));
        if (beta) AUG_CUDA(cudaMemcpy2DAsync(beta + r0, (size_t)(nlat_out > 1 ? ldo : rows) * 8, d + o_b,
                                             (size_t)chunk * 8, (size_t)rows * 8, nlat_out, cudaMemcpyDeviceToHost,
                                             p->s_out));
        if (gamma) AUG_CUDA(cudaMemcpy2DAsync(gamma + r0, (size_t)(nlat_out > 1 ? ldo : rows) * 8, d + o_g,
                                              (size_t)chunk * 8, (size_t)rows * 8, nlat_out,
                                              cudaMemcpyDeviceToHost, p->s_out));
        if (dsc) AUG_CUDA(cudaMemcpyAsync(p->hscal + (size_t)k * AUG_NSCALARS, dsc, sizeof(double) * AUG_NSCALARS,
                                          cudaMemcpyDeviceToHost, p->s_out));
        AUG_CUDA(cudaEventRecord(p->ev_out[s], p->s_out));
    }
    AUG_CUDA(cudaStreamSynchronize(p->s_out));
    AUG_CUDA(cudaStreamSynchronize(c->stream));
    if (scalars_host) {
        for (int64_t k = 0; k < nchunks; ++k) {
            const double* h = p->hscal + (size_t)k * AUG_NSCALARS;
            scalars_host[AUG_S_EXPECTED_LOGTILT] += h[AUG_S_EXPECTED_LOGTILT];
            scalars_host[AUG_S_KL] += h[AUG_S_KL];
            if (cat) scalars_host[AUG_S_FLAGS] += h[AUG_S_FLAGS];
        }
        scalars_host[AUG_S_EXPECTED_AUGLL] = scalars_host[AUG_S_EXPECTED_LOGTILT] + scalars_host[AUG_S_KL];
    }
    return AUG_OK;
}

int32_t aug_aux_sample_host(aug_ctx* c, const aug_lik* lik, int64_t n, int64_t i0, const void* y, const double* f,
                            int64_t ld, double* omega, int64_t* nvar) {
    if (!c) return AUG_ERR_NOT_INIT;
    if (!lik || n < 0 || !f || !omega) return AUG_ERR_BAD_ARG;
    if (lik->kind < 0 || lik->kind >= AUG_NKINDS) return AUG_ERR_BAD_KIND;
    AUG_CUDA(cudaSetDevice(c->device));
    const int kind = lik->kind;
    const bool cat = is_cat(kind), het = kind == AUG_HETERO;
    const int64_t per = cat ? lik->nlatent : 1;
    const size_t ysz = y_size(kind);
    const bool needs_y = kind != AUG_BERNOULLI;
    const bool needs_n = kind == AUG_POISSON || het || cat;
    if ((needs_y && !y) || (needs_n && !nvar) || (het && ld < n)) return AUG_ERR_BAD_ARG;
    const uint64_t offset = c->offset++;   // one RNG tick for the whole call, whatever the chunking
    if (n == 0) return AUG_OK;
    int64_t chunk = host_chunk_elems() / per;
    if (chunk < 2) chunk = 2;
    chunk &= ~(int64_t)1;
    if (chunk > n) chunk = n;
    const int64_t nchunks = (n + chunk - 1) / chunk;
    size_t off = 0;
    const size_t o_y = off; off += up256((size_t)chunk * per * ysz);
    const size_t o_f = off; off += up256((size_t)chunk * per * 8 * (het ? 2 : 1));
    const size_t o_w = off; off += up256((size_t)chunk * per * 8);
    const size_t o_n = off; off += needs_n ? up256((size_t)chunk * per * 8) : 0;
    int32_t rc = pipe_get(c, off, 1);
    if (rc) return rc;
    aug_pipe* p = c->pipe;
    for (int64_t k = 0; k < nchunks; ++k) {
        const int s = (int)(k % PIPE_SLOTS);
        const int64_t r0 = k * chunk;
        const int64_t rows = (n - r0 < chunk) ? n - r0 : chunk;
        unsigned char* d = p->dbuf[s];
        if (k >= PIPE_SLOTS) AUG_CUDA(cudaStreamWaitEvent(p->s_in, p->ev_out[s], 0));
        if (needs_y)
            AUG_CUDA(cudaMemcpyAsync(d + o_y, (const unsigned char*)y + (size_t)r0 * per * ysz,
                                     (size_t)rows * per * ysz, cudaMemcpyHostToDevice, p->s_in));
        if (het)
            AUG_CUDA(cudaMemcpy2DAsync(d + o_f, (size_t)chunk * 8, f + r0, (size_t)ld * 8, (size_t)rows * 8, 2,
                                       cudaMemcpyHostToDevice, p->s_in));
        else
            AUG_CUDA(cudaMemcpyAsync(d + o_f, f + r0 * per, (size_t)rows * per * 8, cudaMemcpyHostToDevice, p->s_in));
        AUG_CUDA(cudaEventRecord(p->ev_in[s], p->s_in));
        AUG_CUDA(cudaStreamWaitEvent(c->stream, p->ev_in[s], 0));
        if (k >= PIPE_SLOTS) AUG_CUDA(cudaStreamWaitEvent(c->stream, p->ev_out[s], 0));
        rc = aug_aux_sample_dev(c, lik, rows, i0 + r0, d + o_y, (const double*)(d + o_f), chunk,
                                (double*)(d + o_w), needs_n ? (int64_t*)(d + o_n) : nullptr, offset);
        if (rc) return rc;
        AUG_CUDA(cudaEventRecord(p->ev_k[s], c->stream));
        AUG_CUDA(cudaStreamWaitEvent(p->s_out, p->ev_k[s], 0));
        AUG_CUDA(cudaMemcpyAsync(omega + r0 * per, d + o_w, (size_t)rows * per * 8, cudaMemcpyDeviceToHost,
                                 p->s_out));
        if (needs_n)
            AUG_CUDA(cudaMemcpyAsync(nvar + r0 * per, d + o_n, (size_t)rows * per * 8, cudaMemcpyDeviceToHost,
                                     p->s_out));
        AUG_CUDA(cudaEventRecord(p->ev_out[s], p->s_out));
    }
    AUG_CUDA(cudaStreamSynchronize(p->s_out));
    AUG_CUDA(cudaStreamSynchronize(c->stream));
    return AUG_OK;
}

}  // extern "C"
