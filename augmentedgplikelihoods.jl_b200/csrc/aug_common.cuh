// aug_common.cuh — context, launch and reduction plumbing shared by all kernels of libaugcuda.so
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/augcuda.h"

#define AUG_BLOCK 256          // threads per CTA for the map/reduce kernels
#define AUG_MAX_GRID 4096      // upper bound on CTAs of a reducing kernel (partials buffer rows)
#define AUG_NRED 4             // doubles reduced per kernel: elt/logtilt, kl/logprior, flags, spare
#define AUG_TABLE_N 512        // entries of the per-likelihood integer-y constant table
#define AUG_PGTAB_H 0.125       // interval width of the r(z) table
#define AUG_PGTAB_N 160        // intervals: z in [0, 20)
#define AUG_PGTAB_DEG 8        // coefficients per interval (degree 7)
#define AUG_MAX_RANKS 8        // GPUs of one NVSwitch box that can share a peer-memory mailbox
#define AUG_XCH_SLOT 16        // 64-bit words per mailbox slot: up to 7 values, two (32 data bits | 32-bit epoch flag) words each
#define AUG_XCH_NVAL 7         // values every exchange carries (unused ones are zeros)
#define AUG_XCH_WORDS (2 * AUG_MAX_RANKS * AUG_XCH_SLOT)   // [parity][source rank][slot]
// bulk area behind the slots: [parity][source rank][AUG_XCH_BULK doubles] — the P / rhs sums of the sparse-GP sweep
// (m*m + m <= 128*128 + 128 doubles), pushed by the finalise kernel and published with the slot's epoch flag
#define AUG_XCH_BULK (128 * 128 + 128 + 8)
#define AUG_XCH_TOTAL_WORDS (AUG_XCH_WORDS + 2 * AUG_MAX_RANKS * AUG_XCH_BULK)

#define AUG_CUDA(x)                                   \
    do {                                              \
        cudaError_t e__ = (x);                        \
        if (e__ != cudaSuccess) return (int32_t)e__;  \
    } while (0)

struct aug_pipe;  // host-buffer pipeline state (aug_host.cu)

// Peer-memory exchange of the scalar block (aug_ctx.cu: aug_comm_p2p_*).  Lives in device memory of the owning rank;
// box[r] is rank r's mailbox mapped into this process (cudaIpc, or a plain peer pointer inside one process).
// The finalising thread of a reducing kernel pushes its partial sums into slot [epoch & 1][rank] of EVERY rank's
// mailbox over NVLink, publishes the epoch, waits for the nranks slots of its own mailbox and adds them in rank
// order: an all-reduce fused into the kernel that produced the sums, bit-identical on all ranks.  `epoch` is a
// device-side counter (one tick per exchanging launch), so the launch is graph-replayable.
struct AugXchDev {
    unsigned long long* box[AUG_MAX_RANKS];
    int nranks, rank;
    unsigned long long epoch;
    unsigned int* err;                 // ctx error flag word: bit 1 = a peer did not arrive within timeout_ns
    unsigned long long timeout_ns;
    // split-phase exchange (aug_comm_set_deferred): the reducing kernel only PUBLISHES its sums and leaves this note;
    // a later launch on the same stream (an extra CTA of the next sampling kernel, or a 1-thread flush kernel) gathers
    // the ranks' sums and completes the scalar block — the wait for the slowest rank is off the reducing kernel
    double* pend_scalars;              // nullptr: nothing pending
    int pend_nv, pend_first;           // values published (2 or 3), first slot of the verb's triple
};

struct aug_ctx {
    int device;
    cudaStream_t stream;
    bool own_stream;
    int sms;
    int smem_per_sm, smem_optin;   // shared memory per SM / per CTA (opt-in) in bytes
    uint64_t seed, offset;
    uint64_t launches;
    // reduction scratch
    double* partials;      // [AUG_MAX_GRID][AUG_NRED]
    unsigned int* counter; // last-block ticket
    double* dscalars;      // [AUG_NSCALARS] scratch block used by the host-buffer calls
    unsigned int* dflag;   // device-side error flag word
    // integer-y constant table (log y!, negbin log-constants), cached by likelihood parameters
    double* table;
    int table_kind;
    int table_r_is_int;
    double table_param;
    double* pgtab;         // piecewise polynomial table of the PG(1,z) proposal mass r(z) (aug_pg.cuh)
    // categorical logθ-derived constants
    double* dtheta;        // exp(logθ_j)/Σθ  (device, capacity dtheta_cap)
    int dtheta_cap;
    double* htheta;        // host copy of what dtheta holds (htheta_n values; 0: nothing cached): a verb whose logθ
    int htheta_n;          // matches skips the upload and its stream synchronisation
    // NCCL (resolved with dlopen at aug_comm_init)
    void* nccl_lib;
    void* nccl_comm;
    int nranks, rank;
    // peer-memory mailbox (aug_comm_p2p_*): the all-reduce of the scalar block fused into the reducing kernels
    unsigned long long* mailbox;                   // this rank's mailbox [AUG_XCH_WORDS] (cudaMalloc, IPC-exported)
    unsigned long long* peer_box[AUG_MAX_RANKS];   // mapped mailboxes of all ranks (peer_box[rank] == mailbox)
    bool peer_ipc[AUG_MAX_RANKS];                  // opened with cudaIpcOpenMemHandle (to be closed)
    AugXchDev* xch;                                // device copy of the exchange descriptor (nullptr: not attached)
    int xch_ranks, xch_rank;
    int fused;                                     // reducing verbs return globally reduced scalars
    int deferred;                                  // fused mode, split-phase: CAVI verbs publish, a later launch gathers
    int pending;                                   // a published exchange has not been gathered yet (host-side mirror)
    aug_pipe* pipe;
    // sparse-GP sweep (aug_sparse.cu): per-CTA partial P / rhs / ELBO sums, summed in a fixed order by a finalise launch
    double* sparse_scratch;
    size_t sparse_scratch_bytes;
    void* cublas;          // cublasHandle_t for the m > 128 composition (created on first use)
};

// Likelihood constants precomputed on the host once per call (never per observation)
struct LikConst {
    int kind, nl, r_is_int, bij;
    int quirks;             // AUG_LIK_FAITHFUL_QUIRKS
    double p0, p1;          // raw parameters (r | λ | β | ν, σ)
    double c0, c1, c2, c3, c4, c5;  // derived constants, meaning per kind (see lik_const())
    const double* table;    // device table or nullptr
    const double* pgtab;    // r(z) table (always set)
    const double* theta;    // CAT: θ_j/Σθ device vector
};

int32_t aug_lik_const(aug_ctx* ctx, const aug_lik* lik, LikConst* out, bool need_table, bool need_theta);
int aug_grid_for(aug_ctx* ctx, const void* kernel, int64_t work_items, int items_per_block);
// exchange descriptor for the FINAL launch of a scalar-producing verb, or nullptr when the ctx is not in fused mode
static inline AugXchDev* aug_xch_for(aug_ctx* ctx) { return (ctx->fused && ctx->xch) ? ctx->xch : nullptr; }
// complete a pending split-phase exchange with a 1-thread kernel on the ctx stream (no-op when nothing is pending)
int32_t aug_xch_flush(aug_ctx* ctx);
// fused mode with an empty shard: the rank still has to take part in the exchange (aug_ctx.cu)
int32_t aug_xch_zero_contribution(aug_ctx* ctx, double* scalars, int first_slot, int nslots);

static inline bool aug_aligned16(const void* p) { return p == nullptr || (((uintptr_t)p) & 15u) == 0; }

#ifdef __CUDACC__
// ---------------------------------------------------------------- device helpers
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Block tree: warp shuffles, then one smem round, then a fixed-order last-block pass over the
// per-CTA partials.  Deterministic for a given grid size (no floating-point atomics).
// Returns true in exactly one thread of the grid (thread 0 of the last CTA), which holds the totals.
// BLOCK = threads per CTA of the calling kernel (a multiple of 32).
template <int NV, int BLOCK = AUG_BLOCK>
__device__ __forceinline__ bool block_reduce_and_finalize(double (&acc)[NV], double* __restrict__ partials,
                                                          unsigned int* __restrict__ counter,
                                                          double* __restrict__ out /* NV results */) {
    __shared__ double sm[NV][BLOCK / 32];
    __shared__ bool is_last;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        double v = warp_sum(acc[k]);
        if (lane == 0) sm[k][warp] = v;
    }
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            double v = lane < BLOCK / 32 ? sm[k][lane] : 0.0;
            v = warp_sum(v);
            if (lane == 0) partials[(size_t)blockIdx.x * AUG_NRED + k] = v;
        }
    }
    if (threadIdx.x == 0) {
        __threadfence();
        unsigned int t = atomicInc(counter, gridDim.x - 1);  // wraps to 0 -> reusable / graph-replayable
        is_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (!is_last) return false;
    __threadfence();
    double tot[NV];
#pragma unroll
    for (int k = 0; k < NV; ++k) tot[k] = 0.0;
    for (unsigned int b = threadIdx.x; b < gridDim.x; b += BLOCK) {
#pragma unroll
        for (int k = 0; k < NV; ++k) tot[k] += __ldcg(&partials[(size_t)b * AUG_NRED + k]);
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        double v = warp_sum(tot[k]);
        if (lane == 0) sm[k][warp] = v;
    }
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            double v = lane < BLOCK / 32 ? sm[k][lane] : 0.0;
            v = warp_sum(v);
            if (lane == 0) out[k] = v;
        }
    }
    return threadIdx.x == 0;
}

// ---------------------------------------------------------------- all-reduce over peer memory, fused into the finaliser
__device__ __forceinline__ unsigned long long xch_globaltimer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
// Slot protocol (the one NCCL's LL protocol uses): every 64-bit word of a slot carries 32 data bits and the low 32 bits of the
// epoch.  An aligned 8-byte store is single-copy atomic, so a word whose flag matches holds valid data whatever the order the
// words arrive in: NO fence between data and flag, nothing to wait for on the publishing side — the stores are posted and the
// kernel moves on (round-2 8-GPU probe: a release store per peer cost +32 us on the CAVI launch, one fence + relaxed flags
// +13 us, this +a few).  A double takes two words; a slot holds up to 7 values.
// Called by ONE thread of the grid (the finaliser).  Phase 1: push this rank's sums v[0..NV) into every rank's mailbox.
template <int NV>
__device__ __forceinline__ void xch_publish(AugXchDev* __restrict__ x, const double (&v)[NV]) {
    static_assert(NV <= AUG_XCH_NVAL && 2 * AUG_XCH_NVAL <= AUG_XCH_SLOT, "slot holds 7 values, two words each");
    const int nr = x->nranks, me = x->rank;
    const unsigned long long ep = x->epoch + 1ull;
    x->epoch = ep;
    const unsigned long long flag = (ep & 0xffffffffull) << 32;
    const size_t half = (size_t)(ep & 1ull) * AUG_MAX_RANKS * AUG_XCH_SLOT;
    const size_t mine = half + (size_t)me * AUG_XCH_SLOT;
    for (int r = 0; r < nr; ++r) {                         // posted stores over NVLink; local for r == me
        unsigned long long* p = x->box[r] + mine;
        // ALL seven value positions are written (zeros beyond NV) and a gather waits for all of them: ranks may enter one
        // exchange through different code paths with different NV (an empty shard's zero contribution, a deferred verb's
        // triple against a pair) and must still complete each other's slots
#pragma unroll
        for (int k = 0; k < AUG_XCH_NVAL; ++k) {
            const unsigned long long bits = k < NV ? (unsigned long long)__double_as_longlong(v[k < NV ? k : 0]) : 0ull;
            asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p + 2 * k), "l"((bits & 0xffffffffull) | flag) : "memory");
            asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p + 2 * k + 1), "l"((bits >> 32) | flag) : "memory");
        }
    }
}
// Phase 2: wait for the current epoch's words of all ranks in this rank's own mailbox and add the values IN RANK ORDER
// (bit-identical on every rank, independent of arrival order).  A peer that does not arrive within timeout_ns raises
// bit 1 of the error flag and yields NaN.
template <int NV>
__device__ __forceinline__ void xch_gather(AugXchDev* __restrict__ x, double (&v)[NV]) {
    const int nr = x->nranks, me = x->rank;
    const unsigned long long ep = x->epoch;
    const unsigned long long flag = ep & 0xffffffffull;
    const size_t half = (size_t)(ep & 1ull) * AUG_MAX_RANKS * AUG_XCH_SLOT;
    double tot[NV];
#pragma unroll
    for (int k = 0; k < NV; ++k) tot[k] = 0.0;
    const unsigned long long t0 = xch_globaltimer();
    bool ok = true;
    for (int r = 0; r < nr && ok; ++r) {
        const unsigned long long* p = x->box[me] + half + (size_t)r * AUG_XCH_SLOT;
        unsigned long long w[2 * AUG_XCH_NVAL];
        for (;;) {
            bool all = true;
#pragma unroll
            for (int k = 0; k < 2 * AUG_XCH_NVAL; ++k) {   // all loads in flight together
                asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(w[k]) : "l"(p + k) : "memory");
            }
#pragma unroll
            for (int k = 0; k < 2 * AUG_XCH_NVAL; ++k) all = all && (w[k] >> 32) == flag;
            if (all) break;
            if (xch_globaltimer() - t0 > x->timeout_ns) { ok = false; break; }
            __nanosleep(64);
        }
        if (!ok) break;
#pragma unroll
        for (int k = 0; k < NV; ++k)
            tot[k] += __longlong_as_double((long long)((w[2 * k] & 0xffffffffull) | (w[2 * k + 1] << 32)));
    }
    if (!ok) {
        atomicOr(x->err, 2u);
#pragma unroll
        for (int k = 0; k < NV; ++k) tot[k] = __longlong_as_double(0x7ff8000000000000ll);
    }
#pragma unroll
    for (int k = 0; k < NV; ++k) v[k] = tot[k];
}
// v[0..NV) in: this rank's sums; out: the sums over all ranks (both phases back to back: the kernel waits for its peers).
template <int NV>
__device__ __forceinline__ void xch_allreduce(AugXchDev* __restrict__ x, double (&v)[NV]) {
    xch_publish<NV>(x, v);
    xch_gather<NV>(x, v);
}
// Deferred mode, phase 1 from a verb's finaliser: publish, leave the note, and put this rank's LOCAL sums in the block
// until the gather replaces them.  scalars[first .. first+2] = (a, b, a + b); nv == 3 also carries the flag count.
__device__ __forceinline__ void xch_publish_deferred(AugXchDev* __restrict__ x, double* __restrict__ scalars, int first,
                                                     const double (&v)[3], int nv) {
    xch_publish<3>(x, v);
    x->pend_scalars = scalars;
    x->pend_nv = nv;
    x->pend_first = first;
    __threadfence();
}
// Deferred mode, phase 2 (one thread): complete the scalar block the note points to.
__device__ __forceinline__ void xch_finish_pending(AugXchDev* __restrict__ x) {
    double* s = *(double* volatile*)&x->pend_scalars;
    if (s == nullptr) return;
    double v[3];
    xch_gather<3>(x, v);
    const int first = x->pend_first;
    s[first] = v[0];
    s[first + 1] = v[1];
    s[first + 2] = v[0] + v[1];
    if (x->pend_nv == 3) s[AUG_S_FLAGS] = v[2];
    x->pend_scalars = nullptr;
    __threadfence();
}

// Every scalar-producing verb owns the WHOLE 8-slot block of its call: the finaliser writes the verb's slots and zeroes
// the rest, so a caller never has to clear the block (and an 8-slot all-reduce never sums stale values).
__device__ __forceinline__ void scal_zero_except(double* __restrict__ s, unsigned keep_mask) {
#pragma unroll
    for (int k = 0; k < AUG_NSCALARS; ++k)
        if (!((keep_mask >> k) & 1u)) s[k] = 0.0;
}

// 128-bit streaming loads/stores (read-once / write-once data: keep it out of L1)
__device__ __forceinline__ double2 ld_stream2(const double* p) {
    double2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
    return r;
}
__device__ __forceinline__ double ld_stream1(const double* p) {
    double r;
    asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(r) : "l"(p));
    return r;
}
__device__ __forceinline__ void st_stream2(double* p, double a, double b) {
    asm volatile("st.global.cs.v2.f64 [%0], {%1, %2};" ::"l"(p), "d"(a), "d"(b) : "memory");
}
__device__ __forceinline__ void st_stream1(double* p, double a) {
    asm volatile("st.global.cs.f64 [%0], %1;" ::"l"(p), "d"(a) : "memory");
}
__device__ __forceinline__ void st_stream2_i64(int64_t* p, int64_t a, int64_t b) {
    asm volatile("st.global.cs.v2.s64 [%0], {%1, %2};" ::"l"(p), "l"(a), "l"(b) : "memory");
}
// ---------------------------------------------------------------- bulk-async (TMA 1-D) + mbarrier helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// pull a span into L2 ahead of the bulk copy that will stage it (no completion tracking)
__device__ __forceinline__ void bulk_prefetch_l2(const void* src, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!done);
}
#endif  // __CUDACC__
