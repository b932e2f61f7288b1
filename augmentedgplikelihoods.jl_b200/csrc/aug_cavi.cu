// aug_cavi.cu — fused single-pass CAVI kernels for the scalar-latent likelihoods and the
// two-latent heteroscedastic Gaussian.
//
// One launch = aux_posterior! + expected_auglik_potential_and_precision (+ the per-observation
// expected_logtilt / aux_kldivergence terms reduced in the same pass): y, mu, var are read once
// with 128-bit loads, state / beta / gamma are written once with 128-bit streaming stores, and the
// ELBO partials go through warp shuffles, a block tree and a fixed-order last-block pass.
// The same template, instantiated with FROM_STATE, serves the stand-alone verbs that start from an
// existing qΩ (expected_auglik_*, expected_logtilt, aux_kldivergence).
//
// Reference behaviour replaced (paths relative to /root/reference/src):
//   likelihoods/bernoulli.jl:17-65, negativebinomial.jl:24-73, poisson.jl:30-85, laplace.jl:44-104,
//   studentt.jl:50-91, heteroscedasticgaussian.jl:34-145, utils.jl:1-14,
//   SpecialDistributions/polyagamma.jl:25-31,99-110, polyagammapoisson.jl:35-51, api.jl:219-223,
//   generic.jl:52-62.
#include <stdlib.h>

#include "aug_common.cuh"
#include "aug_math.cuh"
#include "aug_cavi_eval.cuh"

#ifndef CAVI_MIN_BLOCKS
#define CAVI_MIN_BLOCKS 4
#endif
#ifndef CAVI_U
#define CAVI_U 1
#endif
#ifndef CAVI_PREFETCH
#define CAVI_PREFETCH 0
#endif

namespace {

struct CaviArgs {
    int64_t n;
    const void* y;
    const double* mu;    // latent 0 (f)
    const double* var;
    const double* mu_g;  // latent 1 (HETERO g)
    const double* var_g;
    // state outputs (fused / aux_posterior!) — any may be null
    double* s0;
    double* s1;
    void* s2;
    // state inputs (FROM_STATE)
    const double* rs0;
    const double* rs1;
    const void* rs2;
    double* beta;
    double* gamma;
    double* beta_g;
    double* gamma_g;
    double* partials;
    unsigned int* counter;
    double* scalars;
    int accumulate;       // add to the scalars already in memory (second launch of one verb)
    AugXchDev* xch;       // non-null on the final launch of a verb in fused multi-GPU mode: all-reduce over peer memory
    int xch_defer;        // split-phase: publish only, a later launch gathers (aug_comm_set_deferred)
    LikConst L;
};


// ------------------------------------------------------------------ loads / stores
template <typename T>
__device__ __forceinline__ void load_y2(const void* y, int64_t p, double& a, double& b);
template <>
__device__ __forceinline__ void load_y2<uint8_t>(const void* y, int64_t p, double& a, double& b) {
    const uchar2 v = __ldg(reinterpret_cast<const uchar2*>(y) + p);
    a = (double)v.x;
    b = (double)v.y;
}
template <>
__device__ __forceinline__ void load_y2<int64_t>(const void* y, int64_t p, double& a, double& b) {
    const longlong2 v = __ldg(reinterpret_cast<const longlong2*>(y) + p);
    a = (double)v.x;
    b = (double)v.y;
}
template <>
__device__ __forceinline__ void load_y2<double>(const void* y, int64_t p, double& a, double& b) {
    const double2 v = ld_stream2(reinterpret_cast<const double*>(y) + 2 * p);
    a = v.x;
    b = v.y;
}
template <typename T>
__device__ __forceinline__ double load_y1(const void* y, int64_t i) {
    return (double)__ldg(reinterpret_cast<const T*>(y) + i);
}
template <typename T>
__device__ __forceinline__ void store_y2(void* s2, int64_t p, double a, double b);
template <>
__device__ __forceinline__ void store_y2<int64_t>(void* s2, int64_t p, double a, double b) {
    reinterpret_cast<longlong2*>(s2)[p] = make_longlong2((long long)a, (long long)b);
}
template <>
__device__ __forceinline__ void store_y2<double>(void* s2, int64_t p, double a, double b) {
    st_stream2(reinterpret_cast<double*>(s2) + 2 * p, a, b);
}
template <>
__device__ __forceinline__ void store_y2<uint8_t>(void*, int64_t, double, double) {}
template <typename T>
__device__ __forceinline__ void store_y1(void* s2, int64_t i, double a) {
    reinterpret_cast<T*>(s2)[i] = (T)a;
}

__device__ __forceinline__ void write_scalars(const CaviArgs& a, const double (&out)[2]) {
    double e = out[0], k = out[1];
    if (a.accumulate) {
        e += a.scalars[AUG_S_EXPECTED_LOGTILT];
        k += a.scalars[AUG_S_KL];
    }
    if (a.xch && a.xch_defer) {               // the block holds the local sums until the deferred gather replaces them
        const double v[3] = {e, k, 0.0};
        xch_publish_deferred(a.xch, a.scalars, AUG_S_EXPECTED_LOGTILT, v, 2);
    } else if (a.xch) {
        double v[2] = {e, k};
        xch_allreduce<2>(a.xch, v);
        e = v[0];
        k = v[1];
    }
    a.scalars[AUG_S_EXPECTED_LOGTILT] = e;
    a.scalars[AUG_S_KL] = k;
    a.scalars[AUG_S_EXPECTED_AUGLL] = e + k;   // generic.jl:52-54 ("+")
    scal_zero_except(a.scalars, 0x07u);
}

template <int KIND, bool FROM_STATE, bool ELBO, bool VEC>
__global__ void __launch_bounds__(AUG_BLOCK, CAVI_MIN_BLOCKS) cavi_kernel(const CaviArgs a) {
    typedef typename YT<KIND>::T yt;
    typedef typename S2T<KIND>::T s2t;
    constexpr bool HET = KIND == AUG_HETERO;
    constexpr bool YSTATE = KIND == AUG_NEGBIN || KIND == AUG_POISSON;
    constexpr bool HAS_S1 = KIND == AUG_POISSON || KIND == AUG_HETERO;
    constexpr int U = HET ? 1 : CAVI_U;  // independent pairs in flight per thread
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t nth = (int64_t)gridDim.x * blockDim.x;
    double acc[2] = {0.0, 0.0};
    const bool need_mv = !FROM_STATE || ELBO;
    const int64_t npairs = VEC ? (a.n >> 1) : 0;

    // One observation through the IEEE / libdevice instantiation (any input).  Used for the odd tail
    // element, for unaligned arrays, and to redo the pairs of a thread that met an out-of-range value.
    auto scalar_obs = [&](int64_t i) {
        Obs o;
        o.y = load_y1<yt>(a.y, i);
        o.ys = o.y;
        o.m = o.v = o.mg = o.vg = 0.0;
        o.s0 = o.s1 = o.s2 = 0.0;
        if (need_mv) { o.m = a.mu[i]; o.v = a.var[i]; }
        if (HET) { o.mg = a.mu_g[i]; if (need_mv) o.vg = a.var_g[i]; }
        if (FROM_STATE) {
            o.s0 = a.rs0[i];
            if (HAS_S1) o.s1 = a.rs1[i];
            if (HET) o.s2 = reinterpret_cast<const double*>(a.rs2)[i];
            if (YSTATE && a.rs2 != nullptr) o.ys = (double)reinterpret_cast<const int64_t*>(a.rs2)[i];
        }
        // the result is a function of the observation alone: straight-line math when it is in range, IEEE / libdevice
        // otherwise — the same choice every kernel makes, so an observation gets the same bits on every route
        if (fast_ok<KIND, FROM_STATE, ELBO>(o)) eval<KIND, FROM_STATE, ELBO, false>(a.L, o);
        else eval<KIND, FROM_STATE, ELBO, true>(a.L, o);
        if (!FROM_STATE) {
            if (a.s0) a.s0[i] = o.s0;
            if (HAS_S1 && a.s1) a.s1[i] = o.s1;
            if (HET && a.s2) reinterpret_cast<double*>(a.s2)[i] = o.s2;
            if (YSTATE && a.s2) store_y1<s2t>(a.s2, i, o.y);
        }
        if (a.beta) a.beta[i] = o.b0;
        if (a.gamma) a.gamma[i] = o.g0;
        if (HET) {
            if (a.beta_g) a.beta_g[i] = o.b1;
            if (a.gamma_g) a.gamma_g[i] = o.g1;
        }
        if (ELBO) { acc[0] += o.elt; acc[1] += o.kl; }
    };

    // loads of one pair (two consecutive observations) — all 128-bit except the 16-bit Bool pair
    auto load_pair = [&](int64_t p, Obs& o0, Obs& o1) {
        load_y2<yt>(a.y, p, o0.y, o1.y);
        if (need_mv) {
            const double2 m = ld_stream2(a.mu + 2 * p);
            const double2 v = ld_stream2(a.var + 2 * p);
            o0.m = m.x; o1.m = m.y;
            o0.v = v.x; o1.v = v.y;
        }
        if (HET) {
            const double2 m = ld_stream2(a.mu_g + 2 * p);
            o0.mg = m.x; o1.mg = m.y;
            if (need_mv) {
                const double2 v = ld_stream2(a.var_g + 2 * p);
                o0.vg = v.x; o1.vg = v.y;
            }
        }
        o0.ys = o0.y; o1.ys = o1.y;
        if (FROM_STATE) {
            const double2 c = ld_stream2(a.rs0 + 2 * p);
            o0.s0 = c.x; o1.s0 = c.y;
            if (HAS_S1) {
                const double2 l = ld_stream2(a.rs1 + 2 * p);
                o0.s1 = l.x; o1.s1 = l.y;
            }
            if (HET) {
                const double2 ps = ld_stream2(reinterpret_cast<const double*>(a.rs2) + 2 * p);
                o0.s2 = ps.x; o1.s2 = ps.y;
            }
            if (YSTATE && a.rs2 != nullptr) load_y2<int64_t>(a.rs2, p, o0.ys, o1.ys);
        }
    };
    auto dead_pair = [&](Obs& o0, Obs& o1) {   // harmless inputs for the straight-line math of a dead slot
        o0.y = o1.y = o0.ys = o1.ys = 0.0;
        o0.m = o1.m = o0.mg = o1.mg = 0.0;
        o0.v = o1.v = o0.vg = o1.vg = 1.0;
        o0.s0 = o1.s0 = o0.s1 = o1.s1 = o0.s2 = o1.s2 = 1.0;
    };
    auto store_pair = [&](int64_t p, const Obs& o0, const Obs& o1) {
        if (!FROM_STATE) {
            if (a.s0) st_stream2(a.s0 + 2 * p, o0.s0, o1.s0);
            if (HAS_S1 && a.s1) st_stream2(a.s1 + 2 * p, o0.s1, o1.s1);
            if (HET && a.s2) st_stream2(reinterpret_cast<double*>(a.s2) + 2 * p, o0.s2, o1.s2);
            if (YSTATE && a.s2) store_y2<s2t>(a.s2, p, o0.y, o1.y);
        }
        if (a.beta) st_stream2(a.beta + 2 * p, o0.b0, o1.b0);
        if (a.gamma) st_stream2(a.gamma + 2 * p, o0.g0, o1.g0);
        if (HET) {
            if (a.beta_g) st_stream2(a.beta_g + 2 * p, o0.b1, o1.b1);
            if (a.gamma_g) st_stream2(a.gamma_g + 2 * p, o0.g1, o1.g1);
        }
        if (ELBO) {
            acc[0] += o0.elt + o1.elt;
            acc[1] += o0.kl + o1.kl;
        }
    };

    if (VEC) {
        bool bad = false;   // some observation of this thread was outside the fast-math range
#if CAVI_PREFETCH
        // software pipeline: the loads of the next pair are in flight while this one is evaluated
        Obs c0, c1, n0, n1;
        if (tid < npairs) load_pair(tid, c0, c1);
        for (int64_t p = tid; p < npairs; p += nth) {
            const int64_t pn = p + nth;
            if (pn < npairs) load_pair(pn, n0, n1);
            bad = bad || !fast_ok<KIND, FROM_STATE, ELBO>(c0) || !fast_ok<KIND, FROM_STATE, ELBO>(c1);
            eval<KIND, FROM_STATE, ELBO, false>(a.L, c0);
            eval<KIND, FROM_STATE, ELBO, false>(a.L, c1);
            store_pair(p, c0, c1);
            c0 = n0;
            c1 = n1;
        }
#else
        for (int64_t p0 = tid; p0 < npairs; p0 += nth * U) {
            Obs o[U][2];
            bool live[U];
#pragma unroll
            for (int j = 0; j < U; ++j) {
                const int64_t p = p0 + (int64_t)j * nth;
                live[j] = p < npairs;
                if (live[j]) load_pair(p, o[j][0], o[j][1]);
                else dead_pair(o[j][0], o[j][1]);
            }
            // straight-line evaluation of all 2U observations (no branch: the compiler interleaves them)
#pragma unroll
            for (int j = 0; j < U; ++j) {
                bad = bad || !fast_ok<KIND, FROM_STATE, ELBO>(o[j][0]) || !fast_ok<KIND, FROM_STATE, ELBO>(o[j][1]);
                eval<KIND, FROM_STATE, ELBO, false>(a.L, o[j][0]);
                eval<KIND, FROM_STATE, ELBO, false>(a.L, o[j][1]);
            }
#pragma unroll
            for (int j = 0; j < U; ++j)
                if (live[j]) store_pair(p0 + (int64_t)j * nth, o[j][0], o[j][1]);
        }
#endif
        if (bad) {
            // rare: redo every pair of this thread with the any-input instantiation (outputs are
            // simply overwritten, the thread's partial sums restart from zero)
            acc[0] = acc[1] = 0.0;
            for (int64_t p = tid; p < npairs; p += nth) {
                scalar_obs(2 * p);
                scalar_obs(2 * p + 1);
            }
        }
    }
    // the odd tail element of the vector kernel, or everything when the arrays are not 16-byte aligned
    for (int64_t i = 2 * npairs + tid; i < a.n; i += nth) scalar_obs(i);

    if (ELBO) {
        double out[2];
        if (block_reduce_and_finalize<2>(acc, a.partials, a.counter, out)) write_scalars(a, out);
    }
}


// ------------------------------------------------------------------ bulk-async (TMA) staged variant
// ncu on the direct-load kernel (profiles/r1b): 60% of the stall cycles are long-scoreboard waits on the
// global loads — with 64 registers/thread there are only ~35 KB of loads in flight per SM, below what
// Little's law asks for at 6.4 TB/s.  Here one elected thread streams whole input tiles into a shared-memory
// ring with cp.async.bulk (SASS: UBLKCP) completing on mbarriers, so the bytes in flight are set by the
// ring (STAGES x tile bytes per CTA), not by registers or occupancy; the 256 threads consume a stage with
// conflict-free 128-bit LDS, evaluate the straight-line math, and write state / beta / gamma straight to
// global with 128-bit streaming stores.
#ifndef CAVI_TMA_BLOCKS
#define CAVI_TMA_BLOCKS 2   // CTAs per SM the ring is sized for (~100 KB of ring per CTA)
#endif
// tile (observations per stage) and ring depth per likelihood: ~25-35 KB per stage, ~100 KB per CTA
__host__ __device__ constexpr int cavi_tile(int kind) { return kind == AUG_BERNOULLI ? 2048 : (kind == AUG_HETERO ? 512 : 1024); }
__host__ __device__ constexpr int cavi_stages(int kind) { return kind == AUG_BERNOULLI ? 3 : (kind == AUG_HETERO ? 5 : 4); }
__host__ __device__ constexpr int pow2_floor(int x) { int p = 1; while (2 * p <= x) p *= 2; return p; }
__host__ __device__ constexpr int clamp_int(int x, int lo, int hi) { return x < lo ? lo : (x > hi ? hi : x); }

// What one ring stage holds.  Fused calls stage y, mu, var (+ the second latent); the FROM_STATE verbs
// (expected_auglik_potential_and_precision, expected_logtilt / aux_kldivergence from an existing q(Omega)) stage the
// state arrays next to them and drop mu / var when the verb does not need them.
template <int KIND, bool FS, bool ELBO>
struct TileLayout {
    typedef typename YT<KIND>::T yt;
    static constexpr bool HET = KIND == AUG_HETERO;
    static constexpr bool YSTATE = KIND == AUG_NEGBIN || KIND == AUG_POISSON;
    static constexpr bool HAS_S1 = KIND == AUG_POISSON || KIND == AUG_HETERO;
    static constexpr bool NEED_MV = !FS || ELBO;
    static constexpr bool HAS_MUG = HET, HAS_VARG = HET && NEED_MV;
    static constexpr bool HAS_RS0 = FS, HAS_RS1 = FS && HAS_S1, HAS_RS2 = FS && (HET || YSTATE);
    static constexpr int N_D = (NEED_MV ? 2 : 0) + (HAS_MUG ? 1 : 0) + (HAS_VARG ? 1 : 0) + (HAS_RS0 ? 1 : 0) +
                               (HAS_RS1 ? 1 : 0) + (HAS_RS2 ? 1 : 0);
    static constexpr int OBS_BYTES = (int)sizeof(yt) + 8 * N_D;
    static constexpr int T = FS ? clamp_int(pow2_floor(36864 / OBS_BYTES), 512, 4096) : cavi_tile(KIND);
    static constexpr int Y_BYTES = T * (int)sizeof(yt);
    static constexpr int D_BYTES = T * 8;
    static constexpr int OFF_MU = (Y_BYTES + 127) / 128 * 128;
    static constexpr int OFF_VAR = OFF_MU + (NEED_MV ? D_BYTES : 0);
    static constexpr int OFF_MUG = OFF_VAR + (NEED_MV ? D_BYTES : 0);
    static constexpr int OFF_VARG = OFF_MUG + (HAS_MUG ? D_BYTES : 0);
    static constexpr int OFF_RS0 = OFF_VARG + (HAS_VARG ? D_BYTES : 0);
    static constexpr int OFF_RS1 = OFF_RS0 + (HAS_RS0 ? D_BYTES : 0);
    static constexpr int OFF_RS2 = OFF_RS1 + (HAS_RS1 ? D_BYTES : 0);
    static constexpr int STAGE_BYTES = OFF_RS2 + (HAS_RS2 ? D_BYTES : 0);
    static constexpr int S = FS ? clamp_int(110592 / STAGE_BYTES, 2, 6) : cavi_stages(KIND);
    static constexpr int TX_BYTES = Y_BYTES + D_BYTES * N_D;   // minus D_BYTES when a NegBin / Poisson caller passes no y-state
    static constexpr int SMEM_BYTES = STAGE_BYTES * S + 64;
};

template <typename T>
__device__ __forceinline__ void lds_y2(const unsigned char* ys, int q, double& a, double& b);
template <>
__device__ __forceinline__ void lds_y2<uint8_t>(const unsigned char* ys, int q, double& a, double& b) {
    const uchar2 v = reinterpret_cast<const uchar2*>(ys)[q];
    a = (double)v.x;
    b = (double)v.y;
}
template <>
__device__ __forceinline__ void lds_y2<int64_t>(const unsigned char* ys, int q, double& a, double& b) {
    const longlong2 v = reinterpret_cast<const longlong2*>(ys)[q];
    a = (double)v.x;
    b = (double)v.y;
}
template <>
__device__ __forceinline__ void lds_y2<double>(const unsigned char* ys, int q, double& a, double& b) {
    const double2 v = reinterpret_cast<const double2*>(ys)[q];
    a = v.x;
    b = v.y;
}

// The staged kernel over the full tiles [0, ntiles * tile) of the shard (+ the ragged tail by its last CTA):
// fused CAVI step (FS = false) or one of the verbs that start from an existing q(Omega) (FS = true).
template <int KIND, bool FS, bool ELBO>
__global__ void __launch_bounds__(AUG_BLOCK, CAVI_TMA_BLOCKS) cavi_tma_kernel(const CaviArgs a, const int64_t ntiles) {
    typedef TileLayout<KIND, FS, ELBO> TL;
    typedef typename YT<KIND>::T yt;
    typedef typename S2T<KIND>::T s2t;
    constexpr bool HET = KIND == AUG_HETERO;
    constexpr bool YSTATE = KIND == AUG_NEGBIN || KIND == AUG_POISSON;
    constexpr bool HAS_S1 = KIND == AUG_POISSON || KIND == AUG_HETERO;
    constexpr int T = TL::T, S = TL::S;
    extern __shared__ __align__(128) unsigned char ring[];
    uint64_t* full = reinterpret_cast<uint64_t*>(ring + (size_t)TL::STAGE_BYTES * S);

    const int tid = threadIdx.x;
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < S; ++s) mbar_init(&full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const bool has_rs2 = TL::HAS_RS2 && a.rs2 != nullptr;   // HETERO: always (host-checked); NegBin / Poisson: optional

    auto issue = [&](int64_t tile, int s) {   // one thread: arm the barrier, then the bulk copies of one tile
        unsigned char* st = ring + (size_t)s * TL::STAGE_BYTES;
        const int64_t o = tile * T;
        mbar_expect_tx(&full[s], TL::TX_BYTES - ((TL::HAS_RS2 && !has_rs2) ? TL::D_BYTES : 0));
        bulk_g2s(st, reinterpret_cast<const unsigned char*>(a.y) + o * sizeof(yt), TL::Y_BYTES, &full[s]);
        if (TL::NEED_MV) {
            bulk_g2s(st + TL::OFF_MU, a.mu + o, TL::D_BYTES, &full[s]);
            bulk_g2s(st + TL::OFF_VAR, a.var + o, TL::D_BYTES, &full[s]);
        }
        if (TL::HAS_MUG) bulk_g2s(st + TL::OFF_MUG, a.mu_g + o, TL::D_BYTES, &full[s]);
        if (TL::HAS_VARG) bulk_g2s(st + TL::OFF_VARG, a.var_g + o, TL::D_BYTES, &full[s]);
        if (TL::HAS_RS0) bulk_g2s(st + TL::OFF_RS0, a.rs0 + o, TL::D_BYTES, &full[s]);
        if (TL::HAS_RS1) bulk_g2s(st + TL::OFF_RS1, a.rs1 + o, TL::D_BYTES, &full[s]);
        if (has_rs2) bulk_g2s(st + TL::OFF_RS2, reinterpret_cast<const unsigned char*>(a.rs2) + o * 8, TL::D_BYTES, &full[s]);
    };

    const int64_t first = blockIdx.x, stride = gridDim.x;
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < S; ++s) {
            const int64_t t = first + (int64_t)s * stride;
            if (t < ntiles) issue(t, s);
        }
    }
    double acc[2] = {0.0, 0.0};
    bool bad = false;
    int64_t k = 0;
    for (int64_t tile = first; tile < ntiles; tile += stride, ++k) {
        const int s = (int)(k % S);
        mbar_wait(&full[s], (uint32_t)((k / S) & 1));
        const unsigned char* st = ring + (size_t)s * TL::STAGE_BYTES;
        const double2* smu = reinterpret_cast<const double2*>(st + TL::OFF_MU);
        const double2* svar = reinterpret_cast<const double2*>(st + TL::OFF_VAR);
        const double2* smug = reinterpret_cast<const double2*>(st + TL::OFF_MUG);
        const double2* svarg = reinterpret_cast<const double2*>(st + TL::OFF_VARG);
        const double2* srs0 = reinterpret_cast<const double2*>(st + TL::OFF_RS0);
        const double2* srs1 = reinterpret_cast<const double2*>(st + TL::OFF_RS1);
        const int64_t base = tile * T;
#pragma unroll 1
        for (int q = tid; q < T / 2; q += AUG_BLOCK) {
            Obs o0, o1;
            lds_y2<yt>(st, q, o0.y, o1.y);
            o0.m = o1.m = o0.v = o1.v = 0.0;
            if (TL::NEED_MV) {
                const double2 m = smu[q], v = svar[q];
                o0.m = m.x; o1.m = m.y;
                o0.v = v.x; o1.v = v.y;
            }
            o0.mg = o1.mg = o0.vg = o1.vg = 0.0;
            if (TL::HAS_MUG) {
                const double2 mg = smug[q];
                o0.mg = mg.x; o1.mg = mg.y;
            }
            if (TL::HAS_VARG) {
                const double2 vg = svarg[q];
                o0.vg = vg.x; o1.vg = vg.y;
            }
            o0.ys = o0.y; o1.ys = o1.y;
            if (FS) {
                o0.s1 = o1.s1 = o0.s2 = o1.s2 = 0.0;
                const double2 c = srs0[q];
                o0.s0 = c.x; o1.s0 = c.y;
                if (TL::HAS_RS1) {
                    const double2 l = srs1[q];
                    o0.s1 = l.x; o1.s1 = l.y;
                }
                if (HET) {
                    const double2 ps = reinterpret_cast<const double2*>(st + TL::OFF_RS2)[q];
                    o0.s2 = ps.x; o1.s2 = ps.y;
                }
                if (YSTATE && has_rs2) lds_y2<int64_t>(st + TL::OFF_RS2, q, o0.ys, o1.ys);
            }
            bad = bad || !fast_ok<KIND, FS, ELBO>(o0) || !fast_ok<KIND, FS, ELBO>(o1);
            eval<KIND, FS, ELBO, false>(a.L, o0);
            eval<KIND, FS, ELBO, false>(a.L, o1);
            const int64_t i = base + 2 * q;
            if (!FS) {
                if (a.s0) st_stream2(a.s0 + i, o0.s0, o1.s0);
                if (HAS_S1 && a.s1) st_stream2(a.s1 + i, o0.s1, o1.s1);
                if (HET && a.s2) st_stream2(reinterpret_cast<double*>(a.s2) + i, o0.s2, o1.s2);
                if (YSTATE && a.s2) store_y2<s2t>(a.s2, i >> 1, o0.y, o1.y);
            }
            if (a.beta) st_stream2(a.beta + i, o0.b0, o1.b0);
            if (a.gamma) st_stream2(a.gamma + i, o0.g0, o1.g0);
            if (HET) {
                if (a.beta_g) st_stream2(a.beta_g + i, o0.b1, o1.b1);
                if (a.gamma_g) st_stream2(a.gamma_g + i, o0.g1, o1.g1);
            }
            if (ELBO) {
                acc[0] += o0.elt + o1.elt;
                acc[1] += o0.kl + o1.kl;
            }
        }
        __syncthreads();   // every thread is done reading stage s: it can be refilled
        if (tid == 0) {
            const int64_t nt = tile + (int64_t)S * stride;
            if (nt < ntiles) issue(nt, s);
        }
    }
    // one observation straight from global memory (any input): the redo path below and the ragged tail
    auto scalar_obs = [&](const int64_t i) {
        Obs o;
        o.y = load_y1<yt>(a.y, i);
        o.ys = o.y;
        o.m = o.v = o.mg = o.vg = 0.0;
        o.s0 = o.s1 = o.s2 = 0.0;
        if (TL::NEED_MV) { o.m = a.mu[i]; o.v = a.var[i]; }
        if (HET) { o.mg = a.mu_g[i]; if (TL::NEED_MV) o.vg = a.var_g[i]; }
        if (FS) {
            o.s0 = a.rs0[i];
            if (HAS_S1) o.s1 = a.rs1[i];
            if (HET) o.s2 = reinterpret_cast<const double*>(a.rs2)[i];
            if (YSTATE && has_rs2) o.ys = (double)reinterpret_cast<const int64_t*>(a.rs2)[i];
        }
        if (fast_ok<KIND, FS, ELBO>(o)) eval<KIND, FS, ELBO, false>(a.L, o);
        else eval<KIND, FS, ELBO, true>(a.L, o);
        if (!FS) {
            if (a.s0) a.s0[i] = o.s0;
            if (HAS_S1 && a.s1) a.s1[i] = o.s1;
            if (HET && a.s2) reinterpret_cast<double*>(a.s2)[i] = o.s2;
            if (YSTATE && a.s2) store_y1<s2t>(a.s2, i, o.y);
        }
        if (a.beta) a.beta[i] = o.b0;
        if (a.gamma) a.gamma[i] = o.g0;
        if (HET) {
            if (a.beta_g) a.beta_g[i] = o.b1;
            if (a.gamma_g) a.gamma_g[i] = o.g1;
        }
        if (ELBO) { acc[0] += o.elt; acc[1] += o.kl; }
    };
    if (bad) {
        // rare: this thread met a value outside the fast-math range -> redo its observations of every tile
        // of this CTA from global memory with the any-input instantiation (outputs are overwritten)
        acc[0] = acc[1] = 0.0;
        for (int64_t tile = first; tile < ntiles; tile += stride)
            for (int q = tid; q < T / 2; q += AUG_BLOCK) {
                scalar_obs(tile * T + 2 * q);
                scalar_obs(tile * T + 2 * q + 1);
            }
    }
    // the ragged tail (< one tile) rides in the same launch: the last CTA reads it directly (a.n = all observations)
    if (blockIdx.x == gridDim.x - 1)
        for (int64_t i = ntiles * T + tid; i < a.n; i += AUG_BLOCK) scalar_obs(i);
    if (ELBO) {
        double out[2];
        if (block_reduce_and_finalize<2>(acc, a.partials, a.counter, out)) write_scalars(a, out);
    }
}

template <int KIND, bool FS, bool ELBO>
int32_t launch_tma(aug_ctx* ctx, const CaviArgs& a) {
    typedef TileLayout<KIND, FS, ELBO> TL;
    const int64_t ntiles = a.n / TL::T;
    const void* k = (const void*)cavi_tma_kernel<KIND, FS, ELBO>;
    static bool configured_dev[64] = {false};   // the attribute is per device: one flag per ordinal
    bool& configured = configured_dev[ctx->device & 63];
    static int occ = 1;
    if (!configured) {
        AUG_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, TL::SMEM_BYTES));
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k, AUG_BLOCK, TL::SMEM_BYTES) != cudaSuccess || occ < 1)
            occ = 1;
        configured = true;
    }
    int64_t grid = (int64_t)ctx->sms * occ;
    if (grid > AUG_MAX_GRID) grid = AUG_MAX_GRID;
    if (grid > ntiles) grid = ntiles;
    cavi_tma_kernel<KIND, FS, ELBO><<<(unsigned)grid, AUG_BLOCK, TL::SMEM_BYTES, ctx->stream>>>(a, ntiles);
    ctx->launches++;
    return (int32_t)cudaGetLastError();
}

template <int KIND>
int32_t launch_tma1(aug_ctx* ctx, const CaviArgs& a, bool from_state, bool elbo) {
    if (from_state) return elbo ? launch_tma<KIND, true, true>(ctx, a) : launch_tma<KIND, true, false>(ctx, a);
    return elbo ? launch_tma<KIND, false, true>(ctx, a) : launch_tma<KIND, false, false>(ctx, a);
}

// the staged kernel wants at least four full tiles (largest tile of the kind's instantiations)
__host__ constexpr int64_t cavi_tma_min_n(int kind, bool from_state) { return from_state ? 4 * 4096 : 4 * (int64_t)cavi_tile(kind); }

template <int KIND, bool FROM_STATE, bool ELBO>
int32_t launch2(aug_ctx* ctx, const CaviArgs& a, bool vec) {
    const void* k = vec ? (const void*)cavi_kernel<KIND, FROM_STATE, ELBO, true>
                        : (const void*)cavi_kernel<KIND, FROM_STATE, ELBO, false>;
    const int64_t work = vec ? (a.n + 1) / 2 : a.n;
    const int grid = aug_grid_for(ctx, k, work, AUG_BLOCK * (KIND == AUG_HETERO ? 1 : CAVI_U));
    if (vec) cavi_kernel<KIND, FROM_STATE, ELBO, true><<<grid, AUG_BLOCK, 0, ctx->stream>>>(a);
    else cavi_kernel<KIND, FROM_STATE, ELBO, false><<<grid, AUG_BLOCK, 0, ctx->stream>>>(a);
    ctx->launches++;
    return (int32_t)cudaGetLastError();
}

template <int KIND>
int32_t launch1(aug_ctx* ctx, const CaviArgs& a, bool from_state, bool elbo, bool vec) {
    if (from_state) return elbo ? launch2<KIND, true, true>(ctx, a, vec) : launch2<KIND, true, false>(ctx, a, vec);
    return elbo ? launch2<KIND, false, true>(ctx, a, vec) : launch2<KIND, false, false>(ctx, a, vec);
}

bool getenv_no_tma() {   // AUGCUDA_NO_TMA=1 keeps every call on the direct-load kernel (A/B measurements)
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("AUGCUDA_NO_TMA");
        v = (e && e[0] == '1') ? 1 : 0;
    }
    return v == 1;
}

}  // namespace

// Shared by aug_cavi_step / aug_aux_posterior / aug_expected_potential_precision / aug_expected_elbo_terms
// for the non-categorical kinds (the categorical row kernels live in aug_cat.cu).
int32_t aug_cavi_dispatch(aug_ctx* ctx, const aug_lik* lik, int64_t n, const void* y, const double* mu,
                          const double* var, int64_t ld, void* s0, void* s1, void* s2, const void* rs0,
                          const void* rs1, const void* rs2, double* beta, double* gamma, int64_t ldo,
                          double* scalars, bool from_state) {
    if (n < 0) return AUG_ERR_BAD_ARG;
    if (scalars && aug_xch_for(ctx)) {       // exchanges happen in call order: complete a pending split-phase one first
        int32_t rf = aug_xch_flush(ctx);
        if (rf) return rf;
    }
    if (n == 0) {
        if (scalars) {
            AUG_CUDA(cudaMemsetAsync(scalars, 0, AUG_NSCALARS * sizeof(double), ctx->stream));
            if (aug_xch_for(ctx)) return aug_xch_zero_contribution(ctx, scalars, AUG_S_EXPECTED_LOGTILT, 2);
        }
        return AUG_OK;
    }
    const bool elbo = scalars != nullptr;
    const bool het = lik->kind == AUG_HETERO;
    if (y == nullptr) return AUG_ERR_BAD_ARG;
    if ((!from_state || elbo) && (mu == nullptr || var == nullptr)) return AUG_ERR_BAD_ARG;
    if (het && (mu == nullptr || ld < n)) return AUG_ERR_BAD_ARG;
    if (het && (beta || gamma) && ldo < n) return AUG_ERR_BAD_ARG;
    if (from_state && rs0 == nullptr) return AUG_ERR_BAD_ARG;
    if (from_state && (lik->kind == AUG_POISSON || het) && rs1 == nullptr) return AUG_ERR_BAD_ARG;
    if (from_state && het && rs2 == nullptr) return AUG_ERR_BAD_ARG;

    CaviArgs a{};
    a.n = n;
    a.y = y;
    a.mu = mu;
    a.var = var;
    a.mu_g = het && mu ? mu + ld : nullptr;
    a.var_g = het && var ? var + ld : nullptr;
    a.s0 = (double*)s0;
    a.s1 = (double*)s1;
    a.s2 = s2;
    a.rs0 = (const double*)rs0;
    a.rs1 = (const double*)rs1;
    a.rs2 = rs2;
    a.beta = beta;
    a.gamma = gamma;
    a.beta_g = het && beta ? beta + ldo : nullptr;
    a.gamma_g = het && gamma ? gamma + ldo : nullptr;
    a.partials = ctx->partials;
    a.counter = ctx->counter;
    a.scalars = scalars;
    int32_t rc = aug_lik_const(ctx, lik, &a.L, elbo, false);
    if (rc) return rc;
    // 128-bit path needs 16-byte aligned arrays (and, for HETERO, an even leading dimension)
    bool vec = aug_aligned16(mu) && aug_aligned16(var) && aug_aligned16(s0) && aug_aligned16(s1) &&
               aug_aligned16(s2) && aug_aligned16(rs0) && aug_aligned16(rs1) && aug_aligned16(rs2) &&
               aug_aligned16(beta) && aug_aligned16(gamma) && aug_aligned16(a.mu_g) && aug_aligned16(a.var_g) &&
               aug_aligned16(a.beta_g) && aug_aligned16(a.gamma_g);
    if (lik->kind == AUG_BERNOULLI) vec = vec && ((((uintptr_t)y) & 1u) == 0);
    else vec = vec && aug_aligned16(y);
    // Calls on 16-byte aligned arrays: ONE launch of the bulk-async staged kernel — the full tiles go through its
    // ring, its last CTA reads the ragged remainder (< one tile) directly.  Fused and from-state verbs alike.
    const bool use_tma = vec && !getenv_no_tma() && aug_aligned16(y) && n >= cavi_tma_min_n(lik->kind, from_state);
    if (elbo) a.xch = aug_xch_for(ctx);      // the verb's (only) launch carries the exchange
    if (a.xch && ctx->deferred) {
        a.xch_defer = 1;
        ctx->pending = 1;
    }
    if (use_tma) {
        switch (lik->kind) {
            case AUG_BERNOULLI: return launch_tma1<AUG_BERNOULLI>(ctx, a, from_state, elbo);
            case AUG_NEGBIN: return launch_tma1<AUG_NEGBIN>(ctx, a, from_state, elbo);
            case AUG_POISSON: return launch_tma1<AUG_POISSON>(ctx, a, from_state, elbo);
            case AUG_LAPLACE: return launch_tma1<AUG_LAPLACE>(ctx, a, from_state, elbo);
            case AUG_STUDENTT: return launch_tma1<AUG_STUDENTT>(ctx, a, from_state, elbo);
            case AUG_HETERO: return launch_tma1<AUG_HETERO>(ctx, a, from_state, elbo);
            default: return AUG_ERR_BAD_KIND;
        }
    }
    switch (lik->kind) {
        case AUG_BERNOULLI: rc = launch1<AUG_BERNOULLI>(ctx, a, from_state, elbo, vec); break;
        case AUG_NEGBIN: rc = launch1<AUG_NEGBIN>(ctx, a, from_state, elbo, vec); break;
        case AUG_POISSON: rc = launch1<AUG_POISSON>(ctx, a, from_state, elbo, vec); break;
        case AUG_LAPLACE: rc = launch1<AUG_LAPLACE>(ctx, a, from_state, elbo, vec); break;
        case AUG_STUDENTT: rc = launch1<AUG_STUDENTT>(ctx, a, from_state, elbo, vec); break;
        case AUG_HETERO: rc = launch1<AUG_HETERO>(ctx, a, from_state, elbo, vec); break;
        default: return AUG_ERR_BAD_KIND;
    }
    if (rc) return rc;
    return AUG_OK;
}
