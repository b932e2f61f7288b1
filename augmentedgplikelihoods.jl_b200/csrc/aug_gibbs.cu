// aug_gibbs.cu — Gibbs side: aux_sample!, init_aux_variables, the raw PG sampler, and the sampled
// (non-expected) potential / precision map for the scalar-latent likelihoods and HETERO.
//
// Reference behaviour replaced (paths relative to /root/reference/src):
//   generic.jl:1-20,32-34 (aux_sample!, aux_sample, init_aux_variables),
//   likelihoods/*.jl aux_full_conditional + auglik_potential/precision,
//   SpecialDistributions/polyagamma.jl:112-257, polyagammapoisson.jl:23-27.
// One thread owns one observation and one Philox stream positioned by the GLOBAL observation
// index, so the draws are identical for any sharding of the observation axis.
#include <stdlib.h>

#include "aug_common.cuh"
#include "aug_math.cuh"
#include "aug_pg.cuh"
#include "aug_pgb.cuh"

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
            st_stream1(a.omega + i, augb::pg_draw_stream(g, 1.0, true, f, a.L.pgtab));
        } else if (KIND == AUG_NEGBIN) {                          // PG(y + r, |f|)  negativebinomial.jl:20-22
            const double y = (double)__ldg(reinterpret_cast<const int64_t*>(a.y) + i);
            st_stream1(a.omega + i, augb::pg_draw_stream(g, y + a.L.p0, a.L.r_is_int != 0, f, a.L.pgtab));
        } else if (KIND == AUG_POISSON) {                         // poisson.jl:26-28, polyagammapoisson.jl:23-27
            const int64_t y = __ldg(reinterpret_cast<const int64_t*>(a.y) + i);
            const int64_t nn = augr::poisson_rand(g, a.L.p0 * augm::logistic(-f));
            a.nvar[i] = nn;
            st_stream1(a.omega + i, augb::pg_draw_stream(g, (double)(nn + y), true, f, a.L.pgtab));
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
            st_stream1(a.omega + i, augb::pg_draw_stream(g, (double)nn + 0.5, false, gg, a.L.pgtab));
        }
    }
}

// ------------------------------------------------------------------ PG(1, c): warp-compacted sampler
// (design notes in aug_pg.cuh).  A warp walks chunks of 32 consecutive elements; a FRESH step starts one round for
// each lane's element — the exponential-branch lanes (~58%) finish in it — and everything that needs more work
// (first truncated-IG attempt, rejected attempts, the rare rejected round) is pushed onto the warp's queue; whenever
// the queue holds >= 32 items a full warp pops them and performs one attempt each.  Queue item: element offset in
// the shard (32 bits), round | attempt << 8, the round's accept uniform, z.
struct Pg1Args {
    int64_t n, i0;
    uint64_t seed, offset;
    const double* c;     // tilt per element (aux_sample!: f), or nullptr -> cs
    double cs;
    double* out;
    const double* tab;
    AugXchDev* gx;       // non-null: one extra CTA completes a pending split-phase exchange (aug_comm_set_deferred)
    augr::PhiloxKeys keys;   // the ten round keys of (seed, offset): constant-bank operands of the Philox rounds
};

// counters exhausted (probability < 1e-14 per draw): finish one draw on a private sequential stream
__device__ __noinline__ double pg1_finish_sequential(uint64_t seed, uint64_t offset, uint64_t gi, double z, const double* tab) {
    augr::Philox g;
    g.init(seed, offset, gi, 3u);
    const augp::PG1 s = augp::pg1_setup(2.0 * z, tab);
    return augp::pg1_draw(g, s);
}

#define PG1_QCAP 64
#define PG1_MAXCTR 255u

// CTA shape (A/B on B200, gpurun_out/ab_pg1.txt, ms per 1e8 draws): 256 x 3 CTAs/SM at 80 registers (24-byte spill) 1.57;
// 256 x 2 (106 registers) 1.54; 320 x 2 (96 registers, no spill) 1.48; 128 x 5 1.52; 608 x 1 1.46; ONE CTA of 640 threads per SM
// (20 warps, 96 registers, no spill) 1.40.  The alternating-series tail of pg1_accept is out of line (aug_pg.cuh): in line it
// costs the fresh step 10 registers (1.54 -> 1.48 at 320 x 2).
#ifndef PG1_BLOCK
#define PG1_BLOCK 640
#endif
#ifndef PG1_MIN_BLOCKS
#define PG1_MIN_BLOCKS 1
#endif
__global__ void __launch_bounds__(PG1_BLOCK, PG1_MIN_BLOCKS) pg1_compact_kernel(const Pg1Args a) {
    __shared__ __align__(16) double tab_s[AUG_PGTAB_N * AUG_PGTAB_DEG];   // r(z) table, coefficient-major: 10 KB, read by every fresh step
    __shared__ uint32_t qel_s[PG1_BLOCK / 32][PG1_QCAP];
    __shared__ uint32_t qra_s[PG1_BLOCK / 32][PG1_QCAP];
    __shared__ uint32_t quacc_s[PG1_BLOCK / 32][PG1_QCAP];
    __shared__ double qz_s[PG1_BLOCK / 32][PG1_QCAP];
    // the extra CTA of a launch that carries a deferred gather: it is scheduled when a working CTA retires, by when the
    // peers have long published, so the wait for the slowest rank costs this kernel (next to) nothing
    const unsigned nwork = a.gx ? gridDim.x - 1 : gridDim.x;
    if (blockIdx.x == nwork) {
        if (threadIdx.x == 0) xch_finish_pending(a.gx);
        return;
    }
    augp::pg1_load_table_cm(tab_s, a.tab, PG1_BLOCK);
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t* qel = qel_s[warp];
    uint32_t* qra = qra_s[warp];
    uint32_t* quacc = quacc_s[warp];
    double* qz = qz_s[warp];
    const uint32_t lt_mask = (1u << lane) - 1u;
    const uint32_t nchunks = (uint32_t)((a.n + 31) >> 5);                 // n < 2^32 (host-checked)
    const uint32_t W = nwork * (PG1_BLOCK / 32);
    const uint32_t k0 = (uint32_t)a.seed, k1 = (uint32_t)(a.seed >> 32) ^ (uint32_t)(a.offset >> 32);
    const uint32_t c3 = (uint32_t)a.offset;
    int qn = 0;

    // push (warp-uniform call): lanes with `want` append their item
    auto push = [&](bool want, uint32_t el, uint32_t ra, uint32_t uacc, double z) {
        const uint32_t m = __ballot_sync(0xffffffffu, want);
        if (want) {
            const int pos = qn + __popc(m & lt_mask);
            qel[pos] = el;
            qra[pos] = ra;
            quacc[pos] = uacc;
            qz[pos] = z;
        }
        qn += __popc(m);
        __syncwarp();
    };

    uint32_t ch = blockIdx.x * (PG1_BLOCK / 32) + warp;
    // the tilt of the NEXT fresh chunk is loaded one step ahead (its latency hides behind the work steps)
    auto load_c = [&](uint32_t chunk) {
        const uint32_t e = (chunk << 5) + lane;
        return (chunk < nchunks && e < a.n) ? (a.c ? ld_stream1(a.c + e) : a.cs) : 0.0;
    };
    double c_next = load_c(ch);
    for (;;) {
        if (qn < 32 && ch < nchunks) {
            // ---- fresh step: round 0 of 32 consecutive elements
            const uint32_t el = (ch << 5) + lane;
            const bool valid = el < a.n;
            const double c = c_next;
            c_next = load_c(ch + W);
            const augp::PG1 s = augp::pg1_setup_cm(c, tab_s);
            const uint64_t gi = (uint64_t)a.i0 + el;
            const uint32_t e_lo = (uint32_t)gi, e_hi = (uint32_t)(gi >> 32);
            uint32_t w[4];
            AUG_PHILOX_RK(a.keys, e_lo, e_hi, augp::pg1_ctr(0u, 0u, 0u), c3, w);
            const double u0 = augr::u32_mid(w[0]);
            const bool exp_branch = u0 < s.r;                                      // mass_texpon, polyagamma.jl:179-192
            const double E = -augf::log_(augr::u53_open0(w[1], w[2]));
            // exponential branch: x = t + E/K.  The other branch (truncated inverse Gaussian on (0, t]) makes its FIRST
            // attempt right here when mu = 1/z > t: the same E proposes x = t/(1 + tE)^2, and the part of the selector
            // uniform above r, (u0 - r)/(1 - r) ~ U(0,1) given u0 >= r, decides both of its rejection tests at once
            // (aug_pg.cuh: trunc_ig_small_z).  72% of those lanes finish here instead of going through the queue.
            double a_ig;
            const double x_ig = augp::trunc_ig_small_z(E, s.z, a_ig);
            const bool ig_ok = s.z < 1.0 / augp::T && u0 <= fma(1.0 - s.r, augf::exp_(-fmin(a_ig, 700.0)), s.r);
            const double x = exp_branch ? fma(E, s.invK, augp::T) : x_ig;
            bool again = valid && !exp_branch && !ig_ok;
            uint32_t ra = 1u << 8;                                                 // round 0, first queued IG attempt
            if (valid && (exp_branch || ig_ok)) {
                if (augp::pg1_accept(x, w[3], k0, k1, e_lo, e_hi, c3, 0u)) {
                    st_stream1(a.out + el, 0.25 * x);
                } else {
                    again = true;
                    ra = 1u;                                                       // round 1, attempt 0
                }
            }
            push(again, el, ra, w[3], s.z);
            ch += W;
            continue;
        }
        if (qn == 0) break;
        // ---- work step: one attempt for each of the last min(qn, 32) queued items
        const int cnt = qn < 32 ? qn : 32;
        const bool active = lane < cnt;
        uint32_t el = 0, ra = 0, uacc = 0;
        double z = 0.0;
        if (active) {
            const int idx = qn - cnt + lane;
            el = qel[idx];
            ra = qra[idx];
            uacc = quacc[idx];
            z = qz[idx];
        }
        __syncwarp();
        qn -= cnt;
        const uint32_t round = ra & 0xffu, attempt = ra >> 8;
        const uint64_t gi = (uint64_t)a.i0 + el;
        const uint32_t e_lo = (uint32_t)gi, e_hi = (uint32_t)(gi >> 32);
        bool again = false;
        if (active) {
            double x = -1.0;
            if (attempt != 0 && attempt < PG1_MAXCTR && round < PG1_MAXCTR) {
                uint32_t w[4];
                AUG_PHILOX_RK(a.keys, e_lo, e_hi, augp::pg1_ctr(1u, round, attempt), c3, w);
                x = augp::trunc_ig_attempt_w(w, z);
                if (x < 0.0) {
                    again = true;
                    ra += 1u << 8;                                                 // next attempt of this round
                }
            } else if (attempt == 0 && round < PG1_MAXCTR) {
                // rare: a new round after a rejected proposal
                const augp::PG1 s = augp::pg1_setup_cm(2.0 * z, tab_s);
                uint32_t w[4];
                augr::philox4x32_10(k0, k1, e_lo, e_hi, augp::pg1_ctr(0u, round, 0u), c3, w);
                uacc = w[3];
                if (augr::u32_mid(w[0]) < s.r) {
                    x = fma(-augf::log_(augr::u53_open0(w[1], w[2])), s.invK, augp::T);
                } else {
                    again = true;
                    ra |= 1u << 8;                                                 // first IG attempt
                }
            } else {
                st_stream1(a.out + el, pg1_finish_sequential(a.seed, a.offset, gi, z, a.tab));
            }
            if (x > 0.0) {
                if (augp::pg1_accept(x, uacc, k0, k1, e_lo, e_hi, c3, round)) {
                    st_stream1(a.out + el, 0.25 * x);
                } else {
                    again = true;
                    ra = round + 1u;                                               // new round, attempt 0
                }
            }
        }
        push(again, el, ra, uacc, z);
    }
}


// ------------------------------------------------------------------ PG(b, c), general b: warp-compacted sampler
// aux_sample! of the NegBin / Poisson / heteroscedastic likelihoods (negativebinomial.jl:20-22, poisson.jl:26-28 +
// polyagammapoisson.jl:23-27, heteroscedasticgaussian.jl:28-32) and the raw rand(PolyaGamma(b, c)).  Design notes
// and the three pieces of the law (Devroye sum / exact fractional piece / certified Gamma convolution) in
// aug_pgb.cuh.  Like pg1_compact_kernel, a draw is a sequence of uniform STEPS; what cannot finish in the
// straight-line FRESH step of its element is parked in one of two per-warp shared-memory queues — QG: Marsaglia-Tsang
// retries and extra terms of the convolution, QX: Devroye rounds / truncated-IG attempts / fractional-piece attempts —
// and a work step pops up to 32 items of ONE queue, so all lanes of a step run the same code.  An element has at
// most one item in flight and its partial sum travels with the item, so the additions happen in a fixed order:
// the output depends on (seed, offset, global element index) only.
enum { PGB_RAW = 100 };
struct PgbArgs {
    int64_t n, i0;
    uint64_t seed, offset;
    const void* y;
    const double* f;
    const double* g;      // HETERO second latent
    const double* b;      // RAW: per-element b (or nullptr -> bs)
    const double* c;      // RAW: per-element c (or nullptr -> cs)
    double bs, cs;
    int b_is_int;
    double p0;            // r | lambda
    double* omega;
    int64_t* nvar;
    const double* tab;
    AugXchDev* gx;        // non-null: one extra CTA completes a pending split-phase exchange
    augr::PhiloxKeys keys;   // round keys of (seed, offset): constant-bank operands of the Philox rounds
};

#define PGB_T_DEV 0u
#define PGB_T_GAM 1u
#define PGB_T_FRAC 2u
#define PGB_QCAP 64
// CTA shape (A/B on B200, gpurun_out/ab_pgb.txt, ab_pgb2.txt; NegBin / Poisson / Hetero ms per 1e8): 256 x 2 CTAs/SM at 128
// registers 3.35 / 5.44 / 7.6; 512 x 1 3.27 / 5.33 / 7.4; 320 x 2 (96 registers) 3.28 / 5.13 / 7.1; 640 x 1 (96) 3.10 / 4.87 / 6.8;
// ONE CTA of 768 threads per SM (24 warps, 80 registers) 2.95 / 4.71 / 6.6: the kernels wait on fixed-latency dependencies, more
// warps hide them better than more registers, and one large CTA beats two of half the size at equal warps and registers.
#ifndef PGB_BLOCK
#define PGB_BLOCK 768
#endif
#define PGB_WARPS (PGB_BLOCK / 32)
// dynamic shared memory: r(z) table | per warp 3 queues (Devroye, Gamma, fractional) x { lo[QCAP], hi[QCAP] } double2
#define PGB_SMEM_BYTES (AUG_PGTAB_N * AUG_PGTAB_DEG * 8 + PGB_WARPS * 3 * 2 * PGB_QCAP * 16)

#ifndef PGB_MIN_BLOCKS
#define PGB_MIN_BLOCKS 1
#endif
template <int KIND>
__global__ void __launch_bounds__(PGB_BLOCK, PGB_MIN_BLOCKS) pgb_kernel(const PgbArgs a) {
    extern __shared__ __align__(16) unsigned char pgb_smem[];
    const unsigned nwork = a.gx ? gridDim.x - 1 : gridDim.x;
    if (blockIdx.x == nwork) {                 // the extra CTA of a launch that carries a deferred gather
        if (threadIdx.x == 0) xch_finish_pending(a.gx);
        return;
    }
    double* tab_s = reinterpret_cast<double*>(pgb_smem);
    augp::pg1_load_table_cm(tab_s, a.tab, PGB_BLOCK);
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double2* qbase = reinterpret_cast<double2*>(pgb_smem + AUG_PGTAB_N * AUG_PGTAB_DEG * 8) + (size_t)warp * 3 * 2 * PGB_QCAP;
    double2 *qd_lo = qbase, *qd_hi = qbase + PGB_QCAP;                       // Devroye rounds / truncated-IG attempts
    double2 *qg_lo = qbase + 2 * PGB_QCAP, *qg_hi = qbase + 3 * PGB_QCAP;    // Marsaglia-Tsang attempts of the convolution
    double2 *qf_lo = qbase + 4 * PGB_QCAP, *qf_hi = qbase + 5 * PGB_QCAP;    // fractional-piece attempts
    const uint32_t lt_mask = (1u << lane) - 1u;
    const uint32_t nchunks = (uint32_t)((a.n + 31) >> 5);
    const uint32_t W = nwork * PGB_WARPS;
    augb::Key key;
    key.k0 = (uint32_t)a.seed;
    key.k1 = (uint32_t)(a.seed >> 32) ^ (uint32_t)(a.offset >> 32);
    key.c3 = (uint32_t)a.offset;
    constexpr double SCALE = 0.5 / (augp::PI * augp::PI);
    // b <= BX: exact pieces.  HETERO (b = n + 1/2, n mostly 0..3) keeps ALL its common cases on the exact pieces and
    // starts no convolution in the fresh step: its hot code (fresh + fractional + Devroye steps) then fits the 32 KB
    // instruction cache — with the convolution in line ncu showed 44% of the stall samples on instruction fetch.
    constexpr double BX = KIND == AUG_HETERO ? 12.75 : PGB_BX;
    constexpr bool INLINE_CONV = KIND != AUG_HETERO;
    int nd = 0, ng = 0, nf = 0;

    auto push = [&](double2* lo, double2* hi, int& qn, bool want, uint32_t el, uint32_t st, double acc, double b, double c) {
        const uint32_t m = __ballot_sync(0xffffffffu, want);
        if (want) {
            const int pos = qn + __popc(m & lt_mask);
            lo[pos] = make_double2(__hiloint2double((int)st, (int)el), acc);
            hi[pos] = make_double2(b, c);
        }
        qn += __popc(m);
        __syncwarp();
    };
    auto pop = [&](double2* lo, double2* hi, int& qn, bool& active, uint32_t& el, uint32_t& st, double& acc, double& b, double& c) {
        const int cnt = qn < 32 ? qn : 32;
        active = lane < cnt;
        if (active) {
            const double2 l = lo[qn - cnt + lane], h = hi[qn - cnt + lane];
            el = (uint32_t)__double2loint(l.x);
            st = (uint32_t)__double2hiint(l.x);
            acc = l.y;
            b = h.x;
            c = h.y;
        }
        __syncwarp();
        qn -= cnt;
    };
    auto finish = [&](uint32_t el, double v) { st_stream1(a.omega + el, v); };

    uint32_t ch = blockIdx.x * PGB_WARPS + warp;
    for (;;) {
        int mode;                                   // 0 fresh, 1 Devroye queue, 2 gamma queue, 3 fractional queue
        // a fractional step feeds the Devroye queue: it may only run while that queue has room for 32 more items
        if (nd >= 32) mode = 1;
        else if (nf >= 32) mode = 3;
        else if (ng >= 32) mode = 2;
        else if (ch < nchunks) mode = 0;
        else if (nf > 0) mode = 3;
        else if (nd > 0) mode = 1;
        else if (ng > 0) mode = 2;
        else break;

        if (mode == 0) {
            // ---------------------------------------------------------------- fresh step: 32 consecutive elements
            const uint32_t el = (ch << 5) + lane;
            ch += W;
            const bool valid = el < a.n;
            const uint64_t gi = (uint64_t)a.i0 + el;
            const uint32_t e_lo = (uint32_t)gi, e_hi = (uint32_t)(gi >> 32);
            double b = 0.0, c = 0.0;
            bool isint = true;
            if (valid) {
                if (KIND == AUG_NEGBIN) {                               // PG(y + r, |f|)  negativebinomial.jl:20-22
                    c = ld_stream1(a.f + el);
                    b = (double)__ldg(reinterpret_cast<const int64_t*>(a.y) + el) + a.p0;
                    isint = a.b_is_int != 0;
                } else if (KIND == AUG_POISSON) {                       // poisson.jl:26-28, polyagammapoisson.jl:23-27
                    c = ld_stream1(a.f + el);
                    const int64_t y = __ldg(reinterpret_cast<const int64_t*>(a.y) + el);
                    const int64_t nn = augb::poisson_draw(a.keys, key, e_lo, e_hi, a.seed, a.offset, gi, a.p0 * augb::logistic_fast(-c));
                    a.nvar[el] = nn;
                    b = (double)(nn + y);
                } else if (KIND == AUG_HETERO) {                        // heteroscedasticgaussian.jl:28-32
                    const double f = ld_stream1(a.f + el);
                    c = ld_stream1(a.g + el);
                    const double d = f - ld_stream1(reinterpret_cast<const double*>(a.y) + el);
                    const int64_t nn = augb::poisson_draw(a.keys, key, e_lo, e_hi, a.seed, a.offset, gi, a.p0 * augb::logistic_fast(-c) * d * d * 0.5);
                    a.nvar[el] = nn;
                    b = (double)nn + 0.5;
                    isint = false;
                } else {                                                // rand(PolyaGamma(b, c))
                    b = a.b ? __ldg(a.b + el) : a.bs;
                    c = a.c ? __ldg(a.c + el) : a.cs;
                    isint = a.b_is_int != 0;
                }
            }
            if (isint) b = rint(b);
            bool live = valid;
            if (live && !(b > 1e-250)) {                                // Dirac at 0  polyagamma.jl:122-124 (and b below 1e-250)
                finish(el, 0.0);
                live = false;
            }
            const bool conv = live && b > BX;
            uint32_t gst = 0;
            double gacc = 0.0;
            bool gpush = false;
            if (INLINE_CONV) {
                if (__any_sync(0xffffffffu, conv)) {
                    if (conv) {
                        // first attempt of the two leading terms (one Box-Muller pair) and of the tail in line
                        const augb::Conv s = augb::conv_setup(b, c);
                        uint32_t w1[4], w0[4];
                        AUG_PHILOX_RK(a.keys, e_lo, e_hi, augb::ctr(3u, 1u, 0u, 0u), key.c3, w1);
                        AUG_PHILOX_RK(a.keys, e_lo, e_hi, augb::ctr(3u, 0u, 0u, 0u), key.c3, w0);
                        double v1, v2;
                        augb::gamma_pair_attempt(w1, b, v1, v2);
                        const double v0 = augb::gamma_attempt(w0, s.shape);
                        uint32_t mask = 0;
                        gacc = s.loc;
                        if (v1 >= 0.0) gacc = fma(v1, augf::rcp(0.25 + s.w), gacc); else mask |= 1u;
                        if (v2 >= 0.0) gacc = fma(v2, augf::rcp(2.25 + s.w), gacc); else mask |= 2u;
                        if (v0 >= 0.0) gacc = fma(v0, s.theta, gacc); else mask |= 4u;
                        const uint32_t extra = s.kt > 2 ? 3u : 0u;
                        if (mask == 0u && extra == 0u) {
                            finish(el, gacc * SCALE);
                        } else {
                            gpush = true;
                            gst = PGB_T_GAM | (mask << 2) | (extra << 5) | ((uint32_t)s.kt << 11) | ((mask ? 1u : 0u) << 17);
                        }
                    }
                }
            } else if (conv) {
                // every term is drawn by the gamma steps, one attempt per visit; the tail visit adds loc (bit 31)
                const uint32_t kt = (uint32_t)augb::conv_kt(fabs(c));
                gpush = true;
                gst = PGB_T_GAM | (7u << 2) | ((kt > 2u ? 3u : 0u) << 5) | (kt << 11) | 0x80000000u;
            }
            push(qg_lo, qg_hi, ng, gpush, el, gst, gacc, b, c);
            // exact pieces: fractional part first (if any), then floor(b) Devroye draws
            const bool exact = live && !conv;
            const double fl = floor(b);
            double e = b - fl;
            if (e < 1e-250) e = 0.0;
            const uint32_t rem = (uint32_t)fl;
            push(qf_lo, qf_hi, nf, exact && e > 0.0, el, PGB_T_FRAC | (rem << 2), 0.0, e, c);   // rem <= 12: 4 bits
            push(qd_lo, qd_hi, nd, exact && e == 0.0, el, PGB_T_DEV | (rem << 2), 0.0, 0.0, c);
            continue;
        }

        if (mode == 2) {
            // ---------------------------------------------------------------- gamma step: one Marsaglia-Tsang attempt per item
            bool active;
            uint32_t el = 0, st = 0;
            double acc = 0.0, b = 1.0, c = 0.0;
            pop(qg_lo, qg_hi, ng, active, el, st, acc, b, c);
            bool again = false;
            if (active) {
                uint32_t mask = (st >> 2) & 7u, extra = (st >> 5) & 63u, att = (st >> 17) & 0x3fffu;
                const uint32_t kt = (st >> 11) & 63u, locflag = st & 0x80000000u;
                const uint32_t k = mask ? ((mask & 1u) ? 1u : ((mask & 2u) ? 2u : 0u)) : extra;
                const uint64_t gi = (uint64_t)a.i0 + el;
                const double xp = 0.5 * fabs(c) * (1.0 / augp::PI);
                double shape = b, wt;
                double loc = 0.0;
                if (k == 0u) {
                    const augb::Conv s = augb::conv_setup(b, c);
                    shape = s.shape;
                    wt = s.theta;
                    if (locflag) loc = s.loc;
                } else {
                    const double km = (double)k - 0.5;
                    wt = augf::rcp(fma(km, km, xp * xp));
                }
                uint32_t w[4];
                AUG_PHILOX_RK(a.keys, (uint32_t)gi, (uint32_t)(gi >> 32), augb::ctr(3u, k, 0u, att), key.c3, w);
                const double v = augb::gamma_attempt(w, shape);
                again = true;
                if (v >= 0.0) {
                    acc = fma(v, wt, acc) + loc;
                    if (mask) {
                        mask &= mask - 1u;                         // clear the lowest pending bit (k = 1, then 2, then tail)
                        att = mask ? 1u : 0u;
                    } else {
                        extra = extra < kt ? extra + 1u : 0u;
                        att = 0u;
                    }
                    if (mask == 0u && extra == 0u) {
                        finish(el, acc * SCALE);
                        again = false;
                    }
                } else if (++att > PGB_MAXATT) {
                    finish(el, augb::pgb_sequential(a.seed, a.offset, gi, b, false, c, a.tab));
                    again = false;
                }
                st = PGB_T_GAM | (mask << 2) | (extra << 5) | (kt << 11) | (att << 17) | locflag;
            }
            push(qg_lo, qg_hi, ng, again, el, st, acc, b, c);
            continue;
        }

        if (mode == 3) {
            // ---------------------------------------------------------------- fractional step: one IG proposal + series test per item
            bool active;
            uint32_t el = 0, st = 0;
            double acc = 0.0, e = 0.5, c = 0.0;
            pop(qf_lo, qf_hi, nf, active, el, st, acc, e, c);
            bool again = false, to_dev = false;
            const uint32_t rem = (st >> 2) & 15u;
            if (active) {
                const uint64_t gi = (uint64_t)a.i0 + el;
                const uint32_t e_lo = (uint32_t)gi, e_hi = (uint32_t)(gi >> 32);
                uint32_t att = (st >> 18) & 0x3fffu;
                uint32_t w1[4];
                AUG_PHILOX_RK(a.keys, e_lo, e_hi, augb::ctr(4u, 0u, 0u, att), key.c3, w1);
                double uacc;
                const double x = augb::frac_propose1(w1, e, 0.5 * fabs(c), uacc);
                if (augb::frac_accept(x, e, uacc)) {
                    acc = 0.25 * x;
                    if (rem == 0u) finish(el, acc);
                    else to_dev = true;
                } else if (++att > PGB_MAXATT) {
                    finish(el, augb::pgb_sequential(a.seed, a.offset, gi, e + (double)rem, false, c, a.tab));
                } else {
                    again = true;
                    st = PGB_T_FRAC | (rem << 2) | (att << 18);
                }
            }
            push(qf_lo, qf_hi, nf, again, el, st, acc, e, c);
            push(qd_lo, qd_hi, nd, to_dev, el, PGB_T_DEV | (rem << 2), acc, 0.0, c);
            continue;
        }

        // -------------------------------------------------------------------- Devroye step: round start or one truncated-IG attempt
        bool active;
        uint32_t el = 0, st = 0;
        double acc = 0.0, bb = 0.0, c = 0.0;
        pop(qd_lo, qd_hi, nd, active, el, st, acc, bb, c);          // bb: the round's accept uniform in the low word
        bool again = false;
        if (active) {
            const uint64_t gi = (uint64_t)a.i0 + el;
            const uint32_t e_lo = (uint32_t)gi, e_hi = (uint32_t)(gi >> 32);
            const double z = 0.5 * fabs(c);
            uint32_t rem = (st >> 2) & 15u, sub = (st >> 6) & 15u, round = (st >> 10) & 0xffu, att = (st >> 18) & 0x3fffu;
            uint32_t uacc = (uint32_t)__double2loint(bb);
            double x = -1.0;
            bool exhausted = false;
            again = true;
            {
                // ONE code path for a round start (att == 0; sample_pg1, polyagamma.jl:225-257) and for a retry of the truncated
                // inverse-Gaussian proposal (att >= 1): both draw E ~ Exp(1) and a decision uniform u0 from one Philox block and
                // evaluate the same straight-line formulas — the exponential proposal t + E/K, the inverse-Gaussian proposal
                // t/(1 + tE)^2 with its merged acceptance exp(-a) (aug_pg.cuh: trunc_ig_small_z).  A round start picks the
                // exponential branch when u0 < r and otherwise uses the part of u0 above r as the IG decision; a retry is
                // already committed to the IG branch and uses u0 itself.  The lanes of a step differ by selects only.
                const bool retry = att != 0u;
                const augp::PG1 s = augp::pg1_setup_cm(c, tab_s);
                uint32_t w[4];
                AUG_PHILOX_RK(a.keys, e_lo, e_hi, augb::ctr(retry ? 1u : 0u, sub, round, att), key.c3, w);
                if (!retry) uacc = w[3];
                const double u0 = augr::u32_mid(w[0]);
                const double E = -augf::log_(augr::u53_open0(w[1], w[2]));
                double a_ig;
                const double x_ig = augp::trunc_ig_small_z(E, z, a_ig);
                const double p_ig = augf::exp_(-fmin(a_ig, 700.0));
                const bool small_z = z < 1.0 / augp::T;
                if (!retry && u0 < s.r) {
                    x = fma(E, s.invK, augp::T);
                } else if (small_z) {
                    if (u0 <= (retry ? p_ig : fma(1.0 - s.r, p_ig, s.r))) x = x_ig;
                    else if (++att > PGB_MAXATT) exhausted = true;
                } else if (!retry) {
                    att = 1u;
                } else {                                               // mu = 1/z <= t: the other proposal (rare: |c| > 3.1)
                    x = augp::trunc_ig_attempt_w(w, z);
                    if (x < 0.0 && ++att > PGB_MAXATT) exhausted = true;
                }
            }
            if (x > 0.0) {
                if (augb::dev_accept(x, uacc, key, e_lo, e_hi, sub, round)) {
                    acc += 0.25 * x;
                    rem -= 1u;
                    sub += 1u;
                    round = 0u;
                    att = 0u;
                    if (rem == 0u) {
                        finish(el, acc);
                        again = false;
                    }
                } else {
                    att = 0u;
                    if (++round > PGB_MAXROUND) exhausted = true;
                }
            }
            if (exhausted) {           // finish the remaining draws on a private sequential stream (probability < 1e-70)
                finish(el, acc + augb::pgb_sequential(a.seed, a.offset ^ 0x5bd1e995u, gi, (double)rem, true, c, a.tab));
                again = false;
            }
            st = PGB_T_DEV | (rem << 2) | (sub << 6) | (round << 10) | (att << 18);
            bb = __hiloint2double(0, (int)uacc);
        }
        push(qd_lo, qd_hi, nd, again, el, st, acc, bb, c);
    }
}

// ------------------------------------------------------------------ Laplace / StudentT: one Philox block per draw, compacted retries
// laplace.jl:40-42   omega ~ InverseGaussian(mu = 1/(2 beta |y - f|), 2 lambda): Michael-Schucany-Haas out of ONE block, straight-line
//                    (stable form x1 = 4 mu lam w / (w + sqrt(w (4 lam + w)))^2, w = mu N^2; other root mu^2 / x1);
// studentt.jl:46-48  omega ~ Gamma(alpha, 2/(nu/sigma^2 + (y - f)^2)): one Marsaglia-Tsang attempt per block (alpha >= 1); the 4 % of
//                    rejected attempts wait in a per-warp queue and are retried 32 at a time (the per-thread loop of
//                    aux_sample_kernel made 3 of 4 warps run a second attempt for one or two lanes).
// Draws depend on (seed, offset, global index) only: block counter = attempt number.
struct MapSampleArgs {
    int64_t n, i0;
    uint64_t offset;
    const double* y;
    const double* f;
    double* omega;
    double c0, c1;
    augr::PhiloxKeys keys;
};
#define MS_QCAP 64
// CTA shape (A/B on B200, gpurun_out/ab_ms.txt; Laplace / StudentT ms per 1e8 draws): 256 x 4 CTAs/SM 1.15 / 1.59; 512 x 2
// 1.10 / 1.50; 1024 x 1 1.07 / 1.39; ONE CTA of 768 threads per SM (71 registers for StudentT) 1.06 / 1.38
#ifndef MS_BLOCK
#define MS_BLOCK 768
#endif
#ifndef MS_MIN_BLOCKS
#define MS_MIN_BLOCKS 1
#endif
template <int KIND>
__global__ void __launch_bounds__(MS_BLOCK, MS_MIN_BLOCKS) map_sample_kernel(const MapSampleArgs a) {
    __shared__ uint32_t qel_s[MS_BLOCK / 32][MS_QCAP];
    __shared__ uint32_t qat_s[MS_BLOCK / 32][MS_QCAP];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t* qel = qel_s[warp];
    uint32_t* qat = qat_s[warp];
    const uint32_t lt_mask = (1u << lane) - 1u;
    const uint32_t nchunks = (uint32_t)((a.n + 31) >> 5);                 // n < 2^32 (host-checked)
    const uint32_t W = gridDim.x * (MS_BLOCK / 32);
    const uint32_t c3 = (uint32_t)a.offset;
    int qn = 0;
    // one attempt for element el (attempt number att): true = omega[el] written
    auto attempt = [&](uint32_t el, uint32_t att, bool valid) {
        const uint64_t gi = (uint64_t)a.i0 + el;
        uint32_t w[4];
        AUG_PHILOX_RK(a.keys, (uint32_t)gi, (uint32_t)(gi >> 32), att, c3, w);
        const double y = valid ? ld_stream1(a.y + el) : 1.0, f = valid ? ld_stream1(a.f + el) : 0.0;
        const double d = y - f;
        if (KIND == AUG_LAPLACE) {
            const double lam = 2.0 * a.c1;
            const double ad = fabs(d);
            const double mu = ad >= 1e-290 ? a.c0 * augf::rcp(fmin(ad, 1e290)) : a.c0 / ad;
            const double rad2 = -2.0 * augf::log_(augr::u53_open0(w[0], w[1]));
            double cs, sn;
            augb::rand_unit_vector(w[2], cs, sn);
            const double wv = fmax(mu * rad2 * cs * cs, 1e-290);          // mu N^2
            const double s = augb::sqrt_pos(wv * fma(4.0, lam, wv));
            const double den = wv + s;
            const double x1 = 4.0 * mu * lam * wv * augf::rcp(den * den);
            const bool first = augr::u32_mid(w[3]) * (mu + x1) <= mu;      // P(x1) = mu / (mu + x1)
            const double x = first ? x1 : mu * mu * augf::rcp(fmax(x1, 1e-290));
            if (valid) st_stream1(a.omega + el, x);
            return true;
        } else {
            const double v = augb::gamma_attempt(w, a.c1);
            const bool ok = v >= 0.0;
            if (valid && ok) st_stream1(a.omega + el, v * 2.0 * augf::rcp(fma(d, d, a.c0)));
            return ok || !valid;
        }
    };
    uint32_t ch = blockIdx.x * (MS_BLOCK / 32) + warp;
    for (;;) {
        if (qn < 32 && ch < nchunks) {
            const uint32_t el = (ch << 5) + lane;
            const bool valid = el < a.n;
            const bool done = attempt(el, 0u, valid);
            ch += W;
            if (KIND != AUG_LAPLACE) {
                const uint32_t m = __ballot_sync(0xffffffffu, !done);
                if (!done) {
                    const int pos = qn + __popc(m & lt_mask);
                    qel[pos] = el;
                    qat[pos] = 1u;
                }
                qn += __popc(m);
                __syncwarp();
            }
            continue;
        }
        if (qn == 0) break;
        const int cnt = qn < 32 ? qn : 32;
        const bool active = lane < cnt;
        uint32_t el = 0, att = 0;
        if (active) { el = qel[qn - cnt + lane]; att = qat[qn - cnt + lane]; }
        __syncwarp();
        qn -= cnt;
        const bool done = attempt(el, att, active);
        const uint32_t m = __ballot_sync(0xffffffffu, active && !done);
        if (active && !done) {
            const int pos = qn + __popc(m & lt_mask);
            qel[pos] = el;
            qat[pos] = att + 1u;
        }
        qn += __popc(m);
        __syncwarp();
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
            omega[i] = augb::pg_draw_stream(g, 1.0, true, 0.0, pgtab);
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
        out[i] = augb::pg_draw_stream(g, bi, b_is_int != 0, ci, pgtab);
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

// PG(1, c) for n elements through the warp-compacted kernel; false if the shape does not fit its item encoding
// takes over a pending split-phase exchange of the ctx, if any: the launch gets one extra CTA that gathers it
AugXchDev* take_pending(aug_ctx* ctx) {
    if (!ctx->pending || !ctx->xch) return nullptr;
    ctx->pending = 0;
    return ctx->xch;
}

bool launch_pg1_compact(aug_ctx* ctx, int64_t n, int64_t i0, uint64_t off, const double* c, double cs, double* out,
                        int32_t* rc) {
    static int occ = 0;
    if (occ == 0) {
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, pg1_compact_kernel, PG1_BLOCK, 0) != cudaSuccess || occ < 1)
            occ = 1;
    }
    int64_t grid = (int64_t)ctx->sms * occ;
    const int64_t nchunks = (n + 31) / 32;
    const int64_t need = (nchunks + (PG1_BLOCK / 32) - 1) / (PG1_BLOCK / 32);
    if (grid > need) grid = need;
    if (grid < 1) grid = 1;
    if (n >= ((int64_t)1 << 32) - 64) return false;   // element offsets are queued as 32-bit words
    Pg1Args a{};
    a.n = n;
    a.i0 = i0;
    a.seed = ctx->seed;
    a.offset = off;
    a.c = c;
    a.cs = cs;
    a.out = out;
    a.tab = ctx->pgtab;
    a.gx = take_pending(ctx);
    if (a.gx) grid += 1;
    augr::philox_round_keys((uint32_t)a.seed, (uint32_t)(a.seed >> 32) ^ (uint32_t)(a.offset >> 32), &a.keys);
    pg1_compact_kernel<<<(unsigned)grid, PG1_BLOCK, 0, ctx->stream>>>(a);
    ctx->launches++;
    *rc = (int32_t)cudaGetLastError();
    return true;
}


// PG(b, c) for n elements through the warp-compacted general-b kernel
template <int KIND>
int32_t launch_pgb(aug_ctx* ctx, PgbArgs a) {
    static int occ = 0;
    if (occ == 0) {
        if (cudaFuncSetAttribute(pgb_kernel<KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, PGB_SMEM_BYTES) != cudaSuccess)
            return (int32_t)cudaGetLastError();
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, pgb_kernel<KIND>, PGB_BLOCK, PGB_SMEM_BYTES) != cudaSuccess || occ < 1)
            occ = 1;
    }
    int64_t grid = (int64_t)ctx->sms * occ;
    const int64_t nchunks = (a.n + 31) / 32;
    const int64_t need = (nchunks + (PGB_BLOCK / 32) - 1) / (PGB_BLOCK / 32);
    if (grid > need) grid = need;
    if (grid < 1) grid = 1;
    a.gx = take_pending(ctx);
    if (a.gx) grid += 1;
    augr::philox_round_keys((uint32_t)a.seed, (uint32_t)(a.seed >> 32) ^ (uint32_t)(a.offset >> 32), &a.keys);
    pgb_kernel<KIND><<<(unsigned)grid, PGB_BLOCK, PGB_SMEM_BYTES, ctx->stream>>>(a);
    ctx->launches++;
    return (int32_t)cudaGetLastError();
}

bool is_cat(int k) { return k == AUG_CAT || k == AUG_CAT_BIJ; }

bool pg1_no_compact() {   // AUGCUDA_NO_COMPACT=1 keeps PG(1) draws on the one-thread-one-draw kernel (A/B measurements)
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("AUGCUDA_NO_COMPACT");
        v = (e && e[0] == '1') ? 1 : 0;
    }
    return v == 1;
}

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
    if (is_cat(lik->kind)) {
        int32_t rf = aug_xch_flush(c);
        if (rf) return rf;
        return aug_cat_sample(c, lik, n, i0, y, f, omega, nvar, off);
    }
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
    if (lik->kind == AUG_BERNOULLI && !pg1_no_compact()) {      // PG(1, |f|)  bernoulli.jl:13-15
        int32_t r2 = 0;
        if (launch_pg1_compact(c, n, i0, off, f, 0.0, omega, &r2)) return r2;
    }
    if (!pg1_no_compact() && n < ((int64_t)1 << 32) - 64 &&
        (lik->kind == AUG_NEGBIN || lik->kind == AUG_POISSON || lik->kind == AUG_HETERO)) {
        PgbArgs p{};
        p.n = n;
        p.i0 = i0;
        p.seed = c->seed;
        p.offset = off;
        p.y = y;
        p.f = f;
        p.g = a.g;
        p.b_is_int = a.L.r_is_int;
        p.p0 = a.L.p0;
        p.omega = omega;
        p.nvar = nvar;
        p.tab = c->pgtab;
        if (lik->kind == AUG_NEGBIN) return launch_pgb<AUG_NEGBIN>(c, p);
        if (lik->kind == AUG_POISSON) return launch_pgb<AUG_POISSON>(c, p);
        return launch_pgb<AUG_HETERO>(c, p);
    }
    { int32_t rf = aug_xch_flush(c); if (rf) return rf; }     // no gather hook in these kernels: complete a pending exchange first
    if (!pg1_no_compact() && n < ((int64_t)1 << 32) - 64 &&
        (lik->kind == AUG_LAPLACE || (lik->kind == AUG_STUDENTT && a.L.c1 >= 1.0))) {
        MapSampleArgs m{};
        m.n = n;
        m.i0 = i0;
        m.offset = off;
        m.y = reinterpret_cast<const double*>(y);
        m.f = f;
        m.omega = omega;
        m.c0 = a.L.c0;
        m.c1 = a.L.c1;
        augr::philox_round_keys((uint32_t)c->seed, (uint32_t)(c->seed >> 32) ^ (uint32_t)(off >> 32), &m.keys);
        const void* k = lik->kind == AUG_LAPLACE ? (const void*)map_sample_kernel<AUG_LAPLACE> : (const void*)map_sample_kernel<AUG_STUDENTT>;
        int occ = 1;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k, MS_BLOCK, 0) != cudaSuccess || occ < 1) occ = 1;
        int64_t grid = (int64_t)c->sms * occ;
        const int64_t need = ((n + 31) / 32 + (MS_BLOCK / 32) - 1) / (MS_BLOCK / 32);
        if (grid > need) grid = need;
        if (grid < 1) grid = 1;
        if (lik->kind == AUG_LAPLACE) map_sample_kernel<AUG_LAPLACE><<<(unsigned)grid, MS_BLOCK, 0, c->stream>>>(m);
        else map_sample_kernel<AUG_STUDENTT><<<(unsigned)grid, MS_BLOCK, 0, c->stream>>>(m);
        c->launches++;
        return (int32_t)cudaGetLastError();
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

// device-pointer implementation of init_aux_variables with an explicit RNG tick (shared with the host-buffer pipeline)
int32_t aug_init_aux_variables_dev(aug_ctx* c, const aug_lik* lik, int64_t n, int64_t i0, double* omega,
                                   int64_t* nvar, uint64_t off) {
    if (n == 0) return AUG_OK;
    const bool needs_n = lik->kind == AUG_POISSON || lik->kind == AUG_HETERO || is_cat(lik->kind);
    if (needs_n && !nvar) return AUG_ERR_BAD_ARG;
    const int64_t per = is_cat(lik->kind) ? lik->nlatent : 1;
    const int64_t m = n * per;
    if ((lik->kind == AUG_BERNOULLI || lik->kind == AUG_NEGBIN) && !pg1_no_compact()) {   // PG(1, 0) only: the compacted sampler
        int32_t r2 = 0;
        if (launch_pg1_compact(c, n, i0, off, nullptr, 0.0, omega, &r2)) return r2;
    }
    const int grid = aug_grid_for(c, (const void*)init_aux_kernel, m, AUG_BLOCK);
    init_aux_kernel<<<grid, AUG_BLOCK, 0, c->stream>>>(lik->kind, m, i0 * per, c->seed, off, omega,
                                                        needs_n ? nvar : nullptr, c->pgtab);
    c->launches++;
    return (int32_t)cudaGetLastError();
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
    return aug_init_aux_variables_dev(c, lik, n, i0, omega, nvar, off);
}

static int32_t pg_rand_common(aug_ctx* c, int64_t n, int64_t i0, const double* b, const double* cc, double bs,
                              double cs, int32_t b_is_int, double* out) {
    if (!c) return AUG_ERR_NOT_INIT;
    if (n < 0 || !out) return AUG_ERR_BAD_ARG;
    AUG_CUDA(cudaSetDevice(c->device));
    const uint64_t off = c->offset++;
    if (n == 0) return AUG_OK;
    if (!b && b_is_int && bs == 1.0 && !pg1_no_compact()) {     // all draws are PG(1, c)
        int32_t r2 = 0;
        if (launch_pg1_compact(c, n, i0, off, cc, cs, out, &r2)) return r2;
    }
    if (!pg1_no_compact() && n < ((int64_t)1 << 32) - 64) {
        PgbArgs p{};
        p.n = n;
        p.i0 = i0;
        p.seed = c->seed;
        p.offset = off;
        p.b = b;
        p.c = cc;
        p.bs = bs;
        p.cs = cs;
        p.b_is_int = b_is_int;
        p.omega = out;
        p.tab = c->pgtab;
        return launch_pgb<PGB_RAW>(c, p);
    }
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
