// aug_sparse.cu — SURVEY §8(f) rows 1 and 2: the sparse-GP steps either side of the augmentation path, and the
// whole CAVI iteration of a sparse variational GP as ONE pass over κ = K_Z⁻¹ K_{Z,X}.
//
// Reference behaviour replaced (paths relative to /root/reference):
//   examples/bernoulli/script.jl:29-39   cavi!: marginals(post_u(x)) → aux_posterior! → S, m update
//   docs/src/index.md:154-163            sparse form: S = (K_Z⁻¹ + κ Diagonal(r) κᵀ)⁻¹, m = S(κ t + K_Z⁻¹ μ₀(Z))
// For every observation t with column κ_t (M doubles, contiguous: Julia's column-major M×N matrix):
//   producer   μ_t = κ_tᵀ m,  σ²_t = k_tt − κ_tᵀ B κ_t                      (B = K_Z − S, M×M)
//   path       aux_posterior! + E[β_t], E[γ_t] + ELBO terms                  (aug_cavi_eval.cuh)
//   consumer   P += γ_t κ_t κ_tᵀ (lower block triangle),  rhs += β_t κ_t
//
// This IS GEMM-shaped work, in fp64: 2M² + M² flops per observation against 8M bytes, so for M >= 32 the bound is
// the FP64 pipe (measured 37.1 TFLOP/s for DMMA and for DFMA on B200, tools/fp64_peak.cu), not HBM.  tcgen05 has
// no f64 kind; the fp64 tensor path of sm_100a is mma.sync.m8n8k4.f64 (SASS DMMA.8x8x4), which is what the GEMM
// warps issue: same peak as DFMA but 1 operand fetch per 256 FMAs instead of per 32.
//
// One persistent CTA per SM: 16 GEMM warps + TO/32 evaluation warps.
//   * κ tiles of TO observations stream through a 3-stage shared-memory ring (cp.async 16 B, rows padded to
//     MT + 4 doubles so that every fragment load below is bank-conflict-free); B is staged once, padded alike.
//   * iteration j:  phase A   GEMM warps: producer GEMM of tile j+1 (C = B·κ_tile, then q_t = Σ_i κ_it C_it by
//                             in-register products + 3 shuffles)  ‖  evaluation warps: μ_t, σ²_t, CAVI closed
//                             forms of tile j, global stores, γ_t / β_t into shared memory
//                   phase B   GEMM warps: consumer SYRK of tile j into register accumulators (warp p of a k-group
//                             owns block-rows p and MB−1−p of the lower triangle: MB+1 blocks each)
//                             ‖  evaluation warps: rhs += β_t κ_t
//     so the latency-bound scalar math of tile j hides behind the DMMA stream of tile j+1.
//   * P / rhs / ELBO partials of every CTA (and k-group) go to a scratch buffer; a second launch adds them in a
//     fixed order (bit-reproducible, no floating-point atomics), mirrors the lower triangle and adds P0 / r0.
#include "aug_common.cuh"
#include "aug_math.cuh"
#include "aug_cavi_eval.cuh"

namespace {

constexpr int SP_GW = 16;        // GEMM warps per CTA
constexpr int SP_STAGES = 3;     // κ ring depth: tile j (consumer), j+1 (producer), j+2 (in flight)

enum { SP_FUSED = 0, SP_PRODUCER = 1, SP_CONSUMER = 2 };

template <int MT>
struct SpCfg {
    static constexpr int MB = MT / 8;                  // 8-row blocks along the inducing axis
    static constexpr int RB = MT == 128 ? 1 : 2;       // producer: row-blocks per warp
    static constexpr int WR = MB / RB;                 // producer: warps along the rows of B
    static constexpr int WT = SP_GW / WR;              // producer: warps along the observations
    static constexpr int TO = 16 * WT;                 // observations per tile (2 column blocks of 8 per warp)
    static constexpr int PAIRS = MB / 2;               // consumer: warp p owns block-rows p and MB-1-p
    static constexpr int QS = MT == 128 ? 2 : 1;       // consumer: warps sharing one row pair (halves of its MB+1 blocks)
    static constexpr int KG = SP_GW / (PAIRS * QS);    // consumer: groups splitting the observations of a tile
    static constexpr int OG = TO / KG;                 // observations per group and tile (multiple of 4)
    static constexpr int NACC = (MB + QS) / QS;        // 8x8 accumulator blocks per consumer warp (ceil((MB+1)/QS))
    static constexpr int STRIDE = MT + 4;              // padded row length (doubles): ≡ 4 mod 16
    static constexpr int EW = TO >= 32 ? TO / 32 : 1;  // evaluation warps
    static constexpr int NE = EW * 32;
    static constexpr int NT = SP_GW * 32 + NE;         // threads per CTA
    static constexpr int RPT = MT > NE ? MT / NE : 1;  // rhs rows per evaluation thread
    static constexpr int RPARTS = NE > MT ? NE / MT : 1;  // evaluation threads sharing one rhs row
    // shared memory (doubles)
    static constexpr int OFF_B = 0;
    static constexpr int OFF_K = OFF_B + MT * STRIDE;
    static constexpr int OFF_M = OFF_K + SP_STAGES * TO * STRIDE;
    static constexpr int OFF_Q = OFF_M + MT;                       // q partials [2][WR][TO]
    static constexpr int OFF_G = OFF_Q + 2 * WR * TO;              // γ_t [TO]
    static constexpr int OFF_BE = OFF_G + TO;                      // β_t [TO]
    static constexpr int OFF_RED = OFF_BE + TO;                    // end-of-kernel reductions [NE]
    static constexpr int SMEM_DOUBLES = OFF_RED + NE * 2;
    static constexpr int SMEM_BYTES = SMEM_DOUBLES * 8;
    static_assert(OG % 4 == 0 && OG >= 4, "consumer k-steps cover 4 observations");
    static_assert(WR * WT == SP_GW && PAIRS * QS * KG == SP_GW, "warp grids");
    static_assert(STRIDE % 16 == 4, "bank-conflict-free padding");
};

struct SparseArgs {
    int64_t n;
    int m;
    int64_t ntiles;
    const void* y;
    const double* kappa;
    const double* mvec;
    const double* B;
    const double* kdiag;
    const double* gamma_in;   // CONSUMER mode
    const double* beta_in;
    double* mu;
    double* var;
    double* s0;
    double* s1;
    void* s2;
    double* beta;
    double* gamma;
    double* scratch;          // [grid*KG][MT*MT] P partials, then [grid][MT] rhs partials, then [grid][2] ELBO partials
    int vec16;                // kappa is 16-byte aligned and m is even: 16-byte cp.async
    int elbo;
    LikConst L;
};

__device__ __forceinline__ void dmma8x8x4(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}
__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async8(void* dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// issue the copies of one κ tile (rows of observations [t0, t0 + TO) ∩ [0, n), m doubles each) into a ring stage
template <int MT>
__device__ __forceinline__ void sp_issue_tile(const SparseArgs& a, int64_t tile, double* stage) {
    typedef SpCfg<MT> C;
    if (tile < a.ntiles) {
        const int64_t t0 = tile * C::TO;
        const int rows = (int)((a.n - t0) < (int64_t)C::TO ? (a.n - t0) : (int64_t)C::TO);
        const double* src = a.kappa + t0 * a.m;
        const unsigned um = (unsigned)a.m;
        if (a.vec16) {
            const unsigned tot = (unsigned)rows * um / 2u;
            for (unsigned e2 = threadIdx.x; e2 < tot; e2 += C::NT) {
                const unsigned e = 2u * e2;
                const unsigned row = e / um, col = e - row * um;
                cp_async16(stage + row * C::STRIDE + col, src + e);
            }
        } else {
            const unsigned tot = (unsigned)rows * um;
            for (unsigned e = threadIdx.x; e < tot; e += C::NT) {
                const unsigned row = e / um, col = e - row * um;
                cp_async8(stage + row * C::STRIDE + col, src + e);
            }
        }
    }
    cp_async_commit();   // always: keeps the group count uniform across threads and iterations
}

// producer GEMM of one tile: q partial of this warp's row-blocks for its 16 observations → qp[wr][t]
template <int MT>
__device__ __forceinline__ void sp_producer(const double* __restrict__ Bs, const double* __restrict__ kap,
                                            double* __restrict__ qp, int warp, int lane, int kend) {
    typedef SpCfg<MT> C;
    const int wr = warp % C::WR, wt = warp / C::WR;
    const int r = lane >> 2, k = lane & 3;
    double c[C::RB][2][2];
#pragma unroll
    for (int rb = 0; rb < C::RB; ++rb)
#pragma unroll
        for (int tb = 0; tb < 2; ++tb) c[rb][tb][0] = c[rb][tb][1] = 0.0;
    const double* ap = Bs + (8 * (wr * C::RB) + r) * C::STRIDE + k;       // A[row = r][k]     = B[8I + r][k0 + k]
    const double* bp = kap + (8 * (wt * 2) + r) * C::STRIDE + k;          // B[k][col = r]     = κ[k0 + k][t0 + r]
#pragma unroll 4
    for (int k0 = 0; k0 < kend; k0 += 4) {
        double av[C::RB], bv[2];
#pragma unroll
        for (int rb = 0; rb < C::RB; ++rb) av[rb] = ap[rb * 8 * C::STRIDE + k0];
#pragma unroll
        for (int tb = 0; tb < 2; ++tb) bv[tb] = bp[tb * 8 * C::STRIDE + k0];
#pragma unroll
        for (int rb = 0; rb < C::RB; ++rb)
#pragma unroll
            for (int tb = 0; tb < 2; ++tb) dmma8x8x4(c[rb][tb][0], c[rb][tb][1], av[rb], bv[tb]);
    }
    // C[row r][col 2k + e] = (Bκ)[8I + r][t]; q_t = Σ_i κ[i][t] (Bκ)[i][t]: own rows, then the 8 lanes sharing k
#pragma unroll
    for (int tb = 0; tb < 2; ++tb)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const int t = 8 * (wt * 2 + tb) + 2 * k + e;
            double s = 0.0;
#pragma unroll
            for (int rb = 0; rb < C::RB; ++rb)
                s = fma(c[rb][tb][e], kap[t * C::STRIDE + 8 * (wr * C::RB + rb) + r], s);
            s += __shfl_xor_sync(0xffffffffu, s, 4);
            s += __shfl_xor_sync(0xffffffffu, s, 8);
            s += __shfl_xor_sync(0xffffffffu, s, 16);
            if (r == 0) qp[wr * C::TO + t] = s;
        }
}

// consumer SYRK of one tile into this warp's accumulators
template <int MT>
__device__ __forceinline__ void sp_consumer(const double* __restrict__ kap, const double* __restrict__ gs,
                                            double (&acc)[SpCfg<MT>::NACC][2], int warp, int lane) {
    typedef SpCfg<MT> C;
    const int p = warp % C::PAIRS, h = (warp / C::PAIRS) % C::QS, g = warp / (C::PAIRS * C::QS);
    const int r = lane >> 2, k = lane & 3;
#pragma unroll
    for (int ks = 0; ks < C::OG / 4; ++ks) {
        const int t = g * C::OG + 4 * ks + k;
        const double* row = kap + t * C::STRIDE + r;
        const double gm = gs[t];
        const double aP = row[8 * p] * gm;                       // A[row r][k] = γ_t κ[8I + r][t]
        const double aQ = row[8 * (C::MB - 1 - p)] * gm;
#pragma unroll
        for (int qi = 0; qi < C::NACC; ++qi) {
            const int q = h * C::NACC + qi;                       // block index in the concatenated rows p, MB-1-p
            if (C::QS == 1 || q <= C::MB) {
                const bool first = q <= p;
                const int J = first ? q : q - p - 1;
                const double b = row[8 * J];                      // B[k][col r] = κ[8J + r][t]
                dmma8x8x4(acc[qi][0], acc[qi][1], first ? aP : aQ, b);
            }
        }
    }
}

template <int MT, int MODE, int KIND>
__global__ void __launch_bounds__(SpCfg<MT>::NT, 1) sparse_sweep_kernel(const SparseArgs a) {
    typedef SpCfg<MT> C;
    typedef typename YT<KIND>::T yt;
    typedef typename S2T<KIND>::T s2t;
    constexpr bool PROD = MODE != SP_CONSUMER;
    constexpr bool CONS = MODE != SP_PRODUCER;
    constexpr bool YSTATE = KIND == AUG_NEGBIN || KIND == AUG_POISSON;
    extern __shared__ __align__(16) double sm[];
    double* Bs = sm + C::OFF_B;
    double* ring = sm + C::OFF_K;
    double* ms = sm + C::OFF_M;
    double* qs = sm + C::OFF_Q;
    double* gs = sm + C::OFF_G;
    double* bs = sm + C::OFF_BE;
    double* red = sm + C::OFF_RED;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool gemm_warp = warp < SP_GW;
    const int et = tid - SP_GW * 32;                 // evaluation thread index (>= 0 for evaluation warps)
    const int m = a.m;
    const int kend = (m + 3) & ~3;

    // ---- one-time staging: zero the ring (pad columns / rows stay zero or finite for ever), B, m
    for (int e = tid; e < SP_STAGES * C::TO * C::STRIDE; e += C::NT) ring[e] = 0.0;
    if (PROD) {
        for (int e = tid; e < MT * C::STRIDE; e += C::NT) {
            const int i = e / C::STRIDE, j = e - i * C::STRIDE;
            Bs[e] = (i < m && j < m) ? __ldg(a.B + (size_t)i * m + j) : 0.0;
        }
        for (int e = tid; e < MT; e += C::NT) ms[e] = e < m ? __ldg(a.mvec + e) : 0.0;
    }
    for (int e = tid; e < C::TO; e += C::NT) { gs[e] = 0.0; bs[e] = 0.0; }
    __syncthreads();

    // local tile l ↔ global tile blockIdx.x + l * gridDim.x
    const int64_t nloc = a.ntiles > (int64_t)blockIdx.x ? (a.ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    auto gtile = [&](int64_t l) -> int64_t { return l < nloc ? (int64_t)blockIdx.x + l * gridDim.x : a.ntiles; };
    auto stage = [&](int64_t l) -> double* { return ring + (int)(l % SP_STAGES) * (C::TO * C::STRIDE); };

    double acc[C::NACC][2];
#pragma unroll
    for (int q = 0; q < C::NACC; ++q) acc[q][0] = acc[q][1] = 0.0;
    double racc[C::RPT];
#pragma unroll
    for (int i = 0; i < C::RPT; ++i) racc[i] = 0.0;
    double e_elt = 0.0, e_kl = 0.0;

    // ---- prologue
    sp_issue_tile<MT>(a, gtile(0), stage(0));
    sp_issue_tile<MT>(a, gtile(1), stage(1));
    cp_async_wait<1>();
    __syncthreads();                                   // tile 0 visible
    sp_issue_tile<MT>(a, gtile(2), stage(2));
    if (PROD && gemm_warp && nloc > 0) sp_producer<MT>(Bs, stage(0), qs, warp, lane, kend);
    cp_async_wait<1>();
    __syncthreads();                                   // q(0) and tile 1 visible; tile 2 may be in flight

    for (int64_t l = 0; l < nloc; ++l) {
        const double* kap = stage(l);
        // ================= phase A: producer(l+1) ‖ evaluation(l)
        if (gemm_warp) {
            if (PROD && l + 1 < nloc)
                sp_producer<MT>(Bs, stage(l + 1), qs + (int)((l + 1) & 1) * (C::WR * C::TO), warp, lane, kend);
        } else if (et < C::TO) {
            const int t = et;
            const int64_t i = gtile(l) * C::TO + t;
            const bool valid = i < a.n;
            double g0 = 0.0, b0 = 0.0;
            if (MODE == SP_CONSUMER) {
                if (valid) { g0 = __ldg(a.gamma_in + i); b0 = __ldg(a.beta_in + i); }
            } else {
                double kd = 0.0, yv = 0.0;
                if (valid) {
                    kd = __ldg(a.kdiag + i);
                    if (MODE == SP_FUSED) yv = (double)__ldg(reinterpret_cast<const yt*>(a.y) + i);
                }
                const double* row = kap + t * C::STRIDE;
                double mu0 = 0.0, mu1 = 0.0;
                for (int j = 0; j < kend; j += 2) {
                    mu0 = fma(ms[j], row[j], mu0);
                    mu1 = fma(ms[j + 1], row[j + 1], mu1);
                }
                const double mu = mu0 + mu1;
                const double* qp = qs + (int)(l & 1) * (C::WR * C::TO) + t;
                double q = 0.0;
#pragma unroll
                for (int w = 0; w < C::WR; ++w) q += qp[w * C::TO];
                const double var = kd - q;
                if (valid) {
                    if (a.mu) a.mu[i] = mu;
                    if (a.var) a.var[i] = var;
                    if (MODE == SP_FUSED) {
                        Obs o;
                        o.y = yv; o.ys = yv;
                        o.m = mu; o.v = var; o.mg = 0.0; o.vg = 0.0;
                        o.s0 = o.s1 = o.s2 = 0.0;
                        if (a.elbo) eval<KIND, false, true, true>(a.L, o);
                        else eval<KIND, false, false, true>(a.L, o);
                        if (a.s0) a.s0[i] = o.s0;
                        if (KIND == AUG_POISSON && a.s1) a.s1[i] = o.s1;
                        if (YSTATE && a.s2) reinterpret_cast<s2t*>(a.s2)[i] = (s2t)o.y;
                        if (a.beta) a.beta[i] = o.b0;
                        if (a.gamma) a.gamma[i] = o.g0;
                        g0 = o.g0; b0 = o.b0;
                        e_elt += o.elt; e_kl += o.kl;
                    }
                }
            }
            if (CONS) { gs[t] = g0; bs[t] = b0; }
        }
        __syncthreads();
        // ================= phase B: consumer(l) ‖ rhs(l)
        if (CONS) {
            if (gemm_warp) {
                sp_consumer<MT>(kap, gs, acc, warp, lane);
            } else {
                const int part = C::RPARTS > 1 ? et / MT : 0;
                const int i0 = C::RPARTS > 1 ? et - part * MT : et;
#pragma unroll 4
                for (int t = part; t < C::TO; t += C::RPARTS) {
                    const double bt = bs[t];
#pragma unroll
                    for (int rr = 0; rr < C::RPT; ++rr) racc[rr] = fma(bt, kap[t * C::STRIDE + i0 + rr * C::NE], racc[rr]);
                }
            }
        }
        cp_async_wait<0>();                            // tile l+2 landed (the only group in flight)
        __syncthreads();                               // everyone is done with tile l: its stage is free
        sp_issue_tile<MT>(a, gtile(l + 3), stage(l + 3));
    }
    cp_async_wait<0>();

    // ---- epilogue: partials to scratch
    double* Ppart = a.scratch;
    double* rpart = a.scratch + (size_t)gridDim.x * C::KG * (MT * MT);
    double* spart = rpart + (size_t)gridDim.x * MT;
    if (CONS) {
        if (gemm_warp) {
            const int p = warp % C::PAIRS, h = (warp / C::PAIRS) % C::QS, g = warp / (C::PAIRS * C::QS);
            const int r = lane >> 2, k = lane & 3;
            double* dst = Ppart + ((size_t)blockIdx.x * C::KG + g) * (MT * MT);
#pragma unroll
            for (int qi = 0; qi < C::NACC; ++qi) {
                const int q = h * C::NACC + qi;
                if (q <= C::MB) {
                    const bool first = q <= p;
                    const int I = first ? p : C::MB - 1 - p;
                    const int J = first ? q : q - p - 1;
                    double2 v = make_double2(acc[qi][0], acc[qi][1]);
                    *reinterpret_cast<double2*>(dst + (size_t)(8 * I + r) * MT + 8 * J + 2 * k) = v;
                }
            }
        } else {
            // rhs: sum the RPARTS evaluation threads that share a row (fixed order), then one value per row
            if (C::RPARTS > 1) {
                red[et] = racc[0];
            }
        }
    }
    __syncthreads();
    if (CONS && !gemm_warp) {
        if (C::RPARTS > 1) {
            if (et < MT) {
                double s = 0.0;
#pragma unroll
                for (int pp = 0; pp < C::RPARTS; ++pp) s += red[pp * MT + et];
                rpart[(size_t)blockIdx.x * MT + et] = s;
            }
        } else {
#pragma unroll
            for (int rr = 0; rr < C::RPT; ++rr) rpart[(size_t)blockIdx.x * MT + et + rr * C::NE] = racc[rr];
        }
    }
    __syncthreads();
    if (MODE == SP_FUSED && !gemm_warp) {
        // ELBO partials: evaluation threads → warp shuffles → shared → one pair per CTA
        const double s0 = warp_sum(e_elt), s1 = warp_sum(e_kl);
        if (lane == 0) { red[(warp - SP_GW) * 2] = s0; red[(warp - SP_GW) * 2 + 1] = s1; }
    }
    __syncthreads();
    if (MODE == SP_FUSED && tid == 0) {
        double s0 = 0.0, s1 = 0.0;
        for (int w = 0; w < C::EW; ++w) { s0 += red[2 * w]; s1 += red[2 * w + 1]; }
        spart[(size_t)blockIdx.x * 2] = s0;
        spart[(size_t)blockIdx.x * 2 + 1] = s1;
    }
}

// fixed-order sum of the partials: Pr[i*m + j] (mirrored from the lower triangle), Pr[m*m + i], scalars.
// A block = 8 consecutive outputs x 32 slices of the partial index b (slice s adds b = s, s + 32, ... in order, then
// one thread adds the 32 slice sums in order): deterministic, 64-byte coalesced reads, short dependent chains.
// The last block reduces the two ELBO partial columns.
template <int MT>
__global__ void __launch_bounds__(256) sparse_finalize_kernel(const double* __restrict__ scratch, int grid, int m,
                                                              const double* __restrict__ P0, const double* __restrict__ r0,
                                                              double* __restrict__ Pr, double* __restrict__ scalars, int want_P) {
    typedef SpCfg<MT> C;
    __shared__ double part[2][32][8];
    const double* Ppart = scratch;
    const double* rpart = scratch + (size_t)grid * C::KG * (MT * MT);
    const double* spart = rpart + (size_t)grid * MT;
    const int ex = threadIdx.x & 7, sl = threadIdx.x >> 3;
    if (blockIdx.x == gridDim.x - 1) {                        // scalar block
        if (scalars == nullptr) return;
        double s0 = 0.0, s1 = 0.0;
        if (ex == 0)
            for (int b = sl; b < grid; b += 32) { s0 += spart[2 * b]; s1 += spart[2 * b + 1]; }
        part[0][sl][ex] = s0;
        part[1][sl][ex] = s1;
        __syncthreads();
        if (threadIdx.x == 0) {
            double t0 = 0.0, t1 = 0.0;
            for (int q = 0; q < 32; ++q) { t0 += part[0][q][0]; t1 += part[1][q][0]; }
            scalars[AUG_S_EXPECTED_LOGTILT] = t0;
            scalars[AUG_S_KL] = t1;
            scalars[AUG_S_EXPECTED_AUGLL] = t0 + t1;          // generic.jl:52-54 ("+")
        }
        return;
    }
    if (!want_P) return;
    const int e = blockIdx.x * 8 + ex;
    double s = 0.0;
    if (e < m * m) {
        const int i = e / m, j = e - i * m;
        const int hi = i > j ? i : j, lo = i > j ? j : i;
        const double* src = Ppart + hi * MT + lo;
        for (int b = sl; b < grid * C::KG; b += 32) s += __ldg(src + (size_t)b * (MT * MT));
    } else if (e < m * m + m) {
        const int i = e - m * m;
        for (int b = sl; b < grid; b += 32) s += __ldg(rpart + (size_t)b * MT + i);
    }
    part[0][sl][ex] = s;
    __syncthreads();
    if (sl == 0 && e < m * m + m) {
        double t = 0.0;
        for (int q = 0; q < 32; ++q) t += part[0][q][ex];
        if (e < m * m) t += P0 ? P0[e] : 0.0;
        else t += r0 ? r0[e - m * m] : 0.0;
        Pr[e] = t;
    }
}

__global__ void dense_diag_kernel(int64_t n, const double* __restrict__ Kinv, const double* __restrict__ gamma,
                                  const double* __restrict__ beta, const double* __restrict__ r0, double* __restrict__ P,
                                  double* __restrict__ rhs) {
    const int64_t nth = (int64_t)gridDim.x * blockDim.x;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (P != Kinv) {
        for (int64_t e = tid; e < n * n; e += nth) {
            const int64_t i = e / n;
            P[e] = Kinv[e] + (e == i * n + i ? gamma[i] : 0.0);
        }
    } else {
        for (int64_t i = tid; i < n; i += nth) P[i * n + i] = Kinv[i * n + i] + gamma[i];
    }
    if (rhs) for (int64_t i = tid; i < n; i += nth) rhs[i] = beta[i] + (r0 ? r0[i] : 0.0);
}

template <int MT, int MODE, int KIND>
int32_t sp_launch(aug_ctx* ctx, SparseArgs& a, const double* P0, const double* r0, double* Pr, double* scalars) {
    typedef SpCfg<MT> C;
    const void* k = (const void*)sparse_sweep_kernel<MT, MODE, KIND>;
    static bool configured = false;
    if (!configured) {
        AUG_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
        configured = true;
    }
    a.ntiles = (a.n + C::TO - 1) / C::TO;
    int64_t grid = ctx->sms;
    if (grid > a.ntiles) grid = a.ntiles;
    if (grid < 1) grid = 1;
    const size_t need = ((size_t)grid * C::KG * (MT * MT) + (size_t)grid * MT + (size_t)grid * 2) * sizeof(double);
    if (ctx->sparse_scratch_bytes < need) {
        if (ctx->sparse_scratch) {
            AUG_CUDA(cudaStreamSynchronize(ctx->stream));
            AUG_CUDA(cudaFree(ctx->sparse_scratch));
            ctx->sparse_scratch = nullptr;
            ctx->sparse_scratch_bytes = 0;
        }
        AUG_CUDA(cudaMalloc(&ctx->sparse_scratch, need));
        ctx->sparse_scratch_bytes = need;
    }
    a.scratch = ctx->sparse_scratch;
    sparse_sweep_kernel<MT, MODE, KIND><<<(unsigned)grid, C::NT, C::SMEM_BYTES, ctx->stream>>>(a);
    ctx->launches++;
    AUG_CUDA(cudaGetLastError());
    const bool want_P = MODE != SP_PRODUCER;
    const bool want_s = MODE == SP_FUSED && scalars != nullptr;
    if (want_P || want_s) {
        const int nout = want_P ? a.m * a.m + a.m : 0;
        sparse_finalize_kernel<MT><<<(nout + 7) / 8 + 1, 256, 0, ctx->stream>>>(a.scratch, (int)grid, a.m, P0, r0, Pr,
                                                                             want_s ? scalars : nullptr, want_P ? 1 : 0);
        ctx->launches++;
        AUG_CUDA(cudaGetLastError());
    }
    return AUG_OK;
}

template <int MODE, int KIND>
int32_t sp_dispatch_m(aug_ctx* ctx, SparseArgs& a, const double* P0, const double* r0, double* Pr, double* scalars) {
    if (a.m <= 16) return sp_launch<16, MODE, KIND>(ctx, a, P0, r0, Pr, scalars);
    if (a.m <= 32) return sp_launch<32, MODE, KIND>(ctx, a, P0, r0, Pr, scalars);
    if (a.m <= 64) return sp_launch<64, MODE, KIND>(ctx, a, P0, r0, Pr, scalars);
    return sp_launch<128, MODE, KIND>(ctx, a, P0, r0, Pr, scalars);
}

int32_t sp_check(aug_ctx* ctx, int64_t n, int32_t m, const double* kappa) {
    if (!ctx) return AUG_ERR_NOT_INIT;
    if (n < 0 || m < 1 || m > 128) return AUG_ERR_BAD_ARG;
    if (n > 0 && kappa == nullptr) return AUG_ERR_BAD_ARG;
    return AUG_OK;
}

}  // namespace

extern "C" {

int32_t aug_sparse_marginals(aug_ctx* ctx, int64_t n, int32_t m, const double* kappa, const double* mvec,
                             const double* B, const double* kdiag, double* mu, double* var) {
    int32_t rc = sp_check(ctx, n, m, kappa);
    if (rc) return rc;
    if (n == 0) return AUG_OK;
    if (!mvec || !B || !kdiag || (!mu && !var)) return AUG_ERR_BAD_ARG;
    AUG_CUDA(cudaSetDevice(ctx->device));
    SparseArgs a{};
    a.n = n; a.m = m; a.kappa = kappa; a.mvec = mvec; a.B = B; a.kdiag = kdiag; a.mu = mu; a.var = var;
    a.vec16 = (aug_aligned16(kappa) && (m % 2 == 0)) ? 1 : 0;
    return sp_dispatch_m<SP_PRODUCER, AUG_BERNOULLI>(ctx, a, nullptr, nullptr, nullptr, nullptr);
}

int32_t aug_sparse_precision_potential(aug_ctx* ctx, int64_t n, int32_t m, const double* kappa,
                                       const double* gamma, const double* beta, const double* P0,
                                       const double* r0, double* Pr) {
    int32_t rc = sp_check(ctx, n, m, kappa);
    if (rc) return rc;
    if (!Pr || (n > 0 && (!gamma || !beta))) return AUG_ERR_BAD_ARG;
    AUG_CUDA(cudaSetDevice(ctx->device));
    SparseArgs a{};
    a.n = n; a.m = m; a.kappa = kappa; a.gamma_in = gamma; a.beta_in = beta;
    a.vec16 = (aug_aligned16(kappa) && (m % 2 == 0)) ? 1 : 0;
    return sp_dispatch_m<SP_CONSUMER, AUG_BERNOULLI>(ctx, a, P0, r0, Pr, nullptr);
}

int32_t aug_sparse_cavi_sweep(aug_ctx* ctx, const aug_lik* lik, int64_t n, int32_t m, const void* y,
                              const double* kappa, const double* mvec, const double* B, const double* kdiag,
                              double* mu, double* var, void* s0, void* s1, void* s2, double* beta,
                              double* gamma, const double* P0, const double* r0, double* Pr, double* scalars) {
    int32_t rc = sp_check(ctx, n, m, kappa);
    if (rc) return rc;
    if (!lik || !Pr || !mvec || !B) return AUG_ERR_BAD_ARG;
    if (n > 0 && (!y || !kdiag)) return AUG_ERR_BAD_ARG;
    AUG_CUDA(cudaSetDevice(ctx->device));
    SparseArgs a{};
    a.n = n; a.m = m; a.y = y; a.kappa = kappa; a.mvec = mvec; a.B = B; a.kdiag = kdiag;
    a.mu = mu; a.var = var; a.s0 = (double*)s0; a.s1 = (double*)s1; a.s2 = s2; a.beta = beta; a.gamma = gamma;
    a.vec16 = (aug_aligned16(kappa) && (m % 2 == 0)) ? 1 : 0;
    a.elbo = scalars != nullptr ? 1 : 0;
    rc = aug_lik_const(ctx, lik, &a.L, scalars != nullptr, false);
    if (rc) return rc;
    switch (lik->kind) {
        case AUG_BERNOULLI: return sp_dispatch_m<SP_FUSED, AUG_BERNOULLI>(ctx, a, P0, r0, Pr, scalars);
        case AUG_NEGBIN: return sp_dispatch_m<SP_FUSED, AUG_NEGBIN>(ctx, a, P0, r0, Pr, scalars);
        case AUG_POISSON: return sp_dispatch_m<SP_FUSED, AUG_POISSON>(ctx, a, P0, r0, Pr, scalars);
        case AUG_LAPLACE: return sp_dispatch_m<SP_FUSED, AUG_LAPLACE>(ctx, a, P0, r0, Pr, scalars);
        case AUG_STUDENTT: return sp_dispatch_m<SP_FUSED, AUG_STUDENTT>(ctx, a, P0, r0, Pr, scalars);
        default: return AUG_ERR_BAD_KIND;   // multi-latent likelihoods: one aug_sparse_* call per latent GP
    }
}

int32_t aug_dense_precision_potential(aug_ctx* ctx, int64_t n, const double* Kinv, const double* gamma,
                                      const double* beta, const double* r0, double* P, double* rhs) {
    if (!ctx) return AUG_ERR_NOT_INIT;
    if (n < 0) return AUG_ERR_BAD_ARG;
    if (n == 0) return AUG_OK;
    if (!Kinv || !gamma || !P || (rhs && !beta)) return AUG_ERR_BAD_ARG;
    AUG_CUDA(cudaSetDevice(ctx->device));
    const int64_t work = P != Kinv ? n * n : n;
    int64_t grid = (work + 255) / 256;
    if (grid > (int64_t)ctx->sms * 8) grid = (int64_t)ctx->sms * 8;
    dense_diag_kernel<<<(unsigned)grid, 256, 0, ctx->stream>>>(n, Kinv, gamma, beta, r0, P, rhs);
    ctx->launches++;
    return (int32_t)cudaGetLastError();
}

}  // extern "C"
