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
// This IS GEMM-shaped work, in fp64: the symmetric forms need 2M² (+O(M)) flops per observation against 8M bytes, so for
// M >= 32 the bound is the FP64 pipe (measured 37.1 TFLOP/s for DMMA and for DFMA on B200, tools/fp64_peak.cu), not
// HBM.  tcgen05 has no f64 kind; the fp64 tensor path of sm_100a is mma.sync.m8n8k4.f64 (SASS DMMA.8x8x4), which is
// what the GEMM warps issue: same peak as DFMA but 1 operand fetch per 256 FMAs instead of per 32.
//
// One persistent CTA per SM (M <= 128; larger M: sparse_general below):
//   * κ tiles of TO observations stream through a 4-stage shared-memory ring (cp.async 16 B, rows padded to MT + 4
//     doubles so that every fragment load is bank-conflict-free); B' = the symmetrised B arranged by cyclic block
//     diagonals (diagonal and half-diagonal blocks halved: κᵀBκ = 2 Σ_I κ_Iᵀ Σ_d B'_{I,d} κ_{(I+d) mod MB}) is staged once.
//   * warp roles: group X (8 warps) = producer GEMM of tile k (C = B'κ_tile by DMMA, q_t by in-register products and
//     a butterfly over the 8 lanes of a row group, μ_t by 16-element dot products); evaluator (warp 0 of X at
//     M = 128, dedicated warps below) = σ²_t and the CAVI closed forms of tile k-1, global stores, γ_t / β_t into shared
//     memory; group Y (8 warps) = consumer SYRK of tile k-2 into register accumulators (the owner of block-row I holds
//     the blocks (I, (I+d) mod MB), d = 0..MB/2) with rhs = κβ as one more DMMA per block-row.  ONE CTA barrier per
//     tile (sp_role_barrier): the two groups share the tensor pipe, so one group's scalar epilogue overlaps the other's
//     DMMA stream.
//   * P / rhs / ELBO partials of every CTA go to a scratch buffer; a second launch adds them in a fixed order
//     (bit-reproducible, no floating-point atomics), mirrors the lower triangle and adds P0 / r0 — and, in fused
//     multi-GPU mode, all-reduces them over the peer-memory mailbox inside that same launch.
// Measured behaviour, dropped variants and what bounds the kernel: DESIGN.md §3.6.
#include <dlfcn.h>

#include "aug_common.cuh"
#include "aug_math.cuh"
#include "aug_cavi_eval.cuh"

int32_t aug_cavi_dispatch(aug_ctx* ctx, const aug_lik* lik, int64_t n, const void* y, const double* mu,
                          const double* var, int64_t ld, void* s0, void* s1, void* s2, const void* rs0,
                          const void* rs1, const void* rs2, double* beta, double* gamma, int64_t ldo,
                          double* scalars, bool from_state);

namespace {

constexpr int SP_GW = 16;        // GEMM warps per CTA: warps 0-7 = producer group X, warps 8-15 = consumer group Y
constexpr int SP_XW = 8;         // warps per group
constexpr int SP_STAGES = 4;     // κ ring: tile k-2 (consumer), k-1 (between), k (producer), k+1 (in flight)

enum { SP_FUSED = 0, SP_PRODUCER = 1, SP_CONSUMER = 2 };

// Both GEMMs walk the symmetric M×M block structure by CYCLIC BLOCK DIAGONALS: the owner of block-row I handles the
// blocks (I, (I+d) mod MB) for d = 0 .. MB/2 (the half diagonal d = MB/2 only for I < MB/2 in the consumer, with
// weight ½ from both sides in the producer).  Every unordered pair of block-rows is met exactly once, every owner has
// the same amount of work, one of the two operands of every block is the owner's own κ rows, and all block offsets
// are (I + d) & (MB-1) with a compile-time d: no selects, no variable trip counts.
template <int MT>
struct SpCfg {
    static constexpr int MB = MT / 8;                  // 8-row blocks along the inducing axis
    static constexpr int ND = MB / 2 + 1;              // block diagonals per owner (d = 0 .. MB/2)
    static constexpr int OPW = MB > SP_XW ? MB / SP_XW : 1;   // owners (block-rows) per warp of a group
    static constexpr int SPL = MB > SP_XW ? 1 : SP_XW / MB;   // warps of a group sharing one owner (split over observations)
    static constexpr int TO = MT == 128 ? 32 : (MT == 64 ? 64 : (MT == 32 ? 128 : 256));   // observations per tile
    static constexpr int TBW = TO / 8 / SPL;           // producer: 8-observation column blocks per warp
    static constexpr int OG = TO / SPL;                // consumer: observations per warp and tile
    static constexpr int KG = SPL;                     // partial P / rhs copies per CTA
    static constexpr int NSLOT = SP_XW / SPL;          // producer: partial quadratic forms per observation
    static constexpr int STRIDE = MT + 4;              // padded κ row length (doubles): ≡ 4 mod 16
    static constexpr int BRS = ND * 8 + 4;             // padded row length of the staged B' (≡ 4 or 12 mod 16)
    static constexpr int EW = TO / 32;                 // evaluating warps: one observation per lane
    // MT = 128: the evaluator is warp 0 of group X (16 warps per CTA keep 128 registers per thread for the 17
    // accumulator blocks of a consumer warp); smaller MT: EW dedicated warps after the GEMM warps
    static constexpr bool EVAL_IN_X = MT == 128;
    static constexpr int NT = SP_GW * 32 + (EVAL_IN_X ? 0 : EW * 32);   // threads per CTA
    static constexpr int LPO = SP_XW * 32 / TO;        // producer lanes per observation for μ_t = κ_tᵀ m
    // shared memory (doubles)
    static constexpr int OFF_B = 0;
    static constexpr int OFF_K = OFF_B + MT * BRS;
    static constexpr int OFF_M = OFF_K + SP_STAGES * TO * STRIDE;
    static constexpr int OFF_Q = OFF_M + MT;                       // half quadratic forms [2][NSLOT][TO]
    static constexpr int OFF_MU = OFF_Q + 2 * NSLOT * TO;          // μ_t [2][TO]
    static constexpr int OFF_G = OFF_MU + 2 * TO;                  // γ_t [2][TO]
    static constexpr int OFF_BE = OFF_G + 2 * TO;                  // β_t [2][TO]
    static constexpr int OFF_RED = OFF_BE + 2 * TO;                // end-of-kernel reductions [2 EW]
    static constexpr int SMEM_DOUBLES = OFF_RED + 2 * EW + 2;
    static constexpr int SMEM_BYTES = SMEM_DOUBLES * 8;
    static_assert(OG % 4 == 0 && OG >= 4, "consumer k-steps cover 4 observations");
    static_assert(STRIDE % 16 == 4 && (BRS % 16 == 4 || BRS % 16 == 12), "bank-conflict-free padding");
    static_assert(LPO >= 1 && MT / LPO == 16, "μ lanes");
    static_assert(OPW * SP_XW == MB * SPL && (OPW == 1 || SPL == 1), "owner grid");
    static_assert(SMEM_BYTES <= 232448, "shared memory");
};

struct SparseArgs {
    int64_t n;
    int m;
    int64_t ntiles;
    int64_t ostride;          // element stride of the mu / var outputs (1, or nlatent for class-fastest [n][nl] arrays)
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
    double* scratch;          // [grid*KG][MT*MT] P partials, then [grid*KG][MT] rhs partials, then [grid][2] ELBO partials
    int vec16;                // kappa is 16-byte aligned and m is even: 16-byte cp.async
    int elbo;                 // ELBO sums requested
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

// The CTA-wide barrier of the role loops.  The three roles run separate copies of the tile loop; the barrier itself is
// ONE instruction (not inlined) that all of them call, so every thread of the CTA meets the same barrier at the same
// address (what compute-sanitizer's synccheck verifies).
template <int NT>
__device__ __noinline__ void sp_role_barrier() {
    asm volatile("bar.sync 1, %0;" ::"n"(NT) : "memory");
}

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
            if (um == (unsigned)MT) {                       // full width: row / column are shifts
                for (unsigned e2 = threadIdx.x; e2 < tot; e2 += C::NT) {
                    const unsigned e = 2u * e2;
                    cp_async16(stage + (e / MT) * C::STRIDE + (e % MT), src + e);
                }
            } else {
                for (unsigned e2 = threadIdx.x; e2 < tot; e2 += C::NT) {
                    const unsigned e = 2u * e2;
                    const unsigned row = e / um, col = e - row * um;
                    cp_async16(stage + row * C::STRIDE + col, src + e);
                }
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

// producer of one tile (group X, gw = 0..7): half quadratic forms q'_t = Σ_{i ∈ own rows} κ_it (B'κ)_it for this
// warp's TBW column blocks → qp[slot][t]  (κᵀBκ = 2 Σ_slots q'), and μ_t = κ_tᵀ m (LPO lanes per observation) → mus[t]
template <int MT>
__device__ __forceinline__ void sp_producer(const double* __restrict__ Bs, const double* __restrict__ kap,
                                            const double* __restrict__ ms, double* __restrict__ qp,
                                            double* __restrict__ mus, int gw, int lane) {
    typedef SpCfg<MT> C;
    const int r = lane >> 2, k = lane & 3;
    const int sp = C::OPW == 1 ? gw / C::MB : 0;                            // which share of the observations
    const int ow0 = C::OPW == 1 ? gw % C::MB : gw;                          // owners: ow0 (+ 8 when OPW == 2)
    const int tb0 = sp * C::TBW;
    double c[C::OPW][C::TBW][2];
#pragma unroll
    for (int o = 0; o < C::OPW; ++o)
#pragma unroll
        for (int tb = 0; tb < C::TBW; ++tb) c[o][tb][0] = c[o][tb][1] = 0.0;
    const double* ap = Bs + (8 * ow0 + r) * C::BRS + k;                     // A[row r][k] = B'[8 I + r][4 ks + k]
    const double* bp = kap + (8 * tb0 + r) * C::STRIDE + k;                 // B[k][col r] = κ[8 x + 4 half + k][t0 + r]
#pragma unroll
    for (int ks = 0; ks < 2 * C::ND; ++ks) {
        const int d = ks >> 1, half = ks & 1;
#pragma unroll
        for (int o = 0; o < C::OPW; ++o) {
            const double av = ap[o * (8 * SP_XW) * C::BRS + 4 * ks];
            const int xb = (ow0 + o * SP_XW + d) & (C::MB - 1);
            const double* bq = bp + 8 * xb + 4 * half;
#pragma unroll
            for (int tb = 0; tb < C::TBW; ++tb) dmma8x8x4(c[o][tb][0], c[o][tb][1], av, bq[tb * 8 * C::STRIDE]);
        }
    }
#ifdef SP_EXPERIMENT_NO_SCALAR   // timing experiment only (wrong results): the DMMA streams without the scalar FP64 work
    {
        int x = 0;   // keep every accumulator alive with integer ops only
#pragma unroll
        for (int o = 0; o < C::OPW; ++o)
#pragma unroll
            for (int tb = 0; tb < C::TBW; ++tb) x ^= __double2hiint(c[o][tb][0]) ^ __double2loint(c[o][tb][1]);
        if (x == 0x12345678) qp[0] = 1.0;
    }
    return;
#endif
    // C[row r][col 2k + e] = (B'κ)[8 I + r][t]: times κ[8 I + r][t], summed over own rows, then over the 8 lanes sharing k
    double s[C::TBW][2];
#pragma unroll
    for (int tb = 0; tb < C::TBW; ++tb)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const double* row = kap + (8 * (tb0 + tb) + 2 * k + e) * C::STRIDE + r;
            double v = c[0][tb][e] * row[8 * ow0];
            if (C::OPW == 2) v = fma(c[C::OPW - 1][tb][e], row[8 * (ow0 + SP_XW)], v);
            s[tb][e] = v;
        }
    // Σ over the 8 lanes r = 0..7 that share k.  With NV = 2 TBW >= 8 values per lane the butterfly keeps half of the
    // values per round (NV - NV/8 adds and shuffles instead of 3 NV): scalar FP64 instructions are served only in the
    // gaps of the DMMA stream (tools/fp64_peak.cu: a dependent DFMA takes 87 clk next to ONE streaming DMMA warp of the
    // same sub-partition and starves next to three), so every one of them counts.  Lane r ends with the sum of value
    // v = r (+ 8 j for NV > 8).
    constexpr int NV = 2 * C::TBW;
    double* sv = &s[0][0];
    if (NV % 8 == 0) {
#pragma unroll
        for (int j = 0; j < NV / 8; ++j) {
            double* v = sv + 8 * j;
            // round 1 (lanes r, r^1 ↔ xor 4): keep v[0..3] if bit0(r) == 0 else v[4..7]
            const bool b0 = r & 1, b1 = r & 2, b2 = r & 4;
            double w4[4], w2[2];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const double send = b0 ? v[i] : v[i + 4];
                const double keep = b0 ? v[i + 4] : v[i];
                w4[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
            }
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const double send = b1 ? w4[i] : w4[i + 2];
                const double keep = b1 ? w4[i + 2] : w4[i];
                w2[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
            }
            const double send = b2 ? w2[0] : w2[1];
            const double keep = b2 ? w2[1] : w2[0];
            v[0] = keep + __shfl_xor_sync(0xffffffffu, send, 16);   // value index 4 b0 + 2 b1 + b2 (+ 8 j)
        }
        // lane (r, k) holds the total of value vi = 4 b0 + 2 b1 + b2 of group j: tb = (8 j + vi) / 2, e = vi & 1
#pragma unroll
        for (int j = 0; j < NV / 8; ++j) {
            const int vi = ((r & 1) << 2) | (r & 2) | ((r >> 2) & 1);
            const int tb = (8 * j + vi) >> 1, e = vi & 1;
            qp[ow0 * C::TO + 8 * (tb0 + tb) + 2 * k + e] = sv[8 * j];
        }
    } else {
#pragma unroll
        for (int o = 4; o <= 16; o <<= 1)
#pragma unroll
            for (int tb = 0; tb < C::TBW; ++tb) {
                s[tb][0] += __shfl_xor_sync(0xffffffffu, s[tb][0], o);
                s[tb][1] += __shfl_xor_sync(0xffffffffu, s[tb][1], o);
            }
        if (r == 0) {
#pragma unroll
            for (int tb = 0; tb < C::TBW; ++tb) {
                double* dst = qp + ow0 * C::TO + 8 * (tb0 + tb) + 2 * k;       // slot = ow0 (0 .. NSLOT-1)
                dst[0] = s[tb][0];
                dst[1] = s[tb][1];
            }
        }
    }
    // μ_t: LPO consecutive lanes per observation, 16 elements each (start rotated by t when one lane owns a row:
    // 32 rows at stride ≡ 4 mod 16 would otherwise hit 4 banks)
    {
        const int idx = gw * 32 + lane;
        const int t = idx / C::LPO, sidx = idx % C::LPO;
        const double* row = kap + t * C::STRIDE;
        const int rot = C::LPO == 1 ? t : 0;
        double m0 = 0.0, m1 = 0.0;
#pragma unroll
        for (int j = 0; j < 16; j += 2) {
            const int i0 = sidx + C::LPO * ((j + rot) & 15), i1 = sidx + C::LPO * ((j + 1 + rot) & 15);
            m0 = fma(ms[i0], row[i0], m0);
            m1 = fma(ms[i1], row[i1], m1);
        }
        double mu = m0 + m1;
#pragma unroll
        for (int o = 1; o < C::LPO; o <<= 1) mu += __shfl_xor_sync(0xffffffffu, mu, o);
        if (sidx == 0) mus[t] = mu;
    }
}

// consumer of one tile (group Y, gw = 0..7) into this warp's accumulators: blocks P_{I,x} += Σ_t γ_t κ_It κ_xtᵀ for
// x = (I + d) mod MB, and the rhs slice (κβ)_I as one more DMMA whose B operand carries β_t in column 0 — a plain
// DFMA loop would starve behind the DMMA stream on the shared FP64 pipe.
template <int MT>
__device__ __forceinline__ void sp_consumer(const double* __restrict__ kap, const double* __restrict__ gs,
                                            const double* __restrict__ bs, double (&acc)[SpCfg<MT>::OPW][SpCfg<MT>::ND][2],
                                            double (&racc)[SpCfg<MT>::OPW][2], int gw, int lane) {
    typedef SpCfg<MT> C;
    const int r = lane >> 2, k = lane & 3;
    const int g = C::OPW == 1 ? gw / C::MB : 0;
    const int ow0 = C::OPW == 1 ? gw % C::MB : gw;
    // the γ_t scaling of step ks+1 is issued before the DMMAs of step ks: the DMUL only has to find a slot on the
    // FP64 pipe between the DMMAs of all warps, nothing waits for its result for a whole step
    double raw[C::OPW], av[C::OPW], bt;
    {
        const int t = g * C::OG + k;
        const double gm = gs[t];
        bt = r == 0 ? bs[t] : 0.0;                               // B[k][col 0] = β_t, other columns 0
#pragma unroll
        for (int o = 0; o < C::OPW; ++o) {
            raw[o] = kap[t * C::STRIDE + r + 8 * (ow0 + o * SP_XW)];
            av[o] = raw[o] * gm;                                 // A[row r][k] = γ_t κ[8 I + r][t]
        }
    }
#pragma unroll 2
    for (int ks = 0; ks < C::OG / 4; ++ks) {
        const int t = g * C::OG + 4 * ks + k;
        const double* row = kap + t * C::STRIDE + r;
        double rawn[C::OPW], avn[C::OPW], btn = 0.0;
        if (ks + 1 < C::OG / 4) {
            const double gm = gs[t + 4];
            btn = r == 0 ? bs[t + 4] : 0.0;
#pragma unroll
            for (int o = 0; o < C::OPW; ++o) {
                rawn[o] = row[4 * C::STRIDE + 8 * (ow0 + o * SP_XW)];
                avn[o] = rawn[o] * gm;
                asm volatile("" : "+d"(avn[o]));                 // keep the product here (one per owner and step)
            }
        } else {
#pragma unroll
            for (int o = 0; o < C::OPW; ++o) rawn[o] = avn[o] = 0.0;
        }
#pragma unroll
        for (int o = 0; o < C::OPW; ++o) {
            const int own = ow0 + o * SP_XW;
            const bool has_half = C::OPW == 2 ? (o == 0) : (own < C::MB / 2);
#pragma unroll
            for (int d = 0; d < C::ND; ++d) {
                if (d < C::ND - 1 || has_half) {
                    const double b = row[8 * ((own + d) & (C::MB - 1))];   // B[k][col r] = κ[8 x + r][t]
                    dmma8x8x4(acc[o][d][0], acc[o][d][1], av[o], b);
                }
            }
            dmma8x8x4(racc[o][0], racc[o][1], raw[o], bt);
        }
#pragma unroll
        for (int o = 0; o < C::OPW; ++o) { raw[o] = rawn[o]; av[o] = avn[o]; }
        bt = btn;
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
    double* mus = sm + C::OFF_MU;
    double* gs = sm + C::OFF_G;
    double* bs = sm + C::OFF_BE;
    double* red = sm + C::OFF_RED;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int role = warp < SP_XW ? 0 : (warp < SP_GW ? 1 : 2);   // 0 = producer group X, 1 = consumer group Y, 2 = evaluation E
    const bool evaluator = C::EVAL_IN_X ? warp < C::EW : role == 2;
    const int et = C::EVAL_IN_X ? tid : tid - SP_GW * 32;         // evaluation lane index: one observation each
    const int m = a.m;

    // ---- one-time staging: zero the ring (pad columns / rows stay zero or finite for ever), B', m
    for (int e = tid; e < SP_STAGES * C::TO * C::STRIDE; e += C::NT) ring[e] = 0.0;
    if (PROD) {
        // B' row (8 I + rr), column 8 d + cc  =  w_d · sym(B)[8 I + rr][8 ((I + d) mod MB) + cc],  w_0 = w_{MB/2} = ½
        for (int e = tid; e < MT * C::BRS; e += C::NT) {
            const int i = e / C::BRS, c = e - i * C::BRS;
            double v = 0.0;
            if (c < 8 * C::ND) {
                const int I = i >> 3, d = c >> 3;
                const int j = 8 * ((I + d) & (C::MB - 1)) + (c & 7);
                if (i < m && j < m) {
                    v = 0.5 * (__ldg(a.B + (size_t)i * m + j) + __ldg(a.B + (size_t)j * m + i));   // symmetrised
                    if (d == 0 || d == C::MB / 2) v *= 0.5;
                }
            }
            Bs[e] = v;
        }
        for (int e = tid; e < MT; e += C::NT) ms[e] = e < m ? __ldg(a.mvec + e) : 0.0;
    }
    for (int e = tid; e < 2 * C::TO; e += C::NT) { gs[e] = 0.0; bs[e] = 0.0; }
    __syncthreads();

    // local tile l ↔ global tile blockIdx.x + l * gridDim.x
    const int64_t nloc = a.ntiles > (int64_t)blockIdx.x ? (a.ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    auto gtile = [&](int64_t l) -> int64_t { return l < nloc ? (int64_t)blockIdx.x + l * gridDim.x : a.ntiles; };
    auto stage = [&](int64_t l) -> double* { return ring + (int)(l & (SP_STAGES - 1)) * (C::TO * C::STRIDE); };

    // per-observation scalars of the evaluation lanes, fetched one tile ahead (latency hidden behind a whole iteration)
    auto prefetch = [&](int64_t l, double& x0, double& x1) {
        x0 = 0.0; x1 = 0.0;
        if (l < nloc) {
            const int64_t i = gtile(l) * C::TO + et;
            if (i < a.n) {
                if (MODE == SP_CONSUMER) { x0 = __ldg(a.gamma_in + i); x1 = __ldg(a.beta_in + i); }
                else {
                    x0 = __ldg(a.kdiag + i);
                    if (MODE == SP_FUSED) x1 = (double)__ldg(reinterpret_cast<const yt*>(a.y) + i);
                }
            }
        }
    };

    // ---- prologue: tiles 0 and 1 land before the first iteration
    sp_issue_tile<MT>(a, gtile(0), stage(0));
    sp_issue_tile<MT>(a, gtile(1), stage(1));
    cp_async_wait<0>();
    __syncthreads();

    // ---- software pipeline, one CTA barrier per tile:  iteration k = producer(k) ‖ evaluation(k-1) ‖ consumer(k-2).
    // The two GEMM groups share the FP64 tensor pipe: the shuffle / reduction epilogue of the producer group and the
    // scalar math of the evaluation warps run while the other group keeps the pipe busy.  Each role runs its own
    // copy of the loop (so the consumer's accumulators are live in the consumer warps only); every copy executes
    // the same nloc + 2 barriers.
    auto tail = [&](int64_t k) {
        cp_async_wait<0>();                            // tile k+1 landed (the only group in flight)
        sp_role_barrier<C::NT>();                      // tile k-2 is free; q(k), μ(k), γ(k-1), β(k-1) are visible
        sp_issue_tile<MT>(a, gtile(k + 2), stage(k + 2));
    };
    double* Ppart = a.scratch;
    double* rpart = a.scratch + (size_t)gridDim.x * C::KG * (MT * MT);
    double* spart = rpart + (size_t)gridDim.x * C::KG * MT;

    double e_elt = 0.0, e_kl = 0.0;
    double pf0 = 0.0, pf1 = 0.0;                       // FUSED / PRODUCER: k_tt, y   CONSUMER: γ, β
    auto evaluate = [&](int64_t k) {
    if (k >= 1 && k - 1 < nloc) {
        const int64_t l = k - 1;
        const int t = et;
        const int64_t i = gtile(l) * C::TO + t;
        const bool valid = i < a.n;
        const double c0 = pf0, c1 = pf1;
        prefetch(l + 1, pf0, pf1);             // next tile's scalars: in flight during this iteration
        double g0 = 0.0, b0 = 0.0;
        const int cb = (int)(l & 1);
        if (MODE == SP_CONSUMER) {
            g0 = c0; b0 = c1;
        } else {
            const double mu = mus[cb * C::TO + t];
            const double* qp = qs + cb * (C::NSLOT * C::TO) + t;
            double q = 0.0;
#pragma unroll
            for (int w = 0; w < C::NSLOT; ++w) q += qp[w * C::TO];
            const double var = fma(-2.0, q, c0);   // k_tt − κᵀBκ,  κᵀBκ = 2 κᵀB'κ
            if (valid) {
                if (a.mu) a.mu[i * a.ostride] = mu;
                if (a.var) a.var[i * a.ostride] = var;
                if (MODE == SP_FUSED) {
                    Obs o;
                    o.y = c1; o.ys = c1;
                    o.m = mu; o.v = var; o.mg = 0.0; o.vg = 0.0;
                    o.s0 = o.s1 = o.s2 = 0.0;
#ifdef SP_EXPERIMENT_NO_SCALAR
                    o.g0 = 0.25; o.b0 = 0.5; o.elt = o.kl = 0.0;
                    if (false)
#endif
                    if (a.elbo) {
                        if (fast_ok<KIND, false, true>(o)) eval<KIND, false, true, false>(a.L, o);
                        else eval<KIND, false, true, true>(a.L, o);
                        e_elt += o.elt; e_kl += o.kl;
                    } else {
                        if (fast_ok<KIND, false, false>(o)) eval<KIND, false, false, false>(a.L, o);
                        else eval<KIND, false, false, true>(a.L, o);
                    }
                    if (a.s0) a.s0[i] = o.s0;
                    if (KIND == AUG_POISSON && a.s1) a.s1[i] = o.s1;
                    if (YSTATE && a.s2) reinterpret_cast<s2t*>(a.s2)[i] = (s2t)o.y;
                    if (a.beta) a.beta[i] = o.b0;
                    if (a.gamma) a.gamma[i] = o.g0;
                    g0 = o.g0; b0 = o.b0;
                }
            }
        }
        if (CONS) { gs[cb * C::TO + t] = g0; bs[cb * C::TO + t] = b0; }
    }
    };
    if (evaluator) prefetch(0, pf0, pf1);

    if (role == 0) {
        // ================= producer group X (+ evaluation of the previous tile in its first EW warps)
        for (int64_t k = 0; k < nloc + 2; ++k) {
            if (PROD && k < nloc) {
                const int nb = (int)(k & 1);
                sp_producer<MT>(Bs, stage(k), ms, qs + nb * (C::NSLOT * C::TO), mus + nb * C::TO, warp, lane);
            }
            if (C::EVAL_IN_X && evaluator) evaluate(k);
            tail(k);
        }
    } else if (role == 1) {
        // ================= consumer group Y
        double acc[C::OPW][C::ND][2];
        double racc[C::OPW][2];
#pragma unroll
        for (int o = 0; o < C::OPW; ++o) {
            racc[o][0] = racc[o][1] = 0.0;
#pragma unroll
            for (int d = 0; d < C::ND; ++d) acc[o][d][0] = acc[o][d][1] = 0.0;
        }
        const int gw = warp - SP_XW;
        for (int64_t k = 0; k < nloc + 2; ++k) {
            if (CONS && k >= 2) {
                const int cb = (int)(k & 1);           // (k-2) & 1
                sp_consumer<MT>(stage(k - 2), gs + cb * C::TO, bs + cb * C::TO, acc, racc, gw, lane);
            }
            tail(k);
        }
        if (CONS) {
            const int r = lane >> 2, k = lane & 3;
            const int g = C::OPW == 1 ? gw / C::MB : 0;
            const int ow0 = C::OPW == 1 ? gw % C::MB : gw;
            double* dst = Ppart + ((size_t)blockIdx.x * C::KG + g) * (MT * MT);
            double* rd = rpart + ((size_t)blockIdx.x * C::KG + g) * MT;
#pragma unroll
            for (int o = 0; o < C::OPW; ++o) {
                const int own = ow0 + o * SP_XW;
                const bool has_half = C::OPW == 2 ? (o == 0) : (own < C::MB / 2);
#pragma unroll
                for (int d = 0; d < C::ND; ++d) {
                    if (d < C::ND - 1 || has_half) {
                        const int x = (own + d) & (C::MB - 1);
                        // C[r][2k + e] = P[8 own + r][8 x + 2k + e]: the lower-triangle copy is this block (x <= own)
                        // or its transpose (x > own)
                        if (x <= own) {
                            *reinterpret_cast<double2*>(dst + (size_t)(8 * own + r) * MT + 8 * x + 2 * k) =
                                make_double2(acc[o][d][0], acc[o][d][1]);
                        } else {
                            dst[(size_t)(8 * x + 2 * k) * MT + 8 * own + r] = acc[o][d][0];
                            dst[(size_t)(8 * x + 2 * k + 1) * MT + 8 * own + r] = acc[o][d][1];
                        }
                    }
                }
                if (k == 0) rd[8 * own + r] = racc[o][0];   // column 0 of the rhs block
            }
        }
    }
    else {
        // ================= dedicated evaluation warps (MT < 128)
        for (int64_t k = 0; k < nloc + 2; ++k) {
            evaluate(k);
            tail(k);
        }
    }
    if (MODE == SP_FUSED && evaluator) {
        // ELBO partials: evaluation lanes → warp shuffles → shared → one pair per CTA
        const int ew = et >> 5;
        const double s0 = warp_sum(e_elt), s1 = warp_sum(e_kl);
        if (lane == 0) { red[ew * 2] = s0; red[ew * 2 + 1] = s1; }
    }
    cp_async_wait<0>();
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
    const double* spart = rpart + (size_t)grid * C::KG * MT;
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
            scal_zero_except(scalars, 0x07u);
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
        for (int b = sl; b < grid * C::KG; b += 32) s += __ldg(rpart + (size_t)b * MT + i);
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

// Fused multi-GPU mode: the same fixed-order sums, then the all-reduce over ranks INSIDE this launch.  Every block
// pushes its 8 outputs into the bulk area [epoch parity][this rank] of EVERY rank's mailbox (peer stores over NVLink),
// the last block to arrive exchanges the two ELBO sums through the slot protocol (xch_allreduce: its release / acquire
// of the epoch flag also publishes the bulk data pushed before it), then adds the ranks' bulk entries in rank order
// (bit-identical on all ranks) plus P0 / r0 and writes Pr.
template <int MT>
__global__ void __launch_bounds__(256) sparse_finalize_xch_kernel(const double* __restrict__ scratch, int grid, int m,
                                                                  const double* __restrict__ P0, const double* __restrict__ r0,
                                                                  double* __restrict__ Pr, double* __restrict__ scalars,
                                                                  int kge, AugXchDev* __restrict__ x,
                                                                  unsigned int* __restrict__ counter) {
    __shared__ double part[32][8];
    __shared__ int s_last, s_ok;
    const double* Ppart = scratch;
    const double* rpart = scratch + (size_t)grid * kge * (MT * MT);
    const double* spart = rpart + (size_t)grid * kge * MT;
    const int ex = threadIdx.x & 7, sl = threadIdx.x >> 3;
    const int nr = x->nranks, me = x->rank;
    const unsigned long long ep = x->epoch + 1ull;            // the epoch this launch will publish (read before anyone bumps it)
    const size_t half = (size_t)(ep & 1ull) * AUG_MAX_RANKS * AUG_XCH_BULK;
    const int nout = m * m + m;
    const int e = blockIdx.x * 8 + ex;
    double s = 0.0;
    if (e < m * m) {
        const int i = e / m, j = e - i * m;
        const int hi = i > j ? i : j, lo = i > j ? j : i;
        const double* src = Ppart + hi * MT + lo;
        for (int b = sl; b < grid * kge; b += 32) s += __ldg(src + (size_t)b * (MT * MT));
    } else if (e < nout) {
        const int i = e - m * m;
        for (int b = sl; b < grid * kge; b += 32) s += __ldg(rpart + (size_t)b * MT + i);
    }
    part[sl][ex] = s;
    __syncthreads();
    if (sl == 0 && e < nout) {
        double t = 0.0;
        for (int q = 0; q < 32; ++q) t += part[q][ex];
        const unsigned long long bits = (unsigned long long)__double_as_longlong(t);
        for (int r = 0; r < nr; ++r)
            asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(x->box[r] + AUG_XCH_WORDS + half + (size_t)me * AUG_XCH_BULK + e),
                         "l"(bits)
                         : "memory");
        __threadfence_system();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned int t = atomicInc(counter, gridDim.x - 1);   // wraps to 0: graph-replayable
        s_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    // ELBO partials of the CTAs: one load pair per thread (in flight together), then a fixed-order sum by thread 0
    __shared__ double se[2][256];
    {
        double a0 = 0.0, a1 = 0.0;
        if (scalars != nullptr)
            for (int b = threadIdx.x; b < grid; b += 256) { a0 += __ldcg(spart + 2 * b); a1 += __ldcg(spart + 2 * b + 1); }
        se[0][threadIdx.x] = a0;
        se[1][threadIdx.x] = a1;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double v[2] = {0.0, 0.0};
        const int nb = grid < 256 ? grid : 256;
        for (int b = 0; b < nb; ++b) { v[0] += se[0][b]; v[1] += se[1][b]; }
        // every block fenced its bulk stores at system scope before taking its ticket; this fence orders all of them (seen
        // through the ticket) before the slot words published next, and the one after the gather orders the peers' bulk
        // stores before the loads below (the slot protocol itself carries no fence)
        __threadfence_system();
        xch_allreduce<2>(x, v);                               // publishes epoch ep, waits for every rank's words
        __threadfence_system();
        s_ok = !(v[0] != v[0]);                               // NaN = a peer did not arrive (error flag bit 1 is set)
        if (scalars != nullptr) {
            scalars[AUG_S_EXPECTED_LOGTILT] = v[0];
            scalars[AUG_S_KL] = v[1];
            scalars[AUG_S_EXPECTED_AUGLL] = v[0] + v[1];      // generic.jl:52-54 ("+")
            scal_zero_except(scalars, 0x07u);
        }
    }
    __syncthreads();
    const bool ok = s_ok != 0;
    const unsigned long long* mine = x->box[me] + AUG_XCH_WORDS + half;
    // gather: 8 outputs per thread and step, all their loads in flight before the first add (system-scope loads are
    // not cached: one dependent load per add would cost ~1 µs each)
    constexpr int GU = 8;
    for (int o0 = threadIdx.x; o0 < nout; o0 += blockDim.x * GU) {
        double t[GU];
#pragma unroll
        for (int u = 0; u < GU; ++u) t[u] = 0.0;
        for (int r = 0; r < nr; ++r) {                        // rank order: identical bits on every rank
            unsigned long long w[GU];
#pragma unroll
            for (int u = 0; u < GU; ++u) {
                const int o = o0 + u * blockDim.x;
                w[u] = 0ull;
                if (o < nout)
                    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(w[u]) : "l"(mine + (size_t)r * AUG_XCH_BULK + o) : "memory");
            }
#pragma unroll
            for (int u = 0; u < GU; ++u) t[u] += __longlong_as_double((long long)w[u]);
        }
#pragma unroll
        for (int u = 0; u < GU; ++u) {
            const int o = o0 + u * blockDim.x;
            if (o < nout) {
                double v = t[u];
                if (o < m * m) v += P0 ? P0[o] : 0.0;
                else v += r0 ? r0[o - m * m] : 0.0;
                Pr[o] = ok ? v : __longlong_as_double(0x7ff8000000000000ll);
            }
        }
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
    static bool configured_dev[64] = {false};   // the attribute is per device: one flag per ordinal
    bool& configured = configured_dev[ctx->device & 63];
    if (!configured) {
        AUG_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
        configured = true;
    }
    a.ntiles = (a.n + C::TO - 1) / C::TO;
    int64_t grid = ctx->sms;
    if (grid > a.ntiles) grid = a.ntiles;
    if (grid < 1) grid = 1;
    const size_t need = ((size_t)grid * C::KG * (MT * MT) + (size_t)grid * C::KG * MT + (size_t)grid * 2) * sizeof(double);
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
    AugXchDev* xch = want_P ? aug_xch_for(ctx) : nullptr;    // fused multi-GPU mode: the verb is collective
    if (xch) { int32_t rf = aug_xch_flush(ctx); if (rf) return rf; }
    if (xch) {
        const int nout = a.m * a.m + a.m;
        sparse_finalize_xch_kernel<MT><<<(nout + 7) / 8, 256, 0, ctx->stream>>>(a.scratch, (int)grid, a.m, P0, r0, Pr,
                                                                               want_s ? scalars : nullptr, C::KG, xch,
                                                                               ctx->counter + 1);
        ctx->launches++;
        AUG_CUDA(cudaGetLastError());
    } else if (want_P || want_s) {
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

// ================================================================== m > 128: chunked library composition
// Same three steps with cuBLAS doing the GEMM-shaped parts (plain library GEMMs, bound with dlopen so that
// libaugcuda.so has no link-time dependency): per chunk of observations
//   T = sym(B)·κ_chunkᵀ (DGEMM) → μ_t, σ²_t = k_tt − κ_t·T_t (row kernel)      [producer]
//   aug_cavi_dispatch over all n (the streaming kernel of §3.1)                  [path]
//   W = κ_chunk scaled by γ_t (row kernel) → Pacc += W·κ_chunkᵀ (DGEMM), racc += κ_chunk·β (DGEMV)   [consumer]
typedef int (*fn_cublas_create)(void**);
typedef int (*fn_cublas_destroy)(void*);
typedef int (*fn_cublas_setstream)(void*, cudaStream_t);
typedef int (*fn_cublas_dgemm)(void*, int, int, int, int, int, const double*, const double*, int, const double*, int,
                               const double*, double*, int);
typedef int (*fn_cublas_dgemv)(void*, int, int, int, const double*, const double*, int, const double*, int,
                               const double*, double*, int);
struct CublasApi {
    void* lib = nullptr;
    fn_cublas_create create = nullptr;
    fn_cublas_destroy destroy = nullptr;
    fn_cublas_setstream setstream = nullptr;
    fn_cublas_dgemm dgemm = nullptr;
    fn_cublas_dgemv dgemv = nullptr;
} g_cublas;

int32_t cublas_load() {
    if (g_cublas.lib) return AUG_OK;
    void* h = dlopen("libcublas.so.12", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);   // the one the process already holds
    if (!h) h = dlopen("libcublas.so.12", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libcublas.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return AUG_ERR_NO_CUBLAS;
    g_cublas.create = (fn_cublas_create)dlsym(h, "cublasCreate_v2");
    g_cublas.destroy = (fn_cublas_destroy)dlsym(h, "cublasDestroy_v2");
    g_cublas.setstream = (fn_cublas_setstream)dlsym(h, "cublasSetStream_v2");
    g_cublas.dgemm = (fn_cublas_dgemm)dlsym(h, "cublasDgemm_v2");
    g_cublas.dgemv = (fn_cublas_dgemv)dlsym(h, "cublasDgemv_v2");
    if (!g_cublas.create || !g_cublas.destroy || !g_cublas.setstream || !g_cublas.dgemm || !g_cublas.dgemv)
        return AUG_ERR_NO_CUBLAS;
    g_cublas.lib = h;
    return AUG_OK;
}
#define AUG_CUBLAS(x)                          \
    do {                                       \
        int s__ = (x);                         \
        if (s__ != 0) return 2000 + s__;       \
    } while (0)

// Bs[i*m + j] = (B[i*m + j] + B[j*m + i]) / 2
__global__ void gen_sym_kernel(int m, const double* __restrict__ B, double* __restrict__ Bs) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e < m * m) {
        const int i = e / m, j = e - i * m;
        Bs[e] = 0.5 * (B[e] + B[(size_t)j * m + i]);
    }
}
// one warp per observation: μ_t = κ_t·m, σ²_t = k_tt − κ_t·T_t
__global__ void gen_rowdot_kernel(int64_t rows, int m, const double* __restrict__ kap, const double* __restrict__ T,
                                  const double* __restrict__ mvec, const double* __restrict__ kdiag,
                                  double* __restrict__ mu, double* __restrict__ var, int64_t ostride) {
    const int lane = threadIdx.x & 31;
    const int64_t w0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t t = w0; t < rows; t += nw) {
        const double* k = kap + t * m;
        const double* tt = T + t * m;
        double q = 0.0, a = 0.0;
        for (int i = lane; i < m; i += 32) {
            const double kv = k[i];
            q = fma(kv, tt[i], q);
            a = fma(kv, __ldg(mvec + i), a);
        }
        q = warp_sum(q);
        a = warp_sum(a);
        if (lane == 0) {
            if (mu) mu[t * ostride] = a;
            if (var) var[t * ostride] = kdiag[t] - q;
        }
    }
}
// W[t][i] = γ_t κ[t][i]
__global__ void gen_scale_kernel(int64_t rows, int m, const double* __restrict__ kap, const double* __restrict__ gamma,
                                 double* __restrict__ W) {
    const int64_t tot = rows * m, nth = (int64_t)gridDim.x * blockDim.x;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < tot; e += nth) W[e] = kap[e] * gamma[e / m];
}
// Pr[i*m + j] = Pacc(max(i,j), min(i,j)) + P0, Pr[m*m + i] = racc[i] + r0   (Pacc column-major, lower triangle used)
__global__ void gen_finalize_kernel(int m, const double* __restrict__ Pacc, const double* __restrict__ racc,
                                    const double* __restrict__ P0, const double* __restrict__ r0, double* __restrict__ Pr) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e < m * m) {
        const int i = e / m, j = e - i * m;
        const int hi = i > j ? i : j, lo = i > j ? j : i;
        Pr[e] = Pacc[(size_t)lo * m + hi] + (P0 ? P0[e] : 0.0);
    } else if (e < m * m + m) {
        const int i = e - m * m;
        Pr[e] = racc[i] + (r0 ? r0[i] : 0.0);
    }
}

// mode: SP_PRODUCER / SP_CONSUMER / SP_FUSED.  All per-observation outputs may be null (kept in scratch when a later
// step needs them).
int32_t sparse_general(aug_ctx* ctx, int mode, const aug_lik* lik, int64_t n, int m, const void* y, const double* kappa,
                       const double* mvec, const double* B, const double* kdiag, const double* gamma_in,
                       const double* beta_in, double* mu, double* var, void* s0, void* s1, void* s2, double* beta,
                       double* gamma, const double* P0, const double* r0, double* Pr, double* scalars,
                       int64_t ostride = 1) {
    if (mode != SP_PRODUCER && aug_xch_for(ctx)) return AUG_ERR_PRECONDITION;   // the in-kernel exchange covers m <= 128
    int32_t rc = cublas_load();
    if (rc) return rc;
    if (!ctx->cublas) AUG_CUBLAS(g_cublas.create(&ctx->cublas));
    AUG_CUBLAS(g_cublas.setstream(ctx->cublas, ctx->stream));
    const bool prod = mode != SP_CONSUMER, cons = mode != SP_PRODUCER, fused = mode == SP_FUSED;
    int64_t ch = ((int64_t)1 << 25) / m;             // chunk: 256 MB of T
    if (ch < 1024) ch = 1024;
    if (ch > n) ch = n > 0 ? n : 1;
    // scratch: Bs | Pacc | racc | T[ch*m] | mu[n] var[n] beta[n] gamma[n] (only what the caller did not provide)
    size_t off = 0;
    auto take = [&](size_t doubles) { size_t o = off; off += (doubles + 1) & ~(size_t)1; return o; };
    const size_t o_bs = take((size_t)m * m), o_pacc = take((size_t)m * m), o_racc = take((size_t)m);
    const size_t o_t = take((size_t)ch * m);
    const size_t o_mu = (fused && !mu) ? take((size_t)n) : 0, o_var = (fused && !var) ? take((size_t)n) : 0;
    const size_t o_be = (fused && !beta) ? take((size_t)n) : 0, o_ga = (fused && !gamma) ? take((size_t)n) : 0;
    const size_t need = off * sizeof(double);
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
    double* S = ctx->sparse_scratch;
    double *Bs = S + o_bs, *Pacc = S + o_pacc, *racc = S + o_racc, *T = S + o_t;
    double* mu_w = mu ? mu : (fused ? S + o_mu : nullptr);
    double* var_w = var ? var : (fused ? S + o_var : nullptr);
    double* be_w = fused ? (beta ? beta : S + o_be) : nullptr;
    double* ga_w = fused ? (gamma ? gamma : S + o_ga) : nullptr;
    const double one = 1.0, zero = 0.0;
    const int grid = ctx->sms * 8;
    if (prod && n > 0) {
        gen_sym_kernel<<<(m * m + 255) / 256, 256, 0, ctx->stream>>>(m, B, Bs);
        ctx->launches++;
        for (int64_t r0i = 0; r0i < n; r0i += ch) {
            const int64_t rows = n - r0i < ch ? n - r0i : ch;
            const double* kc = kappa + r0i * m;
            // column-major view: κ_chunkᵀ is m × rows (ld m); T = Bs · κ_chunkᵀ
            AUG_CUBLAS(g_cublas.dgemm(ctx->cublas, 0, 0, m, (int)rows, m, &one, Bs, m, kc, m, &zero, T, m));
            gen_rowdot_kernel<<<grid, 256, 0, ctx->stream>>>(rows, m, kc, T, mvec, kdiag + r0i,
                                                             mu_w ? mu_w + r0i * ostride : nullptr,
                                                             var_w ? var_w + r0i * ostride : nullptr, ostride);
            ctx->launches++;
        }
        AUG_CUDA(cudaGetLastError());
    }
    if (fused) {
        rc = aug_cavi_dispatch(ctx, lik, n, y, mu_w, var_w, 0, s0, s1, s2, nullptr, nullptr, nullptr, be_w, ga_w, n, scalars,
                               false);
        if (rc) return rc;
    }
    if (cons) {
        const double* g = fused ? ga_w : gamma_in;
        const double* b = fused ? be_w : beta_in;
        AUG_CUDA(cudaMemsetAsync(Pacc, 0, sizeof(double) * ((size_t)m * m), ctx->stream));
        AUG_CUDA(cudaMemsetAsync(racc, 0, sizeof(double) * (size_t)m, ctx->stream));
        for (int64_t r0i = 0; r0i < n; r0i += ch) {
            const int64_t rows = n - r0i < ch ? n - r0i : ch;
            const double* kc = kappa + r0i * m;
            gen_scale_kernel<<<grid, 256, 0, ctx->stream>>>(rows, m, kc, g + r0i, T);
            ctx->launches++;
            // Pacc (m × m, column-major) += W · κ_chunk  with W = T viewed as m × rows
            AUG_CUBLAS(g_cublas.dgemm(ctx->cublas, 0, 1, m, m, (int)rows, &one, T, m, kc, m, &one, Pacc, m));
            AUG_CUBLAS(g_cublas.dgemv(ctx->cublas, 0, m, (int)rows, &one, kc, m, b + r0i, 1, &one, racc, 1));
        }
        gen_finalize_kernel<<<(m * m + m + 255) / 256, 256, 0, ctx->stream>>>(m, Pacc, racc, P0, r0, Pr);
        ctx->launches++;
        AUG_CUDA(cudaGetLastError());
    }
    return AUG_OK;
}

int32_t sp_check(aug_ctx* ctx, int64_t n, int32_t m, const double* kappa) {
    if (!ctx) return AUG_ERR_NOT_INIT;
    if (n < 0 || m < 1) return AUG_ERR_BAD_ARG;
    if (n > 0 && kappa == nullptr) return AUG_ERR_BAD_ARG;
    return AUG_OK;
}

}  // namespace

void aug_cublas_destroy(aug_ctx* c) {
    if (c->cublas && g_cublas.destroy) g_cublas.destroy(c->cublas);
    c->cublas = nullptr;
}

extern "C" {

int32_t aug_sparse_marginals_strided(aug_ctx* ctx, int64_t n, int32_t m, const double* kappa, const double* mvec,
                                     const double* B, const double* kdiag, double* mu, double* var, int64_t stride) {
    int32_t rc = sp_check(ctx, n, m, kappa);
    if (rc) return rc;
    if (n == 0) return AUG_OK;
    if (!mvec || !B || !kdiag || (!mu && !var) || stride < 1) return AUG_ERR_BAD_ARG;
    AUG_CUDA(cudaSetDevice(ctx->device));
    if (m > 128)
        return sparse_general(ctx, SP_PRODUCER, nullptr, n, m, nullptr, kappa, mvec, B, kdiag, nullptr, nullptr, mu, var,
                              nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, stride);
    SparseArgs a{};
    a.n = n; a.m = m; a.kappa = kappa; a.mvec = mvec; a.B = B; a.kdiag = kdiag; a.mu = mu; a.var = var;
    a.ostride = stride;
    a.vec16 = (aug_aligned16(kappa) && (m % 2 == 0)) ? 1 : 0;
    return sp_dispatch_m<SP_PRODUCER, AUG_BERNOULLI>(ctx, a, nullptr, nullptr, nullptr, nullptr);
}

int32_t aug_sparse_marginals(aug_ctx* ctx, int64_t n, int32_t m, const double* kappa, const double* mvec,
                             const double* B, const double* kdiag, double* mu, double* var) {
    return aug_sparse_marginals_strided(ctx, n, m, kappa, mvec, B, kdiag, mu, var, 1);
}

int32_t aug_sparse_precision_potential(aug_ctx* ctx, int64_t n, int32_t m, const double* kappa,
                                       const double* gamma, const double* beta, const double* P0,
                                       const double* r0, double* Pr) {
    int32_t rc = sp_check(ctx, n, m, kappa);
    if (rc) return rc;
    if (!Pr || (n > 0 && (!gamma || !beta))) return AUG_ERR_BAD_ARG;
    AUG_CUDA(cudaSetDevice(ctx->device));
    if (m > 128)
        return sparse_general(ctx, SP_CONSUMER, nullptr, n, m, nullptr, kappa, nullptr, nullptr, nullptr, gamma, beta, nullptr,
                              nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, P0, r0, Pr, nullptr);
    SparseArgs a{};
    a.n = n; a.m = m; a.kappa = kappa; a.gamma_in = gamma; a.beta_in = beta;
    a.ostride = 1;
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
    if (m > 128) {
        if (lik->kind == AUG_HETERO || lik->kind == AUG_CAT || lik->kind == AUG_CAT_BIJ) return AUG_ERR_BAD_KIND;
        return sparse_general(ctx, SP_FUSED, lik, n, m, y, kappa, mvec, B, kdiag, nullptr, nullptr, mu, var, s0, s1, s2, beta,
                              gamma, P0, r0, Pr, scalars);
    }
    SparseArgs a{};
    a.n = n; a.m = m; a.y = y; a.kappa = kappa; a.mvec = mvec; a.B = B; a.kdiag = kdiag;
    a.mu = mu; a.var = var; a.s0 = (double*)s0; a.s1 = (double*)s1; a.s2 = s2; a.beta = beta; a.gamma = gamma;
    a.ostride = 1;
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
