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
#include "aug_pgb.cuh"

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
    AugXchDev* xch;          // final launch of a verb in fused multi-GPU mode (aug_common.cuh)
    int xch_defer;           // split-phase exchange: publish only (aug_comm_set_deferred)
    unsigned int* dflag;
    LikConst L;
};

// inputs for which the straight-line instantiation is valid (exp_ needs |(-m-c)/2| <= 708)
__device__ __forceinline__ bool cat_fast_ok(double m, double v) {
    const double s2 = fma(m, m, v);   // s2 <= 2.4e5 bounds |m| and c by 490: |(-m-c)/2| <= 490
    return s2 >= 1e-290 && s2 <= 2.4e5;
}

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
            } else if (cat_fast_ok(m, v)) {
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
                    else if (a.L.quirks) sx3 += __longlong_as_double(0x7ff8000000000000ll);          // 0 * -Inf as the reference writes it
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
            if (a.xch && a.xch_defer) xch_publish_deferred(a.xch, a.scalars, AUG_S_EXPECTED_LOGTILT, out, 3);
            else if (a.xch) xch_allreduce<3>(a.xch, out);
            a.scalars[AUG_S_EXPECTED_LOGTILT] = out[0];
            a.scalars[AUG_S_KL] = out[1];
            a.scalars[AUG_S_EXPECTED_AUGLL] = out[0] + out[1];
            a.scalars[AUG_S_FLAGS] = out[2];
            scal_zero_except(a.scalars, 0x47u);
            if (out[2] > 0.0) atomicOr(a.dflag, 1u);
        }
    } else if (acc[2] > 0.0) {
        atomicOr(a.dflag, 1u);
    }
}

// ------------------------------------------------------------------ bulk-async (TMA) staged variant
// ncu on cat_cavi_kernel (profiles/r1d): 490 issued instructions per (obs, class) element — an integer
// division per element in phases 1 and 3, a log per element in phase 2, scalar 8-byte loads whose latency
// (long-scoreboard = 30% of the stall cycles) only occupancy could hide — so the kernel was issue-bound at
// 2.3-3.5 TB/s.  This variant, for the full tiles of a fused call on 16-byte aligned arrays:
//   * an elected thread streams each tile (R rows = ONE contiguous span of y, mu, var) into a shared-memory ring
//     with cp.async.bulk completing on mbarriers (SASS: UBLKCP / SYNCS), 2 CTAs per SM;
//   * phase 1 handles two consecutive elements per thread per iteration (128-bit LDS / st.global.v2, the two
//     straight-line evaluations interleave), with the (row, class) index advanced incrementally;
//   * log p_ij comes from quantities already at hand (log σ̃ = (−m−c)/2 − log(1+e^{-c})), so phase 2 is pure
//     additions: Σp, Σ p(−ln2 − m/2 − s2 h/2), Σ p(lch − c²h/2 + log p − log p_prior) per row;
//   * phase 3 walks 16-row groups (one 128-byte segment per class and group) with shifts instead of divisions.
struct CatTmaArgs {
    CatArgs a;               // a.n = rows covered by full tiles
    int64_t ntiles;
    int E;                   // elements per tile = R * nl (even; a multiple of 16 bytes of y)
    int S;                   // ring depth
    int off_mu, off_var, stage_bytes;
    int di, dj;              // (2*AUG_BLOCK) / nl and % nl: row/class step of a thread between phase-1 iterations
    int rs;                  // row stride of the staged tiles in doubles: nl (odd nl) or nl + 2 (even nl)
    int accumulate;          // add onto the scalars the ragged-tail launch left in memory
    double log_inv_denom;    // log(1/(D + nl)) resp. log(1/nl)
};

// One (obs, class) element.  SAFE = false is straight-line code valid for 2^-8 <= m² + v <= 2.4e5 (so that
// c >= 1/16: no series branch, |(-m-c)/2| <= 490: exp_ in range); SAFE = true takes any input.
template <bool ELBO, bool SAFE>
__device__ __forceinline__ void cat_elem(const double m, const double v, const bool yb, const double inv_denom,
                                         const double log_inv_denom, const double c2, double& c, double& p,
                                         double& h, double& x1, double& x23, double& a0, double& a1,
                                         const bool quirks = false) {
    using namespace augm;
    const double s2m = fma(m, m, v);
    PGTerms t;
    double sig, logp;
    if (SAFE) {
        c = sqrt(s2m);                                                    // categorical.jl:88,105
        t = pg_terms_ic<ELBO, true>(c, 0.0);
        sig = approx_expected_logistic<true>(-m, c, t);                   // :90-92, :107
        p = sig * inv_denom;
        // p == 0 (σ̃ saturated): the limit p log p = 0, or the reference's 0 * -Inf = NaN (negativemultinomial.jl:80)
        logp = p > 0.0 ? log(p) : (quirks ? __longlong_as_double(0x7ff8000000000000ll) : 0.0);
    } else {
        double ic;
        augf::sqrt_inv(s2m, c, ic);
        t = pg_terms_ic<ELBO, false, false>(c, ic);
        const double ah = 0.5 * (-m - c);
        const double w = augf::exp_(ah) * t.inv1pe;                       // exp(-m/2) sech(c/2)/2
        const bool hi = -m > LOGISTIC_HI;                                 // utils.jl:12-13 (|m| <= 490: never below LO)
        sig = hi ? 1.0 : w;
        p = sig * inv_denom;
        logp = (hi ? 0.0 : ah - t.l1pe) + log_inv_denom;                  // log p without a log
    }
    h = t.h;
    if (ELBO) {
        const double hb = 0.5 * s2m * t.h;
        const double d = fma(-0.5 * c * c, t.h, t.lch);
        x1 = p * (-LN2 - 0.5 * m - hb);
        x23 = p * (d + (logp - c2));                                      // + q_j (log q_j − log p_j), negativemultinomial.jl:79-81
        // row-independent parts: y_j (−ln2 + m/2 − s2 h/2) and y_j (lch − c² h/2); y is 0/1 (selects, no I2F)
        a0 += yb ? (-LN2 + 0.5 * m - hb) : 0.0;
        a1 += yb ? d : 0.0;
    }
}

// out-of-range test of the straight-line instantiation on the high word of m² + v (integer pipe; NaN, negative
// and infinite values land outside too): not in [2^-8, 2.4e5)
__device__ __forceinline__ bool cat_slow(double m, double v) {
    const unsigned hi = (unsigned)__double2hiint(fma(m, m, v));
    return hi - 0x3f700000u >= 0x410d4c00u - 0x3f700000u;
}

// The element of cat_row_kernel's straight-line pass: cat_slow(m, v) false and the logistic not saturated (-m <= LOGISTIC_HI).
// c, p, h carry the same operations as cat_elem<., false>; the ELBO terms are regrouped so that no logarithm is left in
// the per-element work.  With A = -ln2 - s2 h/2 (s2 = m^2 + v = c^2), lid = log(1/denominator):
//   expected_logtilt:  sum_j nbar_j (A - m/2) + y_j (A + m/2)                                  (nbar_j = p_j / p0)
//   PG KL:             sum_j (y_j + nbar_j) (logcosh(c/2) - c^2 h/2),  logcosh(c/2) = c/2 + log(1 + e^-c) - ln2
//   NM KL:             sum_j nbar_j (log p_j - log prior_j),           log p_j = (-m - c)/2 - log(1 + e^-c) + lid
// the log(1 + e^-c) of the nbar-proportional parts cancel: sum_j nbar_j [KL terms] = sum_j nbar_j (A - m/2) + (lid - log prior) sum_j nbar_j,
// i.e. the row sum the expected_logtilt needs anyway plus a multiple of sum_j p_j.  What is left of the logarithm is
// sum_{j: y_j = 1} log(1 + e^-c_j): the factors (1 + e^-c_j) are multiplied up (one per row for one-hot y, at most 25 per lane and
// tile) and ONE log of the product is taken per lane and tile.
template <bool ELBO>
__device__ __forceinline__ void cat_elem_row(const double m, const double v, const bool yb, const double inv_denom, double& c,
                                             double& p, double& h, double& A1, double& t0, double& t1, double& prod) {
    const double s2m = fma(m, m, v);
    double ic;
    augf::sqrt_inv(s2m, c, ic);                                           // categorical.jl:88,105
    const double e = augf::exp_(-fmin(c, 708.0));
    const double inv = augf::rcp(1.0 + e);
    h = (1.0 - e) * inv * (0.5 * ic);                                     // tanh(c/2)/(2c)
    const double w = augf::exp_(0.5 * (-m - c)) * inv;                    // :90-92, :107 (utils.jl:11-14, not saturated)
    p = w * inv_denom;
    if (ELBO) {
        const double A = fma(-0.5 * s2m, h, -augm::LN2);
        const double hm = 0.5 * m;
        const double yd = yb ? 1.0 : 0.0;
        A1 = fma(p, A - hm, A1);
        t0 = fma(yd, A + hm, t0);
        t1 = fma(yd, fma(0.5, c, A), t1);
        prod *= fma(yd, e, 1.0);
    }
}
// the saturated logistic (-m > LOGISTIC_HI: sigma~ = 1, utils.jl:12-13) goes with the any-input instantiation in cat_row_kernel
__device__ __forceinline__ bool cat_row_slow(double m, double v) { return cat_slow(m, v) || -m > augm::LOGISTIC_HI; }

// EVEN: nl is even -> element pairs never straddle rows and the staged rows are padded by two doubles
// (conflict-free-enough transposed reads); odd nl -> the staged tiles are the plain linear span.
template <bool ELBO, bool EVEN>
__global__ void __launch_bounds__(AUG_BLOCK, 2) cat_tma_kernel(const CatTmaArgs ta) {
    const CatArgs& a = ta.a;
    extern __shared__ __align__(128) unsigned char cat_ring[];
    const int nl = a.nl, R = a.R, E = ta.E, S = ta.S, rs = ta.rs;
    const int tile_sz = R * rs;
    unsigned char* ring = cat_ring;
    double* P = reinterpret_cast<double*>(ring + (size_t)S * ta.stage_bytes);
    double* H = P + tile_sz;
    double* X1 = ELBO ? H + tile_sz : nullptr;
    double* X23 = ELBO ? X1 + tile_sz : nullptr;
    double* rinv = ELBO ? X23 + tile_sz : H + tile_sz;
    uint64_t* full = reinterpret_cast<uint64_t*>(rinv + R);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        for (int s = 0; s < S; ++s) mbar_init(&full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    auto issue = [&](int64_t tile, int s) {
        unsigned char* st = ring + (size_t)s * ta.stage_bytes;
        const int64_t o = tile * E;
        mbar_expect_tx(&full[s], (uint32_t)E * 17u);
        bulk_g2s(st, a.y + o, (uint32_t)E, &full[s]);
        bulk_g2s(st + ta.off_mu, a.mu + o, (uint32_t)E * 8u, &full[s]);
        bulk_g2s(st + ta.off_var, a.var + o, (uint32_t)E * 8u, &full[s]);
    };
    const int64_t first = blockIdx.x, stride = gridDim.x;
    if (tid == 0) {
        for (int s = 0; s < S; ++s) {
            const int64_t t = first + (int64_t)s * stride;
            if (t < ta.ntiles) issue(t, s);
        }
    }
    const int i_first = (2 * tid) / nl, j_first = 2 * tid - i_first * nl;   // row, class of element 2*tid
    const int l8 = lane & 7, l16 = lane & 15;
    const int jj_first = 4 * warp + (lane >> 3);                            // phase 3: class of this thread
    const double inv_denom = 1.0 / a.L.c0;
    const bool vec_out = ((a.ldo & 1) == 0) && ((((uintptr_t)a.beta) | ((uintptr_t)a.gamma)) & 15u) == 0;
    const bool both_out = a.beta != nullptr && a.gamma != nullptr && vec_out;
    double acc[3] = {0.0, 0.0, 0.0};
    int s = 0;
    uint32_t parity = 0;
    for (int64_t tile = first; tile < ta.ntiles; tile += stride) {
        mbar_wait(&full[s], parity);
        const unsigned char* st = ring + (size_t)s * ta.stage_bytes;
        const uchar2* sy = reinterpret_cast<const uchar2*>(st);
        const double2* smu = reinterpret_cast<const double2*>(st + ta.off_mu);
        const double2* svar = reinterpret_cast<const double2*>(st + ta.off_var);
        // ---- phase 1: element pairs, straight-line
        {
            double* ps0 = a.s0 ? a.s0 + tile * E + 2 * tid : nullptr;
            double* ps1 = a.s1 ? a.s1 + tile * E + 2 * tid : nullptr;
            uint8_t* ps2 = a.s2 ? a.s2 + tile * E + 2 * tid : nullptr;
            int i = i_first, j = j_first;
            bool bad = false;
            double t0 = 0.0, t1 = 0.0;   // this tile's ELBO terms of the thread (discarded if the tile is redone)
#pragma unroll 1
            for (int q = tid; q < (E >> 1); q += AUG_BLOCK) {
                const uchar2 yy = sy[q];
                const double2 m = smu[q], v = svar[q];
                const bool y0 = yy.x != 0, y1 = yy.y != 0;
                bad = bad || cat_slow(m.x, v.x) || cat_slow(m.y, v.y);
                double c0, p0, h0, c1, p1, h1, xa0 = 0.0, xb0 = 0.0, xa1 = 0.0, xb1 = 0.0;
                cat_elem<ELBO, false>(m.x, v.x, y0, inv_denom, ta.log_inv_denom, a.L.c2, c0, p0, h0, xa0, xb0, t0, t1);
                cat_elem<ELBO, false>(m.y, v.y, y1, inv_denom, ta.log_inv_denom, a.L.c2, c1, p1, h1, xa1, xb1, t0, t1);
                if (ps0) { st_stream2(ps0, c0, c1); ps0 += 2 * AUG_BLOCK; }
                if (ps1) { st_stream2(ps1, p0, p1); ps1 += 2 * AUG_BLOCK; }
                if (ps2) { *reinterpret_cast<uchar2*>(ps2) = yy; ps2 += 2 * AUG_BLOCK; }   // φᵢ.y .= y[i]  categorical.jl:89,106
                const int se = 2 * q + (EVEN ? 2 * i : 0);
                *reinterpret_cast<double2*>(P + se) = make_double2(p0, p1);
                // h = tanh(c/2)/(2c) > 0 always: its sign bit carries y_ij to phase 3 (no separate y tile, no I2F)
                *reinterpret_cast<double2*>(H + se) = make_double2(y0 ? -h0 : h0, y1 ? -h1 : h1);
                if (ELBO) {
                    *reinterpret_cast<double2*>(X1 + se) = make_double2(xa0, xa1);
                    *reinterpret_cast<double2*>(X23 + se) = make_double2(xb0, xb1);
                }
                if (EVEN) {
                    i += ta.di;
                    j += ta.dj;
                    if (j >= nl) { j -= nl; ++i; }
                }
            }
            if (!bad) {
                acc[0] += t0;
                acc[1] += t1;
            } else {
                // rare: some element of this thread is outside the straight-line range -> redo the thread's pairs of
                // this tile with the any-input instantiation (the stage is still resident; outputs are overwritten)
                i = i_first;
                j = j_first;
                for (int q = tid; q < (E >> 1); q += AUG_BLOCK) {
                    const uchar2 yy = sy[q];
                    const double2 m = smu[q], v = svar[q];
                    const bool y0 = yy.x != 0, y1 = yy.y != 0;
                    double c0, p0, h0, c1, p1, h1, xa0 = 0.0, xb0 = 0.0, xa1 = 0.0, xb1 = 0.0;
                    cat_elem<ELBO, true>(m.x, v.x, y0, inv_denom, ta.log_inv_denom, a.L.c2, c0, p0, h0, xa0, xb0, acc[0], acc[1], a.L.quirks != 0);
                    cat_elem<ELBO, true>(m.y, v.y, y1, inv_denom, ta.log_inv_denom, a.L.c2, c1, p1, h1, xa1, xb1, acc[0], acc[1], a.L.quirks != 0);
                    const int64_t o = tile * E + 2 * q;
                    if (a.s0) st_stream2(a.s0 + o, c0, c1);
                    if (a.s1) st_stream2(a.s1 + o, p0, p1);
                    const int se = 2 * q + (EVEN ? 2 * i : 0);
                    *reinterpret_cast<double2*>(P + se) = make_double2(p0, p1);
                    *reinterpret_cast<double2*>(H + se) = make_double2(y0 ? -h0 : h0, y1 ? -h1 : h1);
                    if (ELBO) {
                        *reinterpret_cast<double2*>(X1 + se) = make_double2(xa0, xa1);
                        *reinterpret_cast<double2*>(X23 + se) = make_double2(xb0, xb1);
                    }
                    if (EVEN) {
                        i += ta.di;
                        j += ta.dj;
                        if (j >= nl) { j -= nl; ++i; }
                    }
                }
            }
        }
        __syncthreads();   // P/H/X complete; every thread is done reading stage s
        if (tid == 0) {
            const int64_t nt = tile + (int64_t)S * stride;
            if (nt < ta.ntiles) issue(nt, s);
        }
        if (++s == S) { s = 0; parity ^= 1u; }
        // ---- phase 2: 16 lanes per row (two rows per warp step): Σ_j p_ij -> 1/p0, log p0
        for (int r = 2 * warp + (lane >> 4); r < R; r += 2 * (AUG_BLOCK / 32)) {
            const double* Pr = P + r * rs;
            double sp = 0.0;
            for (int jj = l16; jj < nl; jj += 16) sp += Pr[jj];
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) sp += __shfl_xor_sync(0xffffffffu, sp, o);
            if (l16 == 0) {
                const double p0 = 1.0 - sp;                               // _p₀ negativemultinomial.jl:27
                const bool ok = p0 >= 1e-290;
                rinv[r] = ok ? augf::rcp(p0) : 1.0 / p0;
                if (!(sp < 1.0)) acc[2] += 1.0;                           // ctor precondition :18-22
                // KL(NM(1,q)||NM(1,p)) = log p0q − log p0p + (1/p0q) Σ q_j (log q_j − log p_j)  :72-82
                if (ELBO) acc[1] += (ok ? augf::log_(p0) : log(p0)) - a.L.c3;
            }
        }
        __syncthreads();
        // ---- phase 3: class-major; 8 lanes x 2 rows = 16 consecutive rows = one 128-byte segment per array;
        //      a warp covers 4 classes, the CTA 32 classes per step
        if (ELBO || a.beta || a.gamma) {
            for (int g = 0; g < (R >> 4); ++g) {
                const int r = (g << 4) + 2 * l8;
                const double2 ri = *reinterpret_cast<const double2*>(rinv + r);
                const double* Pp = P + r * rs + jj_first;
                const double* Hp = H + r * rs + jj_first;
                int64_t o = (int64_t)jj_first * a.ldo + (tile * R + r);
                const int64_t ostep = 4 * (AUG_BLOCK / 32) * a.ldo;
#pragma unroll 2
                for (int jj = jj_first; jj < nl; jj += 4 * (AUG_BLOCK / 32)) {
                    const double hs0 = Hp[0], hs1 = Hp[rs];
                    const double y0 = __double2hiint(hs0) < 0 ? 1.0 : 0.0, y1 = __double2hiint(hs1) < 0 ? 1.0 : 0.0;
                    const double n0 = Pp[0] * ri.x, n1 = Pp[rs] * ri.y;                  // mean(NM(1,p)) :54
                    const double b0 = 0.5 * (y0 - n0), b1 = 0.5 * (y1 - n1);             // categorical.jl:124,135
                    const double g0 = (y0 + n0) * fabs(hs0), g1 = (y1 + n1) * fabs(hs1); // :128; pgnm.jl:41-54
                    if (both_out) {
                        st_stream2(a.beta + o, b0, b1);
                        st_stream2(a.gamma + o, g0, g1);
                    } else if (vec_out) {
                        if (a.beta) st_stream2(a.beta + o, b0, b1);
                        if (a.gamma) st_stream2(a.gamma + o, g0, g1);
                    } else {
                        if (a.beta) { st_stream1(a.beta + o, b0); st_stream1(a.beta + o + 1, b1); }
                        if (a.gamma) { st_stream1(a.gamma + o, g0); st_stream1(a.gamma + o + 1, g1); }
                    }
                    if (ELBO) {   // the n̄-proportional parts of expected_logtilt and KL
                        const double* X1p = X1 + (Pp - P);
                        const double* X23p = X23 + (Pp - P);
                        acc[0] = fma(ri.x, X1p[0], fma(ri.y, X1p[rs], acc[0]));
                        acc[1] = fma(ri.x, X23p[0], fma(ri.y, X23p[rs], acc[1]));
                    }
                    Pp += 4 * (AUG_BLOCK / 32);
                    Hp += 4 * (AUG_BLOCK / 32);
                    o += ostep;
                }
            }
        }
        __syncthreads();   // P/H/rinv are rewritten by the next tile
    }
    if (ELBO) {
        double out[3];
        if (block_reduce_and_finalize<3>(acc, a.partials, a.counter, out)) {
            if (ta.accumulate) {
                out[0] += a.scalars[AUG_S_EXPECTED_LOGTILT];
                out[1] += a.scalars[AUG_S_KL];
                out[2] += a.scalars[AUG_S_FLAGS];
            }
            if (a.xch && a.xch_defer) xch_publish_deferred(a.xch, a.scalars, AUG_S_EXPECTED_LOGTILT, out, 3);
            else if (a.xch) xch_allreduce<3>(a.xch, out);
            a.scalars[AUG_S_EXPECTED_LOGTILT] = out[0];
            a.scalars[AUG_S_KL] = out[1];
            a.scalars[AUG_S_EXPECTED_AUGLL] = out[0] + out[1];
            a.scalars[AUG_S_FLAGS] = out[2];
            scal_zero_except(a.scalars, 0x47u);
            if (out[2] > 0.0) atomicOr(a.dflag, 1u);
        }
    } else if (acc[2] > 0.0) {
        atomicOr(a.dflag, 1u);
    }
}

// ------------------------------------------------------------------ row-aligned staged variant (round 1e)
// ncu on cat_tma_kernel (profiles/r1d): issue slots 53 % busy, stalls on the CTA-wide barriers between the three
// phases; with K = 100 a 256-thread CTA also leaves 22 % of phase 1 and phase 3 idle (800 pairs / 256 threads =
// 3.125 -> 4 iterations).  This variant removes both:
//   * a CTA is TWO warps and owns ONE tile of 16 rows; 7-8 CTAs are resident per SM, so a barrier only ever stalls
//     two warps and the other CTAs of the SM are in different phases;
//   * phase 1 is row-aligned: 4 lanes per row walk the row's element pairs, so Σ_j p_ij and the two row-wise ELBO
//     sums are thread-private and finish with two shuffles - phase 2 and the X1/X23 staging arrays disappear, and
//     the ELBO instantiation needs no more shared memory than the plain one;
//   * P and ±h overwrite mu and var IN PLACE in the staged tile (17 B of shared memory per element in flight
//     instead of 50-66), which is what lets 8 tiles be resident per SM; the stage is refilled (cp.async.bulk on an
//     mbarrier, one 800-byte row per copy into padded rows when nl is even) as soon as phase 3 has read it.
#ifndef CAT_PIPE_DEFAULT
#define CAT_PIPE_DEFAULT 322     // rows per tile * 10 + stages per CTA
#endif
#ifndef CAT_ROW_ILP4
#define CAT_ROW_ILP4 1
#endif
struct CatRowArgs {
    CatArgs a;               // a.n = rows covered by full tiles
    int64_t ntiles;
    int E;                   // elements per tile = 16 * nl
    int rs;                  // row stride of the staged rows in doubles: nl (odd nl) or nl + 2 (even nl)
    int off_mu, off_var, off_rinv;   // within a stage; off_rinv = bytes of one stage, rinv[16] and the mbarriers follow the last stage
    int accumulate;
    double log_inv_denom;
};

// R / 8 warps per CTA work on one R-row tile at a time (R = 16, 32, 64: the class-major stores of phase 3 are R * 8-byte
// segments); NS = 1: the tile is refilled in place after its phase 3 (the
// round-1 kernel: 8 CTAs of two warps per SM hide each other's refill), NS = 2: the CTA owns two stages and the refill of
// one has the whole processing time of the other to land (ncu, profiles/r2l: 26 % of the stall samples of the NS = 1
// kernel sit in the mbarrier wait of the refill).
template <bool ELBO, bool EVEN, int CAT_ROW_R, int NS>
__global__ void __launch_bounds__(CAT_ROW_R * 4, 128 / (CAT_ROW_R * NS) < 1 ? 1 : 128 / (CAT_ROW_R * NS)) cat_row_kernel(const CatRowArgs ta) {
    constexpr int NW = CAT_ROW_R / 8;                                      // 4 lanes per row (8 per row pair when nl is odd)
    constexpr int CAT_ROW_BLOCK = NW * 32;
    const CatArgs& a = ta.a;
    extern __shared__ __align__(128) unsigned char cat_stage0[];
    const int nl = a.nl, E = ta.E, rs = ta.rs;
    double* rinv = reinterpret_cast<double*>(cat_stage0 + (size_t)NS * ta.off_rinv);
    uint64_t* fulls = reinterpret_cast<uint64_t*>(rinv + CAT_ROW_R);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < NS; ++s) mbar_init(&fulls[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const int64_t first = blockIdx.x, stride = gridDim.x;
    // warp 0 streams a tile into a stage: y as one span; mu/var as one span (odd nl) or one padded row per lane
    auto issue = [&](int64_t tile, int s) {
        const int64_t o = tile * E;
        unsigned char* cat_stage = cat_stage0 + (size_t)s * ta.off_rinv;
        double* P = reinterpret_cast<double*>(cat_stage + ta.off_mu);
        double* H = reinterpret_cast<double*>(cat_stage + ta.off_var);
        uint64_t* full = &fulls[s];
        if (lane == 0) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the stage was written through the generic proxy
            mbar_expect_tx(full, (uint32_t)E * 17u);
            bulk_g2s(cat_stage, a.y + o, (uint32_t)E, full);
            if (!EVEN) {
                bulk_g2s(P, a.mu + o, (uint32_t)E * 8u, full);
                bulk_g2s(H, a.var + o, (uint32_t)E * 8u, full);
            }
            // (pulling the CTA's next tile into L2 here with cp.async.bulk.prefetch.L2 shortens the wait on this
            //  single-buffered stage but was measured 12 % SLOWER overall: the prefetched lines compete with the
            //  streaming stores for L2)
        }
        if (EVEN) {
            __syncwarp();
#pragma unroll
            for (int i = lane; i < 2 * CAT_ROW_R; i += 32) {
                const int r = i % CAT_ROW_R;
                const bool is_mu = i < CAT_ROW_R;
                const double* src = (is_mu ? a.mu : a.var) + o + (int64_t)r * nl;
                double* dst = (is_mu ? P : H) + r * rs;
                bulk_g2s(dst, src, (uint32_t)nl * 8u, full);
            }
        }
    };
    if (warp == 0) {
#pragma unroll
        for (int s = 0; s < NS; ++s)
            if (first + s * stride < ta.ntiles) issue(first + s * stride, s);
    }

    constexpr int LPC = CAT_ROW_R / 2;                                     // phase 3: lanes per class (two rows each)
    const int l8 = lane % LPC;
    const int jj_first = (32 / LPC) * warp + lane / LPC;                   // phase 3: class of this thread
    const double inv_denom = 1.0 / a.L.c0;
    const bool vec_out = ((a.ldo & 1) == 0) && ((((uintptr_t)a.beta) | ((uintptr_t)a.gamma)) & 15u) == 0;
    const bool both_out = a.beta != nullptr && a.gamma != nullptr && vec_out;
    double acc[3] = {0.0, 0.0, 0.0};
    uint32_t kt = 0;
    for (int64_t tile = first; tile < ta.ntiles; tile += stride, ++kt) {
        const int st = NS == 1 ? 0 : (int)(kt % NS);
        unsigned char* cat_stage = cat_stage0 + (size_t)st * ta.off_rinv;
        const uint8_t* Y = cat_stage;
        double* P = reinterpret_cast<double*>(cat_stage + ta.off_mu);      // mu, then p
        double* H = reinterpret_cast<double*>(cat_stage + ta.off_var);     // var, then ±h (sign = y)
        mbar_wait(&fulls[st], (kt / NS) & 1u);
        // ---- phase 1: straight-line element pairs, in place.  Even nl: 4 lanes per row (a row is nl/2 aligned pairs).
        //      Odd nl: 8 lanes per ROW PAIR - rows 2k and 2k+1 are one 16-byte aligned span of nl pairs whose pair
        //      h = (nl-1)/2 straddles the two rows; a lane walks the span upwards, so its row sums switch from the
        //      first to the second row exactly once.
        {
            constexpr int LQ = EVEN ? 4 : 8;                               // lanes per row (even nl) / row pair (odd nl)
            const int sub = tid / LQ;                                       // row (even nl) / row pair (odd nl)
            const int kq = tid % LQ;                                        // first pair of this lane; LQ = its pair step
            const int npairs = EVEN ? (nl >> 1) : nl, hq = nl >> 1;         // odd nl: hq = the straddling pair
            const int so = EVEN ? sub * rs : 2 * sub * nl;
            double* Pr = P + so;
            double* Hr = H + so;
            const uint8_t* Yr = Y + (EVEN ? sub * nl : 2 * sub * nl);
            const int64_t gro = tile * E + (int64_t)(EVEN ? sub : 2 * sub) * nl;   // global element offset of the span
            double sp = 0.0, A1 = 0.0, A23 = 0.0, t0 = 0.0, t1 = 0.0;      // running row sums / row-independent ELBO terms
            double spA = 0.0, A1A = 0.0, A23A = 0.0;                        // odd nl: sums of the first row once passed
            bool bad = false, sw = false;
            struct PairIn { uchar2 yy; double2 m, v; };
            struct PairOut { double c0, c1, p0, p1, h0, h1; };
            auto load = [&](const int q) {
                PairIn in;
                in.yy = *reinterpret_cast<const uchar2*>(Yr + 2 * q);
                in.m = *reinterpret_cast<const double2*>(Pr + 2 * q);
                in.v = *reinterpret_cast<const double2*>(Hr + 2 * q);
                return in;
            };
            auto next_row = [&]() {
                spA = sp; sp = 0.0;
                if (ELBO) { A1A = A1; A23A = A23; A1 = 0.0; A23 = 0.0; }
                sw = true;
            };
            double prod = 1.0;                                              // ELBO: product of (1 + e^-c) over this lane's y = 1 elements
            auto eval = [&](const int q, const PairIn& in) {
                PairOut o;
                bad = bad || cat_row_slow(in.m.x, in.v.x) || cat_row_slow(in.m.y, in.v.y);
                if (EVEN) {
                    cat_elem_row<ELBO>(in.m.x, in.v.x, in.yy.x != 0, inv_denom, o.c0, o.p0, o.h0, A1, t0, t1, prod);
                    cat_elem_row<ELBO>(in.m.y, in.v.y, in.yy.y != 0, inv_denom, o.c1, o.p1, o.h1, A1, t0, t1, prod);
                    sp += o.p0 + o.p1;
                } else {
                    if (q > hq && !sw) next_row();
                    cat_elem_row<ELBO>(in.m.x, in.v.x, in.yy.x != 0, inv_denom, o.c0, o.p0, o.h0, A1, t0, t1, prod);
                    sp += o.p0;
                    if (q == hq) next_row();
                    cat_elem_row<ELBO>(in.m.y, in.v.y, in.yy.y != 0, inv_denom, o.c1, o.p1, o.h1, A1, t0, t1, prod);
                    sp += o.p1;
                }
                return o;
            };
            auto store = [&](const int q, const PairIn& in, const PairOut& o) {
                if (a.s0) st_stream2(a.s0 + gro + 2 * q, o.c0, o.c1);
                if (a.s1) st_stream2(a.s1 + gro + 2 * q, o.p0, o.p1);
                if (a.s2) *reinterpret_cast<uchar2*>(a.s2 + gro + 2 * q) = in.yy;    // φᵢ.y .= y[i]  categorical.jl:89,106
                *reinterpret_cast<double2*>(Pr + 2 * q) = make_double2(o.p0, o.p1);
                // h = tanh(c/2)/(2c) > 0 always: its sign bit carries y_ij to phase 3
                *reinterpret_cast<double2*>(Hr + 2 * q) = make_double2(in.yy.x ? -o.h0 : o.h0, in.yy.y ? -o.h1 : o.h1);
            };
            int q = kq;
#if CAT_ROW_ILP4
#pragma unroll 1
            for (; q + LQ < npairs; q += 2 * LQ) {      // two pairs = four independent evaluations in flight
                const PairIn i0 = load(q), i1 = load(q + LQ);
                const PairOut o0 = eval(q, i0), o1 = eval(q + LQ, i1);
                store(q, i0, o0);
                store(q + LQ, i1, o1);
            }
#endif
#pragma unroll 1
            for (; q < npairs; q += LQ) {
                const PairIn i0 = load(q);
                const PairOut o0 = eval(q, i0);
                store(q, i0, o0);
            }
            if (bad) {
                // rare: an element of this thread is outside the straight-line range -> redo the thread's share of the
                // span with the any-input instantiation; the staged inputs are gone (in place), so re-read them
                sp = A1 = A23 = t0 = t1 = spA = A1A = A23A = 0.0;
                sw = false;
                for (q = kq; q < npairs; q += LQ) {
                    for (int u = 0; u < 2; ++u) {
                        const int64_t o = gro + 2 * q + u;
                        const bool yb = a.y[o] != 0;
                        double c, pp, h, xa = 0.0, xb = 0.0;
                        cat_elem<ELBO, true>(a.mu[o], a.var[o], yb, inv_denom, ta.log_inv_denom, a.L.c2, c, pp, h, xa, xb, t0, t1, a.L.quirks != 0);
                        if (a.s0) a.s0[o] = c;
                        if (a.s1) a.s1[o] = pp;
                        Pr[2 * q + u] = pp;
                        Hr[2 * q + u] = yb ? -h : h;
                        if (!EVEN && !sw && 2 * q + u >= nl) next_row();
                        sp += pp;
                        if (ELBO) { A1 += xa; A23 += xb; }
                    }
                }
            }
            if (!EVEN && !sw) next_row();
            if (ELBO && !bad) {
                // the straight-line pass left the NM-KL row sums implicit and the log(1 + e^-c) of the y = 1 elements as a product
                const double K2 = ta.log_inv_denom - a.L.c2;
                A23 = fma(K2, sp, A1);
                if (!EVEN) A23A = fma(K2, spA, A1A);
                t1 += augf::log_(prod);
            }
            acc[0] += t0;
            acc[1] += t1;
            // row sums over the lanes of the row (pair)
#pragma unroll
            for (int o = 1; o < LQ; o <<= 1) {
                sp += __shfl_xor_sync(0xffffffffu, sp, o);
                if (ELBO) {
                    A1 += __shfl_xor_sync(0xffffffffu, A1, o);
                    A23 += __shfl_xor_sync(0xffffffffu, A23, o);
                }
                if (!EVEN) {
                    spA += __shfl_xor_sync(0xffffffffu, spA, o);
                    if (ELBO) {
                        A1A += __shfl_xor_sync(0xffffffffu, A1A, o);
                        A23A += __shfl_xor_sync(0xffffffffu, A23A, o);
                    }
                }
            }
            // even nl: lane 0 of the row; odd nl: lane 0 finishes the first row of the pair, lane 1 the second
            if (kq == 0 || (!EVEN && kq == 1)) {
                const bool second = !EVEN && kq == 1;
                const double srow = (EVEN || second) ? sp : spA;
                const double a1 = (EVEN || second) ? A1 : A1A, a23 = (EVEN || second) ? A23 : A23A;
                const double p0 = 1.0 - srow;                             // _p₀ negativemultinomial.jl:27
                const bool ok = p0 >= 1e-290;
                const double ri = ok ? augf::rcp(p0) : 1.0 / p0;
                rinv[EVEN ? sub : 2 * sub + (second ? 1 : 0)] = ri;
                if (!(srow < 1.0)) acc[2] += 1.0;                         // ctor precondition :18-22
                if (ELBO) {
                    // the n̄-proportional parts of expected_logtilt and of the PG KL, and
                    // KL(NM(1,q)||NM(1,p)) = log p0q − log p0p + (1/p0q) Σ q_j (log q_j − log p_j)  :72-82
                    acc[0] = fma(ri, a1, acc[0]);
                    acc[1] += fma(ri, a23, (ok ? augf::log_(p0) : log(p0)) - a.L.c3);
                }
            }
        }
        __syncthreads();
        // ---- phase 3: class-major; R/2 lanes x 2 rows = the tile's R rows = one R*8-byte segment per array;
        //      the CTA covers 8 classes per step
        if (a.beta || a.gamma) {
            const int r = 2 * l8;
            const double2 ri = *reinterpret_cast<const double2*>(rinv + r);
            const double* Pp = P + r * rs + jj_first;
            const double* Hp = H + r * rs + jj_first;
            // (measured: the same stores to contiguous addresses would make the fused call 9 % faster - the class-major
            //  result layout of the reference, one 128-byte segment per class and tile, is what costs the DRAM pages)
            int64_t o = (int64_t)jj_first * a.ldo + (tile * CAT_ROW_R + r);
            const int64_t ostep = 8 * a.ldo;
#pragma unroll 2
            for (int jj = jj_first; jj < nl; jj += 8) {
                const double hs0 = Hp[0], hs1 = Hp[rs];
                const double y0 = __double2hiint(hs0) < 0 ? 1.0 : 0.0, y1 = __double2hiint(hs1) < 0 ? 1.0 : 0.0;
                const double n0 = Pp[0] * ri.x, n1 = Pp[rs] * ri.y;                  // mean(NM(1,p)) :54
                const double b0 = 0.5 * (y0 - n0), b1 = 0.5 * (y1 - n1);             // categorical.jl:124,135
                const double g0 = (y0 + n0) * fabs(hs0), g1 = (y1 + n1) * fabs(hs1); // :128; pgnm.jl:41-54
                if (both_out) {
                    st_stream2(a.beta + o, b0, b1);
                    st_stream2(a.gamma + o, g0, g1);
                } else if (vec_out) {
                    if (a.beta) st_stream2(a.beta + o, b0, b1);
                    if (a.gamma) st_stream2(a.gamma + o, g0, g1);
                } else {
                    if (a.beta) { st_stream1(a.beta + o, b0); st_stream1(a.beta + o + 1, b1); }
                    if (a.gamma) { st_stream1(a.gamma + o, g0); st_stream1(a.gamma + o + 1, g1); }
                }
                Pp += 8;
                Hp += 8;
                o += ostep;
            }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes of P/H before the async refill
        __syncthreads();   // every thread is done with the stage: refill it
        if (warp == 0 && tile + NS * stride < ta.ntiles) issue(tile + NS * stride, st);
    }
    if (ELBO) {
        double out[3];
        if (block_reduce_and_finalize<3, CAT_ROW_BLOCK>(acc, a.partials, a.counter, out)) {
            if (ta.accumulate) {
                out[0] += a.scalars[AUG_S_EXPECTED_LOGTILT];
                out[1] += a.scalars[AUG_S_KL];
                out[2] += a.scalars[AUG_S_FLAGS];
            }
            if (a.xch && a.xch_defer) xch_publish_deferred(a.xch, a.scalars, AUG_S_EXPECTED_LOGTILT, out, 3);
            else if (a.xch) xch_allreduce<3>(a.xch, out);
            a.scalars[AUG_S_EXPECTED_LOGTILT] = out[0];
            a.scalars[AUG_S_KL] = out[1];
            a.scalars[AUG_S_EXPECTED_AUGLL] = out[0] + out[1];
            a.scalars[AUG_S_FLAGS] = out[2];
            scal_zero_except(a.scalars, 0x47u);
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
            st_stream1(a.omega + base + e, augb::pg_draw_stream(g, (double)(nn + yv), true, a.f[base + e], a.L.pgtab));
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------ Gibbs for CAT, warp-compacted
// ncu on cat_sample_kernel (profiles/r1d): 640 issue slots per (obs, class) element at 20.8 active lanes of 32 —
// ~97% of the elements have b = y + n = 0 (omega = 0) but their warps wait for the few lanes that run Devroye's
// PG(1) sampler; libdevice exp / IEEE division in logistic and in the Poisson inversion; an integer division per
// element.  Here the element pass is straight-line (fast exp / rcp, Poisson by an unrolled chop-down on a single
// 53-bit uniform) and every element that needs a Polya-Gamma draw is QUEUED per warp:
//   F: fresh PG(1) rounds (b == 1, or a new round after a rejected proposal)   -> pg1 round start
//   G: truncated-inverse-Gaussian proposal attempts                            -> one attempt per step
//   B: b >= 2 (about 1e-3 of the elements)                                     -> Gamma convolution
// and a queue is worked on by a full warp as soon as it holds 32 items (see aug_pg.cuh for the PG(1) step machine,
// which is the one of pg1_compact_kernel).  Streams are keyed by the GLOBAL row / element index.
struct CatGibbsArgs {
    int64_t n, i0, ntiles;
    int nl, R, E, di, dj;        // di, dj = AUG_BLOCK / nl, AUG_BLOCK % nl
    uint64_t seed, offset;
    const uint8_t* y;
    const double* f;
    double* omega;
    int64_t* nvar;
    unsigned int* dflag;
    int bulk_ok;                 // f and y are 16-byte aligned and R * nl is a multiple of 16
    int vec_ok;                  // omega and nvar are 16-byte aligned: the element pass stores two elements per lane
    LikConst L;
    augr::PhiloxKeys keys;       // round keys of (seed, offset): constant-bank operands of the per-element Philox block
};

#ifndef CG_BLOCK
#define CG_BLOCK 256
#endif
#ifndef CG_MIN_BLOCKS
#define CG_MIN_BLOCKS 2
#endif
#define CG_STAGES 3
#define CG_QCAP 96
#define CG_BCAP 96           // a two-element step starts with < 32 items and appends up to 64
#define CG_DENSE_MARK 255    // count byte of every element of a dense row (sparse rows hold at most 250 picks)
// Round 2: the counts of a ROW are drawn from the NegativeMultinomial law directly instead of one Poisson draw per element.
// The reference samples NM(x0 = 1, p) as a Gamma-Poisson mixture (negativemultinomial.jl:35-45: tau ~ Exp(1)/p0, n_ij ~
// Poisson(p_ij tau) independent).  The same law in two steps (superposition / splitting of Poisson counts, then the mixture
// over tau in closed form): the row total N_i = sum_j n_ij is GEOMETRIC, P(N_i = k) = p0 (1 - p0)^k, and given N_i the counts are
// Multinomial(N_i; p_ij / (1 - p0)).  E[N_i] = (1 - p0)/p0 is of order one, so a row needs ONE Philox block — a geometric
// draw by chop-down (no exp, no log) and the first class pick — and N_i picks by inversion over the row's cumulative p (16 lanes
// cooperating) instead of K = 100 Philox blocks, exps and searches; the element pass only reads its count from a byte array.
// Rows with p0 < CG_DENSE_P0 (long geometric tails) keep the per-element mixture draws: exact for any input, the choice depends
// on p0 only.  Streams: the row block (tag 7, counter 0) gives the geometric uniform and the first pick (dense rows: Exp(1));
// blocks 1, 2, ... two further picks each.
#define CG_DENSE_P0 0.2

// logistic(x) with the LogExpFunctions saturation (categorical.jl:23), straight-line
__device__ __forceinline__ double logistic_fast(double x) {
    const double ax = fmin(fabs(x), 708.0);
    const double e = augf::exp_(-ax);
    const double r = augf::rcp(1.0 + e);
    const double v = x >= 0.0 ? r : e * r;
    return x < augm::LOGISTIC_LO ? 0.0 : (x > augm::LOGISTIC_HI ? 1.0 : v);
}

__device__ __noinline__ int64_t poisson_slow(uint64_t seed, uint64_t offset, uint64_t gi, double lam) {
    augr::Philox g;
    g.init(seed, offset, gi, 4u);
    return augr::poisson_rand(g, lam);
}
__device__ __noinline__ double pg_slow(uint64_t seed, uint64_t offset, uint64_t gi, double b, double c, const double* tab) {
    augr::Philox g;
    g.init(seed, offset, gi, 6u);
    return augb::pg_draw_stream(g, b, true, c, tab);
}
__device__ __noinline__ double pg1_finish_sequential_cat(uint64_t seed, uint64_t offset, uint64_t gi, double z, const double* tab) {
    augr::Philox g;
    g.init(seed, offset, gi, 3u);
    const augp::PG1 s = augp::pg1_setup(2.0 * z, tab);
    return augp::pg1_draw(g, s);
}

__global__ void __launch_bounds__(CG_BLOCK, CG_MIN_BLOCKS) cat_gibbs_kernel(const CatGibbsArgs a) {
    extern __shared__ __align__(128) unsigned char cg_smem[];
    const int nl = a.nl, R = a.R, E = a.E;
    // ring of CG_STAGES input tiles (f: E doubles, y: E bytes), filled by cp.async.bulk two tiles ahead
    unsigned char* ring = cg_smem;
    const int off_y = E * 8, stage_bytes = (E * 9 + 127) & ~127;
    // (the r(z) table stays in global memory: only the ~3 % of the elements that need a PG draw read it, through L1)
    double* Pb = reinterpret_cast<double*>(ring + CG_STAGES * stage_bytes);     // [E]   p_ij of the tile
    double* rsc = Pb + E;                                                    // [R]   Exp(1)/p0 per row (dense rows)
    double* qz_all = rsc + R;                                            // per warp: F and G items
    uint32_t* qw_all = reinterpret_cast<uint32_t*>(qz_all + (CG_BLOCK / 32) * 2 * CG_QCAP);
    uint64_t* full = reinterpret_cast<uint64_t*>(qw_all + (CG_BLOCK / 32) * (5 * CG_QCAP + 2 * CG_BCAP));
    uint64_t* empty = full + CG_STAGES;                                           // a stage is free again: one arrival per thread
    unsigned char* Ncb = reinterpret_cast<unsigned char*>(empty + CG_STAGES);     // [E]  n_ij of the tile (CG_DENSE_MARK: dense row)
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double* fz = qz_all + warp * 2 * CG_QCAP;
    double* gz = fz + CG_QCAP;
    uint32_t* fel = qw_all + warp * (5 * CG_QCAP + 2 * CG_BCAP);
    uint32_t* fra = fel + CG_QCAP;
    uint32_t* gel = fra + CG_QCAP;
    uint32_t* gra = gel + CG_QCAP;
    uint32_t* gua = gra + CG_QCAP;
    uint32_t* bel = gua + CG_QCAP;
    uint32_t* bbv = bel + CG_BCAP;
    if (tid == 0) {
        for (int s = 0; s < CG_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], CG_BLOCK); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // a tile is bulk-copied when it is a full one (its y span is then a multiple of 16 bytes at a 16-byte aligned
    // offset); the ragged last tile is loaded by the threads themselves
    auto issue = [&](int64_t t, int s) {
        if (t >= a.ntiles || (t + 1) * R > a.n || !a.bulk_ok) return;
        unsigned char* st = ring + (size_t)s * stage_bytes;
        const int64_t o = t * E;
        mbar_expect_tx(&full[s], (uint32_t)E * 9u);
        bulk_g2s(st, a.f + o, (uint32_t)E * 8u, &full[s]);
        bulk_g2s(st + off_y, a.y + o, (uint32_t)E, &full[s]);
    };
    __syncthreads();
    if (tid == 0) {
        issue((int64_t)blockIdx.x, 0);
        issue((int64_t)blockIdx.x + gridDim.x, 1);
    }

    const uint32_t lt_mask = (1u << lane) - 1u;
    const uint32_t k0 = (uint32_t)a.seed, k1 = (uint32_t)(a.seed >> 32) ^ (uint32_t)(a.offset >> 32);
    const uint32_t c3 = (uint32_t)a.offset;
    const uint64_t e_base = (uint64_t)a.i0 * (uint64_t)nl;                   // global index of the shard's first element
    int nf = 0, ng = 0, nb = 0;

    auto push_f = [&](bool want, uint32_t el, uint32_t ra, double z) {
        const uint32_t m = __ballot_sync(0xffffffffu, want);
        if (want) {
            const int pos = nf + __popc(m & lt_mask);
            fel[pos] = el; fra[pos] = ra; fz[pos] = z;
        }
        nf += __popc(m);
    };
    auto push_g = [&](bool want, uint32_t el, uint32_t ra, uint32_t ua, double z) {
        const uint32_t m = __ballot_sync(0xffffffffu, want);
        if (want) {
            const int pos = ng + __popc(m & lt_mask);
            gel[pos] = el; gra[pos] = ra; gua[pos] = ua; gz[pos] = z;
        }
        ng += __popc(m);
    };
    // accept/reject of a proposal x of (el, round); on rejection the element starts a new round
    auto finish = [&](bool have_x, double x, uint32_t el, uint32_t round, uint32_t ua, double z, uint32_t e_lo, uint32_t e_hi) {
        bool newround = false;
        if (have_x) {
            if (augp::pg1_accept(x, ua, k0, k1, e_lo, e_hi, c3, round)) st_stream1(a.omega + el, 0.25 * x);
            else newround = true;
        }
        push_f(newround, el, round + 1u, z);
    };

    const int l16 = lane & 15;
    // Every warp owns R / 8 rows of each tile END TO END (p_ij, the row draws, the element pass): no CTA-wide barrier inside a
    // tile — ncu on the barrier version (profiles/r2y): 27 % of the stall samples at the two __syncthreads per tile, issue slots
    // 45 % busy.  The warps only meet at the mbarriers of the input ring: full[s] (the bulk copy landed) and empty[s] (all 8
    // warps are done with the stage; thread 0 waits for it before it refills the stage, two tiles ahead).
    const int rows_w = R / (CG_BLOCK / 32);                                  // host: R is a multiple of the warp count
    const int dj64 = 64 % nl;
    // state of the element pass (phase C) of the current tile; e0 >= ew1: the next tile has to be set up first
    int64_t tile = (int64_t)blockIdx.x - gridDim.x;
    int e0 = 0, ew0 = 0, ew1 = 0;     // this warp's element range [ew0, ew1) of the tile and its position e0 in it
    int stg = CG_STAGES - 1;          // stage of the current tile
    uint32_t par = 1;                 // its mbarrier phase parity (flips when stg wraps to 0)
    uint32_t base = 0;
    uint32_t kt = 0;                  // tiles this warp has started (parity of the empty barriers)
    double* const P = Pb;
    double* const rs = rsc;
    unsigned char* const Nc = Ncb;    // counts of the current tile
    const double* Fs = nullptr;       // staged f and y of the current tile
    const unsigned char* Ys = nullptr;
    bool drain = false;
    for (;;) {
        if (nf >= 32 || (drain && nf > 0)) {
            // ---- F: round start for the last min(nf, 32) items
            const int cnt = nf < 32 ? nf : 32;
            const bool active = lane < cnt;
            uint32_t el = 0, round = 0;
            double z = 0.0;
            if (active) { const int idx = nf - cnt + lane; el = fel[idx]; round = fra[idx]; z = fz[idx]; }
            __syncwarp();
            nf -= cnt;
            const uint64_t gi = e_base + el;
            const uint32_t e_lo = (uint32_t)gi, e_hi = (uint32_t)(gi >> 32);
            bool have_x = false, to_g = false;
            double x = 0.0;
            uint32_t ua = 0;
            if (active) {
                if (round < 255u) {
                    const augp::PG1 s = augp::pg1_setup<false>(2.0 * z, a.L.pgtab);
                    uint32_t w[4];
                    augr::philox4x32_10(k0, k1, e_lo, e_hi, augp::pg1_ctr(0u, round, 0u), c3, w);
                    ua = w[3];
                    // exponential proposal, or the first truncated-IG attempt in line (as in pg1_compact_kernel; aug_pg.cuh)
                    const double u0 = augr::u32_mid(w[0]);
                    const double E = -augf::log_(augr::u53_open0(w[1], w[2]));
                    double a_ig;
                    const double x_ig = augp::trunc_ig_small_z(E, z, a_ig);
                    if (u0 < s.r) { x = fma(E, s.invK, augp::T); have_x = true; }
                    else if (z < 1.0 / augp::T && u0 <= fma(1.0 - s.r, augf::exp_(-fmin(a_ig, 700.0)), s.r)) { x = x_ig; have_x = true; }
                    to_g = !have_x;
                } else {
                    st_stream1(a.omega + el, pg1_finish_sequential_cat(a.seed, a.offset, gi, z, a.L.pgtab));
                }
            }
            push_g(to_g, el, round | (1u << 8), ua, z);
            finish(have_x, x, el, round, ua, z, e_lo, e_hi);
            __syncwarp();
            continue;
        }
        if (ng >= 32 || (drain && ng > 0)) {
            // ---- G: one truncated-IG attempt for the last min(ng, 32) items
            const int cnt = ng < 32 ? ng : 32;
            const bool active = lane < cnt;
            uint32_t el = 0, ra = 0, ua = 0;
            double z = 0.0;
            if (active) { const int idx = ng - cnt + lane; el = gel[idx]; ra = gra[idx]; ua = gua[idx]; z = gz[idx]; }
            __syncwarp();
            ng -= cnt;
            const uint32_t round = ra & 0xffu, attempt = ra >> 8;
            const uint64_t gi = e_base + el;
            const uint32_t e_lo = (uint32_t)gi, e_hi = (uint32_t)(gi >> 32);
            bool have_x = false, again = false;
            double x = 0.0;
            if (active) {
                if (attempt < 255u) {
                    uint32_t w[4];
                    augr::philox4x32_10(k0, k1, e_lo, e_hi, augp::pg1_ctr(1u, round, attempt), c3, w);
                    x = augp::trunc_ig_attempt_w(w, z);
                    have_x = x > 0.0;
                    again = !have_x;
                } else {
                    st_stream1(a.omega + el, pg1_finish_sequential_cat(a.seed, a.offset, gi, z, a.L.pgtab));
                }
            }
            push_g(again, el, ra + (1u << 8), ua, z);
            finish(have_x, x, el, round, ua, z, e_lo, e_hi);
            __syncwarp();
            continue;
        }
        if (nb >= 32 || (drain && nb > 0)) {
            // ---- B: PG(b >= 2, c) by the Gamma convolution (rare)
            const int cnt = nb < 32 ? nb : 32;
            if (lane < cnt) {
                const int idx = nb - cnt + lane;
                const uint32_t el = bel[idx];
                st_stream1(a.omega + el, pg_slow(a.seed, a.offset, e_base + el, (double)bbv[idx], a.f[el], a.L.pgtab));
            }
            __syncwarp();
            nb -= cnt;
            continue;
        }
        if (drain) break;
        if (e0 >= ew1) {
            // ---- next tile: this warp's rows of it
            if (tile >= 0) mbar_arrive(&empty[stg]);                             // done with the stage of the tile just finished (every thread
                                                                                 // arrives for its own reads: one instruction per tile)
            tile += gridDim.x;
            if (tile >= a.ntiles) { drain = true; e0 = 0; ew1 = 0; continue; }
            if (++stg == CG_STAGES) { stg = 0; par ^= 1u; }
            const int64_t row0 = tile * R;
            const int rows = (int)min((int64_t)R, a.n - row0);
            base = (uint32_t)(row0 * nl);                                        // n * nl < 2^32 (host-checked)
            double* Fw = reinterpret_cast<double*>(ring + (size_t)stg * stage_bytes);
            unsigned char* Yw = ring + (size_t)stg * stage_bytes + off_y;
            const bool bulk = rows == R && a.bulk_ok;
            // thread 0 keeps the ring two tiles ahead: the stage of the PREVIOUS tile takes tile + 2 grids once all warps left it
            // (the first tile finds that stage unused; later ones wait for the previous tile's CG_BLOCK arrivals)
            if (tid == 0) {
                const int sp_ = stg == 0 ? CG_STAGES - 1 : stg - 1;
                if (kt >= 1) mbar_wait(&empty[sp_], ((kt - 1) / CG_STAGES) & 1u);
                issue(tile + 2 * (int64_t)gridDim.x, sp_);
            }
            ++kt;                                                                // (every thread counts its warp's tiles)
            const int rbeg = warp * rows_w, rend = min(rows, rbeg + rows_w);
            ew0 = rbeg * nl;
            ew1 = rend > rbeg ? rend * nl : ew0;
            if (bulk) {
                mbar_wait(&full[stg], par);
            } else {
                // the ragged last tile is written by the warps themselves: not before every warp has left the tile that used
                // this stage three tiles ago (bulk tiles get that guarantee through full[], which thread 0 only arms then)
                if (kt > CG_STAGES) mbar_wait(&empty[stg], ((kt - 1) / CG_STAGES - 1) & 1u);
                for (int e = ew0 + lane; e < ew1; e += 32) {
                    Fw[e] = ld_stream1(a.f + base + e);
                    Yw[e] = __ldg(a.y + base + e);
                }
                __syncwarp();
            }
            double* Pw = P;
            double* rw = rs;
            // phase A: p_ij = theta_j logistic(f_ij) / sum(theta)   categorical.jl:72-78
            // (two independent elements per lane and iteration: the exp / reciprocal chains are latency-bound at 4 warps per
            //  scheduler)
            {
                int j = lane % nl, j2 = (lane + 32) % nl;
                int e = ew0 + lane;
                for (; e + 32 < ew1; e += 64) {
                    const double la = logistic_fast(Fw[e]), lb = logistic_fast(Fw[e + 32]);
                    Pw[e] = __ldg(a.L.theta + j) * la;
                    Pw[e + 32] = __ldg(a.L.theta + j2) * lb;
                    j += dj64;
                    if (j >= nl) j -= nl;
                    j2 += dj64;
                    if (j2 >= nl) j2 -= nl;
                }
                if (e < ew1) Pw[e] = __ldg(a.L.theta + j) * logistic_fast(Fw[e]);
            }
            __syncwarp();
            // phase B: 16 lanes per row: p0 = 1 - sum_j p_ij, then the row's counts (see CG_DENSE_P0 above)
            unsigned char* Ncw = Nc;
            const int CH = (nl + 15) >> 4;                                         // classes per lane chunk
            for (int r0 = rbeg; r0 < rend; r0 += 2) {                              // warp-uniform bound (shuffles)
                const int r = r0 + (lane >> 4);
                const bool rv = r < rend;
                const double* Pr = Pw + (rv ? r : rbeg) * nl;
                unsigned char* Nr = Ncw + (rv ? r : rbeg) * nl;
                // chunked sums: lane l16 owns classes [l16*CH, (l16+1)*CH); the inclusive scan over the 16 lanes is the
                // cumulative p the class picks invert, and its total is the row sum (one association for both)
                const int j0 = l16 * CH, j1 = min(nl, j0 + CH);
                double cs = 0.0;
                if (rv) for (int jj = j0; jj < j1; ++jj) { cs += Pr[jj]; Nr[jj] = 0; }
                double inc = cs;
#pragma unroll
                for (int o = 1; o < 16; o <<= 1) {
                    const double up = __shfl_up_sync(0xffffffffu, inc, o, 16);
                    if (l16 >= o) inc += up;
                }
                const double sp = __shfl_sync(0xffffffffu, inc, 15, 16);           // row sum
                double exc = __shfl_up_sync(0xffffffffu, inc, 1, 16);                // the previous lane's inclusive sum: the chunks'
                if (l16 == 0) exc = 0.0;                                             // intervals (exc, inc] tile (0, sp] exactly
                int nrow = 0;
                double t_first = 2.0;                                            // target of the first pick (> any cumulative sum)
                const uint64_t gr = (uint64_t)(a.i0 + row0 + (rv ? r : 0));
                if (rv && l16 == 0) {
                    const double p0 = 1.0 - sp;
                    if (!(sp < 1.0)) atomicOr(a.dflag, 1u);                      // ctor precondition :17-22
                    uint32_t w[4];
                    augr::philox4x32_10(k0, k1, (uint32_t)gr, (uint32_t)(gr >> 32), 7u << 28, c3, w);
                    const bool dense = !(p0 >= CG_DENSE_P0);                     // also NaN
                    nrow = dense ? -1 : 0;
                    if (dense) {
                        rw[r] = -augf::log_(augr::u53_open0(w[0], w[1])) / p0;   // tau / (1 - p0) for the per-element draws
                    } else {
                        double pk = p0, u = augr::u53_open0(w[0], w[1]);         // geometric by chop-down: P(N = k) = p0 sp^k
                        while (u > pk && nrow < 250) {
                            u -= pk;
                            ++nrow;
                            pk *= sp;
                        }
                        t_first = augr::u53_open0(w[2], w[3]) * sp;              // in (0, sp]
                    }
                }
                nrow = __shfl_sync(0xffffffffu, nrow, 0, 16);
                if (nrow < 0) {                                                  // dense row: every count byte holds CG_DENSE_MARK
                    if (rv) for (int jj = j0; jj < j1; ++jj) Nr[jj] = CG_DENSE_MARK;
                    nrow = 0;
                }
                __syncwarp();                                                    // the zeroed counts before the increments
                int maxn = nrow;
                maxn = max(maxn, __shfl_xor_sync(0xffffffffu, maxn, 16));
                for (int ev = 0; ev < maxn; ++ev) {                              // warp-uniform trip count (the two rows differ)
                    double t = 2.0;                                              // > any cumulative sum: no lane matches
                    if (rv && l16 == 0 && ev < nrow) {
                        if (ev == 0) {
                            t = t_first;
                        } else {                                                 // picks 1, 2 from block 1; 3, 4 from block 2; ...
                            uint32_t w[4];
                            augr::philox4x32_10(k0, k1, (uint32_t)gr, (uint32_t)(gr >> 32), (7u << 28) | (uint32_t)((ev + 1) >> 1), c3, w);
                            const double u = (ev & 1) ? augr::u53_open0(w[0], w[1]) : augr::u53_open0(w[2], w[3]);
                            t = u * sp;                                          // in (0, sp]
                        }
                    }
                    t = __shfl_sync(0xffffffffu, t, 0, 16);
                    // the lane whose chunk holds the target scans it; rounding can only push t above the last partial sum of
                    // a chunk by an ulp, in which case the pick is the chunk's last class
                    if (rv && t > exc && (t <= inc || (l16 == 15 && t <= 1.5))) {
                        double acc = exc;
                        int jp = j1 - 1;
                        for (int jj = j0; jj < j1; ++jj) {
                            acc += Pr[jj];
                            if (t <= acc) { jp = jj; break; }
                        }
                        Nr[jp] = (unsigned char)((int)Nr[jp] + 1);              // <= 250 picks per row: never CG_DENSE_MARK
                    }
                    __syncwarp();
                }
            }
            __syncwarp();
            Fs = Fw;
            Ys = Yw;
            e0 = ew0;
            continue;
        }
        // ---- element pass (phase C): n_ij from the row's split counts (dense rows: n_ij ~ Poisson(p_ij tau/(1-p0)) per element),
        //      b = y + n; queue what needs a PG draw.  98 % of the elements have b = 0 (omega = 0, n = 0, nothing to draw), so the
        //      common step takes TWO adjacent elements per lane: 16-byte stores of (0.0, 0.0) and (n, n') and no queue work unless
        //      some lane of the warp has b > 0.  ncu on the one-element step (profiles/r2z): 74 warp-instructions per 32 elements,
        //      26 % of the kernel.
        // one element per lane (any input: dense rows, unaligned outputs, the odd element in front of the pairs)
        auto elem_step = [&](int e, bool valid) {
            const uint32_t el = base + (uint32_t)e;
            int yv = 0;
            int64_t nn = 0;
            bool dense = false;
            if (valid) {
                yv = (int)Ys[e];
                nn = (int64_t)Nc[e];
                dense = nn == CG_DENSE_MARK;
            }
            const uint64_t gi = e_base + el;
            if (__any_sync(0xffffffffu, dense)) {                                // rare: rows with p0 < CG_DENSE_P0
                if (dense) {
                    const double lam = P[e] * rs[e / nl];
                    uint32_t w[4];
                    AUG_PHILOX_RK(a.keys, (uint32_t)gi, (uint32_t)(gi >> 32), 5u << 28, c3, w);
                    nn = 0;
                    if (lam > 0.0) {
                        if (lam < 12.0) {
                            double u = (double)(((((uint64_t)w[1] << 32) | w[0]) >> 11)) * 0x1.0p-53;
                            double p = augf::exp_(-lam);
                            int k = 0;
                            while (u > p && k < 200) {
                                u -= p;
                                ++k;
                                p *= lam / (double)k;
                            }
                            if (k >= 200) k = (int)poisson_slow(a.seed, a.offset, gi, lam);   // round-off tail
                            nn = k;
                        } else {
                            nn = poisson_slow(a.seed, a.offset, gi, lam);
                        }
                    }
                }
            }
            const int b = yv + (int)min(nn, (int64_t)1 << 30);
            if (valid) {
                a.nvar[el] = nn;
                if (b == 0) st_stream1(a.omega + el, 0.0);
            }
            const bool needs = valid && b >= 1;
            const double z = needs ? 0.5 * fabs(Fs[e]) : 0.0;
            push_f(needs && b == 1, el, 0u, z);
            const bool wb = needs && b >= 2;
            if (__any_sync(0xffffffffu, wb)) {                                   // about 1e-3 of the elements
                const uint32_t m = __ballot_sync(0xffffffffu, wb);
                if (wb) {
                    const int pos = nb + __popc(m & lt_mask);
                    bel[pos] = el;
                    bbv[pos] = (uint32_t)b;
                }
                nb += __popc(m);
            }
            __syncwarp();
        };
        if (!a.vec_ok) {
            elem_step(e0 + lane, e0 + lane < ew1);
            e0 += 32;
        } else if (e0 == ew0 && ((base + (uint32_t)e0) & 1u)) {
            elem_step(e0, lane == 0);                                            // pairs start at even element indices
            e0 += 1;
        } else {
            const int ea = e0 + 2 * lane;
            const bool va = ea < ew1, vb = ea + 1 < ew1;
            int na = 0, nbb = 0, ya = 0, yb = 0;
            if (va) { na = (int)Nc[ea]; ya = (int)Ys[ea]; }
            if (vb) { nbb = (int)Nc[ea + 1]; yb = (int)Ys[ea + 1]; }
            if (__any_sync(0xffffffffu, na == CG_DENSE_MARK || nbb == CG_DENSE_MARK)) {
                elem_step(e0 + lane, e0 + lane < ew1);
                elem_step(e0 + 32 + lane, e0 + 32 + lane < ew1);
            } else {
                const uint32_t el = base + (uint32_t)ea;
                const int ba = ya + na, bb = yb + nbb;
                if (vb) {
                    st_stream2_i64(a.nvar + el, (int64_t)na, (int64_t)nbb);
                    if ((ba | bb) == 0) st_stream2(a.omega + el, 0.0, 0.0);
                } else if (va) {
                    a.nvar[el] = (int64_t)na;
                }
                if (__any_sync(0xffffffffu, (ba | bb) != 0)) {                   // (also false for the invalid lanes: all zero)
                    if (vb && (ba | bb) != 0) {
                        if (ba == 0) st_stream1(a.omega + el, 0.0);
                        if (bb == 0) st_stream1(a.omega + el + 1, 0.0);
                    }
                    push_f(ba == 1, el, 0u, ba == 1 ? 0.5 * fabs(Fs[ea]) : 0.0);
                    push_f(bb == 1, el + 1u, 0u, bb == 1 ? 0.5 * fabs(Fs[ea + 1]) : 0.0);
                    const bool wa = ba >= 2, wbb = bb >= 2;
                    if (__any_sync(0xffffffffu, wa || wbb)) {                    // about 1e-3 of the elements
                        const uint32_t m1 = __ballot_sync(0xffffffffu, wa);
                        if (wa) {
                            const int pos = nb + __popc(m1 & lt_mask);
                            bel[pos] = el;
                            bbv[pos] = (uint32_t)ba;
                        }
                        nb += __popc(m1);
                        const uint32_t m2 = __ballot_sync(0xffffffffu, wbb);
                        if (wbb) {
                            const int pos = nb + __popc(m2 & lt_mask);
                            bel[pos] = el + 1u;
                            bbv[pos] = (uint32_t)bb;
                        }
                        nb += __popc(m2);
                    }
                }
                if (va && !vb && ba == 0) st_stream1(a.omega + el, 0.0);         // the odd last element of the warp's rows
                __syncwarp();
            }
            e0 += 64;
        }
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

bool cat_no_tma() {   // AUGCUDA_NO_TMA=1 keeps every call on the direct-load kernel (A/B measurements)
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("AUGCUDA_NO_TMA");
        v = (e && e[0] == '1') ? 1 : 0;
    }
    return v == 1;
}

int cat_row_pipe() {   // AUGCUDA_CAT_PIPE = 161 | 321 | 322: rows per tile and stages per CTA of cat_row_kernel (A/B measurements)
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("AUGCUDA_CAT_PIPE");
        v = e ? atoi(e) : CAT_PIPE_DEFAULT;
        if (v != 161 && v != 321 && v != 322) v = CAT_PIPE_DEFAULT;
    }
    return v;
}

bool cat_row_enabled() {   // AUGCUDA_CAT_ROW=0 keeps wide rows on the CTA-wide staged kernel (A/B measurements)
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("AUGCUDA_CAT_ROW");
        v = (e && e[0] == '0') ? 0 : 1;
    }
    return v == 1;
}

// the direct-load row kernel: any alignment, FROM_STATE verbs, ragged tails, rows too wide for the staged ring
int32_t launch_cat_direct(aug_ctx* ctx, CatArgs a, bool elbo, bool from_state) {
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
    const int64_t ntiles = (a.n + a.R - 1) / a.R;
    int64_t grid = (int64_t)ctx->sms * occ;
    if (grid > AUG_MAX_GRID) grid = AUG_MAX_GRID;
    if (grid > ntiles) grid = ntiles;
    void* args[] = {(void*)&a};
    AUG_CUDA(cudaLaunchKernel(k, dim3((unsigned)grid), dim3(AUG_BLOCK), args, smem, ctx->stream));
    ctx->launches++;
    return AUG_OK;
}

}  // namespace

int32_t aug_cat_dispatch(aug_ctx* ctx, const aug_lik* lik, int64_t n, const void* y, const double* mu,
                         const double* var, void* s0, void* s1, void* s2, const void* rs0, const void* rs1,
                         const void* rs2, double* beta, double* gamma, int64_t ldo, double* scalars,
                         bool from_state) {
    if (n < 0 || (!y && n > 0)) return AUG_ERR_BAD_ARG;
    const bool elbo = scalars != nullptr;
    if (elbo && lik->kind == AUG_CAT) return AUG_ERR_PRECONDITION;   // categorical.jl:165-170
    if (elbo && aug_xch_for(ctx)) {          // exchanges happen in call order: complete a pending split-phase one first
        int32_t rf = aug_xch_flush(ctx);
        if (rf) return rf;
    }
    if (n == 0) {
        if (scalars) {
            AUG_CUDA(cudaMemsetAsync(scalars, 0, AUG_NSCALARS * sizeof(double), ctx->stream));
            if (aug_xch_for(ctx)) return aug_xch_zero_contribution(ctx, scalars, AUG_S_EXPECTED_LOGTILT, 3);
        }
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
    a.xch = elbo ? aug_xch_for(ctx) : nullptr;
    if (a.xch && ctx->deferred) {
        a.xch_defer = 1;
        ctx->pending = 1;
    }
    int32_t rc = aug_lik_const(ctx, lik, &a.L, false, false);
    if (rc) return rc;
    // ---- full tiles of a fused call on 16-byte aligned arrays: the bulk-async staged kernel
    int64_t n0 = 0;   // rows it covers; the ragged tail [n0, n) goes through the direct-load kernel below
    // (a) wide rows (K of the order of 100): the row-aligned two-warp kernel, 7-8 tiles resident per SM
    const int pipe = cat_row_pipe();                                       // rows per tile * 10 + stages per CTA
    const int CAT_ROW_R = pipe / 10, NSr = pipe % 10;
    if (!from_state && !cat_no_tma() && cat_row_enabled() && a.nl >= 32 && aug_aligned16(y) && aug_aligned16(mu) &&
        aug_aligned16(var) && aug_aligned16(s0) && aug_aligned16(s1) && aug_aligned16(s2) && n >= CAT_ROW_R) {
        CatRowArgs ta{};
        ta.a = a;
        const int nl = a.nl;
        const bool even = (nl & 1) == 0;
        ta.a.R = CAT_ROW_R;
        ta.E = CAT_ROW_R * nl;
        ta.rs = even ? nl + 2 : nl;
        ta.off_mu = (ta.E + 127) & ~127;
        ta.off_var = (ta.off_mu + CAT_ROW_R * ta.rs * 8 + 127) & ~127;
        ta.off_rinv = (ta.off_var + CAT_ROW_R * ta.rs * 8 + 127) & ~127;   // = bytes of one stage
        ta.log_inv_denom = -log(a.L.c0);
        const int CAT_ROW_BLOCK = CAT_ROW_R * 4;
        const size_t smem = (size_t)NSr * ta.off_rinv + CAT_ROW_R * sizeof(double) + 8 * NSr + 16;
        if ((smem + 1024 + 128) * (256 / CAT_ROW_BLOCK) <= (size_t)ctx->smem_per_sm) {   // at least 8 warps per SM
            ta.ntiles = n / CAT_ROW_R;
            n0 = ta.ntiles * CAT_ROW_R;
            ta.a.n = n0;
#define CAT_ROW_PICK(R_, NS_)                                                                                                \
    (elbo ? (even ? (const void*)cat_row_kernel<true, true, R_, NS_> : (const void*)cat_row_kernel<true, false, R_, NS_>)      \
          : (even ? (const void*)cat_row_kernel<false, true, R_, NS_> : (const void*)cat_row_kernel<false, false, R_, NS_>))
            const void* k = pipe == 321 ? CAT_ROW_PICK(32, 1) : pipe == 322 ? CAT_ROW_PICK(32, 2) : CAT_ROW_PICK(16, 1);
#undef CAT_ROW_PICK
            AUG_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            int occ = 1;
            if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k, CAT_ROW_BLOCK, smem) != cudaSuccess || occ < 1) occ = 1;
            int64_t grid = (int64_t)ctx->sms * occ;
            if (grid > AUG_MAX_GRID) grid = AUG_MAX_GRID;
            if (grid > ta.ntiles) grid = ta.ntiles;
            ta.accumulate = (elbo && n0 < n) ? 1 : 0;
            if (n0 < n) {   // the tail launch goes first and the staged launch accumulates onto its scalars
                const int64_t eo = n0 * nl;
                CatArgs t = a;
                t.xch = nullptr;
                t.n = n - n0;
                t.y = a.y + eo;
                t.mu = a.mu + eo;
                t.var = a.var + eo;
                if (a.s0) t.s0 = a.s0 + eo;
                if (a.s1) t.s1 = a.s1 + eo;
                if (a.s2) t.s2 = a.s2 + eo;
                if (a.beta) t.beta = a.beta + n0;
                if (a.gamma) t.gamma = a.gamma + n0;
                rc = launch_cat_direct(ctx, t, elbo, false);
                if (rc) return rc;
            }
            void* args[] = {(void*)&ta};
            AUG_CUDA(cudaLaunchKernel(k, dim3((unsigned)grid), dim3(CAT_ROW_BLOCK), args, smem, ctx->stream));
            ctx->launches++;
            return AUG_OK;
        }
    }
    // (b) narrow rows: CTA-wide tiles of ~1792 elements
    if (!from_state && !cat_no_tma() && aug_aligned16(y) && aug_aligned16(mu) && aug_aligned16(var) &&
        aug_aligned16(s0) && aug_aligned16(s1) && aug_aligned16(s2)) {
        CatTmaArgs ta{};
        ta.a = a;
        const int nl = a.nl;
        int R = (1792 / nl) & ~15;
        if (R < 16) R = 16;
        if (R > 1024) R = 1024;
        ta.a.R = R;
        ta.E = R * nl;
        ta.off_mu = (ta.E + 127) & ~127;
        ta.off_var = ta.off_mu + ta.E * 8;
        ta.stage_bytes = (ta.off_var + ta.E * 8 + 127) & ~127;
        ta.di = (2 * AUG_BLOCK) / nl;
        ta.dj = (2 * AUG_BLOCK) % nl;
        ta.log_inv_denom = -log(a.L.c0);
        const bool even = (nl & 1) == 0;
        ta.rs = even ? nl + 2 : nl;
        const size_t fixed = (size_t)R * ta.rs * sizeof(double) * (elbo ? 4 : 2) + (size_t)R * sizeof(double) +
                             8 * sizeof(uint64_t) + 128;
        const size_t per_cta2 = (size_t)(ctx->smem_per_sm - 2 * 1024) / 2, per_cta1 = (size_t)ctx->smem_optin;
        int S = 0;
        if (fixed + 2 * (size_t)ta.stage_bytes <= per_cta2) S = (int)((per_cta2 - fixed) / ta.stage_bytes);
        else if (fixed + 2 * (size_t)ta.stage_bytes <= per_cta1) S = (int)((per_cta1 - fixed) / ta.stage_bytes);
        if (S > 4) S = 4;
        ta.ntiles = n / R;
        if (S >= 2 && ta.ntiles >= 1) {
            ta.S = S;
            n0 = ta.ntiles * R;
            ta.a.n = n0;
            const size_t smem = fixed + (size_t)S * ta.stage_bytes;
            const void* k = elbo ? (even ? (const void*)cat_tma_kernel<true, true> : (const void*)cat_tma_kernel<true, false>)
                                 : (even ? (const void*)cat_tma_kernel<false, true> : (const void*)cat_tma_kernel<false, false>);
            AUG_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            int occ = 1;
            if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k, AUG_BLOCK, smem) != cudaSuccess || occ < 1) occ = 1;
            int64_t grid = (int64_t)ctx->sms * occ;
            if (grid > AUG_MAX_GRID) grid = AUG_MAX_GRID;
            if (grid > ta.ntiles) grid = ta.ntiles;
            // the tail launch (if any) goes first and this one accumulates onto its scalars
            ta.accumulate = (elbo && n0 < n) ? 1 : 0;
            if (n0 < n) {
                const int64_t eo = n0 * nl;
                CatArgs t = a;
                t.xch = nullptr;                 // the staged launch below is the verb's last one
                t.n = n - n0;
                t.y = a.y + eo;
                t.mu = a.mu + eo;
                t.var = a.var + eo;
                if (a.s0) t.s0 = a.s0 + eo;
                if (a.s1) t.s1 = a.s1 + eo;
                if (a.s2) t.s2 = a.s2 + eo;
                if (a.beta) t.beta = a.beta + n0;
                if (a.gamma) t.gamma = a.gamma + n0;
                rc = launch_cat_direct(ctx, t, elbo, false);
                if (rc) return rc;
            }
            void* args[] = {(void*)&ta};
            AUG_CUDA(cudaLaunchKernel(k, dim3((unsigned)grid), dim3(AUG_BLOCK), args, smem, ctx->stream));
            ctx->launches++;
            return AUG_OK;
        }
    }
    return launch_cat_direct(ctx, a, elbo, from_state);
}

int32_t aug_cat_sample(aug_ctx* ctx, const aug_lik* lik, int64_t n, int64_t i0, const void* y, const double* f,
                       double* omega, int64_t* nvar, uint64_t offset) {
    if (!y || !f || !omega || !nvar) return AUG_ERR_BAD_ARG;
    const int nl = lik->nlatent;
    // the warp-compacted kernel queues 32-bit element offsets and keeps two tiles of p in shared memory
    if (!cat_no_tma() && n * (int64_t)nl < ((int64_t)1 << 32) - 4096 && nl <= 4096) {
        CatGibbsArgs g{};
        g.n = n;
        g.i0 = i0;
        g.nl = nl;
        const int W = CG_BLOCK / 32;                 // every warp owns R / W rows of a tile, two at a time when it can
        int R = (7 * CG_BLOCK) / nl;
        if (R > 1024) R = 1024;
        if (R >= 2 * W) R -= R % (2 * W);
        else R = W;
        g.R = R;
        g.bulk_ok = (((int64_t)R * nl) % 16 == 0) && aug_aligned16(y) && aug_aligned16(f);
        g.vec_ok = aug_aligned16(omega) && aug_aligned16(nvar);
        g.E = R * nl;
        g.ntiles = (n + R - 1) / R;
        g.di = AUG_BLOCK / nl;
        g.dj = AUG_BLOCK % nl;
        g.seed = ctx->seed;
        g.offset = offset;
        augr::philox_round_keys((uint32_t)g.seed, (uint32_t)(g.seed >> 32) ^ (uint32_t)(g.offset >> 32), &g.keys);
        g.y = (const uint8_t*)y;
        g.f = f;
        g.omega = omega;
        g.nvar = nvar;
        g.dflag = ctx->dflag;
        int32_t rc = aug_lik_const(ctx, lik, &g.L, false, true);
        if (rc) return rc;
        const size_t smem = (size_t)CG_STAGES * (((size_t)g.E * 9 + 127) & ~(size_t)127) +
                            sizeof(double) * ((size_t)g.E + (size_t)R) +
                            (CG_BLOCK / 32) * (2 * CG_QCAP * sizeof(double) + (5 * CG_QCAP + 2 * CG_BCAP) * sizeof(uint32_t)) +
                            2 * CG_STAGES * sizeof(uint64_t) + (size_t)g.E + (size_t)R + 128;
        if (smem <= (size_t)ctx->smem_optin) {
            AUG_CUDA(cudaFuncSetAttribute(cat_gibbs_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            int occ = 1;
            if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, cat_gibbs_kernel, CG_BLOCK, smem) != cudaSuccess || occ < 1)
                occ = 1;
            int64_t grid = (int64_t)ctx->sms * occ;
            if (grid > g.ntiles) grid = g.ntiles;
            cat_gibbs_kernel<<<(unsigned)grid, CG_BLOCK, smem, ctx->stream>>>(g);
            ctx->launches++;
            return (int32_t)cudaGetLastError();
        }
    }
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
