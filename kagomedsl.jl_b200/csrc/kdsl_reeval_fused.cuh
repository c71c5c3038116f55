// kdsl_reeval_fused.cuh -- reevaluateW! (src/MonteCarlo.jl:55-66) as ONE kernel: no explicit inverse, no gather
// kernel, no separate GEMM.
//
// W = U inv(tilde_U) has unit rows on the occupied sites (W[R_l, l'] = delta) and the rows V inv(tilde_U) on the
// unoccupied ones (V = U[unoccupied sites, :]).  Transposed:  inv(tilde_U)^T [tilde_U^T | V^T] = [I | W_unocc^T], i.e.
// Gauss-Jordan ROW elimination with partial (row) pivoting on the N x (N + M) matrix B = [tilde_U^T | V^T],
// B[j, c] = U[site(c), j], carries the V^T columns straight to the non-trivial part of W.  Only the columns that are
// not finished yet are touched, so the work is N^3 + 2 N^2 M = 3 N^3 flop per matrix at half filling, where inverse
// (2 N^3) + product (2 N^2 M) need 4 N^3; and the 2 N^2 M flop of the V part are pure tensor-pipe streaming that runs
// UNDER the latency-bound pivot chain of the next panel (look-ahead, warp specialisation as in k_inverse_v5).
//
// 512 threads, one CTA per SM, persistent over the (walker, species) items of the refresh list.
//   team P (warps 0-7, thread j owns row j = orbital j): panel load, pivot loop, operands R - E to shared memory
//   team G (warps 8-15): DMMA update  B[:, J] += (R - E) B_old[(p_q), J]  of every unfinished column outside panels
//   s and s + 1.
// Rows are never exchanged (implicit pivoting): after the last step row p_t of the V part is column t of W.
// Data movement per matrix: the columns of B are first read from a transposed copy of U (UT[site][:], contiguous,
// shared by all walkers, L2 resident), live in a per-CTA workspace of Np x (Np + Mp) doubles that is reused for
// every item (148 x 746 KB at 432 sites, L2 resident), and the last block step writes W itself: every warp owns a
// contiguous range of sites and stores full, coalesced segments of the columns of W (unit rows of the occupied
// sites included) with streaming stores.
// The pivot rule is k_inverse_v4's (largest |x| by its top 32 bits over the rows that were never a pivot, ties to
// the lowest row; zero or non-finite pivot => singular, W of that species is left untouched).
#pragma once
#include "kdsl_common.cuh"
#include "kdsl_refresh.cuh"
#include "kdsl_refresh_fast.cuh"
#include "kdsl_inverse_v5.cuh"
#include <type_traits>


// UT[site][j] = U[site, j] for j < N (0 up to Np); rows ns .. ns+7: unit vectors e_{N+i} (identity padding of
// tilde_U up to Np); row ns+8: zero (padding of the V part up to Mp)
__global__ void k_build_UT(const double *__restrict__ U, double *__restrict__ UT, int ns, int N, int Np) {
    const size_t total = (size_t)(ns + 9) * Np;
    for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
        const int site = (int)(e / Np), j = (int)(e - (size_t)site * Np);
        double v = 0.0;
        if (site < ns) v = j < N ? U[(size_t)j * ns + site] : 0.0;
        else if (site < ns + 8) v = (j == N + (site - ns)) ? 1.0 : 0.0;
        UT[e] = v;
    }
}

// Per-species bookkeeping after k_reeval_fused: a species whose tilde_U was fine has a fresh W0 and drops its pending
// updates; a walker with a singular species is flagged (its other species is still refreshed: the two W are independent).
__global__ void k_refresh_status_fused(DevState S, const int *__restrict__ list, const int *__restrict__ status) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= batch_count(S, list)) return;
    const int w = list ? list[b] : b;
    if (!status[2 * b]) S.fcnt[2 * w] = 0;
    if (!status[2 * b + 1]) S.fcnt[2 * w + 1] = 0;
    if (status[2 * b] | status[2 * b + 1]) {
        atomicOr(&S.flags[w], KDSL_FLAG_SINGULAR_DEV);
        atomicAdd(&S.cnt[3], 1);
    } else {
        atomicAnd(&S.flags[w], ~KDSL_FLAG_SINGULAR_DEV);
        S.n_refresh[w] += 1ull;
    }
}


// Workspace traffic carries an L2 evict_last policy (the per-CTA workspace is re-read and re-written at every block step
// and should outlive the streams passing through the L2: W output, transposed U, the other kernels' data) and does not
// allocate in L1.
#ifndef KDSL_FUSED_L2HINT
#define KDSL_FUSED_L2HINT 1
#endif
__device__ __forceinline__ unsigned long long fused_policy_keep() {
    unsigned long long p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ double2 ws_ld2(const double2 *a, unsigned long long p) {
#if KDSL_FUSED_L2HINT
    double2 v;
    asm volatile("ld.global.L1::no_allocate.L2::cache_hint.v2.f64 {%0, %1}, [%2], %3;" : "=d"(v.x), "=d"(v.y) : "l"(a), "l"(p));
    return v;
#else
    return __ldcg(a);
#endif
}
__device__ __forceinline__ double ws_ld1(const double *a, unsigned long long p) {
#if KDSL_FUSED_L2HINT
    double v;
    asm volatile("ld.global.L1::no_allocate.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(a), "l"(p));
    return v;
#else
    return __ldcg(a);
#endif
}
__device__ __forceinline__ void ws_st2(double2 *a, double2 v, unsigned long long p) {
#if KDSL_FUSED_L2HINT
    asm volatile("st.global.L1::no_allocate.L2::cache_hint.v2.f64 [%0], {%1, %2}, %3;" ::"l"(a), "d"(v.x), "d"(v.y), "l"(p) : "memory");
#else
    __stcg(a, v);
#endif
}

// split barrier for the pivot search (mbarrier in shared memory): every thread of team P arrives as soon as its warp's
// candidate key is published and waits only after it has applied the pending panel update, so the 22 DFMAs per pivot
// (which queue behind the other team's DMMAs) no longer sit between the candidates and the decision
__device__ __forceinline__ void mbar_init(unsigned addr, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(addr), "r"(count) : "memory");
}
__device__ __forceinline__ unsigned long long mbar_arrive(unsigned addr) {
    unsigned long long tok;
    asm volatile("mbarrier.arrive.shared::cta.b64 %0, [%1];" : "=l"(tok) : "r"(addr) : "memory");
    return tok;
}
__device__ __forceinline__ void mbar_wait(unsigned addr, unsigned long long tok) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "MBAR_WAIT_%=:\n\t"
        "mbarrier.try_wait.shared::cta.b64 p, [%0], %1;\n\t"
        "@!p bra MBAR_WAIT_%=;\n\t"
        "}" ::"r"(addr), "l"(tok) : "memory");
}

// Shared-memory carve-up and the per-item context.  The phases below are separate device functions that rebuild
// their pointers from the kernel parameters (constant bank) and re-read the item context from shared memory, so that
// nothing but the loop counters is live across the latency-critical pivot loop (one monolithic scope cost the pivot
// chain its registers: ptxas spilled inside it).
struct FusedCtx {
    const double *UT;       // transposed U of this species
    double *W;              // W of this walker and species
    int N, Np, M, Cp;       // particles, padded; unoccupied sites; columns of B (Np + Mp)
    int status_idx;
};
template <int NB, int NWARPS>
struct FusedSmem {
    double *sMb, *sX, *sStg, *sRow, *sRinv;
    unsigned *sKey;
    int *sIdx, *sPivRow, *sScan, *sStep, *sColSite, *sSiteInfo;
    FusedCtx *ctx;
    __device__ __forceinline__ FusedSmem(double *sm, int NpMax, int CpMax, int ns) {
        sMb = sm;                                           // [2][NpMax x NB] frag-major (r = row, k = q): R - E
        sX = sMb + (size_t)2 * NB * NpMax;                  // [CpMax x NB] frag-major (r = column, k = q): B[p_q, column]
        sRow = sX + (size_t)NB * CpMax;                     // [NB] the pivot row of the current step
        sRinv = sRow + 2 * NB;                              // [2]   (sRow is double buffered: [2][NB])
        ctx = reinterpret_cast<FusedCtx *>(sRinv + 2);      // (48 bytes reserved)
        sKey = reinterpret_cast<unsigned *>(sRinv + 2 + 6); // [8] per-warp candidate keys
        sIdx = reinterpret_cast<int *>(sKey + 8);           // [4]: [0] pivot row, [1] singular flag
        sPivRow = sIdx + 4;                                 // [NB] pivot row of each step of the panel being factored
        sScan = sPivRow + NB;                               // [NWARPS + 1 (+ padding to 32)]
        sStep = sScan + 32;                                 // [NpMax] elimination step at which row j was the pivot
        sColSite = sStep + NpMax;                           // [CpMax] row of UT behind column c of B
        sSiteInfo = sColSite + CpMax;                       // [ns] >= 0: index u of an unoccupied site, < 0: -label
        // [8][ns] output staging of the last step (8 columns of W), only allocated when the idle R - E buffer is
        // too small for it; last so that no other offset depends on it
        sStg = reinterpret_cast<double *>(sSiteInfo + ns + ((NB + NpMax + CpMax + ns) & 1));
    }
};
static_assert(sizeof(FusedCtx) <= 48, "FusedCtx must fit its reserved slot");
// what a phase function needs to rebuild the carve-up (passed by value: the phases are NOT inlined, each gets its own
// register allocation)
struct FusedArgs {
    double *sm;
    int NpMax, CpMax, ns;
};

// ---- team P (256 threads, thread j owns row j): factor the panel of columns [k0, k0 + kw); operands R - E -> sM.
//      Returns true when this thread's row became a pivot in this panel.  FIRST: the columns still live in UT. ----
template <int NB, bool FIRST, int TP = 256>
__device__ __noinline__ bool fused_factor_panel(const FusedArgs FA, const double *__restrict__ ws, int k0, int kw,
                                                   double *sM, bool pivoted) {
    const FusedSmem<NB, 16> L(FA.sm, FA.NpMax, FA.CpMax, FA.ns);
    const unsigned long long pol = fused_policy_keep();
    (void)pol;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int Np = L.ctx->Np, Cp = L.ctx->Cp;
    const bool has_row = tid < Np;
    double a[NB];
    int mypiv = -1;
    if (FIRST) {
        const double *UT = L.ctx->UT;
#pragma unroll
        for (int c = 0; c < NB; c++) a[c] = (has_row && c < kw) ? __ldcg(UT + (size_t)L.sColSite[k0 + c] * Np + tid) : 0.0;
    } else {
        const double2 *src = reinterpret_cast<const double2 *>(ws + (size_t)tid * Cp + k0);
#pragma unroll
        for (int c = 0; c < NB; c += 2) {
            const double2 v = (has_row && c < kw) ? ws_ld2(src + (c >> 1), pol) : make_double2(0.0, 0.0);
            a[c] = v.x;
            a[c + 1] = v.y;
        }
    }
    double *sRowB = L.sRow, *sRinv = L.sRinv;
    unsigned *sKey = L.sKey;
    const unsigned mbar = (unsigned)__cvta_generic_to_shared(L.sScan + 24);
    int *sIdx = L.sIdx, *sPivRow = L.sPivRow, *sStep = L.sStep;
    double tail = 0.0;
    bool pend = false, pend_p = false;
    auto apply_pending = [&](const double *sRow) {   // columns 2.. of the pending step (sRow still holds its pivot row)
        if (!has_row) return;
        const double2 *prow2 = reinterpret_cast<const double2 *>(sRow);
#pragma unroll
        for (int j = 2; j < NB; j += 2) {
            const double2 pv = prow2[j >> 1];
            if (pend_p) {
                a[j - 1] = a[j] * tail;
                a[j] = a[j + 1 < NB ? j + 1 : j] * tail;
            } else {
                a[j - 1] = fma(tail, pv.x, a[j]);
                a[j] = fma(tail, pv.y, a[j + 1 < NB ? j + 1 : j]);
            }
        }
        a[NB - 1] = tail;
    };
    bool singular = false;
#pragma unroll 1
    for (int k = 0; k < kw; k++) {
        const bool valid = has_row && !pivoted;
        const unsigned hi = valid ? ((unsigned)__double2hiint(a[0]) & 0x7fffffffu) : 0u;
        const unsigned mhi = __reduce_max_sync(0xffffffffu, hi);
        double my_rinv = rcp_fast(valid ? a[0] : 1.0);      // speculative reciprocal of my candidate
        asm volatile("" : "+d"(my_rinv));
        const unsigned win = __ballot_sync(0xffffffffu, valid && hi == mhi);
        const bool leader = win != 0u && lane == __ffs(win) - 1;
        if (lane == 0) sKey[warp] = (win != 0u) ? ((mhi & 0xfffffff8u) | (unsigned)(7 - warp)) : 0u;
        const unsigned long long tok = mbar_arrive(mbar);   // candidates published ...
        if (pend) apply_pending(sRowB + ((k + 1) & 1) * NB);  // ... the pending update of step k-1 runs while the others arrive
        double my_s = my_rinv * a[1];                        // scaled next-column entry of my row (now up to date)
        asm volatile("" : "+d"(my_s));
        mbar_wait(mbar, tok);
        unsigned bk;
        {
            const uint4 k0v = *reinterpret_cast<const uint4 *>(sKey);
            const uint4 k1v = *reinterpret_cast<const uint4 *>(sKey + 4);
            bk = max(max(max(k0v.x, k0v.y), max(k0v.z, k0v.w)), max(max(k1v.x, k1v.y), max(k1v.z, k1v.w)));
        }
        if ((bk >> 3) == 0u || bk >= 0x7ff00000u) {  // zero (below 2^-1039) / non-finite pivot: singular
            singular = true;                         // (uniform over team P)
            break;
        }
        const int wq = 7 - (int)(bk & 7u);
        if (warp == wq && leader) {
            sIdx[0] = tid;
            sPivRow[k] = tid;
            sStep[tid] = k0 + k;
            sRinv[0] = my_rinv;
            sRinv[1] = my_s;
            double2 *dst = reinterpret_cast<double2 *>(sRowB + (k & 1) * NB);
#pragma unroll
            for (int j = 0; j < NB; j += 2) dst[j >> 1] = make_double2(a[j], a[j + 1]);
        }
        bar_team<TP>();
        const int p = sIdx[0];
        const double2 rs = *reinterpret_cast<const double2 *>(sRinv);   // (1 / pivot, pivot row's next entry / pivot)
        if (has_row && tid == p) {
            pivoted = true;
            pend_p = true;
            mypiv = k;
            tail = rs.x;
            a[0] = rs.y;
        } else {
            pend_p = false;
            const double a0 = a[0];
            a[0] = fma(-a0, rs.y, a[1]);             // the ONE FP64 instruction between the barrier and the next search
            tail = -(a0 * rs.x);
        }
        pend = true;
    }
    if (singular) {
        if (tid == 0) { sIdx[1] = 1; }
        return pivoted;
    }
    if (pend) apply_pending(sRowB + ((kw + 1) & 1) * NB);
    // publish R - E in fragment order.  After kw rotations register slot cs holds panel column (cs + kw) mod NB
    // (columns >= kw are zero padding).
    if (has_row) {
#pragma unroll
        for (int cs = 0; cs < NB; cs += 4) {
            int col = cs + kw;
            if (col >= NB) col -= NB;
            double v[4];
#pragma unroll
            for (int e = 0; e < 4; e++) v[e] = a[cs + e];
#pragma unroll
            for (int e = 0; e < 4; e++) if (col + e == mypiv) v[e] -= 1.0;
            double2 *dst = reinterpret_cast<double2 *>(sM + frag_idx(tid, col, NB));
            dst[0] = make_double2(v[0], v[1]);
            dst[1] = make_double2(v[2], v[3]);
        }
    }
    return pivoted;
}

// ---- all T threads: raw pivot rows (kw of them, sPivRow) of the unfinished columns [c_lo, Cp): whole rows of the
//      row-major workspace, read coalesced; item = (column, four consecutive q) ----
template <int NB, int T, bool FIRST>
__device__ __noinline__ void fused_gather_X(const FusedArgs FA, const double *__restrict__ ws, int c_lo, int kw) {
    const FusedSmem<NB, 16> L(FA.sm, FA.NpMax, FA.CpMax, FA.ns);
    const unsigned long long pol = fused_policy_keep();
    (void)pol;
    constexpr int KS = NB / 4;
    const int Np = L.ctx->Np, Cp = L.ctx->Cp;
    const double *UT = L.ctx->UT;
    const int ncols = Cp - c_lo;
    for (int idx = threadIdx.x; idx < ncols * KS; idx += T) {
        const int qg = idx / ncols, j = c_lo + (idx - qg * ncols), q = qg << 2;
        double v[4];
        if (FIRST) {
            const double *col = UT + (size_t)L.sColSite[j] * Np;
#pragma unroll
            for (int e = 0; e < 4; e++) v[e] = (q + e < kw) ? __ldcg(col + L.sPivRow[q + e]) : 0.0;
        } else {
#pragma unroll
            for (int e = 0; e < 4; e++) v[e] = (q + e < kw) ? ws_ld1(ws + (size_t)L.sPivRow[q + e] * Cp + j, pol) : 0.0;
        }
        double2 *dst = reinterpret_cast<double2 *>(L.sX + frag_idx(j, q, NB));
        dst[0] = make_double2(v[0], v[1]);
        dst[1] = make_double2(v[2], v[3]);
    }
}

// ---- DMMA update B[:, J] += (R - E) X[:, J] of the column tiles [t_lo, t_hi); groups of CT tiles are dealt round-robin
//      to `nw_team` warps (this warp is `w_team`); D row tiles of loads in flight per warp.
//      FIRST: the columns still live in UT;  LAST: the results are the columns of W and leave through the staging
//      buffer as full coalesced segments (unit rows of the occupied sites included) ----
template <int NB, int CT, int D, bool FIRST>
__device__ __noinline__ void fused_update_cols(const FusedArgs FA, double *__restrict__ ws, const double *sM,
                                                  int t_lo, int t_hi, int w_team, int nw_team, int ns) {
    const FusedSmem<NB, 16> L(FA.sm, FA.NpMax, FA.CpMax, FA.ns);
    const unsigned long long pol = fused_policy_keep();
    (void)pol;
    constexpr int KS = NB / 4;
    const int lane = threadIdx.x & 31;
    const int gr = lane >> 2, tg = lane & 3;
    const int Np = L.ctx->Np, Cp = L.ctx->Cp;
    const int nrt = Np >> 3;
    const double *UT = L.ctx->UT;
    const double *sX = L.sX;
    const int *sColSite = L.sColSite;
    const int nct = t_hi - t_lo;
    const int groups = (nct + CT - 1) / CT;
    for (int g = w_team; g < groups; g += nw_team) {
        const int t0 = t_lo + g * CT;
        double xf[CT][KS];
        bool cv[CT];
        const double *u0p[CT], *u1p[CT];
#pragma unroll
        for (int c = 0; c < CT; c++) {
            cv[c] = g * CT + c < nct;
            const int ct = cv[c] ? t0 + c : t_lo;
#pragma unroll
            for (int s = 0; s < KS; s++) xf[c][s] = cv[c] ? sX[(((ct * KS) + s) << 5) + lane] : 0.0;
            if (FIRST) {
                const int col = (ct << 3) + 2 * tg;
                u0p[c] = UT + (size_t)sColSite[col] * Np + gr;
                u1p[c] = UT + (size_t)sColSite[col + 1] * Np + gr;
            }
        }
        double *rp = ws + (size_t)gr * Cp + (t0 << 3) + 2 * tg;
        // D row tiles of loads in flight in statically indexed registers.  The pipelined loop is branch free: it
        // covers the full blocks of D row tiles, its refill loads clamp the row tile to the last one (a few redundant
        // loads) and the remaining nrt % D tiles are finished from the registers afterwards.  (With `if (rt < nrt)`
        // inside the unrolled body ptxas keeps the queue in local memory: every load is followed by a spill store
        // that waits for it.  A rotating queue with register moves fails the same way: a move of a pending load's
        // destination waits for the load.  Processing the row tiles in pairs -- four accumulation chains per warp --
        // was measured 4 % slower: the two warps per SM sub-partition already cover the DMMA latency.)
        double2 buf[D][CT];
#define FUSED_LD(rt_, d_)                                                                                      \
    do {                                                                                                       \
        const int rl_ = min((rt_), nrt - 1);                                                                   \
        _Pragma("unroll") for (int c = 0; c < CT; c++) {                                                       \
            if (!cv[c]) (d_)[c] = make_double2(0.0, 0.0);                                                      \
            else if (FIRST) (d_)[c] = make_double2(__ldcg(u0p[c] + (rl_ << 3)), __ldcg(u1p[c] + (rl_ << 3)));   \
            else (d_)[c] = ws_ld2(reinterpret_cast<const double2 *>(rp + (size_t)(rl_ << 3) * Cp + c * 8), pol);     \
        }                                                                                                      \
    } while (0)
#define FUSED_MMA(rt_, d_)                                                                                     \
    do {                                                                                                       \
        double mf[KS];                                                                                         \
        _Pragma("unroll") for (int s = 0; s < KS; s++) mf[s] = sM[((((rt_) * KS) + s) << 5) + lane];           \
        _Pragma("unroll") for (int s = 0; s < KS; s++)                                                         \
            _Pragma("unroll") for (int c = 0; c < CT; c++) dmma_8x8x4((d_)[c].x, (d_)[c].y, mf[s], xf[c][s]);   \
    } while (0)
#define FUSED_ST(rt_, d_)                                                                                      \
    do {                                                                                                       \
        _Pragma("unroll") for (int c = 0; c < CT; c++)                                                         \
            if (cv[c]) ws_st2(reinterpret_cast<double2 *>(rp + (size_t)((rt_) << 3) * Cp + c * 8), (d_)[c], pol); \
    } while (0)
#define FUSED_TILE(rt_, d_) do { FUSED_MMA(rt_, d_); FUSED_ST(rt_, d_); } while (0)
#pragma unroll
        for (int i = 0; i < D; i++) FUSED_LD(i, buf[i]);
        int rt0 = 0;
#pragma unroll 1
        for (; rt0 + D <= nrt; rt0 += D) {
            // the store of a tile (and the refill of its registers) is issued after the DMMAs of the NEXT tile: the warp
            // issues in order, a store right behind its last DMMA would hold the next tile back for the DMMA latency
#pragma unroll
            for (int i = 0; i < D; i++) {
                FUSED_MMA(rt0 + i, buf[i]);
                if (i > 0) {
                    FUSED_ST(rt0 + i - 1, buf[i - 1]);
                    FUSED_LD(rt0 + i - 1 + D, buf[i - 1]);
                }
            }
            FUSED_ST(rt0 + D - 1, buf[D - 1]);
            FUSED_LD(rt0 + 2 * D - 1, buf[D - 1]);
        }
        const int rem = nrt - rt0;
#pragma unroll
        for (int i = 0; i < D - 1; i++)
            if (i < rem) FUSED_TILE(rt0 + i, buf[i]);
#undef FUSED_MMA
#undef FUSED_ST
#undef FUSED_LD
#undef FUSED_TILE
    }
}

// ---- the last block step, all 16 warps in lock step over the row tiles: warp g updates the CT2 = 2 column tiles of
//      group g of the V part (sites u0 .. u0 + 15 of the unoccupied list), deposits its 8 x 16 results in the staging
//      buffer stage[r][site]; after a barrier the CTA writes the 8 finished columns of W -- row (rt*8 + r) of B is
//      column sStep[..] of W -- as full, aligned, coalesced segments (thread = site; unit rows of the occupied sites
//      filled in on the way; every 32-byte sector is written whole, so the L2 never has to fetch one to merge). ----
template <int NB, int D, bool FIRST>
__device__ __noinline__ void fused_last_step(const FusedArgs FA, const double *__restrict__ ws, const double *sM, double *stage, int stage_cap, int ns) {
    const FusedSmem<NB, 16> L(FA.sm, FA.NpMax, FA.CpMax, FA.ns);
    const unsigned long long pol = fused_policy_keep();
    (void)pol;
    constexpr int KS = NB / 4, CT = 2;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int gr = lane >> 2, tg = lane & 3;
    const int Np = L.ctx->Np, Cp = L.ctx->Cp, N = L.ctx->N;
    const int nrt = Np >> 3;
    const double *UT = L.ctx->UT;
    double *W = L.ctx->W;
    const int *sColSite = L.sColSite, *sStep = L.sStep;
    const int t_lo = Np >> 3, nct = (Cp >> 3) - t_lo;        // V part: nct column tiles, <= 32 (host check)
    const int t0 = t_lo + warp * CT;
    double xf[CT][KS];
    bool cv[CT];
    const double *u0p[CT], *u1p[CT];
    const int Mp = Cp - Np;
    // the stage holds 8 rows x Mp unoccupied sites; two of them when they fit (then one barrier per row tile is
    // enough: the stores of tile rt overlap the DMMAs of tile rt + 1)
    const int nbuf = stage_cap >= 16 * Mp ? 2 : 1;
#pragma unroll
    for (int c = 0; c < CT; c++) {
        cv[c] = warp * CT + c < nct;
        const int ct = cv[c] ? t0 + c : t_lo;
#pragma unroll
        for (int s = 0; s < KS; s++) xf[c][s] = cv[c] ? L.sX[(((ct * KS) + s) << 5) + lane] : 0.0;
        if (FIRST) {
            const int col = (ct << 3) + 2 * tg;
            u0p[c] = UT + (size_t)sColSite[col] * Np + gr;
            u1p[c] = UT + (size_t)sColSite[col + 1] * Np + gr;
        }
    }
    double *dep = stage + gr * Mp + ((t0 - t_lo) << 3) + 2 * tg;   // my two accumulator columns = unoccupied sites u, u + 1
    const double *rp = ws + (size_t)gr * Cp + (t0 << 3) + 2 * tg;
    // the writer side: thread = site
    const bool wv = tid < ns;
    const int info = wv ? L.sSiteInfo[tid] : 0;              // >= 0: unoccupied, < 0: -label
    double2 buf[D][CT];
#define FUSED_LD(rt_, d_)                                                                                      \
    do {                                                                                                       \
        const int rl_ = min((rt_), nrt - 1);                                                                   \
        _Pragma("unroll") for (int c = 0; c < CT; c++) {                                                       \
            if (!cv[c]) (d_)[c] = make_double2(0.0, 0.0);                                                      \
            else if (FIRST) (d_)[c] = make_double2(__ldcg(u0p[c] + (rl_ << 3)), __ldcg(u1p[c] + (rl_ << 3)));   \
            else (d_)[c] = ws_ld2(reinterpret_cast<const double2 *>(rp + (size_t)(rl_ << 3) * Cp + c * 8), pol);     \
        }                                                                                                      \
    } while (0)
#define FUSED_TILE(rt_, d_)                                                                                    \
    do {                                                                                                       \
        double mf[KS];                                                                                         \
        _Pragma("unroll") for (int s = 0; s < KS; s++) mf[s] = sM[((((rt_) * KS) + s) << 5) + lane];           \
        _Pragma("unroll") for (int s = 0; s < KS; s++)                                                         \
            _Pragma("unroll") for (int c = 0; c < CT; c++) dmma_8x8x4((d_)[c].x, (d_)[c].y, mf[s], xf[c][s]);   \
        const int bo_ = (nbuf == 2 && ((rt_) & 1)) ? 8 * Mp : 0;                                               \
        _Pragma("unroll") for (int c = 0; c < CT; c++)                                                         \
            if (cv[c]) *reinterpret_cast<double2 *>(dep + bo_ + c * 8) = (d_)[c];                              \
        __syncthreads();                                                                                       \
        if (wv) {                                                                                              \
            _Pragma("unroll") for (int r = 0; r < 8; r++) {                                                    \
                const int t = sStep[((rt_) << 3) + r];         /* row (rt*8 + r) of B is column t of W */      \
                if (t < N)                                     /* (else: identity padding) */                  \
                    __stcs(W + (size_t)t * ns + tid, info >= 0 ? stage[bo_ + r * Mp + info] : ((-info - 1 == t) ? 1.0 : 0.0)); \
            }                                                                                                  \
        }                                                                                                      \
        if (nbuf == 1) __syncthreads();                                                                        \
    } while (0)
#pragma unroll
    for (int i = 0; i < D; i++) FUSED_LD(i, buf[i]);
    int rt0 = 0;
#pragma unroll 1
    for (; rt0 + D <= nrt; rt0 += D) {
#pragma unroll
        for (int i = 0; i < D; i++) {
            FUSED_TILE(rt0 + i, buf[i]);
            FUSED_LD(rt0 + i + D, buf[i]);
        }
    }
    const int rem = nrt - rt0;
#pragma unroll
    for (int i = 0; i < D - 1; i++)
        if (i < rem) FUSED_TILE(rt0 + i, buf[i]);
#undef FUSED_LD
#undef FUSED_TILE
}

// ---- phase A: the kn columns of the next panel (from column c0), all 16 warps: item = row tile (at most two per warp,
//      both loaded up front) ----
template <int NB, bool FIRST, int NWARPS = 16>
__device__ __noinline__ void fused_update_next_panel(const FusedArgs FA, double *__restrict__ ws, const double *sM,
                                                        int c0, int kn) {
    const FusedSmem<NB, 16> L(FA.sm, FA.NpMax, FA.CpMax, FA.ns);
    const unsigned long long pol = fused_policy_keep();
    (void)pol;
    constexpr int KS = NB / 4, NT = NB / 8;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int gr = lane >> 2, tg = lane & 3;
    const int Np = L.ctx->Np, Cp = L.ctx->Cp;
    const int nrt = Np >> 3;
    const double *UT = L.ctx->UT;
    const int nt = kn >> 3, t0 = c0 >> 3;
    double xf[NT][KS];
    const double *u0p[NT], *u1p[NT];
#pragma unroll
    for (int c = 0; c < NT; c++) {
        const int ct = c < nt ? t0 + c : t0;
#pragma unroll
        for (int s = 0; s < KS; s++) xf[c][s] = c < nt ? L.sX[(((ct * KS) + s) << 5) + lane] : 0.0;
        if (FIRST) {
            const int col = (ct << 3) + 2 * tg;
            u0p[c] = UT + (size_t)L.sColSite[col] * Np + gr;
            u1p[c] = UT + (size_t)L.sColSite[col + 1] * Np + gr;
        }
    }
    double *rp = ws + (size_t)gr * Cp + c0 + 2 * tg;
    double2 d[2][NT];
#pragma unroll
    for (int i = 0; i < 2; i++) {
        const int rt = warp + i * NWARPS;
#pragma unroll
        for (int c = 0; c < NT; c++) {
            if (rt >= nrt || c >= nt) d[i][c] = make_double2(0.0, 0.0);
            else if (FIRST) d[i][c] = make_double2(__ldcg(u0p[c] + (rt << 3)), __ldcg(u1p[c] + (rt << 3)));
            else d[i][c] = ws_ld2(reinterpret_cast<const double2 *>(rp + (size_t)(rt << 3) * Cp + c * 8), pol);
        }
    }
#pragma unroll
    for (int i = 0; i < 2; i++) {
        const int rt = warp + i * NWARPS;
        if (rt < nrt) {
            double mf[KS];
#pragma unroll
            for (int s = 0; s < KS; s++) mf[s] = sM[(((rt * KS) + s) << 5) + lane];
#pragma unroll
            for (int s = 0; s < KS; s++)
#pragma unroll
                for (int c = 0; c < NT; c++) dmma_8x8x4(d[i][c].x, d[i][c].y, mf[s], xf[c][s]);
#pragma unroll
            for (int c = 0; c < NT; c++)
                if (c < nt) ws_st2(reinterpret_cast<double2 *>(rp + (size_t)(rt << 3) * Cp + c * 8), d[i][c], pol);
        }
    }
}

// ---- item set-up, all T threads: context and the column -> site tables (tilde_U^T columns in label order, then the
//      unoccupied sites in ascending order) ----
template <int NB, int T>
__device__ __noinline__ void fused_item_setup(const FusedArgs FA, const DevState &S, const int *__restrict__ list,
                                                 const double *UT_up, const double *UT_dn, int *__restrict__ status,
                                                 int Np_up, int Np_dn, int item) {
    const FusedSmem<NB, 16> L(FA.sm, FA.NpMax, FA.CpMax, FA.ns);
    const unsigned long long pol = fused_policy_keep();
    (void)pol;
    constexpr int NWARPS = T / 32;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = item >> 1, spin = item & 1;
    const int w = list ? list[b] : b;
    const int ns = S.ns;
    const int N = spin ? S.n_dn : S.n_up, Np = spin ? Np_dn : Np_up;
    const int M = ns - N, Mp = (M + 7) & ~7;
    const int *kap = (spin ? S.kdn : S.kup) + (size_t)w * ns;
    if (tid == 0) {
        FusedCtx c;
        c.UT = spin ? UT_dn : UT_up;
        c.W = (spin ? S.W_dn : S.W_up) + (size_t)w * ns * N;
        c.N = N; c.Np = Np; c.M = M; c.Cp = Np + Mp;
        c.status_idx = 2 * b + spin;
        *L.ctx = c;
        L.sScan[NWARPS] = 0;
        L.sIdx[1] = 0;
        status[2 * b + spin] = 0;
    }
    for (int c = N + tid; c < Np; c += T) L.sColSite[c] = ns + (c - N);
    for (int u = M + tid; u < Mp; u += T) L.sColSite[Np + u] = ns + 8;
    __syncthreads();
    for (int s0 = 0; s0 < ns; s0 += T) {
        const int site = s0 + tid;
        const int l = site < ns ? kap[site] : -1;
        const bool un = l == 0;
        const unsigned m = __ballot_sync(0xffffffffu, un);
        if (lane == 0) L.sScan[warp] = __popc(m);
        __syncthreads();
        if (site < ns) {
            if (un) {
                int off = L.sScan[NWARPS];
                for (int q = 0; q < warp; q++) off += L.sScan[q];
                const int u = off + __popc(m & ((1u << lane) - 1u));
                L.sColSite[Np + u] = site;
                L.sSiteInfo[site] = u;
            } else {
                L.sColSite[l - 1] = site;
                L.sSiteInfo[site] = -l;
            }
        }
        __syncthreads();
        if (tid == 0) {
            int t = 0;
            for (int q = 0; q < NWARPS; q++) t += L.sScan[q];
            L.sScan[NWARPS] += t;
        }
        __syncthreads();
    }
}

// T = 512 (one CTA per SM; teams of 8 warps; N <= 256, ns <= 512) is the production instantiation; T = 256 (two CTAs per
// SM; teams of 4 warps; N <= 128, M <= 128, ns <= 256) doubles the matrices in flight on small lattices, where the kernel
// is bound by the latency of one pivot chain per SM (192 sites: 96 pivots per matrix).
template <int NB, int CT, int DG = 6, int DBG = 0, int T = 512>
__global__ void __launch_bounds__(T, T == 512 ? 1 : 2)
k_reeval_fused(DevState S, const int *__restrict__ list, double *__restrict__ ws_base, size_t ws_stride,
               const double *__restrict__ UT_up, const double *__restrict__ UT_dn, int *__restrict__ status,
               int Np_up, int Np_dn, int NpMax, int CpMax, int stage_doubles) {
    constexpr int NWARPS = T / 32, GW = NWARPS / 2, TP = T / 2;
    static_assert(NB % 8 == 0 && (T == 512 || T == 256), "panel width must be a multiple of 8");
    extern __shared__ double sm[];
    const FusedSmem<NB, NWARPS> L(sm, NpMax, CpMax, S.ns);
    const FusedArgs FA{sm, NpMax, CpMax, S.ns};
    const int n_items = 2 * batch_count(S, list);
    double *ws = ws_base + (size_t)blockIdx.x * ws_stride;
    if (threadIdx.x == 0) mbar_init((unsigned)__cvta_generic_to_shared(L.sScan + 24), TP);
    if (threadIdx.x < 8) L.sKey[threadIdx.x] = 0u;          // (keys of warps that do not exist never win)
    __syncthreads();

    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        fused_item_setup<NB, T>(FA, S, list, UT_up, UT_dn, status, Np_up, Np_dn, item);
        const bool teamP = threadIdx.x < TP;
        bool pivoted = false;                               // my row has been a pivot
        const int Np = L.ctx->Np;

        // ---- step 0 panel ----
        const int kw0 = min(NB, Np);
        if (teamP) pivoted = fused_factor_panel<NB, true, TP>(FA, ws, 0, kw0, L.sMb, pivoted);
        __syncthreads();
        if (!L.sIdx[1]) {
            fused_gather_X<NB, T, true>(FA, ws, kw0, kw0);
            __syncthreads();
            long long t_phase = PHASE_CLOCK();
#define FUSED_TICK(idx, thr) PHASE_TICK_AT(idx, thr)
            for (int k0 = 0, s = 0; k0 < Np; k0 += NB, s++) {
                const int kw = min(NB, Np - k0);            // multiple of 8
                const int k1 = k0 + kw, kn = min(NB, Np - k1);  // next panel (kn <= 0: none)
                const bool first = s == 0;
                const double *sM = L.sMb + (size_t)(s & 1) * NB * Np;
                double *sMn = L.sMb + (size_t)((s + 1) & 1) * NB * Np;
                const int ntc = L.ctx->Cp >> 3;
                PHASE_RESTART(TP);
                FUSED_TICK(4, 0);                           // (loop overhead)
                if (kn > 0) {
                    if (first) fused_update_next_panel<NB, true, NWARPS>(FA, ws, sM, k1, kn);
                    else fused_update_next_panel<NB, false, NWARPS>(FA, ws, sM, k1, kn);
                    __syncthreads();
                    FUSED_TICK(0, 0);
                    PHASE_RESTART(TP);
                    if (teamP) {
                        if (DBG != 2) pivoted = fused_factor_panel<NB, false, TP>(FA, ws, k1, kn, sMn, pivoted);
                        else if (threadIdx.x < kn) { L.sPivRow[threadIdx.x] = k1 + threadIdx.x; L.sStep[k1 + threadIdx.x] = k1 + threadIdx.x; }
                        FUSED_TICK(1, 0);
                    } else if (DBG != 1) {
                        const int wg = (threadIdx.x >> 5) - GW;
                        if (first) fused_update_cols<NB, CT, DG, true>(FA, ws, sM, (k1 + kn) >> 3, ntc, wg, GW, S.ns);
                        else fused_update_cols<NB, CT, DG, false>(FA, ws, sM, (k1 + kn) >> 3, ntc, wg, GW, S.ns);
                        FUSED_TICK(2, TP);
                    }
                    __syncthreads();
                    FUSED_TICK(5, 0);                       // team P waiting for team G
                    if (L.sIdx[1]) break;                   // singular: W of this species stays as it was
                    fused_gather_X<NB, T, false>(FA, ws, k1 + kn, kn);
                    __syncthreads();
                    FUSED_TICK(3, 0);
                } else {
                    // output staging: the idle R - E buffer when it holds 8 columns of W, else the dedicated area
                    double *stage = stage_doubles == 0 ? sMn : L.sStg;
                    const int stage_cap = stage_doubles == 0 ? NB * Np : stage_doubles;
                    if (first) fused_last_step<NB, 4, true>(FA, ws, sM, stage, stage_cap, S.ns);
                    else fused_last_step<NB, 4, false>(FA, ws, sM, stage, stage_cap, S.ns);
                    FUSED_TICK(6, 0);
                }
            }
        }
        __syncthreads();
        if (threadIdx.x == 0 && L.sIdx[1]) status[L.ctx->status_idx] = 1;
        __syncthreads();                                    // the tables and the workspace are reused by the next item
    }
}
