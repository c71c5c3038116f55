// kdsl_reeval_cl.cuh -- reevaluateW! (src/MonteCarlo.jl:55-66) + tilde_U (src/MonteCarlo.jl:92-115) for Np <= 512
// (built for 972 sites: N = 486; inverse_variant 8) as ONE kernel over THREAD-BLOCK CLUSTERS.  Measured:
// slower than cluster inverse + product at 972 sites (68 ms vs 59 ms per 2048-walker bin) and slower than the one-CTA
// k_reeval_fused at 432 sites (25.8 ms vs 16.3 ms with two 256-thread CTAs per SM), so it is a selectable variant only: the mathematics of k_reeval_fused (Gauss-Jordan with implicit
// row pivoting on B = [tilde_U^T | V^T], N x (N + M); the V^T columns end as the non-trivial rows of W; finished tilde_U^T
// columns are dropped: N^3 + 2 N^2 M = 3 N^3 flop instead of inverse + product = 4 N^3) with the work split of k_inverse_cl:
//
//   one matrix per cluster of CL CTAs, 148 / CL matrices live, so the row-major workspaces (3.8 MB each at 972 sites) stay in
//   the 126 MB L2 where one CTA per matrix would keep 560 MB live;
//   CTA 0 ("P", thread = row of B = orbital): brings the columns of panel s+1 up to date with the operands of step s (DMMA,
//       all row tiles of a warp loaded up front), then factors panel s+1 -- the latency-bound pivot chain has an SM (and its
//       FP64 pipe) to itself;
//   CTAs 1 .. CL-1 ("G"): block step s, B[:, J] += (R_s - E) B_old[(p_q), J], for the unfinished column-tile pairs dealt
//       round-robin to them; a CTA gathers the raw pivot rows of ITS columns (coalesced row segments of the row-major workspace)
//       before it updates them; items = (pair, row slice) over the 16 warps, D row tiles of loads in flight per warp;
//   first step: the columns are read straight from the transposed copy of U (UT[site][:], shared by all walkers) -- there is
//       no gather pass; last step: every CTA of the cluster (P included) takes a contiguous block of the V columns = a
//       contiguous range of sites, computes the final rows and stores W itself: 8 columns of W per row tile go through a
//       double-buffered shared-memory stage and leave as full coalesced segments (thread = site, unit rows of the occupied
//       sites filled in on the way);
//   exchange per block step through a double-buffered per-cluster scratch in global memory (operands R_s - E in DMMA fragment
//       order, NB pivot rows, singular flag, the row -> elimination-step map) and ONE barrier.cluster.
// The pivot rule is k_inverse_v4's (largest |x| by its top 28 bits over the rows that were never a pivot, ties to the lowest
// row; zero or non-finite pivot => singular, W of that species is left untouched).
#pragma once
#include "kdsl_common.cuh"
#include "kdsl_refresh.cuh"
#include "kdsl_refresh_fast.cuh"
#include "kdsl_inverse_v5.cuh"
#include "kdsl_inverse_cl.cuh"
#include "kdsl_reeval_fused.cuh"

// Header of the dynamic shared memory: kernel-wide layout, the per-item context and the small exchange arrays of the pivot
// loop.  The phases are separate NOT-inlined functions that take this one pointer (a by-value argument struct above 20 bytes
// goes through local memory; one monolithic scope costs the pivot chain its registers: see kdsl_reeval_fused.cuh).
struct RclShared {
    int NpMax, CpMax, ns, XT, SW, NG, rank, RS;     // XT: X slots (column tiles) per CTA; SW: stage width (columns)
    const double *UT;       // transposed U of this species
    double *W;              // W of this walker and species
    double *ws;             // row-major workspace of this cluster [Np][Cp]
    double *Mg;             // [2][NB * NpMax] operands in fragment order (global scratch)
    int *Pg;                // [2][KDSL_CL_PGS] pivot rows + singular flag
    int *Sg;                // [NpMax] row -> elimination step
    int N, Np, M, Cp, status_idx, pad0, pad1, pad2;
    double sRow[32];
    double sRinv[2];
    unsigned sKey[16];
    int sIdx[4];
    int sPivRow[32];
    int sScan[20];
};
static_assert(sizeof(RclShared) % 16 == 0, "arrays behind the header are 16-byte aligned");

template <int NB>
struct RclArrays {
    double *sM, *sX, *stage;
    int *sColSite, *sSiteInfo, *sStep;
    __device__ __forceinline__ RclArrays(RclShared *H) {
        sM = reinterpret_cast<double *>(H + 1);             // [NpMax x NB] frag-major: R - E of the current step
        sX = sM + (size_t)NB * H->NpMax;                    // [XT tiles x 8 x NB] frag-major by LOCAL column: B[p_q, column]
        stage = sX + (size_t)NB * 8 * H->XT;                // [2][8][SW] output stage of the last step
        sColSite = reinterpret_cast<int *>(stage + (size_t)16 * H->SW);   // [CpMax] row of UT behind column c of B
        sSiteInfo = sColSite + H->CpMax;                    // [ns] >= 0: index u of an unoccupied site, < 0: -label
        sStep = sSiteInfo + H->ns;                          // [NpMax] elimination step at which row j was the pivot
    }
};
inline size_t reeval_cl_smem(int NB, int NpMax, int CpMax, int ns, int XT, int SW) {
    return sizeof(RclShared) + ((size_t)NB * NpMax + (size_t)NB * 8 * XT + (size_t)16 * SW) * sizeof(double) +
           ((size_t)CpMax + ns + NpMax + 4) * sizeof(int);
}
// doubles of per-cluster scratch: workspace, two operand buffers, pivot-row buffers, step map
__host__ __device__ inline size_t reeval_cl_scratch_doubles(int NB, int NpMax, int CpMax) {
    return (size_t)NpMax * CpMax + (size_t)2 * NB * NpMax + (2 * KDSL_CL_PGS * sizeof(int) + 7) / 8 + ((size_t)NpMax * sizeof(int) + 7) / 8 + 2;
}

// ---- item set-up, all threads of every CTA: context and the column -> site tables (tilde_U^T columns in label order, then
//      the unoccupied sites in ascending order) ----
template <int NB, int T>
__device__ __noinline__ void rcl_setup(RclShared *H, const DevState &S, const int *__restrict__ list, const double *UT_up,
                                        const double *UT_dn, int *__restrict__ status, int Np_up, int Np_dn, int item) {
    const RclArrays<NB> L(H);
    constexpr int NWARPS = T / 32;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = item >> 1, spin = item & 1;
    const int w = list ? list[b] : b;
    const int ns = S.ns;
    const int N = spin ? S.n_dn : S.n_up, Np = spin ? Np_dn : Np_up;
    const int M = ns - N, Mp = (M + 7) & ~7;
    const int *kap = (spin ? S.kdn : S.kup) + (size_t)w * ns;
    if (tid == 0) {
        H->UT = spin ? UT_dn : UT_up;
        H->W = (spin ? S.W_dn : S.W_up) + (size_t)w * ns * N;
        H->N = N; H->Np = Np; H->M = M; H->Cp = Np + Mp;
        H->status_idx = 2 * b + spin;
        H->sScan[NWARPS] = 0;
        if (H->rank == 0) status[2 * b + spin] = 0;
    }
    for (int c = N + tid; c < Np; c += T) L.sColSite[c] = ns + (c - N);
    for (int u = M + tid; u < Mp; u += T) L.sColSite[Np + u] = ns + 8;
    __syncthreads();
    for (int s0 = 0; s0 < ns; s0 += T) {
        const int site = s0 + tid;
        const int l = site < ns ? kap[site] : -1;
        const bool un = l == 0;
        const unsigned m = __ballot_sync(0xffffffffu, un);
        if (lane == 0) H->sScan[warp] = __popc(m);
        __syncthreads();
        if (site < ns) {
            if (un) {
                int off = H->sScan[NWARPS];
                for (int q = 0; q < warp; q++) off += H->sScan[q];
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
            for (int q = 0; q < NWARPS; q++) t += H->sScan[q];
            H->sScan[NWARPS] += t;
        }
        __syncthreads();
    }
}

// ---- P (512 threads, thread j owns row j): factor the panel of columns [k0, k0 + kw); operands R - E -> sM and the
//      scratch buffer `buf`, pivot rows and the singular flag -> Pg[buf], row -> step map -> Sg.  Returns the updated
//      `pivoted` flag of this thread's row.  FIRST: the columns still live in UT. ----
template <int NB, bool FIRST>
__device__ __noinline__ bool rcl_factor_panel(RclShared *H, int k0, int kw, int buf, bool pivoted) {
    const RclArrays<NB> L(H);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int Np = H->Np, Cp = H->Cp;
    const bool has_row = tid < Np;
    double *Mgb = H->Mg + (size_t)buf * NB * H->NpMax;
    int *Pgb = H->Pg + buf * KDSL_CL_PGS;
    double a[NB];
    int mypiv = -1;
    if (FIRST) {
        const double *UT = H->UT;
#pragma unroll
        for (int c = 0; c < NB; c++) a[c] = (has_row && c < kw) ? __ldcg(UT + (size_t)L.sColSite[k0 + c] * Np + tid) : 0.0;
    } else {
        const double2 *src = reinterpret_cast<const double2 *>(H->ws + (size_t)tid * Cp + k0);
#pragma unroll
        for (int c = 0; c < NB; c += 2) {
            const double2 v = (has_row && c < kw) ? __ldcg(src + (c >> 1)) : make_double2(0.0, 0.0);
            a[c] = v.x;
            a[c + 1] = v.y;
        }
    }
    double *sRow = H->sRow, *sRinv = H->sRinv;
    unsigned *sKey = H->sKey;
    int *sIdx = H->sIdx, *sPivRow = H->sPivRow;
    double tail = 0.0;
    bool pend = false, pend_p = false;
    auto apply_pending = [&]() {                    // columns 2.. of the pending step (sRow still holds its pivot row)
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
        double my_rinv = rcp_fast(valid ? a[0] : 1.0);
        asm volatile("" : "+d"(my_rinv));
        if (pend) apply_pending();
        double my_s = my_rinv * a[1];
        asm volatile("" : "+d"(my_s));
        const unsigned win = __ballot_sync(0xffffffffu, valid && hi == mhi);
        const bool leader = win != 0u && lane == __ffs(win) - 1;
        if (lane == 0) sKey[warp] = (win != 0u) ? ((mhi & 0xfffffff0u) | (unsigned)(15 - warp)) : 0u;
        __syncthreads();
        unsigned bk;
        {
            const uint4 v0 = *reinterpret_cast<const uint4 *>(sKey);
            const uint4 v1 = *reinterpret_cast<const uint4 *>(sKey + 4);
            const uint4 v2 = *reinterpret_cast<const uint4 *>(sKey + 8);
            const uint4 v3 = *reinterpret_cast<const uint4 *>(sKey + 12);
            const unsigned m0 = max(max(v0.x, v0.y), max(v0.z, v0.w)), m1 = max(max(v1.x, v1.y), max(v1.z, v1.w));
            const unsigned m2 = max(max(v2.x, v2.y), max(v2.z, v2.w)), m3 = max(max(v3.x, v3.y), max(v3.z, v3.w));
            bk = max(max(m0, m1), max(m2, m3));
        }
        if ((bk >> 4) == 0u || bk >= 0x7ff00000u) {      // zero (below 2^-1038) / non-finite pivot: singular
            singular = true;                             // (uniform over the CTA)
            break;
        }
        const int wq = 15 - (int)(bk & 15u);
        if (warp == wq && leader) {
            sIdx[0] = tid;
            sPivRow[k] = tid;
            L.sStep[tid] = k0 + k;
            __stcg(H->Sg + tid, k0 + k);
            sRinv[0] = my_rinv;
            sRinv[1] = my_s;
            double2 *dst = reinterpret_cast<double2 *>(sRow);
#pragma unroll
            for (int j = 0; j < NB; j += 2) dst[j >> 1] = make_double2(a[j], a[j + 1]);
        }
        __syncthreads();
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
            a[0] = fma(-a0, rs.y, a[1]);
            tail = -(a0 * rs.x);
        }
        pend = true;
    }
    if (singular) {
        if (tid == 0) __stcg(Pgb + NB, 1);
        return pivoted;
    }
    if (pend) apply_pending();
    // publish R - E in fragment order (shared memory + scratch).  After kw rotations register slot cs holds panel column
    // (cs + kw) mod NB (columns >= kw are zero padding).
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
            const int fi = frag_idx(tid, col, NB);
            double2 *dst = reinterpret_cast<double2 *>(L.sM + fi);
            dst[0] = make_double2(v[0], v[1]);
            dst[1] = make_double2(v[2], v[3]);
            double2 *dg = reinterpret_cast<double2 *>(Mgb + fi);
            __stcg(dg, make_double2(v[0], v[1]));
            __stcg(dg + 1, make_double2(v[2], v[3]));
        }
    }
    if (tid < kw) __stcg(Pgb + tid, sPivRow[tid]);
    if (tid == NB) __stcg(Pgb + NB, 0);
    return pivoted;
}

// raw pivot-row entries X[q .. q+3, column j] (kw pivots in sPivRow) -> sX slot `slot_col` (local column index)
template <int NB, bool FIRST>
__device__ __forceinline__ void rcl_gather_one(RclShared *H, const RclArrays<NB> &L, int j, int slot_col, int q, int kw) {
    double v[4];
    if (FIRST) {
        const double *col = H->UT + (size_t)L.sColSite[j] * H->Np;
#pragma unroll
        for (int e = 0; e < 4; e++) v[e] = (q + e < kw) ? __ldcg(col + H->sPivRow[q + e]) : 0.0;
    } else {
        const double *ws = H->ws;
        const int Cp = H->Cp;
#pragma unroll
        for (int e = 0; e < 4; e++) v[e] = (q + e < kw) ? __ldcg(ws + (size_t)H->sPivRow[q + e] * Cp + j) : 0.0;
    }
    double2 *dst = reinterpret_cast<double2 *>(L.sX + frag_idx(slot_col, q, NB));
    dst[0] = make_double2(v[0], v[1]);
    dst[1] = make_double2(v[2], v[3]);
}

// ---- P, all 16 warps: the kn columns of the next panel (from column k1) with the operands of step s (sM, sPivRow, kw pivots).
//      Item = row tile; a warp owns row tiles warp, warp + 16, ... (at most NI = 4: Np <= 512) and loads them all up front. ----
template <int NB, bool FIRST, int T>
__device__ __noinline__ void rcl_next_panel(RclShared *H, int k1, int kn, int kw) {
    const RclArrays<NB> L(H);
    constexpr int KS = NB / 4, NT = NB / 8, NWARPS = T / 32, NI = 4;   // NI * NWARPS row tiles: Np <= 512 (T = 512), 256 (T = 256)
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int gr = lane >> 2, tg = lane & 3;
    const int Np = H->Np, Cp = H->Cp;
    const int nrt = Np >> 3;
    for (int idx = tid; idx < kn * KS; idx += T) {
        const int qg = idx / kn, jl = idx - qg * kn;
        rcl_gather_one<NB, FIRST>(H, L, k1 + jl, jl, qg << 2, kw);
    }
    __syncthreads();
    const double *UT = H->UT;
    double *ws = H->ws;
    const int nt = kn >> 3;
    double xf[NT][KS];
    const double *u0p[NT], *u1p[NT];
#pragma unroll
    for (int c = 0; c < NT; c++) {
#pragma unroll
        for (int s = 0; s < KS; s++) xf[c][s] = c < nt ? L.sX[(((c * KS) + s) << 5) + lane] : 0.0;
        if (FIRST) {
            const int col = k1 + (c < nt ? c : 0) * 8 + 2 * tg;
            u0p[c] = UT + (size_t)L.sColSite[col] * Np + gr;
            u1p[c] = UT + (size_t)L.sColSite[col + 1] * Np + gr;
        }
    }
    double *rp = ws + (size_t)gr * Cp + k1 + 2 * tg;
    double2 d[NI][NT];
#pragma unroll
    for (int i = 0; i < NI; i++) {
        const int rt = warp + i * NWARPS;
#pragma unroll
        for (int c = 0; c < NT; c++) {
            if (rt >= nrt || c >= nt) d[i][c] = make_double2(0.0, 0.0);
            else if (FIRST) d[i][c] = make_double2(__ldcg(u0p[c] + (rt << 3)), __ldcg(u1p[c] + (rt << 3)));
            else d[i][c] = __ldcg(reinterpret_cast<const double2 *>(rp + (size_t)(rt << 3) * Cp + c * 8));
        }
    }
#pragma unroll
    for (int i = 0; i < NI; i++) {
        const int rt = warp + i * NWARPS;
        if (rt < nrt) {
            double mf[KS];
#pragma unroll
            for (int s = 0; s < KS; s++) mf[s] = L.sM[(((rt * KS) + s) << 5) + lane];
#pragma unroll
            for (int s = 0; s < KS; s++)
#pragma unroll
                for (int c = 0; c < NT; c++) dmma_8x8x4(d[i][c].x, d[i][c].y, mf[s], xf[c][s]);
#pragma unroll
            for (int c = 0; c < NT; c++)
                if (c < nt) __stcg(reinterpret_cast<double2 *>(rp + (size_t)(rt << 3) * Cp + c * 8), d[i][c]);
        }
    }
}

// ---- G: operands and pivot rows of step s: scratch -> shared memory ----
template <int NB, int T>
__device__ __noinline__ void rcl_load_operands(RclShared *H, int s) {
    const RclArrays<NB> L(H);
    const int tid = threadIdx.x;
    const double2 *src = reinterpret_cast<const double2 *>(H->Mg + (size_t)(s & 1) * NB * H->NpMax);
    double2 *dst = reinterpret_cast<double2 *>(L.sM);
    const int n2 = (NB * H->Np) >> 1;
    for (int i = tid; i < n2; i += T) dst[i] = __ldcg(src + i);
    if (tid < NB) H->sPivRow[tid] = __ldcg(H->Pg + (s & 1) * KDSL_CL_PGS + tid);
}

// ---- G: block step with the operands in sM: the unfinished column tiles [tf, Cp / 8) in pairs; pair pi (tiles 2 pi, 2 pi + 1)
//      belongs to G CTA (pi - tf / 2) mod NG and sits in X slots 2 li, 2 li + 1 (li = its local index).  First the raw pivot rows
//      of my columns, then items = (pair, row slice) over the 16 warps.  D row tiles of loads in flight per warp. ----
template <int NB, int D, bool FIRST, int T>
__device__ __noinline__ void rcl_update(RclShared *H, int tf, int kw) {
    const RclArrays<NB> L(H);
    constexpr int KS = NB / 4, CT = 2, NWARPS = T / 32;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int gr = lane >> 2, tg = lane & 3;
    const int Np = H->Np, Cp = H->Cp, NG = H->NG, grank = H->rank - 1, RS = H->RS;
    const int nrt = Np >> 3, ntc = Cp >> 3;
    const int pf = tf >> 1, PA = ((ntc + 1) >> 1) - pf;           // first active pair, active pairs
    const int n_my = PA > grank ? (PA - grank + NG - 1) / NG : 0;
    // raw pivot rows of my columns: item = (local pair, q group, column within the pair)
    for (int idx = tid; idx < n_my * KS * 16; idx += T) {
        const int jl = idx & 15, r = idx >> 4;
        const int li = r % n_my, qg = r / n_my;
        const int j = ((pf + grank + li * NG) << 4) + jl;
        if (j >= (tf << 3) && j < Cp) rcl_gather_one<NB, FIRST>(H, L, j, (li << 4) + jl, qg << 2, kw);
    }
    __syncthreads();
    const double *UT = H->UT;
    double *ws = H->ws;
    const double *sM = L.sM, *sX = L.sX;
    const int *sColSite = L.sColSite;
    const int items = n_my * RS;
    for (int it = warp; it < items; it += NWARPS) {
        const int li = it / RS, rs = it - li * RS;
        const int t0 = (pf + grank + li * NG) << 1;
        const int r_lo = rs * nrt / RS, r_hi = (rs + 1) * nrt / RS;
        if (r_lo >= r_hi) continue;
        double xf[CT][KS];
        bool cv[CT];
        const double *u0p[CT], *u1p[CT];
#pragma unroll
        for (int c = 0; c < CT; c++) {
            cv[c] = t0 + c >= tf && t0 + c < ntc;
#pragma unroll
            for (int s = 0; s < KS; s++) xf[c][s] = cv[c] ? sX[(((((li << 1) + c) * KS) + s) << 5) + lane] : 0.0;
            if (FIRST) {
                const int col = ((cv[c] ? t0 + c : tf) << 3) + 2 * tg;
                u0p[c] = UT + (size_t)sColSite[col] * Np + gr;
                u1p[c] = UT + (size_t)sColSite[col + 1] * Np + gr;
            }
        }
        double *rp = ws + (size_t)gr * Cp + (t0 << 3) + 2 * tg;
        double2 buf[D][CT];
#define RCL_LD(rt_, d_)                                                                                        \
    do {                                                                                                       \
        const int rl_ = min((rt_), r_hi - 1);                                                                  \
        _Pragma("unroll") for (int c = 0; c < CT; c++) {                                                       \
            if (!cv[c]) (d_)[c] = make_double2(0.0, 0.0);                                                      \
            else if (FIRST) (d_)[c] = make_double2(__ldcg(u0p[c] + (rl_ << 3)), __ldcg(u1p[c] + (rl_ << 3)));   \
            else (d_)[c] = __ldcg(reinterpret_cast<const double2 *>(rp + (size_t)(rl_ << 3) * Cp + c * 8));    \
        }                                                                                                      \
    } while (0)
#define RCL_MMA(rt_, d_)                                                                                       \
    do {                                                                                                       \
        double mf[KS];                                                                                         \
        _Pragma("unroll") for (int s = 0; s < KS; s++) mf[s] = sM[((((rt_) * KS) + s) << 5) + lane];           \
        _Pragma("unroll") for (int s = 0; s < KS; s++)                                                         \
            _Pragma("unroll") for (int c = 0; c < CT; c++) dmma_8x8x4((d_)[c].x, (d_)[c].y, mf[s], xf[c][s]);   \
    } while (0)
#define RCL_ST(rt_, d_)                                                                                        \
    do {                                                                                                       \
        _Pragma("unroll") for (int c = 0; c < CT; c++)                                                         \
            if (cv[c]) __stcg(reinterpret_cast<double2 *>(rp + (size_t)((rt_) << 3) * Cp + c * 8), (d_)[c]);   \
    } while (0)
#pragma unroll
        for (int i = 0; i < D; i++) RCL_LD(r_lo + i, buf[i]);
        int rt0 = r_lo;
#pragma unroll 1
        for (; rt0 + D <= r_hi; rt0 += D) {
            // the store of a tile (and the refill of its registers) is issued after the DMMAs of the NEXT tile
#pragma unroll
            for (int i = 0; i < D; i++) {
                RCL_MMA(rt0 + i, buf[i]);
                if (i > 0) {
                    RCL_ST(rt0 + i - 1, buf[i - 1]);
                    RCL_LD(rt0 + i - 1 + D, buf[i - 1]);
                }
            }
            RCL_ST(rt0 + D - 1, buf[D - 1]);
            RCL_LD(rt0 + 2 * D - 1, buf[D - 1]);
        }
        const int rem = r_hi - rt0;
#pragma unroll
        for (int i = 0; i < D - 1; i++)
            if (i < rem) { RCL_MMA(rt0 + i, buf[i]); RCL_ST(rt0 + i, buf[i]); }
#undef RCL_MMA
#undef RCL_ST
#undef RCL_LD
    }
}

// ---- the last block step, every CTA of the cluster (worker wl = cluster rank of nwk): a contiguous block of the V column
//      tiles = a contiguous range of sites.  All 16 warps move in lock step over the row tiles: a warp updates its (at most
//      two) column tiles, deposits the 8 x 16 results in the stage; after a barrier the CTA writes its site range of the 8
//      finished columns of W (row rt*8 + r of B is column sStep[..] of W; thread = site; unit rows of the occupied sites
//      filled in on the way). ----
template <int NB, int D, bool FIRST, int T>
__device__ __noinline__ void rcl_last_step(RclShared *H, int kw, int nwk) {
    const RclArrays<NB> L(H);
    constexpr int KS = NB / 4, CT = 2;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int gr = lane >> 2, tg = lane & 3;
    const int Np = H->Np, Cp = H->Cp, N = H->N, M = H->M, ns = H->ns, wl = H->rank;
    const int nrt = Np >> 3;
    const int tv = Np >> 3, nvt = (Cp >> 3) - tv;                  // V part: nvt column tiles from tile tv
    const int vt0 = tv + wl * nvt / nwk, vt1 = tv + (wl + 1) * nvt / nwk, nv = vt1 - vt0;
    const int u_lo = (vt0 - tv) << 3, u_next = (vt1 - tv) << 3;
    const int site_lo = wl == 0 ? 0 : (u_lo < M ? L.sColSite[Np + u_lo] : ns);
    const int site_hi = (wl == nwk - 1 || u_next >= M) ? ns : L.sColSite[Np + u_next];
    const int SWc = nv << 3;                                       // row stride of the stage
    // row -> step map (written by P's leaders) and the raw pivot rows of my V columns
    for (int i = tid; i < Np; i += T) L.sStep[i] = __ldcg(H->Sg + i);
    for (int idx = tid; idx < SWc * KS; idx += T) {
        const int qg = idx / SWc, jl = idx - qg * SWc;
        rcl_gather_one<NB, FIRST>(H, L, (vt0 << 3) + jl, jl, qg << 2, kw);
    }
    __syncthreads();
    const double *UT = H->UT;
    const double *ws = H->ws;
    double *W = H->W;
    const double *sM = L.sM;
    const int *sStep = L.sStep, *sSiteInfo = L.sSiteInfo;
    double *stage = L.stage;
    double xf[CT][KS];
    bool cv[CT];
    const double *u0p[CT], *u1p[CT];
    const int lt0 = warp * CT;
#pragma unroll
    for (int c = 0; c < CT; c++) {
        cv[c] = lt0 + c < nv;
#pragma unroll
        for (int s = 0; s < KS; s++) xf[c][s] = cv[c] ? L.sX[((((lt0 + c) * KS) + s) << 5) + lane] : 0.0;
        if (FIRST) {
            const int col = ((cv[c] ? vt0 + lt0 + c : tv) << 3) + 2 * tg;
            u0p[c] = UT + (size_t)L.sColSite[col] * Np + gr;
            u1p[c] = UT + (size_t)L.sColSite[col + 1] * Np + gr;
        }
    }
    double *dep = stage + gr * SWc + (lt0 << 3) + 2 * tg;          // my two accumulator columns
    const double *rp = ws + (size_t)gr * Cp + ((vt0 + lt0) << 3) + 2 * tg;
    double2 buf[D][CT];
#define RCL_LD(rt_, d_)                                                                                        \
    do {                                                                                                       \
        const int rl_ = min((rt_), nrt - 1);                                                                   \
        _Pragma("unroll") for (int c = 0; c < CT; c++) {                                                       \
            if (!cv[c]) (d_)[c] = make_double2(0.0, 0.0);                                                      \
            else if (FIRST) (d_)[c] = make_double2(__ldcg(u0p[c] + (rl_ << 3)), __ldcg(u1p[c] + (rl_ << 3)));   \
            else (d_)[c] = __ldcg(reinterpret_cast<const double2 *>(rp + (size_t)(rl_ << 3) * Cp + c * 8));    \
        }                                                                                                      \
    } while (0)
#define RCL_TILE(rt_, d_)                                                                                      \
    do {                                                                                                       \
        double mf[KS];                                                                                         \
        _Pragma("unroll") for (int s = 0; s < KS; s++) mf[s] = sM[((((rt_) * KS) + s) << 5) + lane];           \
        _Pragma("unroll") for (int s = 0; s < KS; s++)                                                         \
            _Pragma("unroll") for (int c = 0; c < CT; c++) dmma_8x8x4((d_)[c].x, (d_)[c].y, mf[s], xf[c][s]);   \
        const int bo_ = ((rt_) & 1) ? 8 * SWc : 0;                                                             \
        _Pragma("unroll") for (int c = 0; c < CT; c++)                                                         \
            if (cv[c]) *reinterpret_cast<double2 *>(dep + bo_ + c * 8) = (d_)[c];                              \
        __syncthreads();                                                                                       \
        for (int site = site_lo + tid; site < site_hi; site += T) {                                            \
            const int info = sSiteInfo[site];                  /* >= 0: unoccupied, < 0: -label */             \
            _Pragma("unroll") for (int r = 0; r < 8; r++) {                                                    \
                const int t = sStep[((rt_) << 3) + r];         /* row (rt*8 + r) of B is column t of W */      \
                if (t < N)                                     /* (else: identity padding) */                  \
                    __stcs(W + (size_t)t * ns + site, info >= 0 ? stage[bo_ + r * SWc + info - u_lo] : ((-info - 1 == t) ? 1.0 : 0.0)); \
            }                                                                                                  \
        }                                                                                                      \
    } while (0)
#pragma unroll
    for (int i = 0; i < D; i++) RCL_LD(i, buf[i]);
    int rt0 = 0;
#pragma unroll 1
    for (; rt0 + D <= nrt; rt0 += D) {
#pragma unroll
        for (int i = 0; i < D; i++) {
            RCL_TILE(rt0 + i, buf[i]);
            RCL_LD(rt0 + i + D, buf[i]);
        }
    }
    const int rem = nrt - rt0;
#pragma unroll
    for (int i = 0; i < D - 1; i++)
        if (i < rem) RCL_TILE(rt0 + i, buf[i]);
#undef RCL_LD
#undef RCL_TILE
}

template <int NB, int DG, int T, int MINB>
__global__ void __launch_bounds__(T, MINB)
k_reeval_cl(DevState S, const int *__restrict__ list, double *__restrict__ scratch, const double *__restrict__ UT_up,
            const double *__restrict__ UT_dn, int *__restrict__ status, int Np_up, int Np_dn, int NpMax, int CpMax,
            int XT, int SW, int RS) {
    static_assert(NB % 8 == 0 && NB < KDSL_CL_PGS, "panel width");
    extern __shared__ double sm[];
    RclShared *H = reinterpret_cast<RclShared *>(sm);
    const int tid = threadIdx.x;
    const int rank = (int)cl_ctarank(), CL = (int)cl_nctarank();
    const bool isP = rank == 0;
    if (tid < 16) H->sKey[tid] = 0u;                        // (keys of warps that do not exist never win)
    if (tid == 0) {
        H->NpMax = NpMax; H->CpMax = CpMax; H->ns = S.ns; H->XT = XT; H->SW = SW; H->NG = CL - 1; H->rank = rank; H->RS = RS;
        double *base = scratch + (size_t)cl_clusterid() * reeval_cl_scratch_doubles(NB, NpMax, CpMax);
        H->ws = base;
        H->Mg = base + (size_t)NpMax * CpMax;
        H->Pg = reinterpret_cast<int *>(H->Mg + (size_t)2 * NB * NpMax);
        H->Sg = H->Pg + 2 * KDSL_CL_PGS;
    }
    __syncthreads();
    const int n_items = 2 * batch_count(S, list);
    for (int item = (int)cl_clusterid(); item < n_items; item += (int)cl_nclusterid()) {
        rcl_setup<NB, T>(H, S, list, UT_up, UT_dn, status, Np_up, Np_dn, item);
        const int Np = H->Np;
        bool pivoted = false;                               // P: my row has been a pivot
        long long t_phase = PHASE_CLOCK();
        if (isP) pivoted = rcl_factor_panel<NB, true>(H, 0, min(NB, Np), 0, pivoted);
        CL_TICK(0, 1);
        cl_sync();
        CL_TICK(0, 2);
        CL_TICK(1, 6);
        bool sing = false;
        for (int k0 = 0, s = 0; k0 < Np; k0 += NB, s++) {
            if (__ldcg(H->Pg + (s & 1) * KDSL_CL_PGS + NB) != 0) { sing = true; break; }   // uniform over the cluster
            const int kw = min(NB, Np - k0);                // multiple of 8
            const int k1 = k0 + kw, kn = min(NB, Np - k1);  // next panel (kn <= 0: none)
            const bool first = s == 0;
            if (kn > 0) {
                if (isP) {
                    if (first) rcl_next_panel<NB, true, T>(H, k1, kn, kw);
                    else rcl_next_panel<NB, false, T>(H, k1, kn, kw);
                    __syncthreads();
                    CL_TICK(0, 0);
                    pivoted = rcl_factor_panel<NB, false>(H, k1, kn, (s + 1) & 1, pivoted);
                    CL_TICK(0, 1);
                } else {
                    rcl_load_operands<NB, T>(H, s);
                    __syncthreads();
                    CL_TICK(1, 3);
                    if (first) rcl_update<NB, DG, true, T>(H, (k1 + kn) >> 3, kw);
                    else rcl_update<NB, DG, false, T>(H, (k1 + kn) >> 3, kw);
                    CL_TICK(1, 5);
                }
            } else {
                if (!isP) {
                    rcl_load_operands<NB, T>(H, s);
                    __syncthreads();
                    CL_TICK(1, 3);
                }
                if (first) rcl_last_step<NB, 4, true, T>(H, kw, CL);
                else rcl_last_step<NB, 4, false, T>(H, kw, CL);
                CL_TICK(0, 0);
                CL_TICK(1, 4);
            }
            cl_sync();
            CL_TICK(0, 2);
            CL_TICK(1, 6);
        }
#ifdef KDSL_PHASE_TICKS
        if (blockIdx.x == 0 && tid == 0) g_inv_phase_cycles[15] += 1;
#endif
        if (sing) {
            if (isP && tid == 0) status[H->status_idx] = 1;
            cl_sync();                                      // nobody re-reads the flag after P has moved on to the next item
        }
        __syncthreads();                                    // the tables are rebuilt by the next item
    }
}
