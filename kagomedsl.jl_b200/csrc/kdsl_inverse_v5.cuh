// kdsl_inverse_v5.cuh -- k_inverse_v4 with look-ahead: the latency-bound pivot loop of panel s+1 runs concurrently
// with the tensor-pipe update of step s (warp specialisation inside one CTA).
//
// 512 threads = 16 warps, one CTA per matrix and per SM (matrices up to 256 x 256; larger ones use k_inverse_v4).
//   team P (warps 0-7, one matrix row per thread): panel load, pivot loop (named barrier 1), publish;
//   team G (warps 8-15): the DMMA update of every column outside panels s and s+1.
// Per block step s (operands R_s - E in sM[s & 1] and the raw pivot rows X_s in sX):
//   A. all 16 warps update the columns of panel s+1 with the operands of step s          -> __syncthreads
//   B. team P factors panel s+1 (writes its final columns and sM[(s + 1) & 1]) WHILE team G updates the other columns
//                                                                                           -> __syncthreads
//   C. all threads gather X_{s+1} = A[p_q, :] (every column is now up to date)             -> __syncthreads
// The mathematics, the pivot rule and the stored layout (implicit row pivoting, colsrc) are those of k_inverse_v4.
// Measured (432 sites, tools/inv_phases.py): team P alone needs 214 k cycles per matrix, 330 k while team G runs --
// DMMA and DFMA share the FP64 pipe, so every FP64 instruction on the pivot chain queues behind 16-cycle DMMAs; the
// overlap is worth 5-9 % over k_inverse_v4, not the 40 % a contention-free overlap would give (a two-level panel with
// 8 instead of NB DFMAs per pivot and a 4-warp team G were tried and did not change that).
#pragma once
#include "kdsl_common.cuh"
#include "kdsl_refresh.cuh"
#include "kdsl_refresh_fast.cuh"

__device__ __forceinline__ void bar_team_p() { asm volatile("bar.sync 1, 256;" ::: "memory"); }
template <int TP> __device__ __forceinline__ void bar_team() { asm volatile("bar.sync 1, %0;" ::"n"(TP) : "memory"); }

// 1 / a to about 1 ulp with THREE dependent FP64 instructions after the MUFU seed x0 (relative error e ~ 2^-20 from
// rcp.approx.ftz.f64):  1/a = x0 (1 + e + e^2 + ...),  x = x0 + x0 (e + e^2), error e^3 ~ 2^-60.  The compiler's IEEE
// division is a longer dependent chain plus a slow-path call; on the pivot chain every FP64 instruction queues behind
// the other team's DMMAs, so the length matters.  (Pivots below 2^-1039 never get here: they flag the matrix singular.)
__device__ __forceinline__ double rcp_fast(double a) {
    double x;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(a));
    const double e = fma(-a, x, 1.0);
    const double t = fma(e, e, e);
    return fma(x, t, x);
}

template <int NB, int CT, int GW = 8>
__global__ void __launch_bounds__(256 + 32 * GW, 1)
k_inverse_v5(DevState S, const int *__restrict__ list, double *__restrict__ A_base, int spin,
             int *__restrict__ status, int *__restrict__ colsrc_base, int Np, int cs_stride) {
    constexpr int T = 256 + 32 * GW, TP = 256, NWARPS = 8 + GW;   // threads, team-P threads (team G: GW warps)
    constexpr int KS = NB / 4;
    static_assert(NB % 8 == 0, "panel width must be a multiple of 8");
    extern __shared__ double sm[];
    const int b = blockIdx.x;
    if (b >= batch_count(S, list)) return;
    double *sMb = sm;                                   // [2][Np x NB] frag-major (r = row, k = q): R - E
    double *sX = sMb + (size_t)2 * NB * Np;             // [Np x NB] frag-major (r = column j, k = q): A[p_q, j]
    double *sRow = sX + (size_t)NB * Np;                // [NB] the pivot row of the current step
    double *sRinv = sRow + NB;                          // [2]
    unsigned *sKey = reinterpret_cast<unsigned *>(sRinv + 2);   // [8] per-warp candidate keys
    int *sIdx = reinterpret_cast<int *>(sKey + 8);      // [4]: [0] pivot row, [1] singular flag
    int *sPivRow = sIdx + 4;                            // [NB] pivot row of each step of the panel being factored

    double *A = A_base + (size_t)b * Np * Np;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int gr = lane >> 2, tg = lane & 3;
    const bool teamP = tid < TP;
    const bool has_row = teamP && tid < Np;             // this thread owns matrix row `tid`
    bool pivoted = false;                               // my row has been a pivot
    int gstep = 0;                                      // ... at this elimination step
    const int nrt = Np >> 3;
    if (tid == 0) sIdx[1] = 0;
    __syncthreads();

    // ---- team P: factor the panel starting at column k0 (kw wide); operands for the update go to sM ----
    auto factor_panel = [&](int k0, int kw, double *sM) {
        double a[NB];
        int mypiv = -1;
#pragma unroll
        for (int c = 0; c < NB; c++) a[c] = (has_row && c < kw) ? A[(size_t)(k0 + c) * Np + tid] : 0.0;
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
            double my_rinv = rcp_fast(valid ? a[0] : 1.0);      // speculative; LAPACK getf2 scales by the reciprocal pivot
            asm volatile("" : "+d"(my_rinv));
            if (pend) apply_pending();
            double my_s = my_rinv * a[1];                        // ... and the scaled next-column entry of my row (now up to date)
            asm volatile("" : "+d"(my_s));
            const unsigned win = __ballot_sync(0xffffffffu, valid && hi == mhi);
            const bool leader = win != 0u && lane == __ffs(win) - 1;
            if (lane == 0) sKey[warp] = (win != 0u) ? ((mhi & 0xfffffff8u) | (unsigned)(7 - warp)) : 0u;
            bar_team_p();
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
                sRinv[0] = my_rinv;
                sRinv[1] = my_s;
                double2 *dst = reinterpret_cast<double2 *>(sRow);
#pragma unroll
                for (int j = 0; j < NB; j += 2) dst[j >> 1] = make_double2(a[j], a[j + 1]);
            }
            bar_team_p();
            const int p = sIdx[0];
            const double2 rs = *reinterpret_cast<const double2 *>(sRinv);   // (1 / pivot, pivot row's next entry / pivot)
            if (has_row && tid == p) {
                pivoted = true;
                pend_p = true;
                gstep = k0 + k;
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
            if (tid == 0) { sIdx[1] = 1; status[2 * b + spin] = 1; }
            return;
        }
        if (pend) apply_pending();
        // publish: final panel columns to global memory, R - E to shared memory (fragment order).  After kw rotations
        // register slot cs holds panel column (cs + kw) mod NB (columns >= kw are zero padding).
        if (has_row) {
#pragma unroll
            for (int cs = 0; cs < NB; cs += 4) {
                int col = cs + kw;
                if (col >= NB) col -= NB;
                double v[4];
#pragma unroll
                for (int e = 0; e < 4; e++) v[e] = a[cs + e];
                if (col < kw) {
#pragma unroll
                    for (int e = 0; e < 4; e++) A[(size_t)(k0 + col + e) * Np + tid] = v[e];
                }
#pragma unroll
                for (int e = 0; e < 4; e++) if (col + e == mypiv) v[e] -= 1.0;
                double2 *dst = reinterpret_cast<double2 *>(sM + frag_idx(tid, col, NB));
                dst[0] = make_double2(v[0], v[1]);
                dst[1] = make_double2(v[2], v[3]);
            }
        }
    };
    // ---- all threads: raw pivot rows of every column outside the panel [k0, k0 + kw) ----
    auto gather_X = [&](int k0, int kw) {
        for (int j = tid; j < Np; j += T) {
            if (!(j >= k0 && j < k0 + kw)) {
                const double *col = A + (size_t)j * Np;
#pragma unroll
                for (int q = 0; q < NB; q += 4) {
                    double v[4];
#pragma unroll
                    for (int e = 0; e < 4; e++) v[e] = (q + e < kw) ? col[sPivRow[q + e]] : 0.0;
                    double2 *dst = reinterpret_cast<double2 *>(sX + frag_idx(j, q, NB));
                    dst[0] = make_double2(v[0], v[1]);
                    dst[1] = make_double2(v[2], v[3]);
                }
            }
        }
    };
    // ---- DMMA update A[:, J] += (R - E) X[:, J] of the column tiles [0, nct) mapped around the excluded tile range
    //      [ex0, ex0 + exn); groups of CT tiles are dealt round-robin to `nw_team` warps (this warp is `w_team`) ----
    auto update_cols = [&](const double *sM, int nct, int ex0, int exn, int w_team, int nw_team) {
        const int groups = (nct + CT - 1) / CT;
        for (int g = w_team; g < groups; g += nw_team) {
            double xf[CT][KS];
            double2 *cp[CT];
            bool cv[CT];
#pragma unroll
            for (int c = 0; c < CT; c++) {
                const int t = g * CT + c;
                cv[c] = t < nct;
                const int ct = cv[c] ? (t < ex0 ? t : t + exn) : 0;
#pragma unroll
                for (int s = 0; s < KS; s++) xf[c][s] = cv[c] ? sX[(((ct * KS) + s) << 5) + lane] : 0.0;
                cp[c] = reinterpret_cast<double2 *>(A + (size_t)((ct << 3) + gr) * Np + 2 * tg);
            }
            double2 cur[2][CT], nxt[2][CT];
            auto load_pair = [&](int rt, double2 (&d)[2][CT]) {
#pragma unroll
                for (int h = 0; h < 2; h++)
#pragma unroll
                    for (int c = 0; c < CT; c++)
                        d[h][c] = (cv[c] && rt + h < nrt) ? cp[c][(rt + h) << 2] : make_double2(0.0, 0.0);
            };
            load_pair(0, cur);
            for (int rt = 0; rt < nrt; rt += 2) {
                if (rt + 2 < nrt) load_pair(rt + 2, nxt);
                double mf[2][KS];
#pragma unroll
                for (int h = 0; h < 2; h++)
#pragma unroll
                    for (int s = 0; s < KS; s++)
                        mf[h][s] = (rt + h < nrt) ? sM[((((rt + h) * KS) + s) << 5) + lane] : 0.0;
#pragma unroll
                for (int s = 0; s < KS; s++)
#pragma unroll
                    for (int h = 0; h < 2; h++)
#pragma unroll
                        for (int c = 0; c < CT; c++) dmma_8x8x4(cur[h][c].x, cur[h][c].y, xf[c][s], mf[h][s]);
#pragma unroll
                for (int h = 0; h < 2; h++)
#pragma unroll
                    for (int c = 0; c < CT; c++)
                        if (cv[c] && rt + h < nrt) cp[c][(rt + h) << 2] = cur[h][c];
#pragma unroll
                for (int h = 0; h < 2; h++)
#pragma unroll
                    for (int c = 0; c < CT; c++) cur[h][c] = nxt[h][c];
            }
        }
    };
    // ---- phase A: the kn column tiles of the next panel (first tile nt0), all 16 warps: item = (tile, row chunk) ----
    auto update_next_panel = [&](const double *sM, int nt0, int kn) {
        const int nchunks = NWARPS / kn;
        const int c = warp % kn, chunk = warp / kn;
        if (chunk >= nchunks) return;
        const int ct = nt0 + c;
        double xf[KS];
#pragma unroll
        for (int s = 0; s < KS; s++) xf[s] = sX[(((ct * KS) + s) << 5) + lane];
        double2 *cp = reinterpret_cast<double2 *>(A + (size_t)((ct << 3) + gr) * Np + 2 * tg);
        const int r_lo = chunk * nrt / nchunks, r_hi = (chunk + 1) * nrt / nchunks;
        for (int rt = r_lo; rt < r_hi; rt++) {
            double2 d = cp[rt << 2];
#pragma unroll
            for (int s = 0; s < KS; s++) dmma_8x8x4(d.x, d.y, xf[s], sM[(((rt * KS) + s) << 5) + lane]);
            cp[rt << 2] = d;
        }
    };

    // ---- step 0 panel ----
    if (teamP) factor_panel(0, min(NB, Np), sMb);
    __syncthreads();
    if (sIdx[1]) return;                                // singular
    gather_X(0, min(NB, Np));
    __syncthreads();
    long long t_phase = PHASE_CLOCK();
#define V5_TICK(idx, thr) PHASE_TICK_AT(idx, thr)
    for (int k0 = 0, s = 0; k0 < Np; k0 += NB, s++) {
        const int kw = min(NB, Np - k0);                // multiple of 8
        const int k1 = k0 + kw, kn = min(NB, Np - k1);  // next panel (kn <= 0: none)
        const double *sM = sMb + (size_t)(s & 1) * NB * Np;
        double *sMn = sMb + (size_t)((s + 1) & 1) * NB * Np;
        const int exn = (kw + max(kn, 0)) >> 3;
        PHASE_RESTART(256);
        V5_TICK(4, 0);                                  // (loop overhead)
        if (kn > 0) {
            update_next_panel(sM, k1 >> 3, kn >> 3);
            __syncthreads();
        }
        V5_TICK(0, 0);
        PHASE_RESTART(256);
        if (teamP) {
            if (kn > 0) factor_panel(k1, kn, sMn);
            V5_TICK(1, 0);
        } else {
            update_cols(sM, (Np >> 3) - exn, k0 >> 3, exn, warp - 8, GW);
            V5_TICK(2, 256);
        }
        __syncthreads();
        V5_TICK(5, 0);                                  // team P waiting for team G
        if (sIdx[1]) return;
        if (kn > 0) {
            gather_X(k1, kn);
            __syncthreads();
        }
        V5_TICK(3, 0);
    }
    // ---- index map for the consumer: colsrc[i] = elimination step at which row i was the pivot ----
    int *colsrc = colsrc_base + ((size_t)2 * b + spin) * cs_stride;
    if (has_row) colsrc[tid] = gstep;
}
