// kdsl_inverse_v4.cuh -- batched in-place inversion of tilde_U (reference reevaluateW!, src/MonteCarlo.jl:59-60:
// `tilde_U \ I`), fourth generation: blocked in-place Gauss-Jordan with IMPLICIT partial pivoting.
//
// One CTA of T threads per matrix and one CTA per SM (148 live matrices x 8 Np^2 bytes stay in the 126 MB L2).
// Per block step over NB columns:
//   1. the NB panel columns go to registers, one matrix row per (thread, r);
//   2. kw in-place rank-1 Gauss-Jordan steps on the panel.  The pivot of step k is the largest |.| of the current
//      column over the rows that were never a pivot (LAPACK's rule; an exactly zero / non-finite pivot => singular);
//      rows are NOT exchanged.  The register file rotates by one column per step so the loop body is identical
//      for every k (rolled loop, ~NB DFMAs); two block barriers per pivot (keys, then the winning row), see below;
//      the search compares the top 32 bits of |x| (ties within 2^-17 relative go to the lowest row);
//   3. the finished panel columns R (= the final values of these columns for this block step) are written back,
//      R - E (E = 1 at (pivot row of step q, column q)) goes to shared memory in DMMA fragment order, and the raw
//      pivot rows X = A[p_q, :] of all other columns are gathered;
//   4. every other column is updated on the FP64 tensor pipe:  A[:, J] += (R - E) X[:, J]   (all rows, uniform).
// With p_k the pivot row of elimination step k the stored result is S[p_k, c] = inv(A)[k, p_c]; the consumer
// (k_gemm_W_dmma) reads both index maps from colsrc[i] = "step at which row i was the pivot".
#pragma once
#include "kdsl_common.cuh"
#include "kdsl_refresh.cuh"
#include "kdsl_refresh_fast.cuh"

// T threads per CTA; the first TP of them own RPT matrix rows each during the panel factorisation (a pivot row is
// broadcast through shared memory: fewer, fatter threads read it less often).
template <int NB, int RPT, int T, int TP, int MINB = 1, int CT = 3>
__global__ void __launch_bounds__(T, MINB)
k_inverse_v4(DevState S, const int *__restrict__ list, double *__restrict__ A_base, int spin,
             int *__restrict__ status, int *__restrict__ colsrc_base, int Np, int cs_stride) {
    constexpr int NWARP = T / 32, PWARP = TP / 32;
    constexpr int KS = NB / 4;                          // DMMA k-steps per panel
    static_assert(NB % 8 == 0, "panel width must be a multiple of 8");
    extern __shared__ double sm[];
    const int b = blockIdx.x;
    if (b >= batch_count(S, list)) return;
    double *sM = sm;                                    // [Np x NB] frag-major (r = row, k = q): R - E
    double *sX = sM + (size_t)NB * Np;                  // [Np x NB] frag-major (r = column j, k = q): A[p_q, j]
    double *sRow = sX + (size_t)NB * Np;                // [NB] the pivot row of the current step
    double *sRinv = sRow + NB;                          // [2] reciprocal pivot
    unsigned *sKey = reinterpret_cast<unsigned *>(sRinv + 2);   // [8] per-warp candidate keys (16-byte aligned)
    int *sIdx = reinterpret_cast<int *>(sKey + 8);      // [4] pivot row
    int *sPivRow = sIdx + 4;                            // [NB] pivot row of each step of the current panel
    static_assert(PWARP <= 8 && TP <= T, "the warp index is folded into 3 key bits");

    double *A = A_base + (size_t)b * Np * Np;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int gr = lane >> 2, tg = lane & 3;
    const bool owner = tid < TP;                        // this thread owns rows tid + TP r
    unsigned pivmask = 0u;                              // bit r: my row (tid + TP r) has been a pivot
    int gstep[RPT];                                     // ... and at which elimination step
#pragma unroll
    for (int r = 0; r < RPT; r++) gstep[r] = 0;

    long long t_phase = PHASE_CLOCK();
    for (int k0 = 0; k0 < Np; k0 += NB) {
        const int kw = min(NB, Np - k0);                // multiple of 8
        // ---- 1. panel -> registers ----
        double a[RPT][NB];
        int mypiv[RPT];
#pragma unroll
        for (int r = 0; r < RPT; r++) {
            const int i = tid + TP * r;
            mypiv[r] = -1;
#pragma unroll
            for (int c = 0; c < NB; c++) a[r][c] = (owner && i < Np && c < kw) ? A[(size_t)(k0 + c) * Np + i] : 0.0;
        }
        PHASE_TICK(0);
        // ---- 2. kw in-place Gauss-Jordan steps; the current column is always a[.][0].
        // Two barriers per pivot: (1) every warp posts a 32-bit key of its best candidate, (2) the owner of the
        // winning row posts that row.  Only the NEXT column is updated right after the second barrier; the other
        // NB-2 columns of step k are updated in iteration k+1, where their DFMAs overlap the latency of the warp
        // reduction and of the reciprocal (software pipelining by hand; a warp issues in order). ----
        double tail[RPT];                               // pending: -l of my row (pivot row: 1/pivot)
        bool pend = false;
        unsigned pend_p = 0u;                           // bit r: my row r was the pivot of the pending step
        auto apply_pending = [&]() {                    // columns 2.. of the pending step (sRow still holds its pivot row)
            if (!owner) return;
            const double2 *prow2 = reinterpret_cast<const double2 *>(sRow);
#pragma unroll
            for (int j = 2; j < NB; j += 2) {
                const double2 pv = prow2[j >> 1];
#pragma unroll
                for (int r = 0; r < RPT; r++) {
                    if ((pend_p >> r) & 1u) {
                        a[r][j - 1] = a[r][j] * tail[r];
                        a[r][j] = a[r][j + 1 < NB ? j + 1 : j] * tail[r];
                    } else {
                        a[r][j - 1] = fma(tail[r], pv.x, a[r][j]);
                        a[r][j] = fma(tail[r], pv.y, a[r][j + 1 < NB ? j + 1 : j]);
                    }
                }
            }
#pragma unroll
            for (int r = 0; r < RPT; r++) a[r][NB - 1] = tail[r];
        };
#pragma unroll 1
        for (int k = 0; k < kw; k++) {
            // candidate of this thread: largest |a[r][0]| over its never-pivoted rows (top 32 bits of |x|)
            unsigned hi = 0u;
            int br = 0;
            bool valid = false;
            double bval = 1.0;
#pragma unroll
            for (int r = 0; r < RPT; r++) {
                const int i = tid + TP * r;
                if (owner && i < Np && !((pivmask >> r) & 1u)) {
                    const double v = a[r][0];
                    const unsigned h = (unsigned)__double2hiint(v) & 0x7fffffffu;
                    if (!valid || h > hi) { hi = h; br = r; bval = v; }
                    valid = true;
                }
            }
            const unsigned mhi = __reduce_max_sync(0xffffffffu, valid ? hi : 0u);
            double my_rinv = 1.0 / bval;                // speculative; LAPACK getf2 scales by the reciprocal pivot
            asm volatile("" : "+d"(my_rinv));           // keep it here (ptxas would sink it behind the first barrier)
            if (pend) apply_pending();
            const unsigned win = __ballot_sync(0xffffffffu, valid && hi == mhi);
            const bool leader = win != 0u && lane == __ffs(win) - 1;      // lowest row among the warp's maxima
            if (lane == 0 && warp < 8) sKey[warp] = (win != 0u) ? ((mhi & 0xfffffff8u) | (unsigned)(7 - warp)) : 0u;
            __syncthreads();
            unsigned bk = 0u;
            {
                const uint4 k0v = *reinterpret_cast<const uint4 *>(sKey);
                const uint4 k1v = *reinterpret_cast<const uint4 *>(sKey + 4);
                bk = max(max(max(k0v.x, k0v.y), max(k0v.z, k0v.w)), max(max(k1v.x, k1v.y), max(k1v.z, k1v.w)));
            }
            if ((bk >> 3) == 0u || bk >= 0x7ff00000u) {  // zero (below 2^-1039) / non-finite pivot: singular
                if (tid == 0) status[2 * b + spin] = 1;
                return;
            }
            const int wq = 7 - (int)(bk & 7u);           // ties within 2^-17 relative: the lowest warp
            if (warp == wq && leader) {
                const int i = tid + TP * br;
                sIdx[0] = i;
                sPivRow[k] = i;
                sRinv[0] = my_rinv;
                double2 *dst = reinterpret_cast<double2 *>(sRow);
#pragma unroll
                for (int j = 0; j < NB; j += 2) {
                    double x0 = a[0][j], x1 = a[0][j + 1];
#pragma unroll
                    for (int r = 1; r < RPT; r++) if (br == r) { x0 = a[r][j]; x1 = a[r][j + 1]; }
                    dst[j >> 1] = make_double2(x0, x1);
                }
            }
            __syncthreads();
            const int p = sIdx[0];
            const double rinv = sRinv[0];
            const double prow1 = sRow[1];
            pend_p = 0u;
#pragma unroll
            for (int r = 0; r < RPT; r++) {
                const int i = tid + TP * r;
                if (owner && i == p) {
                    pivmask |= 1u << r;
                    pend_p |= 1u << r;
                    gstep[r] = k0 + k;
                    mypiv[r] = k;
                    tail[r] = rinv;
                    a[r][0] = a[r][1] * rinv;
                } else {
                    tail[r] = -(a[r][0] * rinv);
                    a[r][0] = fma(tail[r], prow1, a[r][1]);
                }
            }
            pend = true;
        }
        if (pend) apply_pending();
        PHASE_TICK(1);
        // ---- 3. write the finished panel columns; R - E to shared memory (fragment order) ----
        // after kw rotations register slot cs holds panel column (cs + kw) mod NB (columns >= kw are zero padding)
#pragma unroll
        for (int r = 0; r < RPT; r++) {
            const int i = tid + TP * r;
            if (owner && i < Np) {
#pragma unroll
                for (int cs = 0; cs < NB; cs += 4) {
                    int col = cs + kw;
                    if (col >= NB) col -= NB;           // kw and NB are multiples of 8: groups of 4 stay together
                    double v[4];
#pragma unroll
                    for (int e = 0; e < 4; e++) v[e] = a[r][cs + e];
                    if (col < kw) {
#pragma unroll
                        for (int e = 0; e < 4; e++) A[(size_t)(k0 + col + e) * Np + i] = v[e];
                    }
#pragma unroll
                    for (int e = 0; e < 4; e++) if (col + e == mypiv[r]) v[e] -= 1.0;
                    double2 *dst = reinterpret_cast<double2 *>(sM + frag_idx(i, col, NB));
                    dst[0] = make_double2(v[0], v[1]);
                    dst[1] = make_double2(v[2], v[3]);
                }
            }
        }
        __syncthreads();                                // sPivRow complete
        // ---- 4. raw pivot rows of the other columns ----
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
        __syncthreads();
        PHASE_TICK(2);
        // ---- 5. A[:, J] += (R - E) X[:, J] on the FP64 tensor pipe, transposed form (D[j][i], 128-bit access).
        //         Work item = CT column tiles x all rows; the X fragments stay in registers ----
        {
            const int nrt = Np >> 3, nct = (Np - kw) >> 3, ktiles = kw >> 3, k0t = k0 >> 3;
            const int groups = (nct + CT - 1) / CT;
            for (int g = warp; g < groups; g += NWARP) {
                double xf[CT][KS];
                double2 *cp[CT];
                bool cv[CT];
#pragma unroll
                for (int c = 0; c < CT; c++) {
                    const int t = g * CT + c;
                    cv[c] = t < nct;
                    const int ct = cv[c] ? (t < k0t ? t : t + ktiles) : 0;
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
        }
        __syncthreads();
        PHASE_TICK(3);
    }
    // ---- 6. index map for the consumer: colsrc[i] = elimination step at which row i was the pivot ----
    int *colsrc = colsrc_base + ((size_t)2 * b + spin) * cs_stride;
#pragma unroll
    for (int r = 0; r < RPT; r++) {
        const int i = tid + TP * r;
        if (owner && i < Np) colsrc[i] = gstep[r];
    }
}
