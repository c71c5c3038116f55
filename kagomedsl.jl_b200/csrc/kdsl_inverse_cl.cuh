// kdsl_inverse_cl.cuh -- batched in-place inversion of tilde_U for 256 < Np <= 512 (972 sites: N = 486; the real
// 2N x 2N embedding of the ComplexF64 engine at 432 sites: Np = 432) by THREAD-BLOCK CLUSTERS: one matrix per cluster.
// Reference work: reevaluateW!, src/MonteCarlo.jl:55-66 (`tilde_U \ I`).
//
// Why clusters.  One CTA per matrix and per SM (k_inverse_v4) keeps 148 matrices of 1.9 MB live: 280 MB against a 126 MB L2.
// ncu (profiles/r2_inverse972_ncu_summary.txt): 41 MB of DRAM traffic per matrix instead of 3.8, 55 % of the warp samples wait
// for those loads, tensor pipe 25 % busy.  With CL CTAs per matrix only 148 / CL matrices are live (70 MB at CL = 4) and every
// block step after the first touch is served by the L2.  The second gain is the split of the two kinds of work over
// DIFFERENT SMs: the pivot chain is a latency-bound sequence of FP64 operations that queue behind 16-cycle DMMAs whenever it
// shares an SM with the trailing update (k_inverse_v5: 214 k cycles alone, 330 k beside team G) -- here it has an SM to itself.
//
//   CTA 0 of the cluster ("P"): thread = matrix row.  Factors panel s+1 (implicit-pivot Gauss-Jordan on NB columns held in
//       registers, the pivot rule and arithmetic of k_inverse_v4 / v5) while the other CTAs apply block step s; before that
//       it brings the columns of panel s+1 up to date with the operands of step s (DMMA, all 16 warps).
//   CTAs 1 .. CL-1 ("G"): the trailing update  A[:, J] += (R_s - E) A_old[(p_q), J]  on the FP64 tensor pipe for the column
//       tiles they own (groups of CT tiles dealt round-robin to the G CTAs; inside a CTA the items (group, row slice) are
//       dealt to the 16 warps).  A CTA gathers the raw pivot rows of ITS columns before it updates them, so no CTA ever
//       reads a column another CTA writes in the same step.
//   Exchange per block step: the operands R_s - E (Np x NB doubles, DMMA fragment order), the NB pivot rows and a
//       singular flag go through a double-buffered per-cluster scratch in global memory (L2); ONE barrier.cluster
//       (arrive.release / wait.acquire) per block step orders everything.  Matrix and scratch traffic bypasses L1 (.cg).
// The stored layout is k_inverse_v4's: with p_k the pivot row of elimination step k, S[p_k, c] = inv(A)[k, p_c];
// colsrc[i] = step at which row i was the pivot (read by k_gemm_W_dmma).
#pragma once
#include "kdsl_common.cuh"
#include "kdsl_refresh.cuh"
#include "kdsl_refresh_fast.cuh"
#include "kdsl_inverse_v5.cuh"

__device__ __forceinline__ unsigned cl_ctarank() {
    unsigned r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ unsigned cl_nctarank() {
    unsigned r;
    asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ unsigned cl_clusterid() {
    unsigned r;
    asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r));
    return r;
}
__device__ __forceinline__ unsigned cl_nclusterid() {
    unsigned r;
    asm volatile("mov.u32 %0, %%nclusterid.x;" : "=r"(r));
    return r;
}
// every thread of every CTA of the cluster; orders global and shared::cluster traffic at cluster scope
__device__ __forceinline__ void cl_sync() {
    asm volatile("barrier.cluster.arrive.release;\n\tbarrier.cluster.wait.acquire;" ::: "memory");
}

#ifdef KDSL_PHASE_TICKS
// developer instrumentation: cycles of cluster 0's P CTA (slots 0-2: next panel, factor, barrier wait) and of its first
// G CTA (slots 3-6: operand load, gather, update (warp 0), barrier wait); slot 7 counts the items
#define CL_TICK(cta, idx)                                                \
    do {                                                                 \
        if (blockIdx.x == (cta) && threadIdx.x == 0) {                   \
            const long long now_ = clock64();                            \
            g_inv_phase_cycles[8 + (idx)] += now_ - t_phase;                   \
            t_phase = now_;                                              \
        }                                                                \
    } while (0)
#else
#define CL_TICK(cta, idx) do { (void)t_phase; } while (0)
#endif

#define KDSL_CL_PGS 32          /* ints per pivot-row buffer: [0, NB) pivot rows, [NB] singular flag */

// doubles of per-cluster scratch: two operand buffers + two pivot-row buffers
__host__ __device__ inline size_t inverse_cl_scratch_doubles(int NB, int NpMax) {
    return (size_t)2 * NB * NpMax + (2 * KDSL_CL_PGS * sizeof(int) + 7) / 8;
}
inline size_t inverse_cl_smem(int NB, int NpMax) {
    return ((size_t)2 * NB * NpMax + NB + 2) * sizeof(double) + ((size_t)16 + 4 + NB) * sizeof(int);
}

template <int NB, int CT>
__global__ void __launch_bounds__(512, 1)
k_inverse_cl(DevState S, const int *__restrict__ list, double *__restrict__ A_up, double *__restrict__ A_dn,
             int *__restrict__ status, int *__restrict__ colsrc_base, int Np_up, int Np_dn, int cs_stride,
             double *__restrict__ scratch, int RS) {
    constexpr int T = 512, NWARPS = 16, KS = NB / 4;
    static_assert(NB % 8 == 0 && NB < KDSL_CL_PGS, "panel width");
    extern __shared__ double sm[];
    const int NpMax = max(Np_up, Np_dn);
    double *sM = sm;                                    // [Np x NB] frag-major (r = row, k = q): R - E of the current step
    double *sX = sM + (size_t)NB * NpMax;               // [Np x NB] frag-major (r = column j, k = q): A[p_q, j]
    double *sRow = sX + (size_t)NB * NpMax;             // [NB] the pivot row of the current step
    double *sRinv = sRow + NB;                          // [2]
    unsigned *sKey = reinterpret_cast<unsigned *>(sRinv + 2);   // [16] per-warp candidate keys
    int *sIdx = reinterpret_cast<int *>(sKey + 16);     // [4]: [0] pivot row
    int *sPivRow = sIdx + 4;                            // [NB] pivot rows of the panel (P: being factored; G: of this step)

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int gr = lane >> 2, tg = lane & 3;
    const int rank = (int)cl_ctarank(), NG = (int)cl_nctarank() - 1, grank = rank - 1;
    const bool isP = rank == 0;
    double *Mg = scratch + (size_t)cl_clusterid() * inverse_cl_scratch_doubles(NB, NpMax);
    int *Pg = reinterpret_cast<int *>(Mg + (size_t)2 * NB * NpMax);
    const int count = batch_count(S, list);

    for (int item = (int)cl_clusterid(); item < 2 * count; item += (int)cl_nclusterid()) {
        const int b = item >> 1, spin = item & 1;
        const int Np = spin ? Np_dn : Np_up;
        double *A = (spin ? A_dn : A_up) + (size_t)b * Np * Np;
        const int nrt = Np >> 3;
        const bool has_row = isP && tid < Np;           // P: this thread owns matrix row `tid`
        bool pivoted = false;                           // my row has been a pivot
        int gstep = 0;                                  // ... at this elimination step

        // ---- P: factor the panel [k0, k0 + kw); operands to sM and to the scratch buffer `buf` ----
        auto factor_panel = [&](int k0, int kw, int buf) {
            double *Mgb = Mg + (size_t)buf * NB * NpMax;
            int *Pgb = Pg + buf * KDSL_CL_PGS;
            double a[NB];
            int mypiv = -1;
#pragma unroll
            for (int c = 0; c < NB; c++) a[c] = (has_row && c < kw) ? __ldcg(A + (size_t)(k0 + c) * Np + tid) : 0.0;
            double tail = 0.0;
            bool pend = false, pend_p = false;
            auto apply_pending = [&]() {                // columns 2.. of the pending step (sRow still holds its pivot row)
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
                    gstep = k0 + k;
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
                if (tid == 0) {
                    status[2 * b + spin] = 1;
                    __stcg(Pgb + NB, 1);
                }
                return;
            }
            if (pend) apply_pending();
            // publish: final panel columns to the matrix, R - E to shared memory AND to the scratch (fragment order).
            // After kw rotations register slot cs holds panel column (cs + kw) mod NB (columns >= kw are zero padding).
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
                        for (int e = 0; e < 4; e++) __stcg(A + (size_t)(k0 + col + e) * Np + tid, v[e]);
                    }
#pragma unroll
                    for (int e = 0; e < 4; e++) if (col + e == mypiv) v[e] -= 1.0;
                    const int fi = frag_idx(tid, col, NB);
                    double2 *dst = reinterpret_cast<double2 *>(sM + fi);
                    dst[0] = make_double2(v[0], v[1]);
                    dst[1] = make_double2(v[2], v[3]);
                    double2 *dg = reinterpret_cast<double2 *>(Mgb + fi);
                    __stcg(dg, make_double2(v[0], v[1]));
                    __stcg(dg + 1, make_double2(v[2], v[3]));
                }
            }
            if (tid < kw) __stcg(Pgb + tid, sPivRow[tid]);       // (written before the loop's last barrier)
            if (tid == NB) __stcg(Pgb + NB, 0);
        };
        // ---- raw pivot rows X[q, j] = A[p_q, j] of the columns j = jb + tid, jb + tid + T, ... < je that `pred` selects ----
        auto gather_cols = [&](int kw, auto pred) {
            for (int j = tid; j < Np; j += T) {
                if (pred(j)) {
                    const double *col = A + (size_t)j * Np;
#pragma unroll
                    for (int q = 0; q < NB; q += 4) {
                        double v[4];
#pragma unroll
                        for (int e = 0; e < 4; e++) v[e] = (q + e < kw) ? __ldcg(col + sPivRow[q + e]) : 0.0;
                        double2 *dst = reinterpret_cast<double2 *>(sX + frag_idx(j, q, NB));
                        dst[0] = make_double2(v[0], v[1]);
                        dst[1] = make_double2(v[2], v[3]);
                    }
                }
            }
        };

        // ---- panel 0 ----
        long long t_phase = PHASE_CLOCK();
        if (isP) factor_panel(0, min(NB, Np), 0);
        CL_TICK(0, 1);
        cl_sync();
        CL_TICK(0, 2);
        CL_TICK(1, 6);
        bool sing = false;
        for (int k0 = 0, s = 0; k0 < Np; k0 += NB, s++) {
            const int *Pgs = Pg + (s & 1) * KDSL_CL_PGS;
            if (__ldcg(Pgs + NB) != 0) { sing = true; break; }   // uniform over the cluster
            const int kw = min(NB, Np - k0);            // multiple of 8
            const int k1 = k0 + kw, kn = min(NB, Np - k1);   // next panel (kn <= 0: none)
            const int ex0 = k0 >> 3, exn = (kw + max(kn, 0)) >> 3;
            if (isP) {
                if (kn > 0) {
                    // columns of panel s+1: A[:, J] += (R_s - E) X_s[:, J]  (sM and sPivRow still hold step s)
                    gather_cols(kw, [&](int j) { return j >= k1 && j < k1 + kn; });
                    __syncthreads();
                    {
                        const int ktn = kn >> 3, nchunks = NWARPS / ktn;
                        const int c = warp % ktn, chunk = warp / ktn;
                        if (chunk < nchunks) {
                            const int ct = (k1 >> 3) + c;
                            double xf[KS];
#pragma unroll
                            for (int q = 0; q < KS; q++) xf[q] = sX[(((ct * KS) + q) << 5) + lane];
                            double2 *cp = reinterpret_cast<double2 *>(A + (size_t)((ct << 3) + gr) * Np + 2 * tg);
                            const int r_lo = chunk * nrt / nchunks, r_hi = (chunk + 1) * nrt / nchunks;
                            for (int rt = r_lo; rt < r_hi; rt++) {
                                double2 d = __ldcg(cp + (rt << 2));
#pragma unroll
                                for (int q = 0; q < KS; q++) dmma_8x8x4(d.x, d.y, xf[q], sM[(((rt * KS) + q) << 5) + lane]);
                                __stcg(cp + (rt << 2), d);
                            }
                        }
                    }
                    __syncthreads();
                    CL_TICK(0, 0);
                    factor_panel(k1, kn, (s + 1) & 1);
                    CL_TICK(0, 1);
                }
            } else {
                // operands and pivot rows of step s: scratch -> shared memory
                {
                    const double2 *src = reinterpret_cast<const double2 *>(Mg + (size_t)(s & 1) * NB * NpMax);
                    double2 *dst = reinterpret_cast<double2 *>(sM);
                    const int n2 = (NB * Np) >> 1;
                    for (int i = tid; i < n2; i += T) dst[i] = __ldcg(src + i);
                    if (tid < NB) sPivRow[tid] = __ldcg(Pgs + tid);
                }
                __syncthreads();
                CL_TICK(1, 3);
                // my column-tile groups: g = grank, grank + NG, ... over the tiles outside [ex0, ex0 + exn)
                const int nct = nrt - exn;
                const int groups = (nct + CT - 1) / CT;
                gather_cols(kw, [&](int j) {
                    const int ct = j >> 3;
                    if (ct >= ex0 && ct < ex0 + exn) return false;
                    const int t = ct < ex0 ? ct : ct - exn;
                    return (t / CT) % NG == grank;
                });
                __syncthreads();
                CL_TICK(1, 4);
                const int n_my = groups > grank ? (groups - grank + NG - 1) / NG : 0;
                const int items = n_my * RS;
                for (int it = warp; it < items; it += NWARPS) {
                    const int lg = it / RS, rs = it - lg * RS;
                    const int g = grank + lg * NG;
                    const int r_lo = rs * nrt / RS, r_hi = (rs + 1) * nrt / RS;
                    double xf[CT][KS];
                    double2 *cp[CT];
                    bool cv[CT];
#pragma unroll
                    for (int c = 0; c < CT; c++) {
                        const int t = g * CT + c;
                        cv[c] = t < nct;
                        const int ct = cv[c] ? (t < ex0 ? t : t + exn) : 0;
#pragma unroll
                        for (int q = 0; q < KS; q++) xf[c][q] = cv[c] ? sX[(((ct * KS) + q) << 5) + lane] : 0.0;
                        cp[c] = reinterpret_cast<double2 *>(A + (size_t)((ct << 3) + gr) * Np + 2 * tg);
                    }
                    double2 cur[2][CT], nxt[2][CT];
                    auto load_pair = [&](int rt, double2 (&d)[2][CT]) {
#pragma unroll
                        for (int h = 0; h < 2; h++)
#pragma unroll
                            for (int c = 0; c < CT; c++)
                                d[h][c] = (cv[c] && rt + h < r_hi) ? __ldcg(cp[c] + ((rt + h) << 2)) : make_double2(0.0, 0.0);
                    };
                    load_pair(r_lo, cur);
                    for (int rt = r_lo; rt < r_hi; rt += 2) {
                        if (rt + 2 < r_hi) load_pair(rt + 2, nxt);
                        double mf[2][KS];
#pragma unroll
                        for (int h = 0; h < 2; h++)
#pragma unroll
                            for (int q = 0; q < KS; q++)
                                mf[h][q] = (rt + h < r_hi) ? sM[((((rt + h) * KS) + q) << 5) + lane] : 0.0;
#pragma unroll
                        for (int q = 0; q < KS; q++)
#pragma unroll
                            for (int h = 0; h < 2; h++)
#pragma unroll
                                for (int c = 0; c < CT; c++) dmma_8x8x4(cur[h][c].x, cur[h][c].y, xf[c][q], mf[h][q]);
#pragma unroll
                        for (int h = 0; h < 2; h++)
#pragma unroll
                            for (int c = 0; c < CT; c++)
                                if (cv[c] && rt + h < r_hi) __stcg(cp[c] + ((rt + h) << 2), cur[h][c]);
#pragma unroll
                        for (int h = 0; h < 2; h++)
#pragma unroll
                            for (int c = 0; c < CT; c++) cur[h][c] = nxt[h][c];
                    }
                }
                CL_TICK(1, 5);
            }
            cl_sync();
            CL_TICK(0, 2);
            CL_TICK(1, 6);
        }
#ifdef KDSL_PHASE_TICKS
        if (blockIdx.x == 0 && tid == 0) g_inv_phase_cycles[15] += 1;
#endif
        if (sing) cl_sync();                            // nobody re-reads the flag after P has moved on to the next item
        // ---- index map for the consumer: colsrc[i] = elimination step at which row i was the pivot ----
        if (has_row && !sing) colsrc_base[((size_t)2 * b + spin) * cs_stride + tid] = gstep;
    }
}
