// kdsl_inverse_v3.cuh -- batched in-place inversion of tilde_U, third generation.
//
// Same mathematics as k_inverse_blocked (blocked Gauss-Jordan, partial row pivoting with LAPACK's pivot
// rule and exact-zero singularity test, rank-NB trailing update on the FP64 tensor pipe) but organised for
// a small instruction footprint and L2 residency, the two things the profile of the previous kernel
// pointed at (instruction-cache hit rate 62 %, DRAM traffic 25x the matrix):
//   * ONE CTA of T threads per matrix and per SM: 148 live matrices x 8 N^2 bytes stay in the 126 MB L2;
//   * the NB-column panel lives in shared memory and is factorised with rolled loops, implicit pivoting
//     (rows stay where they are until the panel is done) and ONE block barrier per pivot step;
//   * everything that is not the pivot search is a small DMMA GEMM with fragment-major shared operands:
//       U_K = Linv X_raw,  pivot rows = (Uinv Linv) X_raw,  panel block column = (-L') Linv,
//       trailing update A += (-L') U_K (transposed form: every thread owns two adjacent rows -> 128-bit access).
#pragma once
#include "kdsl_common.cuh"
#include "kdsl_refresh.cuh"
#include "kdsl_refresh_fast.cuh"

template <int NB, int RPT, int T>
__global__ void __launch_bounds__(T, 1)
k_inverse_v3(DevState S, const int *__restrict__ list, double *__restrict__ A_base, int spin,
             int *__restrict__ status, int *__restrict__ colsrc_base, int Np, int cs_stride) {
    constexpr int NWARP = T / 32;
    constexpr int KS = NB / 4;                          // DMMA k-steps per panel
    extern __shared__ double sm[];
    const int b = blockIdx.x;
    if (b >= batch_count(S, list)) return;
    // ---- shared memory carve-up ----
    double *sP = sm;                                    // [NB][Np] panel, column-major; later X_raw (frag-major)
    double *sL = sP + (size_t)NB * Np;                  // [Np x NB] frag-major: -L' by FINAL row position
    double *sU = sL + (size_t)NB * Np;                  // [Np x NB] frag-major: U_K by column
    double *sLU = sU + (size_t)NB * Np;                 // [NB][NB] packed LU of the pivot block (row-major)
    double *sLi = sLU + NB * NB;                        // frag-major (r = k, kk = m): Linv[k][m]
    double *sLiT = sLi + NB * NB;                       // frag-major (r = c, kk = m): Linv[m][c]
    double *sBm = sLiT + NB * NB;                       // frag-major (r = k, kk = m): (Uinv Linv)[k][m]
    double *sUi = sBm + NB * NB;                        // [NB][NB] row-major Uinv
    double *sRinv = sUi + NB * NB;                      // [2][NWARP]
    unsigned long long *sKey = reinterpret_cast<unsigned long long *>(sRinv + 2 * NWARP);   // [2][NWARP]
    int *sKeyI = reinterpret_cast<int *>(sKey + 2 * NWARP);                                 // [2][NWARP]
    int *sPiv = sKeyI + 2 * NWARP;                      // [Np] sequential-swap record (LAPACK ipiv semantics)
    int *sPosOf = sPiv + Np;                            // [Np] current position of an original row (this panel)
    int *sRowAt = sPosOf + Np;                          // [Np] original row at a position (this panel)
    int *sPivRow = sRowAt + Np;                         // [NB] original row chosen at each step of the panel
    int *sMvPos = sPivRow + NB;                         // [2 NB] displaced positions outside the pivot block
    int *sMvSrc = sMvPos + 2 * NB;                      // [2 NB] ... and the original rows that land there
    __shared__ int sNmv;

    double *A = A_base + (size_t)b * Np * Np;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int gr = lane >> 2, tg = lane & 3;

    long long t_phase = PHASE_CLOCK();
    for (int k0 = 0; k0 < Np; k0 += NB) {
        const int kw = min(NB, Np - k0);                // multiple of 8
        // ---- 1. panel columns -> shared memory; identity row maps ----
        for (int x = tid; x < kw * Np; x += T) {
            const int c = x / Np, i = x - c * Np;
            sP[(size_t)c * Np + i] = A[(size_t)(k0 + c) * Np + i];
        }
        for (int i = tid; i < Np; i += T) { sPosOf[i] = i; sRowAt[i] = i; }
        if (tid == 0) sNmv = 0;
        unsigned pivmask = 0u;                          // bit r: my row (tid + T r) was chosen as a pivot in this panel
        __syncthreads();
        PHASE_TICK(0);
        // ---- 2. LU of the panel, implicit partial pivoting, one barrier per pivot step ----
        for (int k = 0; k < kw; k++) {
            const int par = k & 1;
            unsigned long long key = 0ull;
            int bi = 0x7fffffff;
            double bval = 1.0;
#pragma unroll
            for (int r = 0; r < RPT; r++) {
                const int i = tid + T * r;
                if (i < Np && i >= k0 && !((pivmask >> r) & 1u)) {
                    const double v = sP[(size_t)k * Np + i];
                    const unsigned long long kk = (unsigned long long)__double_as_longlong(fabs(v));
                    if (bi == 0x7fffffff || kk > key) { key = kk; bi = i; bval = v; }
                }
            }
            const bool valid = bi != 0x7fffffff;
            const unsigned hi = (unsigned)(key >> 32), lo = (unsigned)key;
            const unsigned mhi = __reduce_max_sync(0xffffffffu, valid ? hi : 0u);
            const bool c1 = valid && hi == mhi;
            const unsigned mlo = __reduce_max_sync(0xffffffffu, c1 ? lo : 0u);
            const unsigned win = __ballot_sync(0xffffffffu, c1 && lo == mlo);
            if (win == 0u) {
                if (lane == 0) sKeyI[par * NWARP + warp] = 0x7fffffff;
            } else if (lane == __ffs(win) - 1) {        // first lane = lowest row among the warp's maxima
                sKey[par * NWARP + warp] = key;
                sKeyI[par * NWARP + warp] = bi;
                sRinv[par * NWARP + warp] = 1.0 / bval;
            }
            __syncthreads();
            unsigned long long bk = 0ull;
            int p = 0x7fffffff;
            double rinv = 0.0;
#pragma unroll 4
            for (int q = 0; q < NWARP; q++) {
                const int oi = sKeyI[par * NWARP + q];
                const unsigned long long ok = sKey[par * NWARP + q];
                if (oi != 0x7fffffff && (p == 0x7fffffff || ok > bk || (ok == bk && oi < p))) {
                    bk = ok; p = oi; rinv = sRinv[par * NWARP + q];
                }
            }
            if (p == 0x7fffffff || bk == 0ull || bk >= 0x7ff0000000000000ull) {   // exact-zero / non-finite pivot
                if (tid == 0) status[2 * b + spin] = 1;
                return;
            }
            if (tid == 0) sPivRow[k] = p;
#pragma unroll
            for (int r = 0; r < RPT; r++) {
                const int i = tid + T * r;
                if (i == p) pivmask |= 1u << r;
                if (i < Np && !((pivmask >> r) & 1u)) {   // every row that is not a pivot of this panel (above or below)
                    const double l = sP[(size_t)k * Np + i] * rinv;     // LAPACK getf2: scale by the reciprocal pivot
                    sP[(size_t)k * Np + i] = l;
                    for (int j = k + 1; j < kw; j++)
                        sP[(size_t)j * Np + i] = fma(-l, sP[(size_t)j * Np + p], sP[(size_t)j * Np + i]);
                }
            }
        }
        __syncthreads();
        PHASE_TICK(1);
        // ---- 3. bookkeeping: packed LU block, sequential-swap record, final row positions ----
        for (int x = tid; x < kw * kw; x += T) {
            const int q = x / kw, c = x - q * kw;
            sLU[q * NB + c] = sP[(size_t)c * Np + sPivRow[q]];
        }
        if (tid == 0) {
            for (int k = 0; k < kw; k++) {
                const int r = sPivRow[k], g = k0 + k, q = sPosOf[r];
                if (q != g) {
                    const int r2 = sRowAt[g];
                    sRowAt[g] = r; sRowAt[q] = r2;
                    sPosOf[r] = g; sPosOf[r2] = q;
                }
                sPiv[g] = q;
            }
        }
        __syncthreads();
        for (int i = tid; i < Np; i += T) {             // displaced positions outside the pivot block
            if (!(i >= k0 && i < k0 + kw) && sRowAt[i] != i) {
                const int e = atomicAdd(&sNmv, 1);
                sMvPos[e] = i; sMvSrc[e] = sRowAt[i];
            }
        }
        // -L' by final position (zero rows for the pivot block), fragment-major
        for (int x = tid; x < Np * NB; x += T) {
            const int i = x / NB, c = x - i * NB;       // original row i, panel column c
            const int f = sPosOf[i];
            const bool pz = (f >= k0 && f < k0 + kw) || c >= kw;
            sL[frag_idx(f, c, NB)] = pz ? 0.0 : -sP[(size_t)c * Np + i];
        }
        // small triangular inverses of the pivot block (rolled loops, one row per thread)
        if (tid < kw) {
            const int q = tid;
            double *x = sUi + q * NB;                   // row q of Uinv:  y U = e_q
            for (int k = 0; k < kw; k++) {
                double y = (k == q) ? 1.0 : 0.0;
                for (int m = q; m < k; m++) y = fma(-x[m], sLU[m * NB + k], y);
                x[k] = (k >= q) ? y / sLU[k * NB + k] : 0.0;
            }
        } else if (tid >= 32 && tid < 32 + kw) {
            const int q = tid - 32;                     // row q of Linv:  x L = e_q  (unit lower), row-major in sBm for now
            double *x = sBm + q * NB;
            for (int k = kw - 1; k > q; k--) x[k] = 0.0;
            for (int k = q; k >= 0; k--) {
                double v = (k == q) ? 1.0 : 0.0;
                for (int m = k + 1; m <= q; m++) v = fma(-x[m], sLU[m * NB + k], v);
                x[k] = v;
            }
        }
        __syncthreads();
        for (int x = tid; x < NB * NB; x += T) {        // scatter Linv into the two fragment-major operand layouts
            const int q = x / NB, k = x - q * NB;
            const double v = (q < kw && k < kw) ? sBm[q * NB + k] : 0.0;
            sLi[frag_idx(q, k, NB)] = v;                // (r = q, kk = k): Linv[q][k]
            sLiT[frag_idx(k, q, NB)] = v;               // (r = k, kk = q): Linv[q][k]
        }
        __syncthreads();
        for (int x = tid; x < NB * NB; x += T) {        // B = Uinv Linv
            const int q = x / NB, c = x - q * NB;
            double acc = 0.0;
            if (q < kw && c < kw)
                for (int m = max(q, c); m < kw; m++) acc = fma(sUi[q * NB + m], sLi[frag_idx(m, c, NB)], acc);
            sBm[frag_idx(q, c, NB)] = acc;
        }
        __syncthreads();                                // sP is free from here on: reuse it for X_raw
        PHASE_TICK(2);
        // ---- 4. columns outside the panel: gather the raw pivot rows, apply the row moves ----
        double *sX = sP;                                // frag-major (r = column j, kk = k): A[p_k, j]
        const int nmv = sNmv;
#pragma unroll
        for (int r = 0; r < RPT; r++) {
            const int j = tid + T * r;
            if (j < Np) {
                const bool inK = j >= k0 && j < k0 + kw;
                double *col = A + (size_t)j * Np;
                for (int k = 0; k < NB; k++)
                    sX[frag_idx(j, k, NB)] = (!inK && k < kw) ? col[sPivRow[k]] : 0.0;
                if (!inK) {
                    for (int e0 = 0; e0 < nmv; e0 += 8) {
                        double t[8];
#pragma unroll
                        for (int e = 0; e < 8; e++) if (e0 + e < nmv) t[e] = col[sMvSrc[e0 + e]];
#pragma unroll
                        for (int e = 0; e < 8; e++) if (e0 + e < nmv) col[sMvPos[e0 + e]] = t[e];
                    }
                }
            }
        }
        __syncthreads();
        PHASE_TICK(3);
        // ---- 5. small DMMA products per column tile: U_K = Linv X_raw (-> sU) and final pivot rows = B X_raw ----
        {
            const int ctiles = Np >> 3;
            for (int ct = warp; ct < ctiles; ct += NWARP) {
                const int c0 = ct << 3;
                if (c0 >= k0 && c0 < k0 + kw) {         // panel columns: no U_K (kept zero so the update skips them)
                    for (int x = lane; x < 8 * NB; x += 32) sU[(size_t)ct * 8 * NB + x] = 0.0;
                    continue;
                }
                double xa[KS];                          // A operand: X_raw^T fragments (row = column j, k = m)
#pragma unroll
                for (int s = 0; s < KS; s++) xa[s] = sX[(((ct * KS) + s) << 5) + lane];
#pragma unroll
                for (int kt = 0; kt < NB / 8; kt++) {   // output rows k = 8 kt .. 8 kt + 7
                    if (8 * kt >= kw) {
#pragma unroll
                        for (int e = 0; e < 2; e++) sU[frag_idx(c0 + gr, 8 * kt + 2 * tg + e, NB)] = 0.0;
                        continue;
                    }
                    double u0 = 0.0, u1 = 0.0, f0 = 0.0, f1 = 0.0;
#pragma unroll
                    for (int s = 0; s < KS; s++) {
                        dmma_8x8x4(u0, u1, xa[s], sLi[(((kt * KS) + s) << 5) + lane]);   // D[j][k] = sum_m X[m][j] Linv[k][m]
                        dmma_8x8x4(f0, f1, xa[s], sBm[(((kt * KS) + s) << 5) + lane]);   // D[j][k] = sum_m X[m][j] B[k][m]
                    }
                    sU[frag_idx(c0 + gr, 8 * kt + 2 * tg, NB)] = u0;
                    sU[frag_idx(c0 + gr, 8 * kt + 2 * tg + 1, NB)] = u1;
                    *reinterpret_cast<double2 *>(A + (size_t)(c0 + gr) * Np + k0 + 8 * kt + 2 * tg) = make_double2(f0, f1);
                }
            }
        }
        // ---- 6. the panel block column of the result: (-L') Linv for the other rows, B for the pivot block ----
        {
            const int rtiles = Np >> 3;
            for (int rt = warp; rt < rtiles; rt += NWARP) {
                const int r0 = rt << 3;
                if (r0 >= k0 && r0 < k0 + kw) continue;
                double la[KS];
#pragma unroll
                for (int s = 0; s < KS; s++) la[s] = sL[(((rt * KS) + s) << 5) + lane];      // row = final position, k = m
#pragma unroll
                for (int ctk = 0; ctk < NB / 8; ctk++) {
                    if (8 * ctk >= kw) continue;
                    double d0 = 0.0, d1 = 0.0;
#pragma unroll
                    for (int s = 0; s < KS; s++) dmma_8x8x4(d0, d1, la[s], sLiT[(((ctk * KS) + s) << 5) + lane]);
                    A[(size_t)(k0 + 8 * ctk + 2 * tg) * Np + r0 + gr] = d0;     // D[f][c] = sum_m (-L')[f][m] Linv[m][c]
                    A[(size_t)(k0 + 8 * ctk + 2 * tg + 1) * Np + r0 + gr] = d1;
                }
            }
            for (int x = tid; x < kw * kw; x += T) {
                const int q = x / kw, c = x - q * kw;
                A[(size_t)(k0 + c) * Np + k0 + q] = sBm[frag_idx(q, c, NB)];
            }
        }
        __syncthreads();
        PHASE_TICK(4);
        // ---- 7. rank-kw update of everything outside the pivot rows / panel columns (transposed DMMA form) ----
        {
            const int strips = (Np + 31) >> 5, ctiles = Np >> 3;
            const int cgroups = (ctiles + 3) >> 2;
            for (int item = warp; item < strips * cgroups; item += NWARP) {
                const int strip = item / cgroups, cg = item - strip * cgroups;
                const int r0 = strip << 5;
                double lf[4][KS];
                bool mval[4];
#pragma unroll
                for (int m = 0; m < 4; m++) {
                    const int rt = r0 + 8 * m;
                    mval[m] = rt < Np && !(rt >= k0 && rt < k0 + kw);
#pragma unroll
                    for (int s = 0; s < KS; s++) lf[m][s] = mval[m] ? sL[((((rt >> 3) * KS) + s) << 5) + lane] : 0.0;
                }
                const int ct_end = min(cg * 4 + 4, ctiles);
                auto tile_ok = [&](int ct) { return ct < ct_end && !((ct << 3) >= k0 && (ct << 3) < k0 + kw); };
                auto load_c = [&](int ct, double2 (&c)[4]) {
                    const double2 *p0 = reinterpret_cast<const double2 *>(A + (size_t)((ct << 3) + gr) * Np + r0 + 2 * tg);
#pragma unroll
                    for (int m = 0; m < 4; m++) c[m] = mval[m] ? p0[4 * m] : make_double2(0.0, 0.0);
                };
                auto mma_store = [&](int ct, double2 (&c)[4]) {
#pragma unroll
                    for (int s = 0; s < KS; s++) {
                        const double uf = sU[(((ct * KS) + s) << 5) + lane];
#pragma unroll
                        for (int m = 0; m < 4; m++) dmma_8x8x4(c[m].x, c[m].y, uf, lf[m][s]);
                    }
                    double2 *p0 = reinterpret_cast<double2 *>(A + (size_t)((ct << 3) + gr) * Np + r0 + 2 * tg);
#pragma unroll
                    for (int m = 0; m < 4; m++) if (mval[m]) p0[4 * m] = c[m];
                };
                double2 c0[4], c1[4];
                const int ctb = cg * 4;
                if (tile_ok(ctb)) load_c(ctb, c0);
                if (tile_ok(ctb + 1)) load_c(ctb + 1, c1);
                if (tile_ok(ctb)) mma_store(ctb, c0);
                if (tile_ok(ctb + 2)) load_c(ctb + 2, c0);
                if (tile_ok(ctb + 1)) mma_store(ctb + 1, c1);
                if (tile_ok(ctb + 3)) load_c(ctb + 3, c1);
                if (tile_ok(ctb + 2)) mma_store(ctb + 2, c0);
                if (tile_ok(ctb + 3)) mma_store(ctb + 3, c1);
            }
        }
        __syncthreads();
        PHASE_TICK(5);
    }
    // ---- 8. inv(A) = R * P: which stored column of R is each column of the inverse ----
    int *colsrc = colsrc_base + ((size_t)2 * b + spin) * cs_stride;
    int *sC = sPosOf;
    for (int j = tid; j < Np; j += T) sC[j] = j;
    __syncthreads();
    if (tid == 0) {
        for (int k = Np - 1; k >= 0; k--) {
            const int p = sPiv[k];
            if (p != k) { const int t0 = sC[k]; sC[k] = sC[p]; sC[p] = t0; }
        }
    }
    __syncthreads();
    for (int j = tid; j < Np; j += T) colsrc[j] = sC[j];
}
