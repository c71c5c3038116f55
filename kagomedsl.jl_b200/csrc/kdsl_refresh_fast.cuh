// kdsl_refresh_fast.cuh -- production kernels of the periodic W re-evaluation
// (reference reevaluateW!, src/MonteCarlo.jl:55-66:  W = U * (tilde_U \ I)).
//
//   k_inverse_blocked<NB, RPT> : batched in-place inversion of tilde_U (one CTA per matrix, matrix
//       resident in L2), blocked Gauss-Jordan with partial (row) pivoting.  Per block step the
//       NB-column panel is factorised in registers with LU arithmetic (one row per thread, same
//       pivot rule and exact-zero singularity test as LAPACK getrf), and the rank-NB update of the
//       rest of the matrix runs on the FP64 tensor pipe (mma.sync m8n8k4.f64 -> SASS DMMA).
//   k_gemm_W_dmma : W = U * X with X = inverse (column permutation of the pivoting folded into
//       the operand load), 72x72 CTA tiles, 9 warps x (24x24) DMMA tiles, double-buffered smem.
//
// The row interchanges are NOT undone in place: colsrc[j] tells which stored column holds column j
// of the true inverse, and the GEMM reads through it.
#pragma once
#include "kdsl_common.cuh"
#include "kdsl_refresh.cuh"

__device__ __forceinline__ void dmma_8x8x4(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// fragment-major shared layout shared by the A operand (row r, k) and the B operand (k, col r):
// element (r, k) of a [R x KT] operand lives at ((r/8)*(KT/4) + k/4)*32 + (r%8)*4 + k%4, i.e. the 32
// values one DMMA needs from a warp are contiguous (conflict-free LDS.64, no padding).
__device__ __forceinline__ int frag_idx(int r, int k, int KT) {
    return (((r >> 3) * (KT >> 2) + (k >> 2)) << 5) + ((r & 7) << 2) + (k & 3);
}

// Build tilde_U padded to Np = roundup(N, 8) with an identity block (so inv(padded) = padded inv).
// grid (nw, 2); dynamic smem N ints.
__global__ void __launch_bounds__(256)
k_gather_tilde_padded(DevState S, const int *__restrict__ list, double *__restrict__ A_up,
                      double *__restrict__ A_dn, int *__restrict__ status, int Np_up, int Np_dn,
                      int *__restrict__ urow_base, int urow_stride) {
    extern __shared__ int s_site[];
    __shared__ int s_wsum[8];
    __shared__ int s_base;
    const int b = blockIdx.x, spin = blockIdx.y;
    if (b >= batch_count(S, list)) return;
    const int w = list ? list[b] : b;
    const int ns = S.ns, N = spin ? S.n_dn : S.n_up, Np = spin ? Np_dn : Np_up;
    const int *kap = (spin ? S.kdn : S.kup) + (size_t)w * ns;
    const double *U = spin ? S.U_dn : S.U_up;
    double *A = (spin ? A_dn : A_up) + (size_t)b * Np * Np;
    for (int R = threadIdx.x; R < ns; R += blockDim.x) {
        const int l = kap[R];
        if (l != 0) s_site[l - 1] = R;
    }
    if (threadIdx.x == 0) { status[2 * b + spin] = 0; s_base = 0; }
    __syncthreads();
    for (int e = threadIdx.x; e < Np * Np; e += blockDim.x) {
        const int c = e / Np, l = e - c * Np;
        double v;
        if (c < N && l < N) v = U[(size_t)c * ns + s_site[l]];      // tilde_U[l, c] = U[R_l, c]
        else v = (c == l) ? 1.0 : 0.0;
        A[e] = v;
    }
    // ordered list of the sites NOT occupied by this species (the non-trivial rows of W)
    int *urow = urow_base + ((size_t)2 * b + spin) * urow_stride;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int s0 = 0; s0 < ns; s0 += 256) {
        const int site = s0 + threadIdx.x;
        const bool un = site < ns && kap[site] == 0;
        const unsigned m = __ballot_sync(0xffffffffu, un);
        if (lane == 0) s_wsum[warp] = __popc(m);
        __syncthreads();
        int off = s_base;
        for (int q = 0; q < warp; q++) off += s_wsum[q];
        if (un) urow[off + __popc(m & ((1u << lane) - 1u))] = site;
        __syncthreads();
        if (threadIdx.x == 0) {
            int t = 0;
            for (int q = 0; q < 8; q++) t += s_wsum[q];
            s_base += t;
        }
        __syncthreads();
    }
}

// developer instrumentation (build with -DKDSL_PHASE_TICKS): SM cycles spent by CTA 0 in each phase of the
// re-evaluation kernels.  Compiled out of the product library.
#ifdef KDSL_PHASE_TICKS
__device__ long long g_inv_phase_cycles[16];   // [8..15]: k_inverse_cl
#define PHASE_CLOCK() clock64()
#define PHASE_TICK_AT(idx, thr)                                          \
    do {                                                                 \
        if (blockIdx.x == 0 && threadIdx.x == (thr)) {                   \
            const long long now_ = clock64();                            \
            g_inv_phase_cycles[idx] += now_ - t_phase;                   \
            t_phase = now_;                                              \
        }                                                                \
    } while (0)
#define PHASE_RESTART(thr) do { if (blockIdx.x == 0 && threadIdx.x == (thr)) t_phase = clock64(); } while (0)
#else
#define PHASE_CLOCK() 0ll
#define PHASE_TICK_AT(idx, thr) do { (void)t_phase; } while (0)
#define PHASE_RESTART(thr) do { } while (0)
#endif
#define PHASE_TICK(idx) PHASE_TICK_AT(idx, 0)

#ifdef KDSL_DEV_VARIANTS   // first blocked inverse (explicit row exchanges): superseded by k_inverse_v4 / v5 and k_reeval_fused
template <int NB, int RPT, int T, int MINB>
__global__ void __launch_bounds__(T, MINB)
k_inverse_blocked(DevState S, const int *__restrict__ list, double *__restrict__ A_base, int spin,
                  int *__restrict__ status, int *__restrict__ colsrc_base, int Np, int cs_stride) {
    constexpr int NWARP = T / 32;
    extern __shared__ double sm[];
    const int b = blockIdx.x;
    if (b >= batch_count(S, list)) return;
    double *sL = sm;                                   // [Np x NB] frag-major, holds -L'
    double *sU = sL + (size_t)Np * NB;                 // [Np x NB] frag-major, holds U_K,:
    double *sLU = sU + (size_t)Np * NB;                // [NB x NB] row-major packed LU of the pivot block
    double *sRow = sLU + NB * NB;                      // [2][NB] row exchange
    double *sRed = sRow + 2 * NB;                      // [NWARP] warp maxima
    int *sRedI = reinterpret_cast<int *>(sRed + NWARP);  // [NWARP] their rows
    int *sPiv = sRedI + NWARP;                         // [Np] pivot row chosen at each elimination step
    int *sMvPos = sPiv + Np;                           // [2*NB] displaced-row moves of the current panel
    int *sMvSrc = sMvPos + 2 * NB;                     // [2*NB]
    int *sSrcK = sMvSrc + 2 * NB;                      // [NB] source row of each pivot position
    __shared__ int sNmv;

    double *A = A_base + (size_t)b * Np * Np;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int e = tid; e < NB * NB; e += T) sLU[e] = 0.0;
    __syncthreads();
    long long t_phase = PHASE_CLOCK();

    for (int k0 = 0; k0 < Np; k0 += NB) {
        const int kw = min(NB, Np - k0);               // multiple of 8
        // ---- 1. panel columns -> registers, one matrix row per (thread, r) ----
        double a[RPT][NB];
#pragma unroll
        for (int r = 0; r < RPT; r++) {
            const int i = tid + T * r;
#pragma unroll
            for (int k = 0; k < NB; k++) a[r][k] = (i < Np && k < kw) ? A[(size_t)(k0 + k) * Np + i] : 0.0;
        }
        PHASE_TICK(0);
        // ---- 2. LU of the panel with partial pivoting over the not-yet-pivoted rows (>= k0+k);
        //         rows above k0 are eliminated as well (Gauss-Jordan) ----
#pragma unroll
        for (int k = 0; k < NB; k++) {
            if (k < kw) {
                const int gk = k0 + k;
                double best = -1.0;
                int bi = 0x7fffffff;
#pragma unroll
                for (int r = 0; r < RPT; r++) {
                    const int i = tid + T * r;
                    if (i < Np && i >= gk) {
                        double v = fabs(a[r][k]);
                        if (!(v == v)) v = INFINITY;
                        if (v > best) { best = v; bi = i; }
                    }
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const double ov = __shfl_xor_sync(0xffffffffu, best, o);
                    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                    if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
                }
                if (lane == 0) { sRed[warp] = best; sRedI[warp] = bi; }
                __syncthreads();
                double bv = sRed[0];
                int p = sRedI[0];
#pragma unroll
                for (int q = 1; q < NWARP; q++) {
                    const double ov = sRed[q];
                    const int oi = sRedI[q];
                    if (ov > bv || (ov == bv && oi < p)) { bv = ov; p = oi; }
                }
                if (!(bv > 0.0) || bv == INFINITY) {           // exact-zero or non-finite pivot: singular
                    if (tid == 0) status[2 * b + spin] = 1;
                    return;
                }
                if (tid == 0) sPiv[gk] = p;
#pragma unroll
                for (int r = 0; r < RPT; r++) {
                    const int i = tid + T * r;
                    if (i == p) {
#pragma unroll
                        for (int j = 0; j < NB; j++) sRow[j] = a[r][j];
                    }
                    if (i == gk) {
#pragma unroll
                        for (int j = 0; j < NB; j++) sRow[NB + j] = a[r][j];
                    }
                }
                __syncthreads();
                if (p != gk) {
#pragma unroll
                    for (int r = 0; r < RPT; r++) {
                        const int i = tid + T * r;
                        if (i == p) {
#pragma unroll
                            for (int j = 0; j < NB; j++) a[r][j] = sRow[NB + j];
                        }
                        if (i == gk) {
#pragma unroll
                            for (int j = 0; j < NB; j++) a[r][j] = sRow[j];
                        }
                    }
                }
                const double pivot = sRow[k];
#pragma unroll
                for (int r = 0; r < RPT; r++) {
                    const int i = tid + T * r;
                    if (i < Np && !(i >= k0 && i <= gk)) {
                        const double l = a[r][k] / pivot;
                        a[r][k] = l;
#pragma unroll
                        for (int j = k + 1; j < NB; j++) a[r][j] = fma(-l, sRow[j], a[r][j]);
                    }
                }
                __syncthreads();
            }
        }
        PHASE_TICK(1);
        // ---- 3. publish -L' (zero on the pivot rows) and the packed LU of the pivot block ----
#pragma unroll
        for (int r = 0; r < RPT; r++) {
            const int i = tid + T * r;
            if (i < Np) {
                const bool is_piv = (i >= k0 && i < k0 + kw);
#pragma unroll
                for (int k = 0; k < NB; k++) sL[frag_idx(i, k, NB)] = is_piv ? 0.0 : -a[r][k];
                if (is_piv) {
#pragma unroll
                    for (int k = 0; k < NB; k++) sLU[(i - k0) * NB + k] = a[r][k];
#pragma unroll
                    for (int k = 0; k < NB; k++) if (k == i - k0) sRow[k] = 1.0 / a[r][k];   // 1 / U_kk
                }
            }
        }
        // composed row interchanges of this panel: which original row lands on each touched position.
        // sIdx (scratch in the not-yet-written sU region) maps a row to its slot in the move list.
        int *sIdx = reinterpret_cast<int *>(sU);
        for (int k = tid; k < kw; k += T) sIdx[sPiv[k0 + k]] = -1;
        __syncthreads();
        if (tid == 0) {
            int n = 0;
            for (int k = 0; k < kw; k++) { sMvPos[n] = k0 + k; sMvSrc[n] = k0 + k; n++; }
            for (int k = 0; k < kw; k++) {
                const int p = sPiv[k0 + k];
                if (p == k0 + k) continue;
                int ip;
                if (p < k0 + kw) ip = p - k0;                      // another pivot position
                else {
                    ip = sIdx[p];
                    if (ip < 0) { sMvPos[n] = p; sMvSrc[n] = p; sIdx[p] = n; ip = n; n++; }
                }
                const int t0 = sMvSrc[k]; sMvSrc[k] = sMvSrc[ip]; sMvSrc[ip] = t0;
            }
            for (int k = 0; k < kw; k++) sSrcK[k] = sMvSrc[k];
            sNmv = n;                                              // entries [kw, n) are displaced rows outside the pivot block
        }
        __syncthreads();
        PHASE_TICK(2);
        // ---- 4. per column outside the panel: apply the interchanges, U_K,j = L_KK^-1 A_K,j (kept for
        //         the update), final pivot rows A_K,j = U_KK^-1 U_K,j ----
        const int nmv = sNmv;
#pragma unroll
        for (int r = 0; r < RPT; r++) {
            const int j = tid + T * r;
            if (j < Np && !(j >= k0 && j < k0 + kw)) {
                double *col = A + (size_t)j * Np;
                double u[NB];
#pragma unroll
                for (int k = 0; k < NB; k++) u[k] = (k < kw) ? col[sSrcK[k]] : 0.0;
                for (int e0 = kw; e0 < nmv; e0 += 8) {                // displaced rows, 8 at a time
                    double t[8];
#pragma unroll
                    for (int e = 0; e < 8; e++) if (e0 + e < nmv) t[e] = col[sMvSrc[e0 + e]];
#pragma unroll
                    for (int e = 0; e < 8; e++) if (e0 + e < nmv) col[sMvPos[e0 + e]] = t[e];
                }
#pragma unroll
                for (int k = 1; k < NB; k++) {
                    if (k < kw) {
#pragma unroll
                        for (int m = 0; m < k; m++) u[k] = fma(-sLU[k * NB + m], u[m], u[k]);
                    }
                }
#pragma unroll
                for (int k = 0; k < NB; k++) sU[frag_idx(j, k, NB)] = u[k];
#pragma unroll
                for (int k = NB - 1; k >= 0; k--) {
                    if (k < kw) {
#pragma unroll
                        for (int m = k + 1; m < NB; m++) u[k] = fma(-sLU[k * NB + m], u[m], u[k]);
                        u[k] = u[k] * sRow[k];
                    }
                }
#pragma unroll
                for (int k = 0; k < NB; k++) if (k < kw) col[k0 + k] = u[k];
            }
        }
        PHASE_TICK(3);
        // ---- 5. the panel columns of the result: A_IK = -L' L_KK^-1 (other rows), A_KK = U_KK^-1 L_KK^-1 ----
#pragma unroll
        for (int r = 0; r < RPT; r++) {
            const int i = tid + T * r;
            if (i < Np) {
                double x[NB];
                if (i >= k0 && i < k0 + kw) {
                    const int q = i - k0;                             // row q of inv(L U): y U = e_q, then x L = y
#pragma unroll
                    for (int k = 0; k < NB; k++) {
                        double y = (k == q) ? 1.0 : 0.0;
#pragma unroll
                        for (int m = 0; m < k; m++) y = fma(-x[m], sLU[m * NB + k], y);
                        x[k] = (k < kw && k >= q) ? y * sRow[k] : 0.0;
                    }
                } else {
#pragma unroll
                    for (int k = 0; k < NB; k++) x[k] = -a[r][k];
                }
#pragma unroll
                for (int k = NB - 1; k >= 0; k--) {
                    if (k < kw) {
#pragma unroll
                        for (int m = k + 1; m < NB; m++) x[k] = fma(-x[m], sLU[m * NB + k], x[k]);
                    }
                }
#pragma unroll
                for (int k = 0; k < NB; k++) if (k < kw) A[(size_t)(k0 + k) * Np + i] = x[k];
            }
        }
        __syncthreads();
        PHASE_TICK(4);
        // ---- 6. rank-kw update of everything outside the pivot rows / panel columns on the FP64
        //         tensor pipe: A_I,J += (-L'_I) U_K,J.  Work item = 32-row strip x group of 4 col tiles ----
        {
            const int strips = (Np + 31) >> 5, ctiles = Np >> 3;
            const int cgroups = (ctiles + 3) >> 2;
            const int gr = lane >> 2, tg = lane & 3;
            for (int item = warp; item < strips * cgroups; item += NWARP) {
                const int strip = item / cgroups, cg = item - strip * cgroups;
                const int r0 = strip << 5;
                double af[4][NB / 4];
                bool mval[4];
#pragma unroll
                for (int m = 0; m < 4; m++) {
                    const int rt = r0 + 8 * m;
                    mval[m] = rt < Np && !(rt >= k0 && rt < k0 + kw);
#pragma unroll
                    for (int s = 0; s < NB / 4; s++)
                        af[m][s] = mval[m] ? sL[((((rt >> 3) * (NB >> 2)) + s) << 5) + lane] : 0.0;
                }
                const int ct_end = min(cg * 4 + 4, ctiles);
                auto tile_ok = [&](int ct) { return ct < ct_end && !((ct << 3) >= k0 && (ct << 3) < k0 + kw); };
                auto load_c = [&](int ct, double (&c)[4][2]) {
                    const double *p0 = A + (size_t)((ct << 3) + 2 * tg) * Np + r0 + gr;
#pragma unroll
                    for (int m = 0; m < 4; m++) {
                        c[m][0] = mval[m] ? p0[8 * m] : 0.0;
                        c[m][1] = mval[m] ? p0[8 * m + Np] : 0.0;
                    }
                };
                auto mma_store = [&](int ct, double (&c)[4][2]) {
#pragma unroll
                    for (int s = 0; s < NB / 4; s++) {
                        const double bf = sU[(((ct * (NB >> 2)) + s) << 5) + lane];
#pragma unroll
                        for (int m = 0; m < 4; m++) dmma_8x8x4(c[m][0], c[m][1], af[m][s], bf);
                    }
                    double *p0 = A + (size_t)((ct << 3) + 2 * tg) * Np + r0 + gr;
#pragma unroll
                    for (int m = 0; m < 4; m++) {
                        if (mval[m]) {
                            p0[8 * m] = c[m][0];
                            p0[8 * m + Np] = c[m][1];
                        }
                    }
                };
                // the (up to) 4 column tiles of this item, two tiles of loads in flight ahead of the DMMAs
                double c0[4][2], c1[4][2];
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
    // ---- 7. inv(A) = R * P: record which stored column of R is each column of the inverse ----
    int *colsrc = colsrc_base + ((size_t)2 * b + spin) * cs_stride;
    int *sC = reinterpret_cast<int *>(sL);
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
#endif  // KDSL_DEV_VARIANTS

// W[w][unoccupied sites, :] = U[unoccupied sites, :] * X,  X[:, j] = R[:, colsrc[j]] (R stored with leading
// dimension Np; with perm_k the stored rows are permuted as well: X[k, j] = R[rho, colsrc[j]], k = colsrc[rho]); rows of W on occupied sites are the unit vectors e_l (SURVEY 8(a) invariant) and are written by
// the CTA whose row tile spans them, so that every 32-byte sector is completed by one CTA.
// grid (tiles_m * tiles_n, nw, 2), 288 threads = 9 warps in a 3x3 arrangement of 24x24 warp tiles.
template <int KT>
__global__ void __launch_bounds__(288, 2)
k_gemm_W_dmma(DevState S, const int *__restrict__ list, const double *__restrict__ X_up,
              const double *__restrict__ X_dn, const int *__restrict__ status,
              const int *__restrict__ colsrc_base, int Np_up, int Np_dn, int cs_stride,
              const int *__restrict__ urow_base, int urow_stride, int perm_k) {
    constexpr int TM = 72, TN = 72, NT = 288;
    extern __shared__ double gsm[];
    double (*sA)[TM * KT] = reinterpret_cast<double (*)[TM * KT]>(gsm);
    double (*sB)[TN * KT] = reinterpret_cast<double (*)[TN * KT]>(gsm + 2 * TM * KT);
    __shared__ int sSrc[TN];
    __shared__ int sRowSite[TM + 1];
    __shared__ int sKp[1024 + 32];                         // contraction index -> column of U (implicit-pivoting inverse)
    const int b = blockIdx.y, spin = blockIdx.z;
    if (b >= batch_count(S, list)) return;
    if (status[2 * b] | status[2 * b + 1]) return;
    const int w = list ? list[b] : b;
    const int ns = S.ns, N = spin ? S.n_dn : S.n_up, Np = spin ? Np_dn : Np_up;
    const int M = ns - N;                                  // unoccupied sites of this species
    const int tiles_m = (M + TM - 1) / TM;
    const int tm = blockIdx.x % tiles_m, tn = blockIdx.x / tiles_m;
    const int m0 = tm * TM, n0 = tn * TN;
    if (n0 >= N) return;
    const double *U = spin ? S.U_dn : S.U_up;
    const double *X = (spin ? X_dn : X_up) + (size_t)b * Np * Np;
    const int *colsrc = colsrc_base + ((size_t)2 * b + spin) * cs_stride;
    const int *urow = urow_base + ((size_t)2 * b + spin) * urow_stride;
    const int *kap = (spin ? S.kdn : S.kup) + (size_t)w * ns;
    double *W = (spin ? S.W_dn : S.W_up) + (size_t)w * ns * N;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp % 3, wn = warp / 3;
    long long t_phase = PHASE_CLOCK();
    if (tid < TN) sSrc[tid] = (n0 + tid < N) ? colsrc[n0 + tid] : -1;
    if (tid <= TM) sRowSite[tid] = (m0 + tid < M) ? urow[m0 + tid] : -1;
    for (int k = tid; k < 1024 + 32; k += NT) sKp[k] = k < N ? (perm_k ? colsrc[k] : k) : 0;
    __syncthreads();

    constexpr int PER = TM * KT / NT;                  // elements of each operand per thread per stage
    static_assert(TM * KT % NT == 0, "stage must divide evenly");
    double ra[PER], rb[PER];
    auto load_stage = [&](int kk) {
#pragma unroll
        for (int q = 0; q < PER; q++) {
            const int e = tid + q * NT;
            const int r = e % TM, k = e / TM;                      // A: gathered rows of U
            const int site = sRowSite[r];
            ra[q] = (site >= 0 && kk + k < N) ? U[(size_t)sKp[kk + k] * ns + site] : 0.0;
            const int kb = e % KT, n = e / KT;                     // B: contiguous along k
            const int src = sSrc[n];
            rb[q] = (src >= 0 && kk + kb < N) ? X[(size_t)src * Np + kk + kb] : 0.0;
        }
    };
    auto store_stage = [&](int buf) {
#pragma unroll
        for (int q = 0; q < PER; q++) {
            const int e = tid + q * NT;
            sA[buf][frag_idx(e % TM, e / TM, KT)] = ra[q];
            sB[buf][frag_idx(e / KT, e % KT, KT)] = rb[q];
        }
    };
    double c[3][3][2];
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) c[i][j][0] = c[i][j][1] = 0.0;

    load_stage(0);
    store_stage(0);
    __syncthreads();
    PHASE_TICK(5);
    int buf = 0;
    for (int kk = 0; kk < N; kk += KT) {
        const bool more = kk + KT < N;
        if (more) load_stage(kk + KT);
#pragma unroll
        for (int s = 0; s < KT / 4; s++) {
            double af[3], bf[3];
#pragma unroll
            for (int i = 0; i < 3; i++) af[i] = sA[buf][((((wm * 3 + i) * (KT >> 2)) + s) << 5) + lane];
#pragma unroll
            for (int j = 0; j < 3; j++) bf[j] = sB[buf][((((wn * 3 + j) * (KT >> 2)) + s) << 5) + lane];
#pragma unroll
            for (int i = 0; i < 3; i++)
#pragma unroll
                for (int j = 0; j < 3; j++) dmma_8x8x4(c[i][j][0], c[i][j][1], af[i], bf[j]);
        }
        if (more) store_stage(buf ^ 1);
        __syncthreads();
        buf ^= 1;
    }
    PHASE_TICK(6);
    const int gr = lane >> 2, tg = lane & 3;
#pragma unroll
    for (int j = 0; j < 3; j++) {
#pragma unroll
        for (int e = 0; e < 2; e++) {
            const int n = n0 + 24 * wn + 8 * j + 2 * tg + e;
            if (n >= N) continue;
#pragma unroll
            for (int i = 0; i < 3; i++) {
                const int site = sRowSite[24 * wm + 8 * i + gr];
                if (site >= 0) W[(size_t)n * ns + site] = c[i][j][e];
            }
        }
    }
    // unit rows of the occupied sites inside this tile's site range [s_lo, s_hi)
    const int s_lo = (tm == 0) ? 0 : sRowSite[0];
    const int s_hi = (tm == tiles_m - 1 || sRowSite[TM] < 0) ? ns : sRowSite[TM];
    const int ncols = min(TN, N - n0);
    for (int site = s_lo + tid; site < s_hi; site += NT) {   // one site per thread: its label is read once
        const int l = kap[site];
        if (l != 0) {
            double *dst = W + (size_t)n0 * ns + site;
            for (int cc = 0; cc < ncols; cc++) dst[(size_t)cc * ns] = (l - 1 == n0 + cc) ? 1.0 : 0.0;
        }
    }
    __syncthreads();
    PHASE_TICK(7);
}

// (A 3-warp, 72 x 24 tile version of this kernel -- the layout of the ComplexF64 product k_gemm_W_dmma_c, which reaches 63 % of
// the DMMA peak -- was measured SLOWER here: 28.1 ms against 22.6 ms per 2048-walker bin at 972 sites.  A real block product
// has a quarter of the complex one's arithmetic per operand byte, and the narrow tile re-reads the gathered rows of U three
// times as often.)

#ifdef KDSL_DEV_VARIANTS   // superseded kernels: built only with `make DEV=1`, not part of the product library
// cp.async version of k_gemm_W_dmma (gemm_variant 0): the operands go global -> shared memory directly (LDGSTS, 8 bytes
// per element with zero fill), STAGES stages deep, so no register staging, one barrier per stage and loads in flight
// across stages.  Same tiling (72x72 per CTA, 3x3 warps of 24x24), fragment-major shared layout, same epilogue.
__device__ __forceinline__ void cp_async_8(unsigned dst_smem, const void *src, bool valid) {
    const int sz = valid ? 8 : 0;                          // src-size 0: the 8 destination bytes are zero filled
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(dst_smem), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N_> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N_) : "memory"); }

template <int KT, int STAGES, int MINB>
__global__ void __launch_bounds__(288, MINB)
k_gemm_W_cpasync(DevState S, const int *__restrict__ list, const double *__restrict__ X_up,
                 const double *__restrict__ X_dn, const int *__restrict__ status,
                 const int *__restrict__ colsrc_base, int Np_up, int Np_dn, int cs_stride,
                 const int *__restrict__ urow_base, int urow_stride, int perm_k) {
    constexpr int TM = 72, TN = 72, NT = 288, KS = KT / 4;
    constexpr int PER = TM * KT / NT;                      // elements of each operand per thread per stage
    static_assert(TM * KT % NT == 0 && NT % TM == 0 && NT % KT == 0 && TM == TN, "stage must divide evenly");
    extern __shared__ double gsm[];                        // [STAGES][A: TM*KT | B: TN*KT]
    __shared__ int sSrc[TN];
    __shared__ int sRowSite[TM + 1];
    __shared__ int sKp[1024 + 32];
    const int b = blockIdx.y, spin = blockIdx.z;
    if (b >= batch_count(S, list)) return;
    if (status[2 * b] | status[2 * b + 1]) return;
    const int w = list ? list[b] : b;
    const int ns = S.ns, N = spin ? S.n_dn : S.n_up, Np = spin ? Np_dn : Np_up;
    const int M = ns - N;
    const int tiles_m = (M + TM - 1) / TM;
    const int tm = blockIdx.x % tiles_m, tn = blockIdx.x / tiles_m;
    const int m0 = tm * TM, n0 = tn * TN;
    if (n0 >= N) return;
    const double *U = spin ? S.U_dn : S.U_up;
    const double *X = (spin ? X_dn : X_up) + (size_t)b * Np * Np;
    const int *colsrc = colsrc_base + ((size_t)2 * b + spin) * cs_stride;
    const int *urow = urow_base + ((size_t)2 * b + spin) * urow_stride;
    const int *kap = (spin ? S.kdn : S.kup) + (size_t)w * ns;
    double *W = (spin ? S.W_dn : S.W_up) + (size_t)w * ns * N;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp % 3, wn = warp / 3;
    if (tid < TN) sSrc[tid] = (n0 + tid < N) ? colsrc[n0 + tid] : -1;
    if (tid <= TM) sRowSite[tid] = (m0 + tid < M) ? urow[m0 + tid] : -1;
    for (int k = tid; k < 1024 + 32; k += NT) sKp[k] = k < N ? (perm_k ? colsrc[k] : k) : 0;
    __syncthreads();

    // element e = tid + q NT of a stage.  A (row r = e % TM, k = e / TM): this thread always serves the same gathered
    // row of U, k = ka + 4 q, shared offset a_so + 32 q.  B (kb = e % KT, n = e / KT): the same kb, column nb + 12 q.
    const int a_site = sRowSite[tid % TM], ka = tid / TM;
    const double *a_ptr = U + (a_site >= 0 ? a_site : 0);
    const unsigned smem0 = (unsigned)__cvta_generic_to_shared(gsm);
    const unsigned a_so = smem0 + 8u * frag_idx(tid % TM, ka, KT);
    const int kb = tid % KT, nb = tid / KT;
    unsigned b_so[PER];
#pragma unroll
    for (int q = 0; q < PER; q++) b_so[q] = smem0 + 8u * (TM * KT + frag_idx(nb + (NT / KT) * q, kb, KT));
    constexpr unsigned STAGE_BYTES = 8u * (TM * KT + TN * KT);
    auto issue_stage = [&](int kk, int buf) {
        const unsigned off = buf * STAGE_BYTES;
#pragma unroll
        for (int q = 0; q < PER; q++) {
            const int k = kk + ka + (NT / TM) * q;
            const bool va = a_site >= 0 && k < N;
            cp_async_8(a_so + off + 256u * q, a_ptr + (size_t)sKp[k] * ns, va);
            const int src = sSrc[nb + (NT / KT) * q];
            const bool vb = src >= 0 && kk + kb < N;
            cp_async_8(b_so[q] + off, X + (size_t)(src >= 0 ? src : 0) * Np + kk + kb, vb);
        }
    };
    double c[3][3][2];
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) c[i][j][0] = c[i][j][1] = 0.0;
    const int nk = (N + KT - 1) / KT;
#pragma unroll
    for (int s = 0; s < STAGES - 1; s++) {
        if (s < nk) issue_stage(s * KT, s);
        cp_async_commit();
    }
    for (int it = 0; it < nk; it++) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        if (it + STAGES - 1 < nk) issue_stage((it + STAGES - 1) * KT, (it + STAGES - 1) % STAGES);
        cp_async_commit();
        const double *sA = gsm + (size_t)(it % STAGES) * (TM * KT + TN * KT);
        const double *sB = sA + TM * KT;
#pragma unroll
        for (int s = 0; s < KS; s++) {
            double af[3], bf[3];
#pragma unroll
            for (int i = 0; i < 3; i++) af[i] = sA[((((wm * 3 + i) * KS) + s) << 5) + lane];
#pragma unroll
            for (int j = 0; j < 3; j++) bf[j] = sB[((((wn * 3 + j) * KS) + s) << 5) + lane];
#pragma unroll
            for (int i = 0; i < 3; i++)
#pragma unroll
                for (int j = 0; j < 3; j++) dmma_8x8x4(c[i][j][0], c[i][j][1], af[i], bf[j]);
        }
    }
    const int gr = lane >> 2, tg = lane & 3;
#pragma unroll
    for (int j = 0; j < 3; j++) {
#pragma unroll
        for (int e = 0; e < 2; e++) {
            const int n = n0 + 24 * wn + 8 * j + 2 * tg + e;
            if (n >= N) continue;
#pragma unroll
            for (int i = 0; i < 3; i++) {
                const int site = sRowSite[24 * wm + 8 * i + gr];
                if (site >= 0) W[(size_t)n * ns + site] = c[i][j][e];
            }
        }
    }
    const int s_lo = (tm == 0) ? 0 : sRowSite[0];
    const int s_hi = (tm == tiles_m - 1 || sRowSite[TM] < 0) ? ns : sRowSite[TM];
    const int ncols = min(TN, N - n0);
    for (int site = s_lo + tid; site < s_hi; site += NT) {   // one site per thread: its label is read once
        const int l = kap[site];
        if (l != 0) {
            double *dst = W + (size_t)n0 * ns + site;
            for (int cc = 0; cc < ncols; cc++) dst[(size_t)cc * ns] = (l - 1 == n0 + cc) ? 1.0 : 0.0;
        }
    }
}

#endif  // KDSL_DEV_VARIANTS