// kdsl_refresh.cuh -- periodic re-evaluation of W from scratch (reference reevaluateW! + tilde_U,
// src/MonteCarlo.jl:55-66, 92-115):  W = U * inv(tilde_U),  tilde_U[l, :] = U[R_l, :].
// Batched over a device-side list of walkers; three stages:
//   k_gather_tilde : build tilde_U (N x N, column-major) per listed walker and species
//   k_inverse_*    : in-place inversion with partial pivoting (LAPACK getrf/getri semantics:
//                    first maximal |.| pivot, exact-zero / non-finite pivot => singular)
//   k_gemm_W_*     : W = U * X
#pragma once
#include "kdsl_common.cuh"

#define KDSL_FLAG_SINGULAR_DEV 1

// batch entry b -> walker id.  list == nullptr means "all walkers" (b is the walker).
__device__ __forceinline__ int batch_count(const DevState &S, const int *list) {
    return list ? S.cnt[2] : S.nw;
}

// grid (nw, 2): blockIdx.x = batch entry, blockIdx.y = species.  dynamic smem: N ints.
__global__ void __launch_bounds__(256)
k_gather_tilde(DevState S, const int *__restrict__ list, double *__restrict__ A_up,
               double *__restrict__ A_dn, int *__restrict__ status) {
    extern __shared__ int s_site[];
    const int b = blockIdx.x, spin = blockIdx.y;
    if (b >= batch_count(S, list)) return;
    const int w = list ? list[b] : b;
    const int ns = S.ns, N = spin ? S.n_dn : S.n_up;
    const int *kap = (spin ? S.kdn : S.kup) + (size_t)w * ns;
    const double *U = spin ? S.U_dn : S.U_up;
    double *A = (spin ? A_dn : A_up) + (size_t)b * N * N;
    for (int R = threadIdx.x; R < ns; R += blockDim.x) {
        const int l = kap[R];
        if (l != 0) s_site[l - 1] = R;                       // site R_l of particle l
    }
    if (threadIdx.x == 0) status[2 * b + spin] = 0;
    __syncthreads();
    for (int e = threadIdx.x; e < N * N; e += blockDim.x) {
        const int c = e / N, l = e - c * N;
        A[e] = U[(size_t)c * ns + s_site[l]];                // tilde_U[l, c] = U[R_l, c]   (:110)
    }
}

// ---- inverse, variant 0: unblocked in-place Gauss-Jordan in global/L2 memory (simple, slow) ----
// one CTA per (batch entry, species).  dynamic smem: 2*N doubles + N ints + reduction scratch.
__global__ void __launch_bounds__(256)
k_inverse_gj(DevState S, const int *__restrict__ list, double *__restrict__ A_base, int spin,
             int *__restrict__ status) {
    extern __shared__ double sm_d[];
    const int b = blockIdx.x;
    if (b >= batch_count(S, list)) return;
    const int N = spin ? S.n_dn : S.n_up;
    double *prow = sm_d, *colk = sm_d + N;
    int *piv = reinterpret_cast<int *>(sm_d + 2 * N);
    __shared__ double r_val[8];
    __shared__ int r_idx[8];
    __shared__ int s_p;
    double *A = A_base + (size_t)b * N * N;
    const int tid = threadIdx.x, T = blockDim.x, lane = tid & 31, warp = tid >> 5;

    for (int k = 0; k < N; k++) {
        // pivot search in column k, rows >= k: first index of the maximum
        double best = -1.0;
        int bi = k;
        for (int i = k + tid; i < N; i += T) {
            const double v = fabs(A[(size_t)k * N + i]);
            if (v > best || !(v == v)) { best = (v == v) ? v : INFINITY; bi = i; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double ov = __shfl_xor_sync(0xffffffffu, best, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
        }
        if (lane == 0) { r_val[warp] = best; r_idx[warp] = bi; }
        __syncthreads();
        if (tid == 0) {
            double bv = r_val[0];
            int bx = r_idx[0];
            for (int q = 1; q < (T >> 5); q++)
                if (r_val[q] > bv || (r_val[q] == bv && r_idx[q] < bx)) { bv = r_val[q]; bx = r_idx[q]; }
            if (!(bv > 0.0) || bv == INFINITY) bx = -1;      // exact zero or non-finite pivot
            s_p = bx;
            if (bx >= 0) piv[k] = bx;
        }
        __syncthreads();
        const int p = s_p;
        if (p < 0) {
            if (tid == 0) status[2 * b + spin] = 1;
            return;
        }
        if (p != k)
            for (int j = tid; j < N; j += T) {
                const double t0 = A[(size_t)j * N + k];
                A[(size_t)j * N + k] = A[(size_t)j * N + p];
                A[(size_t)j * N + p] = t0;
            }
        __syncthreads();
        for (int i = tid; i < N; i += T) colk[i] = A[(size_t)k * N + i];
        __syncthreads();
        const double d = 1.0 / colk[k];
        for (int j = tid; j < N; j += T) prow[j] = (j == k) ? d : A[(size_t)j * N + k] * d;
        __syncthreads();
        for (int e = tid; e < N * N; e += T) {
            const int j = e / N, i = e - j * N;
            double v;
            if (i == k) v = prow[j];
            else if (j == k) v = -colk[i] * d;
            else v = fma(-colk[i], prow[j], A[e]);
            A[e] = v;
        }
        __syncthreads();
    }
    for (int k = N - 1; k >= 0; k--) {                        // undo the row interchanges on the columns
        const int p = piv[k];
        if (p != k)
            for (int i = tid; i < N; i += T) {
                const double t0 = A[(size_t)k * N + i];
                A[(size_t)k * N + i] = A[(size_t)p * N + i];
                A[(size_t)p * N + i] = t0;
            }
        __syncthreads();
    }
}

// Record singular walkers after the inversions (one thread per batch entry).
__global__ void k_refresh_status(DevState S, const int *__restrict__ list, const int *__restrict__ status) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= batch_count(S, list)) return;
    const int w = list ? list[b] : b;
    if (status[2 * b] | status[2 * b + 1]) {
        atomicOr(&S.flags[w], KDSL_FLAG_SINGULAR_DEV);
        atomicAdd(&S.cnt[3], 1);
    } else {
        atomicAnd(&S.flags[w], ~KDSL_FLAG_SINGULAR_DEV);
        S.n_refresh[w] += 1ull;
        S.fcnt[2 * w] = 0;                                    // W0 is now exact: drop the pending updates
        S.fcnt[2 * w + 1] = 0;
    }
}

// ---- GEMM, variant 0: shared-memory tiled FP64 FMA.  W[w] (ns x N) = U (ns x N) * X_b (N x N) ----
// grid (tiles_m * tiles_n, nw, 2).  A walker with a singular species keeps its old W (both species).
template <int BM, int BN, int BK>
__global__ void __launch_bounds__(256)
k_gemm_W_simt(DevState S, const int *__restrict__ list, const double *__restrict__ X_up,
              const double *__restrict__ X_dn, const int *__restrict__ status) {
    __shared__ double As[BK][BM];
    __shared__ double Bs[BK][BN + 1];
    const int b = blockIdx.y, spin = blockIdx.z;
    if (b >= batch_count(S, list)) return;
    if (status[2 * b] | status[2 * b + 1]) return;
    const int w = list ? list[b] : b;
    const int ns = S.ns, N = spin ? S.n_dn : S.n_up;
    const int tiles_m = (ns + BM - 1) / BM;
    const int tm = blockIdx.x % tiles_m, tn = blockIdx.x / tiles_m;
    if (tn * BN >= N) return;
    const double *U = spin ? S.U_dn : S.U_up;
    const double *X = (spin ? X_dn : X_up) + (size_t)b * N * N;
    double *W = (spin ? S.W_dn : S.W_up) + (size_t)w * ns * N;
    const int m0 = tm * BM, n0 = tn * BN;
    const int tid = threadIdx.x;
    constexpr int TM = BM / 16, TN = BN / 16;
    const int tx = tid & 15, ty = tid >> 4;
    double acc[TM][TN];
#pragma unroll
    for (int a = 0; a < TM; a++)
#pragma unroll
        for (int c = 0; c < TN; c++) acc[a][c] = 0.0;
    for (int k0 = 0; k0 < N; k0 += BK) {
        for (int e = tid; e < BK * BM; e += 256) {
            const int kk = e / BM, mm = e - kk * BM;
            const int m = m0 + mm, k = k0 + kk;
            As[kk][mm] = (m < ns && k < N) ? U[(size_t)k * ns + m] : 0.0;
        }
        for (int e = tid; e < BK * BN; e += 256) {
            const int nn = e / BK, kk = e - nn * BK;
            const int n = n0 + nn, k = k0 + kk;
            Bs[kk][nn] = (n < N && k < N) ? X[(size_t)n * N + k] : 0.0;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < BK; kk++) {
            double av[TM], bv[TN];
#pragma unroll
            for (int a = 0; a < TM; a++) av[a] = As[kk][tx + 16 * a];
#pragma unroll
            for (int c = 0; c < TN; c++) bv[c] = Bs[kk][ty + 16 * c];
#pragma unroll
            for (int a = 0; a < TM; a++)
#pragma unroll
                for (int c = 0; c < TN; c++) acc[a][c] = fma(av[a], bv[c], acc[a][c]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int c = 0; c < TN; c++) {
        const int n = n0 + ty + 16 * c;
        if (n >= N) continue;
#pragma unroll
        for (int a = 0; a < TM; a++) {
            const int m = m0 + tx + 16 * a;
            if (m < ns) W[(size_t)n * ns + m] = acc[a][c];
        }
    }
}
