// kdsl_update.cuh -- the accepted-move Sherman-Morrison rank-1 update of W
// (reference update_W!, src/MonteCarlo.jl:279-292:  W += alpha * col * row^T, BLAS geru!),
// batched over the compacted list of accepted walkers.  This is the HBM-bound kernel the
// roofline is quoted on: 16*ns^2 algorithmic bytes per accepted move (both species, read+write).
//
// Work decomposition: a walker's two W matrices are contiguous column-major slabs of ns*N
// doubles; a work item is a slab of CH consecutive columns (CH*ns*8 contiguous bytes).  A
// persistent grid (multiple of the SM count) strides over the items of all accepted walkers;
// the item count is read from device memory (the accepted count written by k_propose), so no
// host synchronisation sits between proposal and update.
#pragma once
#include "kdsl_common.cuh"

__device__ __forceinline__ double2 ldg_stream(const double2 *p) {
    double2 v;
    asm volatile("ld.global.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    return v;
}
__device__ __forceinline__ void stg_stream(double2 *p, double2 v) {
    asm volatile("st.global.L1::no_allocate.v2.f64 [%0], {%1, %2};" ::"l"(p), "d"(v.x), "d"(v.y) : "memory");
}

// Variant 0: 128-bit LDG/STG streaming with the staged column and scaled row in shared memory.
// Dynamic smem: (ns + CH) doubles.
template <int THREADS, int UNROLL>
__global__ void __launch_bounds__(THREADS)
k_update_ldg(DevState S, int parity, int tiles_up, int tiles_dn, int CH) {
    extern __shared__ double smem[];
    double *s_col = smem;
    double *s_trow = smem + S.ns;
    const int tid = threadIdx.x;
    const int ns = S.ns, half = ns >> 1;
    const int n_acc = S.cnt[parity];
    const int tpw = tiles_up + tiles_dn;
    const long long total = (long long)n_acc * tpw;
    if (blockIdx.x == 0 && tid == 0) {
        S.cnt[parity ^ 1] = 0;   // arm the other parity for the next sweep
        *S.upd_moves += (unsigned long long)n_acc;
    }

    for (long long t = blockIdx.x; t < total; t += gridDim.x) {
        const int a = (int)(t / tpw);
        const int rem = (int)(t - (long long)a * tpw);
        const int spin = rem >= tiles_up;
        const int tile = spin ? rem - tiles_up : rem;
        const int w = S.acc_list[(size_t)parity * S.nw + a];
        const int N = spin ? S.n_dn : S.n_up;
        double *W = (spin ? S.W_dn : S.W_up) + (size_t)w * ns * N;
        const double *col = (spin ? S.col_dn : S.col_up) + (size_t)w * ns;
        const double *trow = (spin ? S.trow_dn : S.trow_up) + (size_t)w * N;
        const int j0 = tile * CH;
        const int jn = min(CH, N - j0);

        __syncthreads();
        for (int x = tid; x < half; x += THREADS)
            reinterpret_cast<double2 *>(s_col)[x] = reinterpret_cast<const double2 *>(col)[x];
        for (int x = tid; x < jn; x += THREADS) s_trow[x] = trow[j0 + x];
        __syncthreads();

        double2 *base = reinterpret_cast<double2 *>(W + (size_t)j0 * ns);
        const int n2 = jn * half;
        // the slab of jn columns is contiguous: UNROLL 16-byte elements per thread in flight, all loads of a trip
        // issued before its first store
        for (int q = tid; q < n2; q += THREADS * UNROLL) {
            double2 v[UNROLL];
#pragma unroll
            for (int u = 0; u < UNROLL; u++) {
                const int idx = q + u * THREADS;
                if (idx < n2) v[u] = ldg_stream(base + idx);
            }
#pragma unroll
            for (int u = 0; u < UNROLL; u++) {
                const int idx = q + u * THREADS;
                if (idx < n2) {
                    const int jq = idx / half, pq = idx - jq * half;
                    const double2 c = reinterpret_cast<const double2 *>(s_col)[pq];
                    const double tj = s_trow[jq];
                    v[u].x = fma(c.x, tj, v[u].x);            // A[i,j] += x[i] * (alpha*y[j])
                    v[u].y = fma(c.y, tj, v[u].y);
                    stg_stream(base + idx, v[u]);
                }
            }
        }
    }
}
