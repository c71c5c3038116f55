// kdsl_common.cuh -- shared device-side state and helpers of libkdsl (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define KDSL_WARP 32
#define KDSL_KALLOC 32        /* capacity of the pending-update buffers (warp-wide Woodbury state) */

// Device view of one engine (passed by value to every kernel).
// Layout in HBM (nw = walkers on this GPU, all arrays walker-major):
//   kup, kdn   int32 [nw][ns]      kappa_up / kappa_down (0 = empty, else 1-based label)
//   rng        u64   [nw][4]       Xoshiro256++ state
//   zmu        int32 [nw]          incrementally maintained Z_mu
//   W_up       f64   [nw][n_up][ns]  column-major per walker (column l contiguous)
//   W_dn       f64   [nw][n_dn][ns]
//   col_*      f64   [nw][ns]      staged W[:, l]           (reference's col_cache)
//   trow_*     f64   [nw][N]       staged alpha*(W[K, :] - e_l) (reference's row_cache, pre-scaled)
struct DevState {
    int ns, n_up, n_dn, n_bonds, nw, n_occ;
    const int *bi, *bj;          // bond endpoints, 0-based, reference order
    const int *adj_off, *adj_nbr;  // CSR adjacency of the bond graph
    const double *U_up, *U_dn;   // ns x N column-major
    int *kup, *kdn;
    unsigned long long *rng;
    int *zmu;
    double *W_up, *W_dn;
    double *col_up, *col_dn, *trow_up, *trow_dn;
    int *acc_list;               // [2][nw] accepted-walker lists (double buffered by sweep parity)
    int *cnt;                    // [0],[1] accepted counts per parity; [2] refresh count; [3] singular count
    int *ref_list;               // [nw] walkers to re-evaluate this sweep
    int *flags;                  // [nw] KDSL_FLAG_* bits
    unsigned long long *n_acc;   // [nw] accepted moves
    unsigned long long *n_reach; // [nw] sweeps that reached the refresh block
    unsigned long long *n_refresh;  // [nw]
    double *ol_sum, *ol_sq, *ol_last;  // [nw]
    unsigned long long *ol_n;    // [nw]
    unsigned long long *upd_moves;  // [0] accepted moves streamed by the rank-1 launches, [1] walkers flushed (delayed mode)
    // delayed (rank-k) updates: W = W0 + sum_{m<fcnt} A_m (x) B_m
    double *facA_up, *facA_dn;   // [nw][kmax][ns]
    double *facB_up, *facB_dn;   // [nw][kmax][N]
    int *fcnt;                   // [nw][2] pending update count per walker and species
    double *wbT;                 // [nw][2][kmax*kmax] Woodbury state: T = inv(W0[K_set, L])
    int *wbK, *wbL;              // [nw][2][kmax] displaced particles: label wbL now sits on site wbK
    int *flush_list;             // [nw] walkers that reached kth pending factors (count in cnt[4])
    int *listed;                 // [nw] 1 while the walker sits in flush_list (a walker is listed at most once)
    int kmax, kth;
    // optional extra observables taken with every :OL sample (kdsl_set_observables):
    // obs_w [nw][4 + 2 nq] = n, sum Z_mu, sum OL Z_mu, -, sum S(q) [nq], sum S(q) Z_mu [nq]
    int obs_on, nq;
    const double *q_cos, *q_sin;     // [nq][ns] cos(q . r_i), sin(q . r_i)
    double *obs_w;
};

__device__ __forceinline__ unsigned long long rotl64(unsigned long long x, int k) {
    return (x << k) | (x >> (64 - k));
}

// Julia Random.Xoshiro = xoshiro256++
struct Xoshiro {
    unsigned long long s0, s1, s2, s3;
    __device__ __forceinline__ unsigned long long next() {
        unsigned long long res = rotl64(s0 + s3, 23) + s0;
        unsigned long long t = s1 << 17;
        s2 ^= s0; s3 ^= s1; s1 ^= s2; s0 ^= s3;
        s2 ^= t;
        s3 = rotl64(s3, 45);
        return res;
    }
    // rand(rng)::Float64
    __device__ __forceinline__ double rand_f64() { return (double)(next() >> 11) * 0x1.0p-53; }
    // rand(rng, 1:n) (SamplerRangeNDL); 1-based
    __device__ __forceinline__ long long rand_index(unsigned long long n) {
        unsigned long long x = next();
        unsigned long long hi = __umul64hi(x, n), lo = x * n;
        if (lo < n) {
            unsigned long long t = (0ull - n) % n;
            while (lo < t) {
                x = next();
                hi = __umul64hi(x, n);
                lo = x * n;
            }
        }
        return (long long)hi + 1;
    }
};

__device__ __forceinline__ int warp_sum_int(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum_f64(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
