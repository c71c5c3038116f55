// kdsl_measure.cuh -- Heisenberg local energy O_L per walker (reference getOL / getxprime /
// spinInteraction! / SzInteraction! / Sz, src/Hamiltonian.jl:762-778, 711-720, 636-672, 593-604,
// 531-565) and the full Z recount (src/MonteCarlo.jl:460-474).  One warp per walker, lanes
// stride over the bond list, warp-shuffle reduction.
#pragma once
#include "kdsl_common.cuh"

#define KDSL_FLAG_BAD_SITE_DEV 4

// OL = sum_bonds Sz_i Sz_j + sum_{antiparallel bonds} (-1/2) W_up[K_up, l_up] W_dn[K_dn, l_dn]
//   j up & i down:  (K_up, l_up, K_dn, l_dn) = (i, kup[j], j, kdn[i])     :649-657
//   i up & j down:  (j, kup[i], i, kdn[j])                               :661-669
// accumulate != 0: also feed the per-walker accumulators (Carlo.measure!, src/MonteCarlo.jl:628-634)
__global__ void __launch_bounds__(256)
k_measure(DevState S, double *__restrict__ ol_out, int accumulate) {
    const int w = (int)((blockIdx.x * (unsigned)blockDim.x + threadIdx.x) >> 5);
    const int lane = threadIdx.x & 31;
    if (w >= S.nw) return;
    const int ns = S.ns;
    const int *kup = S.kup + (size_t)w * ns;
    const int *kdn = S.kdn + (size_t)w * ns;
    const double *Wu = S.W_up + (size_t)w * ns * S.n_up;
    const double *Wd = S.W_dn + (size_t)w * ns * S.n_dn;
    double flips = 0.0;
    int diag4 = 0;     // sum of 4*Sz_i*Sz_j (each +-1)
    int bad = 0;
    for (int b = lane; b < S.n_bonds; b += 32) {
        const int i = S.bi[b], j = S.bj[b];
        const int iu = kup[i], ju = kup[j], id = kdn[i], jd = kdn[j];
        if (ju != 0 && id != 0)
            flips += -0.5 * Wu[(size_t)(ju - 1) * ns + i] * Wd[(size_t)(id - 1) * ns + j];
        if (iu != 0 && jd != 0)
            flips += -0.5 * Wu[(size_t)(iu - 1) * ns + j] * Wd[(size_t)(jd - 1) * ns + i];
        // Sz(i) = +1/2 (up only), -1/2 (down only), else ArgumentError (:531-565)
        const int oi = (iu != 0) + (id != 0), oj = (ju != 0) + (jd != 0);
        if (oi != 1 || oj != 1) bad = 1;
        const int si = iu != 0 ? 1 : -1, sj = ju != 0 ? 1 : -1;
        diag4 += si * sj;
    }
    flips = warp_sum_f64(flips);
    diag4 = warp_sum_int(diag4);
    bad = warp_sum_int(bad);
    if (lane == 0) {
        const double OL = flips + 0.25 * (double)diag4;
        if (bad) atomicOr(&S.flags[w], KDSL_FLAG_BAD_SITE_DEV);
        if (ol_out) ol_out[w] = OL;
        if (accumulate && !(S.flags[w] & 1)) {                  // (a walker frozen by a singular re-evaluation gives no sample)
            S.ol_last[w] = OL;
            S.ol_sum[w] += OL;
            S.ol_sq[w] += OL * OL;
            S.ol_n[w] += 1ull;
        }
    }
}

// Extra observables of one walker (whole warp), taken together with an :OL sample:
//   S(q) = |sum_i exp(i q r_i) Sz_i|^2 / ns   (longitudinal spin structure factor, from kappa alone), and the weights that
//   turn chain averages into |psi|^2 averages: the chain's stationary law is |psi|^2 / Z_mu, so
//   <O>_psi = <O Z_mu> / <Z_mu>.  `up(i)` tells whether site i carries an up spin.
template <typename UpFn>
__device__ __forceinline__ void extra_observables_warp(const DevState &S, int w, int lane, int zmu, double OL, UpFn up) {
    double *obs = S.obs_w + (size_t)w * (4 + 2 * S.nq);
    const double z = (double)zmu;
    for (int q = 0; q < S.nq; q++) {
        const double *c = S.q_cos + (size_t)q * S.ns, *s = S.q_sin + (size_t)q * S.ns;
        double re = 0.0, im = 0.0;
        for (int i = lane; i < S.ns; i += 32) {
            const double sz = up(i) ? 0.5 : -0.5;
            re = fma(__ldg(c + i), sz, re);
            im = fma(__ldg(s + i), sz, im);
        }
        re = warp_sum_f64(re);
        im = warp_sum_f64(im);
        if (lane == 0) {
            const double sq = (re * re + im * im) / (double)S.ns;
            obs[4 + q] += sq;
            obs[4 + S.nq + q] += sq * z;
        }
    }
    if (lane == 0) {
        obs[0] += 1.0;
        obs[1] += z;
        obs[2] += OL * z;
    }
}

// ... for every walker, right after the cadence measurement wrote ol_last (the lock-step paths); one warp per walker
__global__ void __launch_bounds__(256) k_measure_extra(DevState S) {
    const int w = (int)((blockIdx.x * (unsigned)blockDim.x + threadIdx.x) >> 5);
    const int lane = threadIdx.x & 31;
    if (w >= S.nw) return;
    if (S.flags[w] & 1) return;                                 // frozen after a singular re-evaluation: no sample
    const int *kup = S.kup + (size_t)w * S.ns;
    extra_observables_warp(S, w, lane, S.zmu[w], S.ol_last[w], [&](int i) { return kup[i] != 0; });
}

// sums of obs_w over the walkers: one thread per observable index
__global__ void k_reduce_obs(DevState S, double *__restrict__ out) {
    const int n = 4 + 2 * S.nq;
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= n) return;
    double a = 0.0;
    for (int w = 0; w < S.nw; w++) a += S.obs_w[(size_t)w * n + x];
    out[x] = a;
}

// Z(nn, kappa_up, kappa_down): full recount.  store != 0 writes it to S.zmu (after set_config),
// otherwise to out (verification of the incremental value).
__global__ void __launch_bounds__(256) k_count_Z(DevState S, int *__restrict__ out, int store) {
    const int w = (int)((blockIdx.x * (unsigned)blockDim.x + threadIdx.x) >> 5);
    const int lane = threadIdx.x & 31;
    if (w >= S.nw) return;
    const int *kup = S.kup + (size_t)w * S.ns;
    const int *kdn = S.kdn + (size_t)w * S.ns;
    int c = 0;
    for (int b = lane; b < S.n_bonds; b += 32) {
        const int s1 = S.bi[b], s2 = S.bj[b];
        if (kup[s1] != 0 && kdn[s2] != 0) c += 1;
        else if (kup[s2] != 0 && kdn[s1] != 0) c += 1;
    }
    c = warp_sum_int(c);
    if (lane == 0) {
        if (store) S.zmu[w] = c;
        if (out) out[w] = c;
    }
}

// Device-side reduction of the per-walker accumulators into out[8] (KDSL_ACC_* order; slot 0 = the host's
// walker-sweep count, slot 7 = the singular counter cnt[3]).  Single block; deterministic order.  The vector is
// complete on the device so that it can go straight into the NCCL all-reduce of kdsl_accumulators_allreduce.
__global__ void __launch_bounds__(1024) k_reduce_acc(DevState S, double *__restrict__ out, double walker_sweeps) {
    __shared__ double sh[32][6];
    double a[6] = {0, 0, 0, 0, 0, 0};
    for (int w = threadIdx.x; w < S.nw; w += blockDim.x) {
        a[0] += (double)S.n_acc[w];
        a[1] += S.ol_sum[w];
        a[2] += S.ol_sq[w];
        a[3] += (double)S.ol_n[w];
        a[4] += (double)S.n_reach[w];
        a[5] += (double)S.n_refresh[w];
    }
#pragma unroll
    for (int k = 0; k < 6; k++) a[k] = warp_sum_f64(a[k]);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0)
        for (int k = 0; k < 6; k++) sh[warp][k] = a[k];
    __syncthreads();
    if (threadIdx.x < 6) {
        double s = 0.0;
        for (int q = 0; q < (int)(blockDim.x >> 5); q++) s += sh[q][threadIdx.x];
        out[1 + threadIdx.x] = s;
    }
    if (threadIdx.x == 32) {
        out[0] = walker_sweeps;
        out[7] = (double)S.cnt[3];
    }
}
