// kdsl_delayed.cuh -- delayed (rank-k) Sherman-Morrison updates.
//
// The reference applies every accepted move to W at once (update_W!, src/MonteCarlo.jl:279-292:
// W += alpha * col * row^T), i.e. it re-reads and re-writes the whole matrix per accepted move.
// Here each walker keeps   W = W0 + sum_{m < cnt} A_m (x) B_m   with
//     A_m = W[:, l_m]                      (the reference's col_cache at move m)
//     B_m = alpha_m * (W[K_m, :] - e_l)    (the reference's row_cache, pre-scaled)
// which is the same rank-1 formula, merely not yet added into W0.  Matrix entries needed by the
// proposal (determinant ratio) and the measurement are evaluated as W0[K,l] + sum_m A_m[K] B_m[l];
// once a walker has accumulated KTH factors its W0 is brought up to date by ONE streaming pass
// (k_flush: W0 += A B^T on the FP64 tensor pipe), cutting the HBM traffic of the update by ~KTH.
// The periodic from-scratch re-evaluation (reevaluateW!) simply discards the pending factors.
#pragma once
#include "kdsl_common.cuh"
#include "kdsl_propose.cuh"
#include "kdsl_refresh_fast.cuh"

// W[K, l] of walker w / species spin, evaluated cooperatively by a full warp (K, l 0-based).
__device__ __forceinline__ double w_entry_warp(const DevState &S, int w, int spin, int K, int l,
                                               int cnt, int lane) {
    const int ns = S.ns, N = spin ? S.n_dn : S.n_up;
    const double *A = (spin ? S.facA_dn : S.facA_up) + (size_t)w * S.kmax * ns;
    const double *B = (spin ? S.facB_dn : S.facB_up) + (size_t)w * S.kmax * N;
    double part = 0.0;
    for (int m = lane; m < cnt; m += 32) part = fma(A[(size_t)m * ns + K], B[(size_t)m * N + l], part);
    part = warp_sum_f64(part);
    const double *W0 = (spin ? S.W_dn : S.W_up) + (size_t)w * ns * N;
    return W0[(size_t)l * ns + K] + part;
}

// same, by a single thread (measurement)
__device__ __forceinline__ double w_entry_thread(const DevState &S, const double *W0, const double *A,
                                                 const double *B, int ns, int N, int K, int l, int cnt) {
    double part = 0.0;
    for (int m = 0; m < cnt; m++) part = fma(A[(size_t)m * ns + K], B[(size_t)m * N + l], part);
    return W0[(size_t)l * ns + K] + part;
}

// append factor number `cnt` for one species with all threads of the CTA:
//   A_new = W[:, l],  B_new = alpha (W[K, :] - e_l),  alpha = -1 / W[K, l]   (src/MonteCarlo.jl:286-290)
// s_bl[m] = B_m[l], s_ak[m] = A_m[K] of the pending factors are staged in shared memory by the caller.
// All global loads of an element are issued before its FMA chain (KMX-way memory parallelism).
template <int NT, int KMX>
__device__ __forceinline__ void build_factor_cta(const DevState &S, int w, int spin, int K, int l, int cnt,
                                                 const double *s_bl, const double *s_ak, int tid) {
    const int ns = S.ns, N = spin ? S.n_dn : S.n_up;
    double *__restrict__ A = (spin ? S.facA_dn : S.facA_up) + (size_t)w * S.kmax * ns;
    double *__restrict__ B = (spin ? S.facB_dn : S.facB_up) + (size_t)w * S.kmax * N;
    const double *__restrict__ W0 = (spin ? S.W_dn : S.W_up) + (size_t)w * ns * N;
    double *__restrict__ An = A + (size_t)cnt * ns;
    double *__restrict__ Bn = B + (size_t)cnt * N;
    const double *__restrict__ W0col = W0 + (size_t)l * ns;
    double piv = W0col[K];                                     // pivot W[K, l], same accumulation order
    for (int m = 0; m < cnt; m++) piv = fma(s_ak[m], s_bl[m], piv);
    const double alpha = -1.0 / piv;
    for (int i = tid; i < ns; i += NT) {                       // column l of the current W
        double av[KMX];
        double acc = W0col[i];
#pragma unroll
        for (int m = 0; m < KMX; m++) av[m] = m < cnt ? A[(size_t)m * ns + i] : 0.0;
#pragma unroll
        for (int m = 0; m < KMX; m++) if (m < cnt) acc = fma(av[m], s_bl[m], acc);
        An[i] = acc;
    }
    for (int j = tid; j < N; j += NT) {                        // row K of the current W
        double bv[KMX];
        double acc = W0[(size_t)j * ns + K];
#pragma unroll
        for (int m = 0; m < KMX; m++) bv[m] = m < cnt ? B[(size_t)m * N + j] : 0.0;
#pragma unroll
        for (int m = 0; m < KMX; m++) if (m < cnt) acc = fma(s_ak[m], bv[m], acc);
        if (j == l) acc -= 1.0;
        Bn[j] = alpha * acc;
    }
}

// Carlo.sweep! proposal (reference src/MonteCarlo.jl:538-607) with delayed W updates, split in two
// short kernels so that every walker's latency chain runs concurrently in a single wave:
//   k_decide        one warp per walker: the Metropolis decision with the reference's exact predicate and
//                   RNG consumption order, the integer state changes (kappa, Z_mu, counters) and an
//                   "accepted" record (walker, K_up, l_up, K_dn, l_dn, pending count);
//   k_build_factors one CTA per accepted (walker, species): appends the factor pair of the move.
template <bool REPLAY>
__global__ void __launch_bounds__(256, 4)
k_decide(DevState S, int parity, int gate_refresh, const double *__restrict__ rp_r,
         const int *__restrict__ rp_bond, const int *__restrict__ rp_pick) {
    const int w = (int)((blockIdx.x * (unsigned)blockDim.x + threadIdx.x) >> 5);
    const int lane = threadIdx.x & 31;
    if (w >= S.nw) return;
    const int ns = S.ns;
    int *kup = S.kup + (size_t)w * ns;
    int *kdn = S.kdn + (size_t)w * ns;
    const int cnt = S.fcnt[2 * w];
    const int zmu = S.zmu[w];
    Xoshiro g;
    if (!REPLAY) {
        const unsigned long long *st = S.rng + (size_t)w * 4;
        g.s0 = st[0]; g.s1 = st[1]; g.s2 = st[2]; g.s3 = st[3];
    }
    const double r = REPLAY ? rp_r[w] : g.rand_f64();               // :546
    const double zr = (double)zmu / (double)S.n_bonds;
    bool accepted = false, reached = false;
    int i = 0, site = 0, flag = 0, l_up = 0, l_dn = 0, K_up = 0, K_dn = 0;
    int ku_i = 0, ku_s = 0, kd_i = 0, kd_s = 0;
    if (!(r > zr)) {                                                // :547-550
        long long b = REPLAY ? (long long)rp_bond[w] : g.rand_index((unsigned long long)S.n_bonds);  // :552
        if (b < 1) b = 1;
        if (b > S.n_bonds) b = S.n_bonds;
        i = S.bi[b - 1];
        site = S.bj[b - 1];
        ku_i = kup[i]; ku_s = kup[site]; kd_i = kdn[i]; kd_s = kdn[site];
        const bool f1 = ku_i != 0 && kd_s != 0;                     // :558-561
        const bool f2 = ku_s != 0 && kd_i != 0;
        if (f1 || f2) {
            const int nm = (int)f1 + (int)f2;
            long long pick;                                         // :569
            if (REPLAY) pick = rp_pick ? (long long)rp_pick[w] : 1;
            else pick = g.rand_index((unsigned long long)nm);
            flag = (f1 && f2) ? (pick == 1 ? 1 : 2) : (f1 ? 1 : 2);
            l_up = flag == 1 ? ku_i : ku_s;                         // :572-573
            l_dn = flag == 1 ? kd_s : kd_i;
            K_up = flag == 1 ? site : i;
            K_dn = flag == 1 ? i : site;
            const double wu = w_entry_warp(S, w, 0, K_up, l_up - 1, cnt, lane);   // :576-580
            const double wd = w_entry_warp(S, w, 1, K_dn, l_dn - 1, cnt, lane);
            const double ratio = wu * wd;
            const double p = ratio * ratio;                         // abs2(ratio)
            if (p >= 1.0 && r < zr) accepted = true;                // :582-587
            else if (p < 1.0 && r < zr * p) accepted = true;
            if (!(p == p) || p > 1.79e308) {
                if (lane == 0) atomicOr(&S.flags[w], KDSL_FLAG_NONFINITE_DEV);
            }
            reached = true;
        }
    }
    if (accepted) {
        const int ui_o = ku_i != 0, di_o = kd_i != 0, us_o = ku_s != 0, ds_o = kd_s != 0;
        const int ui_n = flag == 1 ? 0 : 1, di_n = flag == 1 ? 1 : 0;
        const int us_n = flag == 1 ? 1 : 0, ds_n = flag == 1 ? 0 : 1;
        int delta = 0;
        for (int q = S.adj_off[i] + lane; q < S.adj_off[i + 1]; q += 32) {
            const int n = S.adj_nbr[q];
            if (n == site) continue;
            const int un = kup[n] != 0, dn = kdn[n] != 0;
            delta += bond_is_anti(ui_n, di_n, un, dn) - bond_is_anti(ui_o, di_o, un, dn);
        }
        for (int q = S.adj_off[site] + lane; q < S.adj_off[site + 1]; q += 32) {
            const int n = S.adj_nbr[q];
            if (n == i) continue;
            const int un = kup[n] != 0, dn = kdn[n] != 0;
            delta += bond_is_anti(us_n, ds_n, un, dn) - bond_is_anti(us_o, ds_o, un, dn);
        }
        if (lane == 0)
            delta += bond_is_anti(ui_n, di_n, us_n, ds_n) - bond_is_anti(ui_o, di_o, us_o, ds_o);
        delta = warp_sum_int(delta);
        if (lane == 0) {
            S.zmu[w] = zmu + delta;
            if (flag == 1) {                                        // :502-503
                kup[i] = 0; kup[site] = l_up;
                kdn[i] = l_dn; kdn[site] = 0;
            } else {                                                // :508-509
                kup[i] = l_up; kup[site] = 0;
                kdn[i] = 0; kdn[site] = l_dn;
            }
            S.n_acc[w] += 1ull;
            if (!gate_refresh) {                                    // (a walker re-evaluated this sweep needs no factor)
                const int slot = atomicAdd(&S.cnt[parity], 1);
                int *rec = S.acc_list + ((size_t)parity * S.nw + slot) * 6;
                rec[0] = w; rec[1] = K_up; rec[2] = l_up - 1; rec[3] = K_dn; rec[4] = l_dn - 1; rec[5] = cnt;
                S.fcnt[2 * w] = cnt + 1;
                S.fcnt[2 * w + 1] = cnt + 1;
                if (cnt + 1 == S.kth) {                             // due for a flush (listed exactly once)
                    const int fs = atomicAdd(&S.cnt[4], 1);
                    S.flush_list[fs] = w;
                }
            }
        }
    }
    if (lane == 0) {
        if (reached) {
            S.n_reach[w] += 1ull;
            if (gate_refresh) {                                     // :595
                const int slot = atomicAdd(&S.cnt[2], 1);
                S.ref_list[slot] = w;
            }
        }
        if (!REPLAY) {
            unsigned long long *st = S.rng + (size_t)w * 4;
            st[0] = g.s0; st[1] = g.s1; st[2] = g.s2; st[3] = g.s3;
        }
    }
}

// grid-stride over (accepted record, species); 128 threads per item, 8 CTAs per SM so that all
// (~2 x 0.13 x n_walkers) items of a sweep are resident in one wave
template <int KMX>
__global__ void __launch_bounds__(128, 8)
k_build_factors(DevState S, int parity) {
    constexpr int NT = 128;
    __shared__ double s_bl[32], s_ak[32];
    const int n_items = 2 * S.cnt[parity];
    const int tid = threadIdx.x;
    if (blockIdx.x == 0 && tid == 0) S.cnt[parity ^ 1] = 0;      // arm the other list for the next sweep
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int *rec = S.acc_list + ((size_t)parity * S.nw + (item >> 1)) * 6;
        const int spin = item & 1;
        const int w = rec[0], K = rec[1 + 2 * spin], l = rec[2 + 2 * spin], cnt = rec[5];
        const int N = spin ? S.n_dn : S.n_up;
        __syncthreads();
        if (tid < cnt) {
            s_bl[tid] = (spin ? S.facB_dn : S.facB_up)[((size_t)w * S.kmax + tid) * N + l];
            s_ak[tid] = (spin ? S.facA_dn : S.facA_up)[((size_t)w * S.kmax + tid) * S.ns + K];
        }
        __syncthreads();
        build_factor_cta<NT, KMX>(S, w, spin, K, l, cnt, s_bl, s_ak, tid);
    }
}

// W0 += sum_m A_m (x) B_m for the listed walkers: the HBM-bound pass of the delayed update.
// Work item = (list entry, species, block of 216 rows); 288 threads = 9 warps, each warp owns a strip
// of 24 rows (3 DMMA m-tiles) and walks over all column tiles.  KPAD = padded factor count.
// FROM_W0: the left operand A_m is column l_m (wbL) of W0 itself (Woodbury form) instead of facA[m]; a warp loads
// its rows of those columns into registers before it overwrites them, and no other warp touches these rows.
template <int KPAD, bool FROM_W0>
__global__ void __launch_bounds__(288)
k_flush(DevState S, const int *__restrict__ list, const int *__restrict__ count_ptr, int count_fixed) {
    extern __shared__ double fsm[];                      // B operand, frag-major [Ncols8 x KPAD]
    const int count = count_ptr ? *count_ptr : count_fixed;
    const int ns = S.ns;
    const int nrb = (ns + 215) / 216;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int gr = lane >> 2, tg = lane & 3;
    const long long total = (long long)count * 2 * nrb;
    for (long long item = blockIdx.x; item < total; item += gridDim.x) {
        const int e = (int)(item / (2 * nrb));
        const int rem = (int)(item - (long long)e * 2 * nrb);
        const int spin = rem / nrb, rb = rem - spin * nrb;
        const int w = list ? list[e] : e;
        const int cnt = S.fcnt[2 * w + spin];
        if (cnt == 0) continue;                           // uniform over the block
        const int N = spin ? S.n_dn : S.n_up;
        const double *A = (spin ? S.facA_dn : S.facA_up) + (size_t)w * S.kmax * ns;
        const double *B = (spin ? S.facB_dn : S.facB_up) + (size_t)w * S.kmax * N;
        double *W0 = (spin ? S.W_dn : S.W_up) + (size_t)w * ns * N;
        const int ctiles = (N + 7) >> 3;
        __syncthreads();
        for (int x = tid; x < ctiles * 8 * KPAD; x += 288) {
            const int m = x / (ctiles * 8), j = x - m * (ctiles * 8);      // coalesced along j
            fsm[frag_idx(j, m, KPAD)] = (m < cnt && j < N) ? B[(size_t)m * N + j] : 0.0;
        }
        const int r0 = rb * 216 + warp * 24;
        double af[3][KPAD / 4];
#pragma unroll
        for (int t = 0; t < 3; t++) {
            const int row = r0 + 8 * t + gr;
#pragma unroll
            for (int s = 0; s < KPAD / 4; s++) {
                const int m = 4 * s + tg;
                if (FROM_W0)
                    af[t][s] = (row < ns && m < cnt) ? W0[(size_t)S.wbL[((size_t)w * 2 + spin) * S.kmax + m] * ns + row] : 0.0;
                else
                    af[t][s] = (row < ns && m < cnt) ? A[(size_t)m * ns + row] : 0.0;
            }
        }
        __syncthreads();
        if (r0 >= ns) continue;
        bool rv[3];
#pragma unroll
        for (int t = 0; t < 3; t++) rv[t] = r0 + 8 * t + gr < ns;
        double c0[3][2], c1[3][2];
        auto load_tile = [&](int ct, double (&cc)[3][2]) {
            const int col = (ct << 3) + 2 * tg;
            const double *p0 = W0 + (size_t)col * ns + r0 + gr;
#pragma unroll
            for (int t = 0; t < 3; t++) {
                cc[t][0] = (rv[t] && col < N) ? p0[8 * t] : 0.0;
                cc[t][1] = (rv[t] && col + 1 < N) ? p0[8 * t + ns] : 0.0;
            }
        };
        auto compute_store = [&](int ct, double (&cc)[3][2]) {
#pragma unroll
            for (int s = 0; s < KPAD / 4; s++) {
                const double bf = fsm[(((ct * (KPAD >> 2)) + s) << 5) + lane];
#pragma unroll
                for (int t = 0; t < 3; t++) dmma_8x8x4(cc[t][0], cc[t][1], af[t][s], bf);
            }
            const int col = (ct << 3) + 2 * tg;
            double *p0 = W0 + (size_t)col * ns + r0 + gr;
#pragma unroll
            for (int t = 0; t < 3; t++) {
                if (rv[t] && col < N) p0[8 * t] = cc[t][0];
                if (rv[t] && col + 1 < N) p0[8 * t + ns] = cc[t][1];
            }
        };
        load_tile(0, c0);
        for (int ct = 0; ct < ctiles; ct += 2) {               // two-deep register pipeline
            if (ct + 1 < ctiles) load_tile(ct + 1, c1);
            compute_store(ct, c0);
            if (ct + 2 < ctiles) load_tile(ct + 2, c0);
            if (ct + 1 < ctiles) compute_store(ct + 1, c1);
        }
    }
}

// after k_flush: the listed walkers have no pending factors any more; re-arm the list
__global__ void k_flush_done(DevState S, const int *__restrict__ list, int *count_ptr, int count_fixed) {
    const int count = count_ptr ? *count_ptr : count_fixed;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < count; e += gridDim.x * blockDim.x) {
        const int w = list ? list[e] : e;
        S.fcnt[2 * w] = 0;
        S.fcnt[2 * w + 1] = 0;
    }
    if (count_ptr && blockIdx.x == 0 && threadIdx.x == 0) {
        S.upd_moves[1] += (unsigned long long)count;      // walkers flushed
    }
}
__global__ void k_zero_int(int *p) { *p = 0; }

// O_L with delayed factors (same formula as k_measure)
__global__ void __launch_bounds__(256)
k_measure_delayed(DevState S, double *__restrict__ ol_out, int accumulate) {
    const int w = (int)((blockIdx.x * (unsigned)blockDim.x + threadIdx.x) >> 5);
    const int lane = threadIdx.x & 31;
    if (w >= S.nw) return;
    const int ns = S.ns, cnt = S.fcnt[2 * w];
    const int *kup = S.kup + (size_t)w * ns;
    const int *kdn = S.kdn + (size_t)w * ns;
    const double *Wu = S.W_up + (size_t)w * ns * S.n_up, *Wd = S.W_dn + (size_t)w * ns * S.n_dn;
    const double *Au = S.facA_up + (size_t)w * S.kmax * ns, *Ad = S.facA_dn + (size_t)w * S.kmax * ns;
    const double *Bu = S.facB_up + (size_t)w * S.kmax * S.n_up, *Bd = S.facB_dn + (size_t)w * S.kmax * S.n_dn;
    double flips = 0.0;
    int diag4 = 0, bad = 0;
    for (int b = lane; b < S.n_bonds; b += 32) {
        const int i = S.bi[b], j = S.bj[b];
        const int iu = kup[i], ju = kup[j], id = kdn[i], jd = kdn[j];
        if (ju != 0 && id != 0)
            flips += -0.5 * w_entry_thread(S, Wu, Au, Bu, ns, S.n_up, i, ju - 1, cnt) *
                     w_entry_thread(S, Wd, Ad, Bd, ns, S.n_dn, j, id - 1, cnt);
        if (iu != 0 && jd != 0)
            flips += -0.5 * w_entry_thread(S, Wu, Au, Bu, ns, S.n_up, j, iu - 1, cnt) *
                     w_entry_thread(S, Wd, Ad, Bd, ns, S.n_dn, i, jd - 1, cnt);
        const int oi = (iu != 0) + (id != 0), oj = (ju != 0) + (jd != 0);
        if (oi != 1 || oj != 1) bad = 1;
        diag4 += (iu != 0 ? 1 : -1) * (ju != 0 ? 1 : -1);
    }
    flips = warp_sum_f64(flips);
    diag4 = warp_sum_int(diag4);
    bad = warp_sum_int(bad);
    if (lane == 0) {
        const double OL = flips + 0.25 * (double)diag4;
        if (bad) atomicOr(&S.flags[w], 4);
        if (ol_out) ol_out[w] = OL;
        if (accumulate) {
            S.ol_last[w] = OL;
            S.ol_sum[w] += OL;
            S.ol_sq[w] += OL * OL;
            S.ol_n[w] += 1ull;
        }
    }
}
