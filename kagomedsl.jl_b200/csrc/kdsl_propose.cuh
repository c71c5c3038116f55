// kdsl_propose.cuh -- the Metropolis proposal of Carlo.sweep! (reference src/MonteCarlo.jl:538-607)
// for all walkers of one GPU, one warp per walker.
//
// All 32 lanes evaluate the (scalar) decision redundantly so that control flow is warp-uniform;
// on acceptance the lanes cooperate to stage the Sherman-Morrison operands (the reference's
// col_cache / row_cache, src/MonteCarlo.jl:286-289) and to apply the 4 kappa writes
// (:502-503 / :508-509).  The W matrices themselves are streamed by k_update_* afterwards.
#pragma once
#include "kdsl_common.cuh"

#define KDSL_FLAG_NONFINITE_DEV 2

// occupancy-pair predicate of Z (src/MonteCarlo.jl:467-471)
__device__ __forceinline__ int bond_is_anti(int ua, int da, int ub, int db) {
    return (ua && db) || (ub && da);
}

template <bool REPLAY>
__global__ void __launch_bounds__(256)
k_propose(DevState S, int parity, int gate_refresh, const double *__restrict__ rp_r,
          const int *__restrict__ rp_bond, const int *__restrict__ rp_pick) {
    const int w = (int)((blockIdx.x * (unsigned)blockDim.x + threadIdx.x) >> 5);
    const int lane = threadIdx.x & 31;
    if (w >= S.nw) return;
    if (S.flags[w] & 1) return;      // frozen after a singular re-evaluation (see k_decide_wb; KDSL_FLAG_SINGULAR)
    const int ns = S.ns;
    int *kup = S.kup + (size_t)w * ns;
    int *kdn = S.kdn + (size_t)w * ns;

    const int zmu = S.zmu[w];                                   // :544 (maintained incrementally)
    Xoshiro g;
    if (!REPLAY) {
        const unsigned long long *st = S.rng + (size_t)w * 4;
        g.s0 = st[0]; g.s1 = st[1]; g.s2 = st[2]; g.s3 = st[3];
    }
    const double r = REPLAY ? rp_r[w] : g.rand_f64();           // :546
    const double zr = (double)zmu / (double)S.n_bonds;          // Zmu / Zmax
    bool accepted = false, reached = false;
    int i = 0, site = 0, flag = 0, l_up = 0, l_dn = 0, K_up = 0, K_dn = 0;
    int ku_i = 0, ku_s = 0, kd_i = 0, kd_s = 0;
    double wu = 0.0, wd = 0.0;

    if (!(r > zr)) {                                            // :547-550
        long long b = REPLAY ? (long long)rp_bond[w] : g.rand_index((unsigned long long)S.n_bonds);  // :552
        if (b < 1) b = 1;
        if (b > S.n_bonds) b = S.n_bonds;
        i = S.bi[b - 1];
        site = S.bj[b - 1];
        ku_i = kup[i]; ku_s = kup[site]; kd_i = kdn[i]; kd_s = kdn[site];
        const bool f1 = ku_i != 0 && kd_s != 0;                 // :558-561
        const bool f2 = ku_s != 0 && kd_i != 0;
        if (f1 || f2) {                                         // :563-567
            const int nm = (int)f1 + (int)f2;
            long long pick;                                     // :569 (draw consumed even when nm == 1)
            if (REPLAY) pick = rp_pick ? (long long)rp_pick[w] : 1;
            else pick = g.rand_index((unsigned long long)nm);
            flag = (f1 && f2) ? (pick == 1 ? 1 : 2) : (f1 ? 1 : 2);
            l_up = flag == 1 ? ku_i : ku_s;                     // :572-573
            l_dn = flag == 1 ? kd_s : kd_i;
            K_up = flag == 1 ? site : i;                        // :576-580 / :500,506
            K_dn = flag == 1 ? i : site;
            wu = S.W_up[(size_t)w * ns * S.n_up + (size_t)(l_up - 1) * ns + K_up];
            wd = S.W_dn[(size_t)w * ns * S.n_dn + (size_t)(l_dn - 1) * ns + K_dn];
            const double ratio = wu * wd;
            const double p = ratio * ratio;                     // abs2(ratio), :581
            if (p >= 1.0 && r < zr) accepted = true;            // :582-584
            else if (p < 1.0 && r < zr * p) accepted = true;    // :585-587
            if (!(p == p) || p > 1.79e308) {                    // NaN / Inf: flag it (the reference silently rejects NaN)
                if (lane == 0) atomicOr(&S.flags[w], KDSL_FLAG_NONFINITE_DEV);
            }
            reached = true;                                     // falls through to :594
        }
    }

    if (accepted) {
        // A walker that is re-evaluated from scratch at the end of this very sweep (:595-604)
        // needs no rank-1 update: reevaluateW! overwrites W from kappa alone.
        if (!gate_refresh) {
            const double au = -1.0 / wu, ad = -1.0 / wd;        // :289
            const double *Wu = S.W_up + (size_t)w * ns * S.n_up;
            const double *Wd = S.W_dn + (size_t)w * ns * S.n_dn;
            double *cu = S.col_up + (size_t)w * ns, *cd = S.col_dn + (size_t)w * ns;
            const double *srcu = Wu + (size_t)(l_up - 1) * ns, *srcd = Wd + (size_t)(l_dn - 1) * ns;
            for (int t = lane; t < ns; t += 32) {               // :288 col_cache = W[:, l]
                cu[t] = srcu[t];
                cd[t] = srcd[t];
            }
            double *tu = S.trow_up + (size_t)w * S.n_up, *td = S.trow_dn + (size_t)w * S.n_dn;
            for (int j = lane; j < S.n_up; j += 32) {           // :286-287 row_cache = W[K, :] - e_l, times alpha
                double v = Wu[(size_t)j * ns + K_up];
                if (j == l_up - 1) v -= 1.0;
                tu[j] = au * v;
            }
            for (int j = lane; j < S.n_dn; j += 32) {
                double v = Wd[(size_t)j * ns + K_dn];
                if (j == l_dn - 1) v -= 1.0;
                td[j] = ad * v;
            }
            if (lane == 0) {
                const int slot = atomicAdd(&S.cnt[parity], 1);
                S.acc_list[(size_t)parity * S.nw + slot] = w;
            }
        }
        // Z_mu changes only through the bonds incident to i or site; the bond (i, site) itself
        // stays antiparallel.  Equals the reference's full recount (:460-474).
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
            if (flag == 1) {                                    // :502-503
                kup[i] = 0; kup[site] = l_up;
                kdn[i] = l_dn; kdn[site] = 0;
            } else {                                            // :508-509
                kup[i] = l_up; kup[site] = 0;
                kdn[i] = 0; kdn[site] = l_dn;
            }
            S.n_acc[w] += 1ull;
        }
    }
    if (lane == 0) {
        if (reached) {
            S.n_reach[w] += 1ull;
            if (gate_refresh) {                                 // :595 ctx.sweeps % n_occupied == 0
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
