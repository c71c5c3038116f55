// kdsl_woodbury.cuh -- delayed updates in Woodbury form (update_variant 2, the production path).
//
// Between two passes over a walker's W0 the configuration differs from the one W0 belongs to by k
// particles: label l_n now sits on site K_n (n < k, labels distinct).  With
//     C = W0[:, (l_n)],   S = W0[(K_n), (l_n)]  (k x k),   Rt = W0[(K_n), :] - E   (E[n, :] = e_{l_n}^T)
// the Sherman-Morrison updates of the reference (update_W!, src/MonteCarlo.jl:279-292) telescope into
//     W = W0 - C S^-1 Rt                                   (Woodbury identity for W = U * inv(tilde_U))
// so a matrix entry needs only 2k+1 entries of W0 and T = S^-1, which is maintained per walker and
// species by bordering (new label) or a rank-1 row replacement (label moved again).  No ns-length vector
// is read or written per accepted move; k_flush_prepare materialises C and G = -T Rt only when the walker
// is flushed (W0 += C G by k_flush), and reevaluateW! simply resets k to 0.
#pragma once
#include "kdsl_common.cuh"
#include "kdsl_propose.cuh"
#include "kdsl_refresh_fast.cuh"

struct WbView {
    const double *W0;
    double *rows;    // [kmax][N] raw rows W0[K_n, :] of the displaced particles' new sites (filled on accept)
    double *T;       // [kmax][kmax] row-major
    int *Ks, *Ls;    // [kmax] target sites / labels (0-based)
    int ns, N, k;
};

__device__ __forceinline__ WbView wb_view(const DevState &S, int w, int spin) {
    WbView v;
    v.ns = S.ns;
    v.N = spin ? S.n_dn : S.n_up;
    v.W0 = (spin ? S.W_dn : S.W_up) + (size_t)w * S.ns * v.N;
    const size_t q = (size_t)w * 2 + spin;
    v.rows = (spin ? S.facA_dn : S.facA_up) + (size_t)w * KDSL_KALLOC * S.ns;
    v.T = S.wbT + q * S.kmax * S.kmax;
    v.Ks = S.wbK + q * S.kmax;
    v.Ls = S.wbL + q * S.kmax;
    v.k = S.fcnt[q];
    return v;
}

// Per-warp evaluation of W[K, l] (0-based).  Outputs kept per lane for a subsequent accept:
//   c (lane n: W0[K, l_n]), vv (lane m: (T r)_m), slot j of l in the label list (or -1).
struct WbEval {
    double entry, c, vv;
    int j, Kn, Ln;
};

__device__ __forceinline__ WbEval wb_entry_warp(const DevState &S, const WbView &v, int K, int l, int lane) {
    WbEval e;
    const int k = v.k, ns = v.ns, kmax = S.kmax;
    e.Kn = lane < k ? v.Ks[lane] : 0;
    e.Ln = lane < k ? v.Ls[lane] : -1;
    const unsigned hit = __ballot_sync(0xffffffffu, lane < k && e.Ln == l);
    e.j = hit ? (__ffs(hit) - 1) : -1;
    const double d = v.W0[(size_t)l * ns + K];
    e.c = lane < k ? v.W0[(size_t)e.Ln * ns + K] : 0.0;                      // W0[K, l_n]
    double r = lane < k ? v.W0[(size_t)l * ns + e.Kn] : 0.0;                 // W0[K_n, l] - delta
    if (lane == e.j) r -= 1.0;
    double vv = 0.0;                                                         // (T r)_m on lane m
    for (int n = 0; n < k; n++) {
        const double rn = __shfl_sync(0xffffffffu, r, n);
        if (lane < k) vv = fma(v.T[lane * kmax + n], rn, vv);
    }
    e.vv = vv;
    const double corr = warp_sum_f64(lane < k ? e.c * vv : 0.0);
    e.entry = d - corr;
    return e;
}

// Both species of one proposal at once: the same arithmetic as two calls of wb_entry_warp, with the loads of the two
// evaluations in flight together and the two T r loops merged (the proposal kernel is a chain of dependent memory
// round trips; the up and down halves of a ratio are independent of each other).
__device__ __forceinline__ void wb_entry_warp2(const DevState &S, const WbView &vu, const WbView &vd, int Ku, int lu, int Kd, int ld,
                                               int lane, WbEval &eu, WbEval &ed) {
    const int ku = vu.k, kd = vd.k, ns = vu.ns, kmax = S.kmax;
    eu.Kn = lane < ku ? vu.Ks[lane] : 0;
    eu.Ln = lane < ku ? vu.Ls[lane] : -1;
    ed.Kn = lane < kd ? vd.Ks[lane] : 0;
    ed.Ln = lane < kd ? vd.Ls[lane] : -1;
    const double du = vu.W0[(size_t)lu * ns + Ku];
    const double dd = vd.W0[(size_t)ld * ns + Kd];
    const unsigned hitu = __ballot_sync(0xffffffffu, lane < ku && eu.Ln == lu);
    const unsigned hitd = __ballot_sync(0xffffffffu, lane < kd && ed.Ln == ld);
    eu.j = hitu ? (__ffs(hitu) - 1) : -1;
    ed.j = hitd ? (__ffs(hitd) - 1) : -1;
    eu.c = lane < ku ? vu.W0[(size_t)eu.Ln * ns + Ku] : 0.0;                 // W0[K, l_n]
    ed.c = lane < kd ? vd.W0[(size_t)ed.Ln * ns + Kd] : 0.0;
    double ru = lane < ku ? vu.W0[(size_t)lu * ns + eu.Kn] : 0.0;            // W0[K_n, l] - delta
    double rd = lane < kd ? vd.W0[(size_t)ld * ns + ed.Kn] : 0.0;
    if (lane == eu.j) ru -= 1.0;
    if (lane == ed.j) rd -= 1.0;
    double vvu = 0.0, vvd = 0.0;                                             // (T r)_m on lane m
    const int kk = max(ku, kd);
    for (int n = 0; n < kk; n++) {
        const double tu = (lane < ku && n < ku) ? vu.T[lane * kmax + n] : 0.0;
        const double td = (lane < kd && n < kd) ? vd.T[lane * kmax + n] : 0.0;
        const double rnu = __shfl_sync(0xffffffffu, ru, n);
        const double rnd = __shfl_sync(0xffffffffu, rd, n);
        if (lane < ku && n < ku) vvu = fma(tu, rnu, vvu);
        if (lane < kd && n < kd) vvd = fma(td, rnd, vvd);
    }
    eu.vv = vvu;
    ed.vv = vvd;
    const double corru = warp_sum_f64(lane < ku ? eu.c * vvu : 0.0);
    const double corrd = warp_sum_f64(lane < kd ? ed.c * vvd : 0.0);
    eu.entry = du - corru;
    ed.entry = dd - corrd;
}

// rows[slot][:] = W0[K_slot, :] for the slots in the two masks (one strided gather per displaced particle, done once
// at the end of a launch).  One slot of each species per round, all loads of a round issued before its stores (the
// pointers may alias as far as the compiler knows), so a round costs one memory latency.
__device__ __forceinline__ void wb_copy_rows(const DevState &S, int w, unsigned mask_up, unsigned mask_dn, int lane) {
    const WbView vu = wb_view(S, w, 0), vd = wb_view(S, w, 1);
    while (mask_up | mask_dn) {
        const int su = mask_up ? __ffs(mask_up) - 1 : -1, sd = mask_dn ? __ffs(mask_dn) - 1 : -1;
        mask_up &= mask_up - 1u;
        mask_dn &= mask_dn - 1u;
        const int Ku = su >= 0 ? vu.Ks[su] : 0, Kd = sd >= 0 ? vd.Ks[sd] : 0;
        const int Nmax = max(vu.N, vd.N);
        for (int c0 = 0; c0 < Nmax; c0 += 128) {
            double tu[4], td[4];
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const int c = c0 + 32 * q + lane;
                tu[q] = (su >= 0 && c < vu.N) ? vu.W0[(size_t)c * vu.ns + Ku] : 0.0;
                td[q] = (sd >= 0 && c < vd.N) ? vd.W0[(size_t)c * vd.ns + Kd] : 0.0;
            }
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const int c = c0 + 32 * q + lane;
                if (su >= 0 && c < vu.N) vu.rows[(size_t)su * vu.N + c] = tu[q];
                if (sd >= 0 && c < vd.N) vd.rows[(size_t)sd * vd.N + c] = td[q];
            }
        }
    }
}

// Apply the accepted move "label l -> site K" to the Woodbury state of one species (whole warp).
// Returns the slot of the displaced-particle list that now refers to site K (its row copy is refreshed by the caller).
__device__ __forceinline__ int wb_accept_warp(const DevState &S, const WbView &v, const WbEval &e, int w,
                                              int spin, int K, int l, int lane) {
    const int k = v.k, kmax = S.kmax;
    double y = 0.0;                                                          // (c^T T)_n on lane n
    for (int m = 0; m < k; m++) {
        const double cm = __shfl_sync(0xffffffffu, e.c, m);
        if (lane < k) y = fma(cm, v.T[m * kmax + lane], y);
    }
    if (e.j < 0) {
        // new label: border S with (column W0[K_set, l], row W0[K, L], corner W0[K, l]); Schur complement s = W[K, l]
        const double inv_s = 1.0 / e.entry;
        for (int n = 0; n < k; n++) {
            const double yn = __shfl_sync(0xffffffffu, y, n) * inv_s;
            if (lane < k) v.T[lane * kmax + n] = fma(e.vv, yn, v.T[lane * kmax + n]);
        }
        if (lane < k) {
            v.T[lane * kmax + k] = -e.vv * inv_s;
            v.T[k * kmax + lane] = -y * inv_s;
        }
        if (lane == 0) {
            v.T[k * kmax + k] = inv_s;
            v.Ks[k] = K;
            v.Ls[k] = l;
            S.fcnt[(size_t)w * 2 + spin] = k + 1;
        }
        return k;
    } else {
        // label already displaced: row j of S becomes W0[K, L] = c.  T' = T - (T e_j) w^T / (c^T T)_j, w = c^T T - e_j
        const int j = e.j;
        const double inv_den = 1.0 / __shfl_sync(0xffffffffu, y, j);
        const double tj = lane < k ? v.T[lane * kmax + j] : 0.0;             // (T e_j)_m on lane m
        double wv = y;
        if (lane == j) wv -= 1.0;
        for (int n = 0; n < k; n++) {
            const double wn = __shfl_sync(0xffffffffu, wv, n) * inv_den;
            if (lane < k) v.T[lane * kmax + n] = fma(-tj, wn, v.T[lane * kmax + n]);
        }
        if (lane == 0) v.Ks[j] = K;
        return j;
    }
}

// Carlo.sweep! proposal (reference src/MonteCarlo.jl:538-607), one warp per walker, Woodbury-form W.
// One sweep of one walker (whole warp).  The Xoshiro state g lives in registers across the sweeps of a launch.
template <bool REPLAY>
__device__ __forceinline__ void decide_sweep_wb(const DevState &S, int w, int lane, int gate_refresh, Xoshiro &g,
                                                const double *__restrict__ rp_r, const int *__restrict__ rp_bond,
                                                const int *__restrict__ rp_pick, unsigned &dirty_up, unsigned &dirty_dn) {
    const int ns = S.ns;
    int *kup = S.kup + (size_t)w * ns;
    int *kdn = S.kdn + (size_t)w * ns;
    const int zmu = S.zmu[w];
    const double r = REPLAY ? rp_r[w] : g.rand_f64();               // :546
    const double zr = (double)zmu / (double)S.n_bonds;
    bool accepted = false, reached = false;
    int i = 0, site = 0, flag = 0, l_up = 0, l_dn = 0, K_up = 0, K_dn = 0;
    int ku_i = 0, ku_s = 0, kd_i = 0, kd_s = 0;
    WbView vu, vd;
    WbEval eu, ed;
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
            vu = wb_view(S, w, 0);
            vd = wb_view(S, w, 1);
            wb_entry_warp2(S, vu, vd, K_up, l_up - 1, K_dn, l_dn - 1, lane, eu, ed);   // :576-580
            const double ratio = eu.entry * ed.entry;
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
        if (!gate_refresh) {                                        // (a walker re-evaluated this sweep needs no update)
            // (a merged two-species accept, the analogue of wb_entry_warp2, was measured slower: 6.47 ms against 6.11 ms
            //  per 432 sweeps -- the accept path is taken by one sweep in eight and the merged loops spill)
            dirty_up |= 1u << wb_accept_warp(S, vu, eu, w, 0, K_up, l_up - 1, lane);
            dirty_dn |= 1u << wb_accept_warp(S, vd, ed, w, 1, K_dn, l_dn - 1, lane);
            const int knew = max(vu.k + (eu.j < 0), vd.k + (ed.j < 0)), kold = max(vu.k, vd.k);
            (void)kold;
            if (lane == 0 && knew >= S.kth && !S.listed[w]) {       // due for a flush: listed at most once (the flag is
                S.listed[w] = 1;                                    // cleared by k_flush_finish_wb), so the list cannot
                const int fs = atomicAdd(&S.cnt[4], 1);             // overflow whatever flush_every / flush_threshold are
                S.flush_list[fs] = w;
            }
        }
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
        }
    }
    if (lane == 0 && reached) {
        S.n_reach[w] += 1ull;
        if (gate_refresh) {                                         // :595
            const int slot = atomicAdd(&S.cnt[2], 1);
            S.ref_list[slot] = w;
        }
    }
}

// Carlo.sweep! for every walker, n_sweeps consecutive proposals per launch (the walkers are independent between
// two flushes, so the lock-step loop only has to come back to the host at gate / flush / measurement sweeps).
// Replay inputs of sweep s are at rp_*[s * nw + w].  gate_refresh must be 0 when n_sweeps > 1.
template <bool REPLAY>
__global__ void __launch_bounds__(256, 4)
k_decide_wb(DevState S, int gate_refresh, int n_sweeps, const double *__restrict__ rp_r,
            const int *__restrict__ rp_bond, const int *__restrict__ rp_pick) {
    const int w = (int)((blockIdx.x * (unsigned)blockDim.x + threadIdx.x) >> 5);
    const int lane = threadIdx.x & 31;
    if (w >= S.nw) return;
    // A walker whose last re-evaluation met a singular tilde_U is frozen: the reference throws SingularException out of
    // sweep! at that point (src/MonteCarlo.jl:596-603), here the walker stops proposing, keeps KDSL_FLAG_SINGULAR and is
    // counted in KDSL_ACC_N_SINGULAR (its W / Woodbury state may no longer belong to its kappa).
    if (S.flags[w] & KDSL_FLAG_SINGULAR_DEV) return;
    Xoshiro g;
    g.s0 = g.s1 = g.s2 = g.s3 = 0ull;
    if (!REPLAY) {
        const unsigned long long *st = S.rng + (size_t)w * 4;
        g.s0 = st[0]; g.s1 = st[1]; g.s2 = st[2]; g.s3 = st[3];
    }
    // Pull this walker's small state (kappa, the Woodbury T matrices and lists: ~13 KB) towards the SM before the
    // first proposal needs it: every sweep is a chain of dependent accesses to it, and between two launches the
    // flush kernel has streamed hundreds of MB through the L2.  One prefetch per 128-byte line, no registers held
    // (-4 % on the launch; issuing the T loads of the ratio / accept loops in chunks was tried and lost to the spills
    // under the 64-register cap that keeps all 4096 warps resident).
    {
        const char *pk_up = reinterpret_cast<const char *>(S.kup + (size_t)w * S.ns);
        const char *pk_dn = reinterpret_cast<const char *>(S.kdn + (size_t)w * S.ns);
        for (int off = lane * 128; off < S.ns * 4; off += 32 * 128) {
            asm volatile("prefetch.global.L2 [%0];" ::"l"(pk_up + off));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(pk_dn + off));
        }
        const char *pT = reinterpret_cast<const char *>(S.wbT + (size_t)w * 2 * S.kmax * S.kmax);
        for (int off = lane * 128; off < 2 * S.kmax * S.kmax * 8; off += 32 * 128)
            asm volatile("prefetch.global.L2 [%0];" ::"l"(pT + off));
        if (lane < 4) {
            const char *pL = reinterpret_cast<const char *>((lane & 1 ? S.wbL : S.wbK) + (size_t)w * 2 * S.kmax);
            asm volatile("prefetch.global.L2 [%0];" ::"l"(pL + (lane >> 1) * 128));
        }
    }
    unsigned dirty_up = 0u, dirty_dn = 0u;                          // displaced-particle slots (re)assigned in this launch
    for (int s = 0; s < n_sweeps; s++) {
        const size_t off = (size_t)s * S.nw;
        decide_sweep_wb<REPLAY>(S, w, lane, gate_refresh, g, REPLAY ? rp_r + off : nullptr,
                                REPLAY ? rp_bond + off : nullptr, (REPLAY && rp_pick) ? rp_pick + off : nullptr,
                                dirty_up, dirty_dn);
        __syncwarp();                                               // lane-0 / row-owner writes -> next sweep's reads
    }
    wb_copy_rows(S, w, dirty_up, dirty_dn, lane);                   // raw rows W0[K, :] for the next flush
    if (!REPLAY && lane == 0) {
        unsigned long long *st = S.rng + (size_t)w * 4;
        st[0] = g.s0; st[1] = g.s1; st[2] = g.s2; st[3] = g.s3;
    }
}

// G[m][j] = sign * sum_n T[m][n] (W0[K_n, j] - delta(l_n, j)) for all columns j, by all threads of the CTA.
// sT: T packed [k][k] in shared memory; sK, sL: the displaced sites / labels.  out has leading dimension ldo.
__device__ __forceinline__ void wb_compute_G_cta(const WbView &v, const double *sT, const int *sK, const int *sL,
                                                 double *out, int ldo, double sign, int tid, int nthreads) {
    const int k = v.k, ns = v.ns, N = v.N;
    for (int j = tid; j < N; j += nthreads) {
        double rt[32];
#pragma unroll
        for (int n = 0; n < 32; n++)
            rt[n] = n < k ? v.W0[(size_t)j * ns + sK[n]] - (sL[n] == j ? 1.0 : 0.0) : 0.0;
        for (int m = 0; m < k; m++) {
            double acc = 0.0;
#pragma unroll
            for (int n = 0; n < 32; n++) if (n < k) acc = fma(sT[m * k + n], rt[n], acc);
            out[(size_t)m * ldo + j] = sign * acc;
        }
    }
}

#ifdef KDSL_DEV_VARIANTS   // flush_variant 1 (k_flush_prepare + k_flush): superseded by k_flush_wb
// Materialise the right flush operand of the listed walkers: facB[m] = G[m, :] = -(T Rt)[m, :], so that k_flush's
// W0 += sum_m W0[:, l_m] (x) facB[m]  applies  W = W0 - C T Rt  (the left operand C is read from W0 itself).
// One CTA per (list entry, species).
__global__ void __launch_bounds__(256)
k_flush_prepare(DevState S, const int *__restrict__ list, const int *__restrict__ count_ptr, int count_fixed) {
    extern __shared__ double psm[];                       // T [k*k], then Ks/Ls ints
    const int count = count_ptr ? *count_ptr : count_fixed;
    const int tid = threadIdx.x;
    for (int item = blockIdx.x; item < 2 * count; item += gridDim.x) {
        const int e = item >> 1, spin = item & 1;
        const int w = list ? list[e] : e;
        const WbView v = wb_view(S, w, spin);
        const int k = v.k, kmax = S.kmax;
        if (k == 0) continue;
        double *sT = psm;
        int *sK = reinterpret_cast<int *>(psm + kmax * kmax), *sL = sK + kmax;
        __syncthreads();
        for (int x = tid; x < k * k; x += blockDim.x) sT[x] = v.T[(x / k) * kmax + (x % k)];
        if (tid < k) { sK[tid] = v.Ks[tid]; sL[tid] = v.Ls[tid]; }
        __syncthreads();
        double *B = (spin ? S.facB_dn : S.facB_up) + (size_t)w * kmax * v.N;
        wb_compute_G_cta(v, sT, sK, sL, B, v.N, -1.0, tid, blockDim.x);
    }
}

#endif  // KDSL_DEV_VARIANTS

// W0 += C G for the listed walkers (C = columns l_m of W0 itself, G = -T Rt built in the kernel): the HBM-bound pass
// of the delayed update, Woodbury form.  Persistent CTAs (two per SM) fetch work items from a device counter;
// item = (list entry, species, block of 216 rows); 9 warps, each owns a strip of 24 rows and walks over all
// column tiles with the DMMA in transposed form (D[column][row]: every lane moves two adjacent rows of one column,
// i.e. 128-bit loads / stores, 64 contiguous bytes per column and warp) and three tiles of loads in flight.
template <int KPAD>
__global__ void __launch_bounds__(288, 2)
k_flush_wb(DevState S, const int *__restrict__ list, const int *__restrict__ count_ptr, int count_fixed,
           int *__restrict__ work_counter) {
    constexpr int KS = KPAD / 4, RT = 3, MT = (KPAD + 7) / 8;
    extern __shared__ double fsm[];                      // G, frag-major (r = column j, k = m)
    __shared__ double sT[MT * 8 * KPAD];                 // T, frag-major (r = m, k = n), zero padded
    __shared__ int sL[KPAD];
    __shared__ int s_item;
    const int count = count_ptr ? *count_ptr : count_fixed;
    const int ns = S.ns;
    const int nrb = (ns + 215) / 216;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int gr = lane >> 2, tg = lane & 3;
    const int total = count * 2 * nrb;
    for (;;) {
        __syncthreads();
        if (tid == 0) s_item = atomicAdd(work_counter, 1);
        __syncthreads();
        const int item = s_item;
        if (item >= total) break;
        const int e = item / (2 * nrb);
        const int rem = item - e * 2 * nrb;
        const int spin = rem / nrb, rb = rem - spin * nrb;
        const int w = list ? list[e] : e;
        const int cnt = S.fcnt[2 * w + spin];
        if (cnt == 0) continue;                           // uniform over the block
        const int N = spin ? S.n_dn : S.n_up;
        const int ctiles = (N + 7) >> 3;
        const WbView v = wb_view(S, w, spin);
        const int *Ls = v.Ls;
        double *W0 = (spin ? S.W_dn : S.W_up) + (size_t)w * ns * N;
        // G = -T Rt (KPAD x N) for this walker and species, straight into shared memory in fragment order:
        // Rt[n][j] = W0[K_n, j] - delta(l_n, j) comes from the row copies made when the moves were accepted
        for (int x = tid; x < MT * 8 * KPAD; x += 288) {
            const int m = x / KPAD, n = x - m * KPAD;
            sT[frag_idx(m, n, KPAD)] = (m < cnt && n < cnt) ? v.T[m * S.kmax + n] : 0.0;
        }
        if (tid < KPAD) sL[tid] = tid < cnt ? Ls[tid] : -1;
        __syncthreads();
        for (int ct = warp; ct < ctiles; ct += 9) {
            const int j = (ct << 3) + gr;
            double rt[KS];                                // B operand (k = n, col = j)
#pragma unroll
            for (int q = 0; q < KS; q++) {
                const int n = 4 * q + tg;
                rt[q] = (n < cnt && j < N) ? v.rows[(size_t)n * N + j] - (sL[n] == j ? 1.0 : 0.0) : 0.0;
            }
#pragma unroll
            for (int t = 0; t < MT; t++) {
                double d0 = 0.0, d1 = 0.0;
#pragma unroll
                for (int q = 0; q < KS; q++) dmma_8x8x4(d0, d1, sT[(((t * KS) + q) << 5) + lane], rt[q]);
                const int m = 8 * t + gr;                 // D[m][j = 8 ct + 2 tg + e] -> frag_idx(j, m, KPAD)
                if (m < KPAD) {
                    double *dst = fsm + ((((size_t)ct * KS) + (m >> 2)) << 5) + (m & 3);
                    dst[(2 * tg) << 2] = -d0;
                    dst[(2 * tg + 1) << 2] = -d1;
                }
            }
        }
        const int r0 = rb * 216 + warp * 24;
        double af[RT][KS];                                // C^T fragments (k = m, n = row): W0[row, l_m]
#pragma unroll
        for (int s = 0; s < KS; s++) {
            const int m = 4 * s + tg;
            const int lm = m < cnt ? Ls[m] : 0;
#pragma unroll
            for (int t = 0; t < RT; t++) {
                const int row = r0 + 8 * t + gr;
                af[t][s] = (row < ns && m < cnt) ? W0[(size_t)lm * ns + row] : 0.0;
            }
        }
        __syncthreads();
        if (r0 >= ns) continue;
        bool rv[RT];
#pragma unroll
        for (int t = 0; t < RT; t++) rv[t] = r0 + 8 * t + 2 * tg < ns;     // ns is even: both rows or none
        double2 *base = reinterpret_cast<double2 *>(W0 + (size_t)gr * ns + r0 + 2 * tg);
        const size_t cstride = (size_t)4 * ns;            // 8 columns, in double2 units
        auto load_tile = [&](int ct, double2 (&cc)[RT]) {
            const bool cok = (ct << 3) + gr < N;
#pragma unroll
            for (int t = 0; t < RT; t++)
                cc[t] = (rv[t] && cok) ? base[(size_t)ct * cstride + 4 * t] : make_double2(0.0, 0.0);
        };
        auto compute_store = [&](int ct, double2 (&cc)[RT]) {
#pragma unroll
            for (int s = 0; s < KS; s++) {
                const double gf = fsm[(((ct * KS) + s) << 5) + lane];
#pragma unroll
                for (int t = 0; t < RT; t++) dmma_8x8x4(cc[t].x, cc[t].y, gf, af[t][s]);
            }
            const bool cok = (ct << 3) + gr < N;
#pragma unroll
            for (int t = 0; t < RT; t++)
                if (rv[t] && cok) base[(size_t)ct * cstride + 4 * t] = cc[t];
        };
        double2 c0[RT], c1[RT], c2[RT];
        load_tile(0, c0);
        if (1 < ctiles) load_tile(1, c1);
        for (int ct = 0; ct < ctiles; ct += 3) {           // three-deep register pipeline
            if (ct + 2 < ctiles) load_tile(ct + 2, c2);
            compute_store(ct, c0);
            if (ct + 3 < ctiles) load_tile(ct + 3, c0);
            if (ct + 1 < ctiles) compute_store(ct + 1, c1);
            if (ct + 4 < ctiles) load_tile(ct + 4, c1);
            if (ct + 2 < ctiles) compute_store(ct + 2, c2);
        }
    }
}

// after k_flush_wb (one CTA): the listed walkers have no pending updates; re-arm the list and the work counter
__global__ void __launch_bounds__(1024)
k_flush_finish_wb(DevState S, const int *__restrict__ list, int *count_ptr, int count_fixed, int *work_counter) {
    const int count = count_ptr ? *count_ptr : count_fixed;
    for (int e = threadIdx.x; e < count; e += blockDim.x) {
        const int w = list ? list[e] : e;
        S.fcnt[(size_t)2 * w] = 0;
        S.fcnt[(size_t)2 * w + 1] = 0;
        if (list) S.listed[w] = 0;                          // (flush-all leaves the list alone: its walkers stay listed
    }                                                       //  until the next listed flush finds them empty)
    __syncthreads();
    if (threadIdx.x == 0) {
        if (count_ptr) { S.upd_moves[1] += (unsigned long long)count; *count_ptr = 0; }
        *work_counter = 0;
    }
}

#ifdef KDSL_DEV_VARIANTS
// after k_flush in Woodbury mode: both species of the listed walkers are up to date
__global__ void k_flush_done_wb(DevState S, const int *__restrict__ list, int *count_ptr, int count_fixed) {
    const int count = count_ptr ? *count_ptr : count_fixed;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < count; e += gridDim.x * blockDim.x) {
        const int w = list ? list[e] : e;
        S.fcnt[(size_t)2 * w] = 0;
        S.fcnt[(size_t)2 * w + 1] = 0;
    }
    if (count_ptr && blockIdx.x == 0 && threadIdx.x == 0) S.upd_moves[1] += (unsigned long long)count;
}

#endif  // KDSL_DEV_VARIANTS

// O_L (reference getOL, src/Hamiltonian.jl:762-778) with Woodbury-form W; one CTA (256 threads) per walker.
// Phase 1: G_s = T_s Rt_s (k x N) for both species into shared memory; phase 2: threads stride over the bonds,
// each flip term needs W[K, l] = W0[K, l] - sum_m W0[K, l_m] G[m][l].
__global__ void __launch_bounds__(256)
k_measure_wb(DevState S, double *__restrict__ ol_out, int accumulate) {
    extern __shared__ double msm[];
    const int w = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ns = S.ns, kmax = S.kmax;
    const int *kup = S.kup + (size_t)w * ns;
    const int *kdn = S.kdn + (size_t)w * ns;
    WbView v[2] = {wb_view(S, w, 0), wb_view(S, w, 1)};
    double *sG[2], *sT[2];
    int *sK[2], *sL[2];
    sG[0] = msm;
    sG[1] = sG[0] + (size_t)kmax * S.n_up;
    sT[0] = sG[1] + (size_t)kmax * S.n_dn;
    sT[1] = sT[0] + kmax * kmax;
    sK[0] = reinterpret_cast<int *>(sT[1] + kmax * kmax);
    sL[0] = sK[0] + kmax; sK[1] = sL[0] + kmax; sL[1] = sK[1] + kmax;
    __shared__ double red_f[8];
    __shared__ int red_d[8], red_b[8];
#pragma unroll
    for (int s = 0; s < 2; s++) {
        const int k = v[s].k;
        for (int x = tid; x < k * k; x += 256) sT[s][x] = v[s].T[(x / k) * kmax + (x % k)];
        if (tid < k) { sK[s][tid] = v[s].Ks[tid]; sL[s][tid] = v[s].Ls[tid]; }
    }
    __syncthreads();
#pragma unroll
    for (int s = 0; s < 2; s++)
        if (v[s].k > 0) wb_compute_G_cta(v[s], sT[s], sK[s], sL[s], sG[s], v[s].N, 1.0, tid, 256);
    __syncthreads();
    auto entry = [&](int s, int K, int l) {
        double e = v[s].W0[(size_t)l * ns + K];
        const int k = v[s].k;
        for (int m = 0; m < k; m++) e = fma(-v[s].W0[(size_t)sL[s][m] * ns + K], sG[s][(size_t)m * v[s].N + l], e);
        return e;
    };
    double flips = 0.0;
    int diag4 = 0, bad = 0;
    for (int b = tid; b < S.n_bonds; b += 256) {
        const int i = S.bi[b], j = S.bj[b];
        const int iu = kup[i], ju = kup[j], id = kdn[i], jd = kdn[j];
        if (ju != 0 && id != 0) flips += -0.5 * entry(0, i, ju - 1) * entry(1, j, id - 1);
        if (iu != 0 && jd != 0) flips += -0.5 * entry(0, j, iu - 1) * entry(1, i, jd - 1);
        const int oi = (iu != 0) + (id != 0), oj = (ju != 0) + (jd != 0);
        if (oi != 1 || oj != 1) bad = 1;
        diag4 += (iu != 0 ? 1 : -1) * (ju != 0 ? 1 : -1);
    }
    flips = warp_sum_f64(flips);
    diag4 = warp_sum_int(diag4);
    bad = warp_sum_int(bad);
    if (lane == 0) { red_f[warp] = flips; red_d[warp] = diag4; red_b[warp] = bad; }
    __syncthreads();
    if (tid == 0) {
        double f = 0.0;
        int d4 = 0, bd = 0;
        for (int q = 0; q < 8; q++) { f += red_f[q]; d4 += red_d[q]; bd += red_b[q]; }
        const double OL = f + 0.25 * (double)d4;
        if (bd) atomicOr(&S.flags[w], 4);
        if (ol_out) ol_out[w] = OL;
        if (accumulate && !(S.flags[w] & KDSL_FLAG_SINGULAR_DEV)) {   // (a frozen walker's W is stale: no sample)
            S.ol_last[w] = OL;
            S.ol_sum[w] += OL;
            S.ol_sq[w] += OL * OL;
            S.ol_n[w] += 1ull;
        }
    }
}
