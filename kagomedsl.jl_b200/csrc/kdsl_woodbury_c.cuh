// kdsl_woodbury_c.cuh -- delayed updates in Woodbury form for the ComplexF64 engine (Peierls flux B != 0:
// src/Hamiltonian.jl:201-204, 325-327, scripts/LL.jl).  Same mathematics and the same launch structure as the real
// path (kdsl_woodbury.cuh):  W = W0 - C T Rt,  C = W0[:, (l_n)],  T = inv(W0[(K_n), (l_n)]),  Rt = W0[(K_n), :] - E,
// all products bilinear (update_W! is an unconjugated geru, src/MonteCarlo.jl:279-292), the acceptance uses
// abs2(ratio) (:581) and O_L is real(OL) (src/Hamiltonian.jl:777).  Matrix elements are interleaved (re, im) doubles.
//   k_decide_wb_c   proposals of up to 8 consecutive sweeps per launch, one warp per walker
//   k_flush_dmma2_c W0 += C G (G = -T Rt) for the listed walkers on the FP64 tensor pipe, four real DMMAs per complex block
//                   product on split (re, im) fragment planes, two 256-thread CTAs per SM (production);
//   k_flush_dmma_c  the same with G of the whole species in shared memory and one 512-thread CTA per SM (flush_variant 7, and
//                   the fall-back when two CTAs do not fit);  k_flush_c: the FMA-pipe version, one thread per row (flush_variant 4,
//                   and the fall-back for sizes whose operands do not fit the shared memory)
//   k_measure_wb_c  O_L with Woodbury-form entries
// The real path's k_flush_finish_wb and refresh status kernels are shared (they only touch counters).
#pragma once
#include "kdsl_common.cuh"
#include "kdsl_complex.cuh"
#include "kdsl_woodbury.cuh"
#include "kdsl_refresh_fast.cuh"

struct WbViewC {
    const cplx *W0;
    cplx *rows;      // [kmax][N] raw rows W0[K_n, :]
    cplx *T;         // [kmax][kmax] row-major
    int *Ks, *Ls;
    int ns, N, k;
};

__device__ __forceinline__ WbViewC wb_view_c(const DevState &S, int w, int spin) {
    WbViewC v;
    v.ns = S.ns;
    v.N = spin ? S.n_dn : S.n_up;
    v.W0 = cW(spin ? S.W_dn : S.W_up, (size_t)w * S.ns * v.N);
    const size_t q = (size_t)w * 2 + spin;
    v.rows = cW(spin ? S.facA_dn : S.facA_up, (size_t)w * KDSL_KALLOC * S.ns);
    v.T = cW(S.wbT, q * S.kmax * S.kmax);
    v.Ks = S.wbK + q * S.kmax;
    v.Ls = S.wbL + q * S.kmax;
    v.k = S.fcnt[q];
    return v;
}

__device__ __forceinline__ cplx c_shfl(cplx v, int src) {
    return make_double2(__shfl_sync(0xffffffffu, v.x, src), __shfl_sync(0xffffffffu, v.y, src));
}
__device__ __forceinline__ cplx c_warp_sum(cplx v) { return make_double2(warp_sum_f64(v.x), warp_sum_f64(v.y)); }
__device__ __forceinline__ cplx c_sub(cplx a, cplx b) { return make_double2(a.x - b.x, a.y - b.y); }

struct WbEvalC {
    cplx entry, c, vv;
    int j, Kn, Ln;
};

// W[K, l] (0-based) in Woodbury form, whole warp; outputs kept per lane for a subsequent accept
__device__ __forceinline__ WbEvalC wb_entry_warp_c(const DevState &S, const WbViewC &v, int K, int l, int lane) {
    WbEvalC e;
    const int k = v.k, ns = v.ns, kmax = S.kmax;
    const cplx zero = c_make(0.0, 0.0);
    e.Kn = lane < k ? v.Ks[lane] : 0;
    e.Ln = lane < k ? v.Ls[lane] : -1;
    const unsigned hit = __ballot_sync(0xffffffffu, lane < k && e.Ln == l);
    e.j = hit ? (__ffs(hit) - 1) : -1;
    const cplx d = v.W0[(size_t)l * ns + K];
    e.c = lane < k ? v.W0[(size_t)e.Ln * ns + K] : zero;                     // W0[K, l_n]
    cplx r = lane < k ? v.W0[(size_t)l * ns + e.Kn] : zero;                  // W0[K_n, l] - delta
    if (lane == e.j) r.x -= 1.0;
    cplx vv = zero;                                                          // (T r)_m on lane m
    for (int n = 0; n < k; n++) {
        const cplx rn = c_shfl(r, n);
        if (lane < k) vv = c_fma(v.T[lane * kmax + n], rn, vv);
    }
    e.vv = vv;
    const cplx corr = c_warp_sum(lane < k ? c_mul(e.c, vv) : zero);
    e.entry = c_sub(d, corr);
    return e;
}

__device__ __forceinline__ void wb_copy_rows_c(const DevState &S, int w, unsigned mask_up, unsigned mask_dn, int lane) {
    const WbViewC vu = wb_view_c(S, w, 0), vd = wb_view_c(S, w, 1);
    const cplx zero = c_make(0.0, 0.0);
    while (mask_up | mask_dn) {
        const int su = mask_up ? __ffs(mask_up) - 1 : -1, sd = mask_dn ? __ffs(mask_dn) - 1 : -1;
        mask_up &= mask_up - 1u;
        mask_dn &= mask_dn - 1u;
        const int Ku = su >= 0 ? vu.Ks[su] : 0, Kd = sd >= 0 ? vd.Ks[sd] : 0;
        const int Nmax = max(vu.N, vd.N);
        for (int c0 = 0; c0 < Nmax; c0 += 64) {
            cplx tu[2], td[2];
#pragma unroll
            for (int q = 0; q < 2; q++) {
                const int c = c0 + 32 * q + lane;
                tu[q] = (su >= 0 && c < vu.N) ? vu.W0[(size_t)c * vu.ns + Ku] : zero;
                td[q] = (sd >= 0 && c < vd.N) ? vd.W0[(size_t)c * vd.ns + Kd] : zero;
            }
#pragma unroll
            for (int q = 0; q < 2; q++) {
                const int c = c0 + 32 * q + lane;
                if (su >= 0 && c < vu.N) vu.rows[(size_t)su * vu.N + c] = tu[q];
                if (sd >= 0 && c < vd.N) vd.rows[(size_t)sd * vd.N + c] = td[q];
            }
        }
    }
}

// accepted move "label l -> site K": border S (new label) or replace row j of S (label moved again); returns the slot
__device__ __forceinline__ int wb_accept_warp_c(const DevState &S, const WbViewC &v, const WbEvalC &e, int w, int spin,
                                                int K, int l, int lane) {
    const int k = v.k, kmax = S.kmax;
    const cplx zero = c_make(0.0, 0.0);
    cplx y = zero;                                                           // (c^T T)_n on lane n
    for (int m = 0; m < k; m++) {
        const cplx cm = c_shfl(e.c, m);
        if (lane < k) y = c_fma(cm, v.T[m * kmax + lane], y);
    }
    if (e.j < 0) {
        const cplx inv_s = c_inv(e.entry);                                    // Schur complement s = W[K, l]
        for (int n = 0; n < k; n++) {
            const cplx yn = c_mul(c_shfl(y, n), inv_s);
            if (lane < k) v.T[lane * kmax + n] = c_fma(e.vv, yn, v.T[lane * kmax + n]);
        }
        if (lane < k) {
            v.T[lane * kmax + k] = c_neg(c_mul(e.vv, inv_s));
            v.T[k * kmax + lane] = c_neg(c_mul(y, inv_s));
        }
        if (lane == 0) {
            v.T[k * kmax + k] = inv_s;
            v.Ks[k] = K;
            v.Ls[k] = l;
            S.fcnt[(size_t)w * 2 + spin] = k + 1;
        }
        return k;
    } else {
        const int j = e.j;
        const cplx inv_den = c_inv(c_shfl(y, j));
        const cplx tj = lane < k ? v.T[lane * kmax + j] : zero;             // (T e_j)_m on lane m
        cplx wv = y;
        if (lane == j) wv.x -= 1.0;
        for (int n = 0; n < k; n++) {
            const cplx wn = c_mul(c_shfl(wv, n), inv_den);
            if (lane < k) v.T[lane * kmax + n] = c_fma(c_neg(tj), wn, v.T[lane * kmax + n]);
        }
        if (lane == 0) v.Ks[j] = K;
        return j;
    }
}

// Carlo.sweep! proposal (src/MonteCarlo.jl:538-607), one warp per walker, complex Woodbury-form W
template <bool REPLAY>
__device__ __forceinline__ void decide_sweep_wb_c(const DevState &S, int w, int lane, int gate_refresh, Xoshiro &g,
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
    WbViewC vu, vd;
    WbEvalC eu, ed;
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
            vu = wb_view_c(S, w, 0);
            vd = wb_view_c(S, w, 1);
            eu = wb_entry_warp_c(S, vu, K_up, l_up - 1, lane);      // :576-580
            ed = wb_entry_warp_c(S, vd, K_dn, l_dn - 1, lane);
            const double p = c_abs2(c_mul(eu.entry, ed.entry));     // abs2(ratio), :581
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
            dirty_up |= 1u << wb_accept_warp_c(S, vu, eu, w, 0, K_up, l_up - 1, lane);
            dirty_dn |= 1u << wb_accept_warp_c(S, vd, ed, w, 1, K_dn, l_dn - 1, lane);
            const int knew = max(vu.k + (eu.j < 0), vd.k + (ed.j < 0));
            if (lane == 0 && knew >= S.kth && !S.listed[w]) {       // due for a flush: listed at most once
                S.listed[w] = 1;
                const int fs = atomicAdd(&S.cnt[4], 1);
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

template <bool REPLAY>
__global__ void __launch_bounds__(256)
k_decide_wb_c(DevState S, int gate_refresh, int n_sweeps, const double *__restrict__ rp_r,
              const int *__restrict__ rp_bond, const int *__restrict__ rp_pick) {
    const int w = (int)((blockIdx.x * (unsigned)blockDim.x + threadIdx.x) >> 5);
    const int lane = threadIdx.x & 31;
    if (w >= S.nw) return;
    if (S.flags[w] & KDSL_FLAG_SINGULAR_DEV) return;                // frozen (see k_decide_wb)
    Xoshiro g;
    g.s0 = g.s1 = g.s2 = g.s3 = 0ull;
    if (!REPLAY) {
        const unsigned long long *st = S.rng + (size_t)w * 4;
        g.s0 = st[0]; g.s1 = st[1]; g.s2 = st[2]; g.s3 = st[3];
    }
    unsigned dirty_up = 0u, dirty_dn = 0u;
    for (int s = 0; s < n_sweeps; s++) {
        const size_t off = (size_t)s * S.nw;
        decide_sweep_wb_c<REPLAY>(S, w, lane, gate_refresh, g, REPLAY ? rp_r + off : nullptr,
                                  REPLAY ? rp_bond + off : nullptr, (REPLAY && rp_pick) ? rp_pick + off : nullptr,
                                  dirty_up, dirty_dn);
        __syncwarp();
    }
    wb_copy_rows_c(S, w, dirty_up, dirty_dn, lane);
    if (!REPLAY && lane == 0) {
        unsigned long long *st = S.rng + (size_t)w * 4;
        st[0] = g.s0; st[1] = g.s1; st[2] = g.s2; st[3] = g.s3;
    }
}

// G[m][j] = sign * sum_n T[m][n] (W0[K_n, j] - delta(l_n, j)) for all columns j (all threads of the CTA); sT packed [k][k]
__device__ __forceinline__ void wb_compute_G_cta_c(const WbViewC &v, const cplx *sT, const int *sL, cplx *out, int ldo,
                                                   double sign, int tid, int nthreads) {
    const int k = v.k, N = v.N;
    for (int j = tid; j < N; j += nthreads) {
        for (int m = 0; m < k; m++) {
            cplx acc = c_make(0.0, 0.0);
            for (int n = 0; n < k; n++) {
                cplx rt = v.rows[(size_t)n * N + j];
                if (sL[n] == j) rt.x -= 1.0;
                acc = c_fma(sT[m * k + n], rt, acc);
            }
            out[(size_t)m * ldo + j] = make_double2(sign * acc.x, sign * acc.y);
        }
    }
}

// W0 += C G for the listed walkers; item = (list entry, species, block of RB rows); KMAX = compile-time bound of the pending
// count (register array of the row's C entries).  Dynamic smem: G [KMAX][Np] complex.
template <int KMAX, int RB>
__global__ void __launch_bounds__(RB + 8, 2)
k_flush_c(DevState S, const int *__restrict__ list, const int *__restrict__ count_ptr, int count_fixed,
          int *__restrict__ work_counter, int Np) {
    extern __shared__ __align__(16) unsigned char fcsm[];
    cplx *sG = reinterpret_cast<cplx *>(fcsm);            // [KMAX][Np]
    __shared__ cplx sT[KMAX * KMAX];
    __shared__ int sL[KMAX];
    __shared__ int s_item;
    const int count = count_ptr ? *count_ptr : count_fixed;
    const int ns = S.ns;
    const int nrb = (ns + RB - 1) / RB;
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int total = count * 2 * nrb;
    const cplx zero = c_make(0.0, 0.0);
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
        const WbViewC v = wb_view_c(S, w, spin);
        const int cnt = v.k;
        if (cnt == 0) continue;                           // uniform over the block
        const int N = v.N;
        cplx *W0 = const_cast<cplx *>(v.W0);
        for (int x = tid; x < cnt * cnt; x += nthr) sT[x] = v.T[(x / cnt) * S.kmax + (x % cnt)];
        if (tid < KMAX) sL[tid] = tid < cnt ? v.Ls[tid] : -1;
        __syncthreads();
        wb_compute_G_cta_c(v, sT, sL, sG, Np, -1.0, tid, nthr);
        const int row = rb * RB + tid;
        const bool rok = tid < RB && row < ns;
        cplx c[KMAX];                                     // this row of C = W0[row, l_m]
#pragma unroll
        for (int m = 0; m < KMAX; m++) c[m] = (rok && m < cnt) ? W0[(size_t)sL[m] * ns + row] : zero;
        __syncthreads();                                  // G complete; every C entry is in registers before any store
        if (!rok) continue;
        for (int j0 = 0; j0 < N; j0 += 4) {
            cplx a[4];
#pragma unroll
            for (int q = 0; q < 4; q++) a[q] = j0 + q < N ? W0[(size_t)(j0 + q) * ns + row] : zero;
#pragma unroll
            for (int m = 0; m < KMAX; m++) {
                if (m < cnt) {
#pragma unroll
                    for (int q = 0; q < 4; q++) a[q] = c_fma(c[m], sG[m * Np + min(j0 + q, N - 1)], a[q]);
                }
            }
#pragma unroll
            for (int q = 0; q < 4; q++)
                if (j0 + q < N) W0[(size_t)(j0 + q) * ns + row] = a[q];
        }
    }
}

// k_flush_dmma_c: the same pass W0 += C G on the FP64 TENSOR pipe.  k_flush_c is bound by its shared-memory broadcasts (one
// LDS.128 per complex FMA); here a complex block product is four real DMMAs on split (re, im) operand planes in DMMA
// fragment order,  Dr += Gr Cr - Gi Ci,  Di += Gr Ci + Gi Cr,  in the transposed form of k_flush_wb: the accumulator fragment
// of a lane is (column j, rows 2t, 2t + 1) = 32 contiguous bytes of the interleaved column-major W0, so every lane moves
// two 128-bit words per tile and direction.  Item = (list entry, species, block of RB rows) pulled from a device counter by
// persistent CTAs (one per SM, 16 warps); per item: T and the displaced labels -> G = -T Rt for all columns (FMA, all
// threads) and the rows' entries of C = W0[:, (l_m)] -> shared memory; then warp w streams the column tiles w, w + 16, ...
// over the block's row tiles with its G fragments in registers, D = 4 tiles of loads in flight behind the DMMAs (D = 2: +10 %,
// D = 3: +4 %, D = 6: no better).
// Dynamic smem: G planes 2 [Npad x KP] + C planes 2 [RB x KP] doubles + T [KP x KP] complex + labels.
template <int KP, int D>
__global__ void __launch_bounds__(512, 1)
k_flush_dmma_c(DevState S, const int *__restrict__ list, const int *__restrict__ count_ptr, int count_fixed,
               int *__restrict__ work_counter, int Npad, int RB) {
    constexpr int KS = KP / 4, T_ = 512, NWARPS = 16;
    extern __shared__ __align__(16) unsigned char fdsm[];
    double *sGr = reinterpret_cast<double *>(fdsm);       // [Npad x KP] frag-major (r = column j, k = m): Re G[m][j]
    double *sGi = sGr + (size_t)Npad * KP;
    double *sCr = sGi + (size_t)Npad * KP;                // [RB x KP] frag-major (r = row, k = m): Re W0[row, l_m]
    double *sCi = sCr + (size_t)RB * KP;
    cplx *sT = reinterpret_cast<cplx *>(sCi + (size_t)RB * KP);   // [cnt x cnt] packed
    int *sL = reinterpret_cast<int *>(sT + KP * KP);
    __shared__ int s_item;
    const int count = count_ptr ? *count_ptr : count_fixed;
    const int ns = S.ns;
    const int nrb = (ns + RB - 1) / RB;
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
        const WbViewC v = wb_view_c(S, w, spin);
        const int cnt = v.k;
        if (cnt == 0) continue;                           // uniform over the block
        const int N = v.N;
        const int nks = (cnt + 3) >> 2;                   // k-steps that hold pending updates
        double *W0 = reinterpret_cast<double *>(const_cast<cplx *>(v.W0));
        for (int x = tid; x < cnt * cnt; x += T_) sT[x] = v.T[(x / cnt) * S.kmax + (x % cnt)];
        if (tid < KP) sL[tid] = tid < cnt ? v.Ls[tid] : -1;
        __syncthreads();
        // G[m][j] = -sum_n T[m][n] (W0[K_n, j] - delta(l_n, j)); zero for m >= cnt and for the padding columns.
        // A thread owns column j and half of the m range: every row copy is read once per half.
        for (int idx = tid; idx < 2 * Npad; idx += T_) {
            constexpr int MG = KP / 2;
            const int mg = idx / Npad, j = idx - mg * Npad;
            cplx acc[MG];
#pragma unroll
            for (int q = 0; q < MG; q++) acc[q] = c_make(0.0, 0.0);
            if (j < N && mg * MG < cnt) {
                for (int n = 0; n < cnt; n++) {
                    cplx rt = v.rows[(size_t)n * N + j];
                    if (sL[n] == j) rt.x -= 1.0;
#pragma unroll
                    for (int q = 0; q < MG; q++)
                        if (mg * MG + q < cnt) acc[q] = c_fma(sT[(mg * MG + q) * cnt + n], rt, acc[q]);
                }
            }
#pragma unroll
            for (int q = 0; q < MG; q++) {
                const int m = mg * MG + q;
                if (m < 4 * nks) {
                    const int fi = frag_idx(j, m, KP);
                    sGr[fi] = -acc[q].x;
                    sGi[fi] = -acc[q].y;
                }
            }
        }
        // C[row][m] = W0[row, l_m] for the rows of this block (all of them read before any store of this item)
        const int r0 = rb * RB, nrows = min(RB, ns - r0);
        const int nrt = (nrows + 7) >> 3;
        for (int idx = tid; idx < (nrt << 3) * (4 * nks); idx += T_) {
            const int m = idx / (nrt << 3), r = idx - m * (nrt << 3);
            cplx c = c_make(0.0, 0.0);
            if (m < cnt && r < nrows) c = v.W0[(size_t)sL[m] * ns + r0 + r];
            const int fi = frag_idx(r, m, KP);
            sCr[fi] = c.x;
            sCi[fi] = c.y;
        }
        __syncthreads();
        const int njt = (N + 7) >> 3;
        for (int jt = warp; jt < njt; jt += NWARPS) {
            double g_r[KS], g_i[KS], g_n[KS];
#pragma unroll
            for (int ks = 0; ks < KS; ks++) {
                g_r[ks] = ks < nks ? sGr[(((jt * KS) + ks) << 5) + lane] : 0.0;
                g_i[ks] = ks < nks ? sGi[(((jt * KS) + ks) << 5) + lane] : 0.0;
                g_n[ks] = -g_i[ks];
            }
            const int j = (jt << 3) + gr;
            const bool jok = j < N;
            // lane's 32 bytes of a tile: rows r0 + rt*8 + 2 tg, + 1 of column j (interleaved re, im)
            double *base = W0 + 2 * ((size_t)j * ns + r0 + 2 * tg);
            double2 b0[D], b1[D];
            auto ld = [&](int rt, double2 &x0, double2 &x1) {
                const int rl = min(rt, nrt - 1);
                const int row = (rl << 3) + 2 * tg;
                const double2 *p = reinterpret_cast<const double2 *>(base + 16 * rl);
                x0 = (jok && row < nrows) ? __ldcs(p) : make_double2(0.0, 0.0);
                x1 = (jok && row + 1 < nrows) ? __ldcs(p + 1) : make_double2(0.0, 0.0);
            };
            auto tile = [&](int rt, double2 &x0, double2 &x1) {
                double2 cr = make_double2(x0.x, x1.x), ci = make_double2(x0.y, x1.y);
#pragma unroll
                for (int ks = 0; ks < KS; ks++) {
                    if (ks < nks) {
                        const double br = sCr[(((rt * KS) + ks) << 5) + lane], bi = sCi[(((rt * KS) + ks) << 5) + lane];
                        dmma_8x8x4(cr.x, cr.y, g_r[ks], br);
                        dmma_8x8x4(ci.x, ci.y, g_r[ks], bi);
                        dmma_8x8x4(cr.x, cr.y, g_n[ks], bi);
                        dmma_8x8x4(ci.x, ci.y, g_i[ks], br);
                    }
                }
                const int row = (rt << 3) + 2 * tg;
                double2 *p = reinterpret_cast<double2 *>(base + 16 * rt);
                if (jok && row < nrows) __stcs(p, make_double2(cr.x, ci.x));
                if (jok && row + 1 < nrows) __stcs(p + 1, make_double2(cr.y, ci.y));
            };
#pragma unroll
            for (int i = 0; i < D; i++) ld(i, b0[i], b1[i]);
            int rt0 = 0;
#pragma unroll 1
            for (; rt0 + D <= nrt; rt0 += D) {
#pragma unroll
                for (int i = 0; i < D; i++) {
                    tile(rt0 + i, b0[i], b1[i]);
                    ld(rt0 + i + D, b0[i], b1[i]);
                }
            }
#pragma unroll
            for (int i = 0; i < D - 1; i++)
                if (rt0 + i < nrt) tile(rt0 + i, b0[i], b1[i]);
        }
    }
}
// k_flush_dmma2_c: the same pass with TWO 256-thread CTAs per SM.  k_flush_dmma_c keeps G = -T Rt of the whole species in shared
// memory (83 KB at 432 sites), which leaves room for one CTA per SM -- and nothing streams while that CTA runs an item's prologue.
// Here a warp builds the G fragments of ITS column tile itself, on the tensor pipe, right before it streams the tile:
//   G^T[j][m] = -sum_n Rt[n][j] T[m][n]:  A operand = Rt^T (8 columns j x 24 n, read from the row copies), B operand = T^T (shared
//   memory, fragment order), 72 real DMMAs per column tile (11 % of the tile's 648); the accumulator fragment (j, m = 2t, 2t + 1) is
//   turned into the A-operand fragment (j, m = 4 ks + t) of the streaming loop by two quad shuffles per value.
// Shared memory: C planes [RB x KP] x 2 + T^T planes [KP x KP] x 2 = 92 KB at 432 sites -> two CTAs per SM cover each other's prologues.
template <int KP, int D>
__global__ void __launch_bounds__(256, 2)
k_flush_dmma2_c(DevState S, const int *__restrict__ list, const int *__restrict__ count_ptr, int count_fixed,
                int *__restrict__ work_counter, int RB) {
    constexpr int KS = KP / 4, MT = KP / 8, T_ = 256, NWARPS = 8;
    static_assert(KP % 8 == 0, "pending-update capacity must be a multiple of 8");
    extern __shared__ __align__(16) unsigned char fd2sm[];
    double *sCr = reinterpret_cast<double *>(fd2sm);      // [RB x KP] frag-major (r = row, k = m): Re W0[row, l_m]
    double *sCi = sCr + (size_t)RB * KP;
    double *sTr = sCi + (size_t)RB * KP;                  // [KP x KP] frag-major (r = m, k = n): Re T[m][n] (B operand: k = n, col = m)
    double *sTi = sTr + KP * KP;
    int *sL = reinterpret_cast<int *>(sTi + KP * KP);
    __shared__ int s_item;
    const int count = count_ptr ? *count_ptr : count_fixed;
    const int ns = S.ns;
    const int nrb = (ns + RB - 1) / RB;
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
        const WbViewC v = wb_view_c(S, w, spin);
        const int cnt = v.k;
        if (cnt == 0) continue;                           // uniform over the block
        const int N = v.N;
        const int nks = (cnt + 3) >> 2;                   // k-steps that hold pending updates
        double *W0 = reinterpret_cast<double *>(const_cast<cplx *>(v.W0));
        for (int x = tid; x < KP * KP; x += T_) {
            const int m = x / KP, n = x - m * KP;
            const cplx t = (m < cnt && n < cnt) ? v.T[m * S.kmax + n] : c_make(0.0, 0.0);
            const int fi = frag_idx(m, n, KP);
            sTr[fi] = t.x;
            sTi[fi] = t.y;
        }
        if (tid < KP) sL[tid] = tid < cnt ? v.Ls[tid] : -1;
        __syncthreads();
        // C[row][m] = W0[row, l_m] for the rows of this block (all of them read before any store of this item)
        const int r0 = rb * RB, nrows = min(RB, ns - r0);
        const int nrt = (nrows + 7) >> 3;
        for (int idx = tid; idx < (nrt << 3) * (4 * nks); idx += T_) {
            const int m = idx / (nrt << 3), r = idx - m * (nrt << 3);
            cplx c = c_make(0.0, 0.0);
            if (m < cnt && r < nrows) c = v.W0[(size_t)sL[m] * ns + r0 + r];
            const int fi = frag_idx(r, m, KP);
            sCr[fi] = c.x;
            sCi[fi] = c.y;
        }
        __syncthreads();
        const int njt = (N + 7) >> 3;
        for (int jt = warp; jt < njt; jt += NWARPS) {
            const int j = (jt << 3) + gr;
            const bool jok = j < N;
            // ---- G^T fragments of this column tile: D[j][m] = sum_n Rt[n][j] T[m][n] (complex), then negated ----
            double dr[MT][2], di[MT][2];
#pragma unroll
            for (int mt = 0; mt < MT; mt++) dr[mt][0] = dr[mt][1] = di[mt][0] = di[mt][1] = 0.0;
#pragma unroll
            for (int ks = 0; ks < KS; ks++) {
                if (ks < nks) {
                    const int n = 4 * ks + tg;
                    cplx rt = (jok && n < cnt) ? v.rows[(size_t)n * N + j] : c_make(0.0, 0.0);
                    if (n < cnt && sL[n] == j) rt.x -= 1.0;
                    const double an = -rt.y;
#pragma unroll
                    for (int mt = 0; mt < MT; mt++) {
                        const double br = sTr[(((mt * KS) + ks) << 5) + lane], bi = sTi[(((mt * KS) + ks) << 5) + lane];
                        dmma_8x8x4(dr[mt][0], dr[mt][1], rt.x, br);
                        dmma_8x8x4(di[mt][0], di[mt][1], rt.x, bi);
                        dmma_8x8x4(dr[mt][0], dr[mt][1], an, bi);
                        dmma_8x8x4(di[mt][0], di[mt][1], rt.y, br);
                    }
                }
            }
            // accumulator fragment (j, m = 8 mt + 2 t, + 1)  ->  A-operand fragment (j, m = 4 ks + t), G = -D
            double g_r[KS], g_i[KS], g_n[KS];
#pragma unroll
            for (int mt = 0; mt < MT; mt++) {
#pragma unroll
                for (int h = 0; h < 2; h++) {             // ks = 2 mt + h: m = 8 mt + 4 h + t lives in lane 2 h + (t >> 1), slot t & 1
                    const int src = (lane & ~3) | (2 * h + (tg >> 1));
                    const double r0v = __shfl_sync(0xffffffffu, dr[mt][0], src), r1v = __shfl_sync(0xffffffffu, dr[mt][1], src);
                    const double i0v = __shfl_sync(0xffffffffu, di[mt][0], src), i1v = __shfl_sync(0xffffffffu, di[mt][1], src);
                    g_r[2 * mt + h] = -((tg & 1) ? r1v : r0v);
                    g_i[2 * mt + h] = -((tg & 1) ? i1v : i0v);
                    g_n[2 * mt + h] = -g_i[2 * mt + h];
                }
            }
            // ---- stream the column tile down the block's row tiles ----
            double *base = W0 + 2 * ((size_t)j * ns + r0 + 2 * tg);
            double2 b0[D], b1[D];
            auto ld = [&](int rt, double2 &x0, double2 &x1) {
                const int rl = min(rt, nrt - 1);
                const int row = (rl << 3) + 2 * tg;
                const double2 *p = reinterpret_cast<const double2 *>(base + 16 * rl);
                x0 = (jok && row < nrows) ? __ldcs(p) : make_double2(0.0, 0.0);
                x1 = (jok && row + 1 < nrows) ? __ldcs(p + 1) : make_double2(0.0, 0.0);
            };
            auto tile = [&](int rt, double2 &x0, double2 &x1) {
                double2 cr = make_double2(x0.x, x1.x), ci = make_double2(x0.y, x1.y);
#pragma unroll
                for (int ks = 0; ks < KS; ks++) {
                    if (ks < nks) {
                        const double br = sCr[(((rt * KS) + ks) << 5) + lane], bi = sCi[(((rt * KS) + ks) << 5) + lane];
                        dmma_8x8x4(cr.x, cr.y, g_r[ks], br);
                        dmma_8x8x4(ci.x, ci.y, g_r[ks], bi);
                        dmma_8x8x4(cr.x, cr.y, g_n[ks], bi);
                        dmma_8x8x4(ci.x, ci.y, g_i[ks], br);
                    }
                }
                const int row = (rt << 3) + 2 * tg;
                double2 *p = reinterpret_cast<double2 *>(base + 16 * rt);
                if (jok && row < nrows) __stcs(p, make_double2(cr.x, ci.x));
                if (jok && row + 1 < nrows) __stcs(p + 1, make_double2(cr.y, ci.y));
            };
#pragma unroll
            for (int i = 0; i < D; i++) ld(i, b0[i], b1[i]);
            int rt0 = 0;
#pragma unroll 1
            for (; rt0 + D <= nrt; rt0 += D) {
#pragma unroll
                for (int i = 0; i < D; i++) {
                    tile(rt0 + i, b0[i], b1[i]);
                    ld(rt0 + i + D, b0[i], b1[i]);
                }
            }
#pragma unroll
            for (int i = 0; i < D - 1; i++)
                if (rt0 + i < nrt) tile(rt0 + i, b0[i], b1[i]);
        }
    }
}
inline size_t flush_dmma2_c_smem(int KP, int RB) {
    return ((size_t)2 * RB * KP + (size_t)2 * KP * KP) * sizeof(double) + (size_t)KP * sizeof(int);
}
inline size_t flush_dmma_c_smem(int KP, int Npad, int RB) {
    return ((size_t)2 * Npad * KP + (size_t)2 * RB * KP + (size_t)2 * KP * KP) * sizeof(double) + (size_t)KP * sizeof(int);
}

// O_L (getOL, src/Hamiltonian.jl:762-778) with complex Woodbury-form W; one CTA (256 threads) per walker
__global__ void __launch_bounds__(256)
k_measure_wb_c(DevState S, double *__restrict__ ol_out, int accumulate) {
    extern __shared__ __align__(16) unsigned char mcsm[];
    const int w = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ns = S.ns, kmax = S.kmax;
    const int *kup = S.kup + (size_t)w * ns;
    const int *kdn = S.kdn + (size_t)w * ns;
    WbViewC v[2] = {wb_view_c(S, w, 0), wb_view_c(S, w, 1)};
    cplx *sG[2], *sT[2];
    int *sL[2];
    sG[0] = reinterpret_cast<cplx *>(mcsm);
    sG[1] = sG[0] + (size_t)kmax * S.n_up;
    sT[0] = sG[1] + (size_t)kmax * S.n_dn;
    sT[1] = sT[0] + kmax * kmax;
    sL[0] = reinterpret_cast<int *>(sT[1] + kmax * kmax);
    sL[1] = sL[0] + kmax;
    __shared__ double red_f[8];
    __shared__ int red_d[8], red_b[8];
#pragma unroll
    for (int s = 0; s < 2; s++) {
        const int k = v[s].k;
        for (int x = tid; x < k * k; x += 256) sT[s][x] = v[s].T[(x / k) * kmax + (x % k)];
        if (tid < k) sL[s][tid] = v[s].Ls[tid];
    }
    __syncthreads();
#pragma unroll
    for (int s = 0; s < 2; s++)
        if (v[s].k > 0) wb_compute_G_cta_c(v[s], sT[s], sL[s], sG[s], v[s].N, 1.0, tid, 256);
    __syncthreads();
    auto entry = [&](int s, int K, int l) {
        cplx e = v[s].W0[(size_t)l * ns + K];
        const int k = v[s].k;
        for (int m = 0; m < k; m++) e = c_fma(c_neg(v[s].W0[(size_t)sL[s][m] * ns + K]), sG[s][(size_t)m * v[s].N + l], e);
        return e;
    };
    double flips = 0.0;
    int diag4 = 0, bad = 0;
    for (int b = tid; b < S.n_bonds; b += 256) {
        const int i = S.bi[b], j = S.bj[b];
        const int iu = kup[i], ju = kup[j], id = kdn[i], jd = kdn[j];
        if (ju != 0 && id != 0) flips += -0.5 * c_mul(entry(0, i, ju - 1), entry(1, j, id - 1)).x;
        if (iu != 0 && jd != 0) flips += -0.5 * c_mul(entry(0, j, iu - 1), entry(1, i, jd - 1)).x;
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
        if (accumulate && !(S.flags[w] & KDSL_FLAG_SINGULAR_DEV)) {
            S.ol_last[w] = OL;
            S.ol_sum[w] += OL;
            S.ol_sq[w] += OL * OL;
            S.ol_n[w] += 1ull;
        }
    }
}
