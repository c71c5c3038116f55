// kdsl_resident.cuh -- small lattices (BASELINE config 2: 6x6 DoubleKagome, 108 sites): ONE persistent kernel per
// kdsl_sweep / kdsl_replay call, the whole walker resident in shared memory (update_variant 3).
//
// Walkers are independent Markov chains, so nothing forces lock step: a CTA takes a walker, runs ALL the sweeps of the
// call on it -- proposals (src/MonteCarlo.jl:538-607), the accepted-move Sherman-Morrison updates (:279-292), the
// periodic reevaluateW! (:55-66, :594-604) and the O_L measurements at the Carlo cadence (:628-634,
// src/Hamiltonian.jl:762-778) -- and only then writes W, kappa, Z_mu, the RNG state and the accumulators back.
// HBM sees each walker twice per call; every inner-loop access is shared memory.
//
// Compact W.  In a Mott state a species occupies N of the ns sites and W[R_l, :] = e_l on them, and every
// entry the path ever reads -- the ratio W[K, l] of a proposal, the flip terms of O_L -- has K on a site the species
// does NOT occupy.  So only the M = ns - N rows of the unoccupied sites are kept:  Wc[l][u] = W[site(u), l]
// (N x M doubles per species, 2 N_up N_dn doubles per walker = 46.7 KB at 108 sites instead of 93.3 KB, which is what
// lets three CTAs share an SM).  An accepted move "label l: R -> K" makes K occupied (its row becomes e_l) and R
// unoccupied; the new row of R,  W'[R, j] = delta_lj + 1.0 * temp_j  (its old row was e_l), takes over K's slot.  The
// arithmetic per entry is the reference's geru order (temp_j = alpha (W[K, j] - delta_lj), A += x temp).
//
// reevaluateW! in the same storage: Gauss-Jordan with row pivoting on [tilde_U^T | V^T] (the idea of k_reeval_fused)
// turns the V^T block -- which IS the compact layout -- into W in place; the tilde_U^T block lives in an N x N scratch.
//
// Thread roles: warp 0 runs the proposals (all lanes redundantly, like k_decide_wb; Xoshiro state in registers) and
// keeps going through rejected sweeps without any barrier; the CTA meets at a barrier only for an EVENT: an accepted
// move (all threads apply the update), a re-evaluation, a measurement, or the end of the call.
#pragma once
#include "kdsl_common.cuh"
#include "kdsl_propose.cuh"
#include "kdsl_refresh.cuh"

struct ResParams {
    int n_sweeps;                // sweeps of this call
    long long sweep0;            // ctx.sweeps before the first of them
    long long period;            // re-evaluation cadence (n_occ unless "refresh_every" is set)
    long long therm;             // measure when ctx.sweeps > therm (post-increment), < 0: never
    const double *rp_r;          // replay inputs [n_sweeps][nw] (REPLAY only)
    const int *rp_bond, *rp_pick;
    int *work_counter;           // dynamic walker assignment
};

#define RES_EV_UPDATE 1
#define RES_EV_REFRESH 2
#define RES_EV_MEASURE 4
#define RES_EV_DONE 8

struct ResCtrl {                 // written by warp 0 before the event barrier
    int ev;
    int l[2], q[2];              // per species: moved label (0-based), compact slot of the target site
    int singular;
    int pad;
};

struct ResSmem {
    double *Wc[2];               // compact W: [N][M]
    double *T;                   // [Nmax][Nmax] re-evaluation scratch (tilde_U^T block)
    double *stg;                 // staging: update (temp[N], col[M] per species) / Gauss-Jordan (prow, krow [N+M], fcol [N])
    short *kap[2];               // kappa per species [ns]
    short *slot[2];              // [ns] compact slot of a site (-1: occupied by the species)
    short *site[2];              // [M] site of a slot
    ResCtrl *ctrl;
    double *red;                 // [8] reduction scratch
    __device__ __forceinline__ ResSmem(unsigned char *base, int ns, int n_up, int n_dn) {
        const int Nmax = max(n_up, n_dn);
        double *d = reinterpret_cast<double *>(base);
        Wc[0] = d; d += (size_t)n_up * (ns - n_up);
        Wc[1] = d; d += (size_t)n_dn * (ns - n_dn);
        T = d; d += (size_t)Nmax * Nmax;
        stg = d; d += 2 * ns + Nmax;
        red = d; d += 8;
        ctrl = reinterpret_cast<ResCtrl *>(d); d += 4;
        short *s = reinterpret_cast<short *>(d);
        kap[0] = s; s += ns; kap[1] = s; s += ns;
        slot[0] = s; s += ns; slot[1] = s; s += ns;
        site[0] = s; s += ns - n_up; site[1] = s; s += ns - n_dn;
    }
};
static_assert(sizeof(ResCtrl) <= 32, "ResCtrl must fit its reserved slot");

__host__ __device__ inline size_t resident_smem_bytes(int ns, int n_up, int n_dn) {
    const int Nmax = n_up > n_dn ? n_up : n_dn;
    size_t d = (size_t)n_up * (ns - n_up) + (size_t)n_dn * (ns - n_dn) + (size_t)Nmax * Nmax + 2 * ns + Nmax + 8 + 4;
    size_t s = (size_t)4 * ns + (ns - n_up) + (ns - n_dn);
    return d * sizeof(double) + ((s * sizeof(short) + 15) & ~(size_t)15);
}

// ---- tables of one species from kappa: slots of the unoccupied sites in ascending site order (all T threads) ----
template <int T>
__device__ __forceinline__ void res_build_tables(const ResSmem &L, int sp, int ns, int *s_scan) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NW = T / 32;
    int base = 0;
    for (int s0 = 0; s0 < ns; s0 += T) {
        const int st = s0 + tid;
        const bool un = st < ns && L.kap[sp][st] == 0;
        const unsigned m = __ballot_sync(0xffffffffu, un);
        if (lane == 0) s_scan[warp] = __popc(m);
        __syncthreads();
        int off = base, tot = 0;
        for (int q = 0; q < NW; q++) {
            const int c = s_scan[q];
            if (q < warp) off += c;
            tot += c;
        }
        if (st < ns) {
            if (un) {
                const int u = off + __popc(m & ((1u << lane) - 1u));
                L.slot[sp][st] = (short)u;
                L.site[sp][u] = (short)st;
            } else {
                L.slot[sp][st] = -1;
            }
        }
        base += tot;
        __syncthreads();
    }
}

// ---- reevaluateW! of one species in the compact storage (all T threads).  Returns false when tilde_U is singular. ----
template <int T>
__device__ __noinline__ bool res_reevaluate(const DevState &S, const ResSmem &L, int sp, int *s_scan) {
    const int tid = threadIdx.x, lane = tid & 31;
    const int ns = S.ns, N = sp ? S.n_dn : S.n_up, M = ns - N;
    const double *U = sp ? S.U_dn : S.U_up;
    double *Tm = L.T, *Wc = L.Wc[sp];
    short *lab_site = reinterpret_cast<short *>(L.stg);          // [N] site of label c (scratch, dead before the elimination)
    res_build_tables<T>(L, sp, ns, s_scan);
    for (int st = tid; st < ns; st += T) {
        const int l = L.kap[sp][st];
        if (l != 0) lab_site[l - 1] = (short)st;
    }
    __syncthreads();
    // B = [tilde_U^T | V^T]: B[j][c] = U[site(c), j]
    for (int idx = tid; idx < N * ns; idx += T) {
        const int j = idx / ns, c = idx - j * ns;
        if (c < N) Tm[j * N + c] = __ldg(U + (size_t)j * ns + lab_site[c]);
        else Wc[j * M + (c - N)] = __ldg(U + (size_t)j * ns + L.site[sp][c - N]);
    }
    __syncthreads();
    double *prow = L.stg, *krow = L.stg + ns, *fcol = L.stg + 2 * ns;   // [N+M], [N+M], [N]
    int *s_piv = s_scan + 16;                                   // [0] pivot row, [1] singular
    for (int k = 0; k < N; k++) {
        // pivot search over rows k .. N-1 of column k (warp 0)
        if (tid < 32) {
            double best = -1.0;
            int bi = -1;
            for (int i = k + lane; i < N; i += 32) {
                const double v = fabs(Tm[i * N + k]);
                if (v > best || !(v == v)) { best = (v == v) ? v : 1.0 / 0.0; bi = i; }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const double ob = __shfl_xor_sync(0xffffffffu, best, o);
                const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                if (ob > best || (ob == best && oi >= 0 && (bi < 0 || oi < bi))) { best = ob; bi = oi; }
            }
            if (lane == 0) {
                s_piv[0] = bi;
                s_piv[1] = !(best > 0.0) || best > 1.79e308;
            }
        }
        __syncthreads();
        if (s_piv[1]) return false;
        const int p = s_piv[0];
        // stage: prow = row p / pivot (the new row k), krow = old row k (the new row p before elimination), fcol = column k
        // with rows k and p exchanged.  Unfinished columns only: c > k of the tilde_U^T block, all of the V^T block.
        {
            const double inv = 1.0 / Tm[p * N + k];
            for (int c = k + 1 + tid; c < N + M; c += T) {
                const double pv = c < N ? Tm[p * N + c] : Wc[p * M + (c - N)];
                const double kv = c < N ? Tm[k * N + c] : Wc[k * M + (c - N)];
                prow[c] = pv * inv;
                krow[c] = kv;
            }
            for (int i = tid; i < N; i += T) fcol[i] = Tm[(i == p ? k : i == k ? p : i) * N + k];
        }
        __syncthreads();
        // eliminate: row i <- row i' - fcol[i] * prow (i' = i with k <-> p exchanged); row k <- prow
        {
            const int nc = N + M - (k + 1);
            for (int idx = tid; idx < N * nc; idx += T) {
                const int i = idx / nc, c = k + 1 + (idx - i * nc);
                double *dst = c < N ? Tm + i * N + c : Wc + i * M + (c - N);
                double v;
                if (i == k) v = prow[c];
                else {
                    const double old = (i == p) ? krow[c] : *dst;
                    v = fma(-fcol[i], prow[c], old);
                }
                *dst = v;
            }
        }
        __syncthreads();
    }
    return true;
}

// ---- O_L (getOL) from the compact storage, all T threads; result valid on thread 0 ----
template <int T>
__device__ __forceinline__ double res_measure(const DevState &S, const ResSmem &L, int *bad_out) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int Mu = S.ns - S.n_up, Md = S.ns - S.n_dn;
    double flips = 0.0;
    int diag4 = 0, bad = 0;
    for (int b = tid; b < S.n_bonds; b += T) {
        const int i = __ldg(S.bi + b), j = __ldg(S.bj + b);
        const int iu = L.kap[0][i], ju = L.kap[0][j], id = L.kap[1][i], jd = L.kap[1][j];
        const int oi = (iu != 0) + (id != 0), oj = (ju != 0) + (jd != 0);
        if (oi != 1 || oj != 1) { bad = 1; continue; }
        if (ju != 0 && id != 0)                                    // W_up[i, ju] W_dn[j, id]
            flips += -0.5 * L.Wc[0][(ju - 1) * Mu + L.slot[0][i]] * L.Wc[1][(id - 1) * Md + L.slot[1][j]];
        if (iu != 0 && jd != 0)                                    // W_up[j, iu] W_dn[i, jd]
            flips += -0.5 * L.Wc[0][(iu - 1) * Mu + L.slot[0][j]] * L.Wc[1][(jd - 1) * Md + L.slot[1][i]];
        diag4 += (iu != 0 ? 1 : -1) * (ju != 0 ? 1 : -1);
    }
    flips = warp_sum_f64(flips);
    diag4 = warp_sum_int(diag4);
    bad = warp_sum_int(bad);
    double *red = L.red;
    int *redi = reinterpret_cast<int *>(L.stg);                  // (staging is idle during a measurement)
    if (lane == 0) { red[warp] = flips; redi[warp] = diag4; redi[8 + warp] = bad; }
    __syncthreads();
    double OL = 0.0;
    if (tid == 0) {
        double f = 0.0;
        int d4 = 0, bd = 0;
        for (int q = 0; q < T / 32; q++) { f += red[q]; d4 += redi[q]; bd += redi[8 + q]; }
        OL = f + 0.25 * (double)d4;
        *bad_out = bd;
    }
    __syncthreads();
    return OL;
}

template <bool REPLAY>
__global__ void __launch_bounds__(256, 3)
k_resident(DevState S, ResParams P) {
    constexpr int T = 256;
    extern __shared__ __align__(16) unsigned char res_sm[];
    __shared__ int s_scan[24];
    __shared__ int s_w;
    const ResSmem L(res_sm, S.ns, S.n_up, S.n_dn);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ns = S.ns;
    const int Nn[2] = {S.n_up, S.n_dn}, Mm[2] = {ns - S.n_up, ns - S.n_dn};

    for (;;) {
        __syncthreads();
        if (tid == 0) s_w = atomicAdd(P.work_counter, 1);
        __syncthreads();
        const int w = s_w;
        if (w >= S.nw) break;
        if (S.flags[w] & KDSL_FLAG_SINGULAR_DEV) continue;       // frozen (see k_decide_wb)

        // ---- load the walker ----
        for (int st = tid; st < ns; st += T) {
            L.kap[0][st] = (short)S.kup[(size_t)w * ns + st];
            L.kap[1][st] = (short)S.kdn[(size_t)w * ns + st];
        }
        __syncthreads();
        res_build_tables<T>(L, 0, ns, s_scan);
        res_build_tables<T>(L, 1, ns, s_scan);
#pragma unroll
        for (int sp = 0; sp < 2; sp++) {
            const double *Wg = (sp ? S.W_dn : S.W_up) + (size_t)w * ns * Nn[sp];
            const int M = Mm[sp];
            for (int idx = tid; idx < Nn[sp] * M; idx += T) {
                const int l = idx / M, u = idx - l * M;
                L.Wc[sp][idx] = Wg[(size_t)l * ns + L.site[sp][u]];
            }
        }
        if (tid == 0) { L.ctrl->singular = 0; L.ctrl->ev = 0; }
        __syncthreads();

        // ---- per-walker scalars: live in warp 0 (all lanes identical) ----
        Xoshiro g;
        g.s0 = g.s1 = g.s2 = g.s3 = 0ull;
        int zmu = 0, s_done = 0;
        unsigned long long c_acc = 0ull, c_reach = 0ull, c_refresh = 0ull, c_ol = 0ull;
        double a_ol = 0.0, a_ol2 = 0.0, last_ol = 0.0;
        bool have_ol = false, dead = false;
        if (warp == 0) {
            zmu = S.zmu[w];
            if (!REPLAY) {
                const unsigned long long *st = S.rng + (size_t)w * 4;
                g.s0 = st[0]; g.s1 = st[1]; g.s2 = st[2]; g.s3 = st[3];
            }
        }

        for (;;) {
            if (warp == 0) {
                int ev = 0;
                while (s_done < P.n_sweeps) {
                    const long long sw = P.sweep0 + s_done;
                    const bool gate = (sw % P.period) == 0;                     // :595 (pre-increment)
                    const size_t ro = (size_t)s_done * S.nw + w;
                    const double r = REPLAY ? P.rp_r[ro] : g.rand_f64();        // :546
                    const double zr = (double)zmu / (double)S.n_bonds;
                    bool accepted = false, reached = false;
                    if (!(r > zr)) {                                            // :547-550
                        long long b = REPLAY ? (long long)P.rp_bond[ro] : g.rand_index((unsigned long long)S.n_bonds);
                        if (b < 1) b = 1;
                        if (b > S.n_bonds) b = S.n_bonds;
                        const int i = __ldg(S.bi + (b - 1)), site = __ldg(S.bj + (b - 1));
                        const int ku_i = L.kap[0][i], ku_s = L.kap[0][site], kd_i = L.kap[1][i], kd_s = L.kap[1][site];
                        const bool f1 = ku_i != 0 && kd_s != 0;                 // :558-561
                        const bool f2 = ku_s != 0 && kd_i != 0;
                        if (f1 || f2) {
                            const int nm = (int)f1 + (int)f2;
                            long long pick;                                     // :569
                            if (REPLAY) pick = P.rp_pick ? (long long)P.rp_pick[ro] : 1;
                            else pick = g.rand_index((unsigned long long)nm);
                            const int flag = (f1 && f2) ? (pick == 1 ? 1 : 2) : (f1 ? 1 : 2);
                            const int l_up = flag == 1 ? ku_i : ku_s, l_dn = flag == 1 ? kd_s : kd_i;   // :572-573
                            const int K_up = flag == 1 ? site : i, K_dn = flag == 1 ? i : site;
                            // a doubly occupied target site has no compact row (cannot happen in a Mott state)
                            const int q_up = L.slot[0][K_up], q_dn = L.slot[1][K_dn];
                            const double wu = q_up >= 0 ? L.Wc[0][(l_up - 1) * Mm[0] + q_up] : (L.kap[0][K_up] == l_up ? 1.0 : 0.0);
                            const double wd = q_dn >= 0 ? L.Wc[1][(l_dn - 1) * Mm[1] + q_dn] : (L.kap[1][K_dn] == l_dn ? 1.0 : 0.0);
                            const double ratio = wu * wd;                       // :576-580
                            const double p = ratio * ratio;                     // abs2(ratio)
                            if (p >= 1.0 && r < zr) accepted = true;            // :582-587
                            else if (p < 1.0 && r < zr * p) accepted = true;
                            if (!(p == p) || p > 1.79e308) {
                                if (lane == 0) atomicOr(&S.flags[w], KDSL_FLAG_NONFINITE_DEV);
                            }
                            reached = true;
                            if (accepted) {
                                if (!gate) {
                                    // stage the Sherman-Morrison operands of both species (update_W!, :286-289):
                                    // temp_j = alpha (W[K, j] - delta_lj), col_u = W[site(u), l]
#pragma unroll
                                    for (int sp = 0; sp < 2; sp++) {
                                        const int N = Nn[sp], M = Mm[sp];
                                        const int l = (sp ? l_dn : l_up) - 1, q = sp ? q_dn : q_up;
                                        const double alpha = -1.0 / (sp ? wd : wu);
                                        double *temp = L.stg + (sp ? S.n_up + Mm[0] : 0), *col = temp + N;
                                        for (int j = lane; j < N; j += 32) {
                                            double v = L.Wc[sp][j * M + q];
                                            if (j == l) v -= 1.0;
                                            temp[j] = alpha * v;
                                        }
                                        for (int u = lane; u < M; u += 32) col[u] = L.Wc[sp][l * M + u];
                                        if (lane == 0) { L.ctrl->l[sp] = l; L.ctrl->q[sp] = q; }
                                    }
                                    ev |= RES_EV_UPDATE;
                                }
                                // Z_mu through the bonds incident to i or site (equals the full recount, :460-474)
                                const int ui_o = ku_i != 0, di_o = kd_i != 0, us_o = ku_s != 0, ds_o = kd_s != 0;
                                const int ui_n = flag == 1 ? 0 : 1, di_n = flag == 1 ? 1 : 0;
                                const int us_n = flag == 1 ? 1 : 0, ds_n = flag == 1 ? 0 : 1;
                                int delta = 0;
                                for (int qq = __ldg(S.adj_off + i) + lane; qq < __ldg(S.adj_off + i + 1); qq += 32) {
                                    const int n = __ldg(S.adj_nbr + qq);
                                    if (n == site) continue;
                                    const int un = L.kap[0][n] != 0, dn = L.kap[1][n] != 0;
                                    delta += bond_is_anti(ui_n, di_n, un, dn) - bond_is_anti(ui_o, di_o, un, dn);
                                }
                                for (int qq = __ldg(S.adj_off + site) + lane; qq < __ldg(S.adj_off + site + 1); qq += 32) {
                                    const int n = __ldg(S.adj_nbr + qq);
                                    if (n == i) continue;
                                    const int un = L.kap[0][n] != 0, dn = L.kap[1][n] != 0;
                                    delta += bond_is_anti(us_n, ds_n, un, dn) - bond_is_anti(us_o, ds_o, un, dn);
                                }
                                if (lane == 0)
                                    delta += bond_is_anti(ui_n, di_n, us_n, ds_n) - bond_is_anti(ui_o, di_o, us_o, ds_o);
                                zmu += warp_sum_int(delta);
                                __syncwarp();
                                if (lane == 0) {
                                    if (flag == 1) {                            // :502-503
                                        L.kap[0][i] = 0; L.kap[0][site] = (short)l_up;
                                        L.kap[1][i] = (short)l_dn; L.kap[1][site] = 0;
                                    } else {                                    // :508-509
                                        L.kap[0][i] = (short)l_up; L.kap[0][site] = 0;
                                        L.kap[1][i] = 0; L.kap[1][site] = (short)l_dn;
                                    }
                                    if (!gate) {
                                        // the vacated site takes over the slot of the newly occupied one
                                        const int R_up = flag == 1 ? i : site, R_dn = flag == 1 ? site : i;
                                        L.slot[0][R_up] = (short)q_up; L.slot[0][K_up] = -1; L.site[0][q_up] = (short)R_up;
                                        L.slot[1][R_dn] = (short)q_dn; L.slot[1][K_dn] = -1; L.site[1][q_dn] = (short)R_dn;
                                    }
                                }
                                __syncwarp();
                                c_acc += 1ull;
                            }
                        }
                    }
                    if (reached) {
                        c_reach += 1ull;
                        if (gate) ev |= RES_EV_REFRESH;                         // :594-604
                    }
                    s_done++;
                    if (P.therm >= 0 && sw + 1 > P.therm && ((sw + 1) % S.n_occ) == 0) ev |= RES_EV_MEASURE;   // :630
                    if (ev) break;
                }
                if (s_done >= P.n_sweeps) ev |= RES_EV_DONE;
                if (lane == 0) L.ctrl->ev = ev;
            }
            __syncthreads();
            const int ev = L.ctrl->ev;
            if (ev & RES_EV_UPDATE) {
#pragma unroll
                for (int sp = 0; sp < 2; sp++) {
                    const int N = Nn[sp], M = Mm[sp];
                    const int l = L.ctrl->l[sp], q = L.ctrl->q[sp];
                    const double *temp = L.stg + (sp ? S.n_up + Mm[0] : 0), *col = temp + N;
                    double *Wc = L.Wc[sp];
                    for (int idx = tid; idx < N * M; idx += T) {
                        const int j = idx / M, u = idx - j * M;
                        const double tj = temp[j];
                        // slot q now stands for the vacated site, whose old row was e_l (x = 1, A = delta_lj)
                        Wc[idx] = (u == q) ? fma(1.0, tj, j == l ? 1.0 : 0.0) : fma(col[u], tj, Wc[idx]);
                    }
                }
                __syncthreads();
            }
            if (ev & RES_EV_REFRESH) {
                bool ok = res_reevaluate<T>(S, L, 0, s_scan);
                __syncthreads();
                if (ok) ok = res_reevaluate<T>(S, L, 1, s_scan);
                __syncthreads();
                if (!ok) {
                    if (tid == 0) {
                        atomicOr(&S.flags[w], KDSL_FLAG_SINGULAR_DEV);
                        atomicAdd(&S.cnt[3], 1);
                    }
                    dead = true;
                } else {
                    c_refresh += 1ull;
                }
            }
            if (dead) break;
            if (ev & RES_EV_MEASURE) {
                int bad = 0;
                const double OL = res_measure<T>(S, L, &bad);
                if (tid == 0) {
                    if (bad) atomicOr(&S.flags[w], 4);
                    a_ol += OL; a_ol2 += OL * OL; c_ol += 1ull; last_ol = OL; have_ol = true;
                }
            }
            if (ev & RES_EV_DONE) break;
        }

        // ---- write the walker back ----
        __syncthreads();
        for (int st = tid; st < ns; st += T) {
            S.kup[(size_t)w * ns + st] = L.kap[0][st];
            S.kdn[(size_t)w * ns + st] = L.kap[1][st];
        }
        if (!dead) {
#pragma unroll
            for (int sp = 0; sp < 2; sp++) {
                double *Wg = (sp ? S.W_dn : S.W_up) + (size_t)w * ns * Nn[sp];
                const int M = Mm[sp];
                for (int idx = tid; idx < Nn[sp] * ns; idx += T) {
                    const int l = idx / ns, st = idx - l * ns;
                    const int u = L.slot[sp][st];
                    Wg[idx] = u >= 0 ? L.Wc[sp][l * M + u] : (L.kap[sp][st] == l + 1 ? 1.0 : 0.0);
                }
            }
        }
        if (tid == 0) {
            S.zmu[w] = zmu;
            if (!REPLAY) {
                unsigned long long *st = S.rng + (size_t)w * 4;
                st[0] = g.s0; st[1] = g.s1; st[2] = g.s2; st[3] = g.s3;
            }
            S.n_acc[w] += c_acc;
            S.n_reach[w] += c_reach;
            S.n_refresh[w] += c_refresh;
            if (have_ol) {
                S.ol_sum[w] += a_ol; S.ol_sq[w] += a_ol2; S.ol_n[w] += c_ol; S.ol_last[w] = last_ol;
            }
        }
    }
}
