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
#include "kdsl_measure.cuh"
#include "kdsl_refresh.cuh"
#include <type_traits>

struct ResParams {
    int n_sweeps;                // sweeps of this call
    long long sweep0;            // ctx.sweeps before the first of them
    long long period;            // re-evaluation cadence (n_occ unless "refresh_every" is set)
    long long phase_p, phase_m;  // sweep0 mod period, sweep0 mod n_occ (the kernel only counts from here: no 64-bit division)
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

// Shared-memory carve-up as OFFSETS from the dynamic shared array (file scope, so that every access is derived from the
// __shared__ symbol: with pointers kept in a struct the struct went to the stack under register pressure and every
// access through it became a generic LD.E / ST.E instead of LDS / STS).
extern __shared__ __align__(16) unsigned char res_sm[];
__device__ __forceinline__ double *res_d(int off) { return reinterpret_cast<double *>(res_sm) + off; }
__device__ __forceinline__ short *res_h(int off) { return reinterpret_cast<short *>(res_sm) + off; }
struct ResSmem {
    int oWc[2], oT, oStg, oRed, oCtrl, oBond;    // in doubles
    int hKap[2], hSlot[2], hSite[2], hAdjOff, hAdjNbr;   // in shorts
    __device__ __forceinline__ ResSmem(int ns, int n_up, int n_dn, int n_bonds) {
        const int Nmax = max(n_up, n_dn);
        int d = 0;
        oWc[0] = d; d += n_up * (ns - n_up);
        oWc[1] = d; d += n_dn * (ns - n_dn);
        oT = d; d += Nmax * Nmax;
        oStg = d; d += 2 * ns + Nmax;
        oRed = d; d += 8;
        oCtrl = d; d += 4;
        oBond = d; d += (n_bonds + 1) / 2;
        int h = 4 * d;
        hKap[0] = h; h += ns; hKap[1] = h; h += ns;
        hSlot[0] = h; h += ns; hSlot[1] = h; h += ns;
        hSite[0] = h; h += ns - n_up; hSite[1] = h; h += ns - n_dn;
        hAdjOff = h; h += ns + 1;
        hAdjNbr = h; h += 2 * n_bonds;
    }
    // CSR adjacency of the bond graph (the Z_mu update of an accepted move walks the neighbours of two sites)
    __device__ __forceinline__ short *adj_off() const { return res_h(hAdjOff); }
    __device__ __forceinline__ short *adj_nbr() const { return res_h(hAdjNbr); }
    __device__ __forceinline__ double *Wc(int sp) const { return res_d(oWc[sp]); }   // compact W: [N][M]
    __device__ __forceinline__ double *T() const { return res_d(oT); }               // [Nmax][Nmax] re-evaluation scratch (tilde_U^T block)
    // staging: update (temp[N], col[M] per species) / Gauss-Jordan (prow, krow [N+M], fcol [N])
    __device__ __forceinline__ double *stg() const { return res_d(oStg); }
    __device__ __forceinline__ double *red() const { return res_d(oRed); }           // [8] reduction scratch
    __device__ __forceinline__ ResCtrl *ctrl() const { return reinterpret_cast<ResCtrl *>(res_d(oCtrl)); }
    // [n_bonds] bond endpoints (0-based): the proposal loop never leaves shared memory
    __device__ __forceinline__ short2 *bond() const { return reinterpret_cast<short2 *>(res_d(oBond)); }
    __device__ __forceinline__ short *kap(int sp) const { return res_h(hKap[sp]); }   // kappa per species [ns]
    __device__ __forceinline__ short *slot(int sp) const { return res_h(hSlot[sp]); } // [ns] compact slot of a site (-1: occupied)
    __device__ __forceinline__ short *site(int sp) const { return res_h(hSite[sp]); } // [M] site of a slot
};
static_assert(sizeof(ResCtrl) <= 32, "ResCtrl must fit its reserved slot");

__host__ __device__ inline size_t resident_smem_bytes(int ns, int n_up, int n_dn, int n_bonds) {
    const int Nmax = n_up > n_dn ? n_up : n_dn;
    size_t d = (size_t)n_up * (ns - n_up) + (size_t)n_dn * (ns - n_dn) + (size_t)Nmax * Nmax + 2 * ns + Nmax + 8 + 4 + (n_bonds + 1) / 2;
    size_t s = (size_t)4 * ns + (ns - n_up) + (ns - n_dn) + (ns + 1) + 2 * (size_t)n_bonds;
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
        const bool un = st < ns && L.kap(sp)[st] == 0;
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
                L.slot(sp)[st] = (short)u;
                L.site(sp)[u] = (short)st;
            } else {
                L.slot(sp)[st] = -1;
            }
        }
        base += tot;
        __syncthreads();
    }
}

// ---- reevaluateW! of one species in the compact storage (all T threads).  Returns false when tilde_U is singular.
//      Gauss-Jordan on B = [tilde_U^T | V^T] with IMPLICIT row pivoting: rows never move; step k picks, among the rows
//      that were not a pivot yet, the one with the largest |B[i][k]| (by the top bits of |x|, ties to the lowest row:
//      the rule of k_inverse_v4), scales it into prow and subtracts B[i][k] * prow from EVERY other row.  The only
//      special row is the pivot row itself (<- prow).  After N steps row p_t of the V^T block is column t of W; one
//      pass through the idle tilde_U^T scratch puts the rows in label order.
//      Cost model: the kernel is issue bound, so this loop is written for instruction count -- one warp per row, this
//      lane's columns of prow in registers, NP column passes unrolled, RU rows in flight (loads, FMAs, stores), no
//      branches and no address clamps inside (loads may run past a row end: they stay inside the dynamic allocation
//      and their results are never stored).  The multiplier B[i][k] is read from column k itself, which no step >= k
//      writes.  Pivot search: each lane owns ONE candidate row (row = warp + NWARP * lane); the key packs the top bits of
//      |x| with 127 - row, so that one REDUX.MAX per warp and one LDS.128 + three IMNMX per thread replace a search. ----
template <int T, int NP>
__device__ __forceinline__ bool res_reevaluate(const DevState &S, const ResSmem &L, int sp, int *s_scan) {
    constexpr int NWARP = T / 32;
    static_assert(NWARP == 4, "pivot keys are exchanged as one int4");
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ns = S.ns, N = sp ? S.n_dn : S.n_up, M = ns - N;
    const double *U = sp ? S.U_dn : S.U_up;
    double *Tm = L.T(), *Wc = L.Wc(sp);
    short *lab_site = reinterpret_cast<short *>(L.stg());          // [N] site of label c (scratch, dead before the elimination)
    res_build_tables<T>(L, sp, ns, s_scan);
    for (int st = tid; st < ns; st += T) {
        const int l = L.kap(sp)[st];
        if (l != 0) lab_site[l - 1] = (short)st;
    }
    __syncthreads();
    // B[j][c] = U[site(c), j]; one warp per row j, RB rows of loads in flight (U comes from the L2)
    {
        int sT[NP], sV[NP];
#pragma unroll
        for (int t = 0; t < NP; t++) {
            sT[t] = lane + 32 * t < N ? lab_site[lane + 32 * t] : 0;
            sV[t] = lane + 32 * t < M ? L.site(sp)[lane + 32 * t] : 0;
        }
        constexpr int RB = 4;
        for (int j0 = warp; j0 < N; j0 += RB * NWARP) {
            double vT[RB][NP], vV[RB][NP];
#pragma unroll
            for (int r = 0; r < RB; r++) {
                const double *Uj = U + (size_t)min(j0 + r * NWARP, N - 1) * ns;
#pragma unroll
                for (int t = 0; t < NP; t++) { vT[r][t] = __ldg(Uj + sT[t]); vV[r][t] = __ldg(Uj + sV[t]); }
            }
#pragma unroll
            for (int r = 0; r < RB; r++) {
                const int j = j0 + r * NWARP;
#pragma unroll
                for (int t = 0; t < NP; t++) {
                    if (j < N && lane + 32 * t < N) Tm[j * N + lane + 32 * t] = vT[r][t];
                    if (j < N && lane + 32 * t < M) Wc[j * M + lane + 32 * t] = vV[r][t];
                }
            }
        }
    }
    __syncthreads();
    double *prow = L.stg();                                      // [N + M] the scaled pivot row of the current step
    int *keys = s_scan + 8;                                      // [NWARP] per-warp pivot keys of the next column (16-byte aligned)
    int *pivrow = s_scan + 12;                                   // [1] (+ the row permutation below, in the staging area)
    short *perm = reinterpret_cast<short *>(L.stg() + ns);       // [N] pivot row of step t
    const int my_row = warp + NWARP * lane;                      // the candidate row of this lane
    bool my_used = my_row >= N;
    auto publish_key = [&](int col) {
        unsigned key = 0u;
        if (!my_used) {
            const unsigned hi = (unsigned)__double2hiint(Tm[my_row * N + col]) & 0x7fffffffu;
            key = (hi & 0xffffff80u) | (unsigned)(127 - my_row);
            if (hi >= 0x7ff00000u) key = 0xffffff80u | (unsigned)(127 - my_row);   // Inf / NaN: make it the pivot, then singular
        }
        const unsigned m = __reduce_max_sync(0xffffffffu, key);
        if (lane == 0) keys[warp] = (int)m;
    };
    publish_key(0);
    __syncthreads();
    for (int k = 0; k < N; k++) {
        const int4 kv = *reinterpret_cast<const int4 *>(keys);
        const unsigned best = max(max((unsigned)kv.x, (unsigned)kv.y), max((unsigned)kv.z, (unsigned)kv.w));
        if ((best >> 7) == 0u || best >= 0x7ff00000u) return false;   // zero (below 2^-1022) / non-finite pivot: singular (uniform)
        const int p = 127 - (int)(best & 127u);
        if (my_row == p) my_used = true;
        // stage the scaled pivot row: unfinished columns only (c > k of the tilde_U^T block, all of the V^T block)
        {
            const double inv = 1.0 / Tm[p * N + k];
            const int c = k + 1 + tid;                           // (N + M - k - 1 <= ns - 1 columns: one pass when ns <= T + 1)
            for (int cc = c; cc < N + M; cc += T)
                prow[cc] = (cc < N ? Tm[p * N + cc] : Wc[p * M + (cc - N)]) * inv;
            if (tid == 0) perm[k] = (short)p;
        }
        __syncthreads();
        double pT[NP], pV[NP];
        bool cT[NP], cV[NP];
#pragma unroll
        for (int t = 0; t < NP; t++) {
            cT[t] = k + 1 + lane + 32 * t < N;
            cV[t] = lane + 32 * t < M;
            pT[t] = cT[t] ? prow[k + 1 + lane + 32 * t] : 0.0;
            pV[t] = cV[t] ? prow[N + lane + 32 * t] : 0.0;
        }
        constexpr int RU = 4;
        const bool t1T = k + 1 + 32 < N;                         // (NP <= 2 is the tuned case: a second tilde_U^T pass or not)
        // body(i0, tail): rows i0 + r NWARP.  Full groups need no row predicate; the one partial group per warp does.
        auto body = [&](int i0, auto tail_c) {
            constexpr bool TAIL = decltype(tail_c)::value;
            double f[RU], vT[RU][NP], vV[RU][NP];
            const double *pt[RU];
            const double *pw[RU];
#pragma unroll
            for (int r = 0; r < RU; r++) {
                const int i = i0 + r * NWARP;
                const int ic = (TAIL && i >= N) ? i0 : i;        // (a row past the end re-reads row i0; nothing of it is stored)
                pt[r] = Tm + ic * N + k;
                pw[r] = Wc + ic * M + lane;
                f[r] = pt[r][0];
#pragma unroll
                for (int t = 0; t < NP; t++) {
                    // The loads are NOT predicated on the column masks of the stores: a lane past the row end reads entries
                    // of the next row (inside the allocation) that are never used -- neither stored nor fed into a stored
                    // value.  Their owner warp may update them at the same time, which compute-sanitizer racecheck reports as
                    // a read-write hazard; it is benign, and predicating the loads costs 9 % of the kernel (146 -> 132.5 M
                    // walker-sweeps/s at 108 sites: two more instructions per load in an issue-bound loop).  `make STRICT=1`
                    // builds the predicated form, on which racecheck is clean (profiles/r4_sanitizer.txt).
#ifdef KDSL_STRICT_LOADS
                    if (t == 0 || (t == 1 ? t1T : k + 1 + 32 * t < N)) vT[r][t] = cT[t] ? pt[r][1 + lane + 32 * t] : 0.0;
                    vV[r][t] = cV[t] ? pw[r][32 * t] : 0.0;
#else
                    if (t == 0 || (t == 1 ? t1T : k + 1 + 32 * t < N)) vT[r][t] = pt[r][1 + lane + 32 * t];
                    vV[r][t] = pw[r][32 * t];
#endif
                }
            }
#pragma unroll
            for (int r = 0; r < RU; r++)
#pragma unroll
                for (int t = 0; t < NP; t++) {
                    if (t == 0 || (t == 1 ? t1T : k + 1 + 32 * t < N)) vT[r][t] = fma(-f[r], pT[t], vT[r][t]);
                    vV[r][t] = fma(-f[r], pV[t], vV[r][t]);
                }
#pragma unroll
            for (int r = 0; r < RU; r++) {
                if (TAIL && i0 + r * NWARP >= N) break;
                double *qt = const_cast<double *>(pt[r]) + 1 + lane, *qw = const_cast<double *>(pw[r]);
#pragma unroll
                for (int t = 0; t < NP; t++) {
                    if (cT[t]) qt[32 * t] = vT[r][t];
                    if (cV[t]) qw[32 * t] = vV[r][t];
                }
            }
        };
        {
            int i0 = warp;
            for (; i0 + (RU - 1) * NWARP < N; i0 += RU * NWARP) body(i0, std::false_type{});
            if (i0 < N) body(i0, std::true_type{});
        }
        if (warp == (p & (NWARP - 1))) {                         // the pivot row itself (its owner just wrote zeros there)
#pragma unroll
            for (int t = 0; t < NP; t++) {
                if (cT[t]) Tm[p * N + k + 1 + lane + 32 * t] = pT[t];
                if (cV[t]) Wc[p * M + lane + 32 * t] = pV[t];
            }
        }
        __syncwarp();
        if (k + 1 < N) publish_key(k + 1);                       // (a warp only reads rows it wrote itself)
        __syncthreads();
    }
    // rows into label order: row perm[t] of the V^T block is column t of W
    for (int j = warp; j < N; j += NWARP)
#pragma unroll
        for (int t = 0; t < NP; t++)
            if (lane + 32 * t < M) Tm[j * M + lane + 32 * t] = Wc[j * M + lane + 32 * t];
    __syncthreads();
    for (int j = warp; j < N; j += NWARP) {
        const int src = perm[j];
#pragma unroll
        for (int t = 0; t < NP; t++)
            if (lane + 32 * t < M) Wc[j * M + lane + 32 * t] = Tm[src * M + lane + 32 * t];
    }
    (void)pivrow;
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
        const int iu = L.kap(0)[i], ju = L.kap(0)[j], id = L.kap(1)[i], jd = L.kap(1)[j];
        const int oi = (iu != 0) + (id != 0), oj = (ju != 0) + (jd != 0);
        if (oi != 1 || oj != 1) { bad = 1; continue; }
        if (ju != 0 && id != 0)                                    // W_up[i, ju] W_dn[j, id]
            flips += -0.5 * L.Wc(0)[(ju - 1) * Mu + L.slot(0)[i]] * L.Wc(1)[(id - 1) * Md + L.slot(1)[j]];
        if (iu != 0 && jd != 0)                                    // W_up[j, iu] W_dn[i, jd]
            flips += -0.5 * L.Wc(0)[(iu - 1) * Mu + L.slot(0)[j]] * L.Wc(1)[(jd - 1) * Md + L.slot(1)[i]];
        diag4 += (iu != 0 ? 1 : -1) * (ju != 0 ? 1 : -1);
    }
    flips = warp_sum_f64(flips);
    diag4 = warp_sum_int(diag4);
    bad = warp_sum_int(bad);
    double *red = L.red();
    int *redi = reinterpret_cast<int *>(L.stg());                  // (staging is idle during a measurement)
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

// NP = ceil(max(N_up, N_dn, M_up, M_dn) / 32): the column passes of the update / elimination loops are unrolled
#define KDSL_RES_THREADS 128     /* 3 CTAs x 128 threads per SM leave 168 registers per thread: nothing spills (256 threads did) */
template <bool REPLAY, int NP>
__global__ void __launch_bounds__(KDSL_RES_THREADS, NP <= 2 ? 3 : 1)
k_resident(DevState S, ResParams P) {
    constexpr int T = KDSL_RES_THREADS;
    __shared__ __align__(16) int s_scan[24];
    __shared__ int s_w;
    constexpr int NWARP = T / 32;
    const ResSmem L(S.ns, S.n_up, S.n_dn, S.n_bonds);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ns = S.ns;
    const int Nn[2] = {S.n_up, S.n_dn}, Mm[2] = {ns - S.n_up, ns - S.n_dn};
    for (int b = tid; b < S.n_bonds; b += T) L.bond()[b] = make_short2((short)S.bi[b], (short)S.bj[b]);   // (once per CTA)
    for (int x = tid; x <= ns; x += T) L.adj_off()[x] = (short)S.adj_off[x];
    for (int x = tid; x < 2 * S.n_bonds; x += T) L.adj_nbr()[x] = (short)S.adj_nbr[x];

    for (;;) {
        __syncthreads();
        if (tid == 0) s_w = atomicAdd(P.work_counter, 1);
        __syncthreads();
        const int w = s_w;
        if (w >= S.nw) break;
        if (S.flags[w] & KDSL_FLAG_SINGULAR_DEV) continue;       // frozen (see k_decide_wb)

        // ---- load the walker ----
        for (int st = tid; st < ns; st += T) {
            L.kap(0)[st] = (short)S.kup[(size_t)w * ns + st];
            L.kap(1)[st] = (short)S.kdn[(size_t)w * ns + st];
        }
        __syncthreads();
        res_build_tables<T>(L, 0, ns, s_scan);
        res_build_tables<T>(L, 1, ns, s_scan);
#pragma unroll
        for (int sp = 0; sp < 2; sp++) {                         // one warp per column l of W: the unoccupied rows only
            const double *Wg = (sp ? S.W_dn : S.W_up) + (size_t)w * ns * Nn[sp];
            const int M = Mm[sp];
            int su[NP];                                          // this lane's unoccupied sites
#pragma unroll
            for (int t = 0; t < NP; t++) su[t] = lane + 32 * t < M ? L.site(sp)[lane + 32 * t] : 0;
            constexpr int LU = 4;                                // columns in flight per warp (all loads issued before the stores)
            for (int l0 = warp; l0 < Nn[sp]; l0 += LU * NWARP) {
                double v[LU][NP];
#pragma unroll
                for (int r = 0; r < LU; r++)
#pragma unroll
                    for (int t = 0; t < NP; t++) {
                        const int l = l0 + r * NWARP;
                        v[r][t] = (l < Nn[sp] && lane + 32 * t < M) ? Wg[(size_t)l * ns + su[t]] : 0.0;
                    }
#pragma unroll
                for (int r = 0; r < LU; r++)
#pragma unroll
                    for (int t = 0; t < NP; t++) {
                        const int l = l0 + r * NWARP;
                        if (l < Nn[sp] && lane + 32 * t < M) L.Wc(sp)[l * M + lane + 32 * t] = v[r][t];
                    }
            }
        }
        if (tid == 0) { L.ctrl()->singular = 0; L.ctrl()->ev = 0; }
        __syncthreads();

        // ---- per-walker scalars: live in warp 0 (all lanes identical) ----
        Xoshiro g;
        g.s0 = g.s1 = g.s2 = g.s3 = 0ull;
        int zmu = 0, s_done = 0;
        long long pp = P.phase_p, pm = P.phase_m, sw = P.sweep0; // sweep counter and its phases (no 64-bit modulo per sweep)
        double zr = 0.0;
        unsigned long long c_acc = 0ull, c_reach = 0ull, c_refresh = 0ull, c_ol = 0ull;
        double a_ol = 0.0, a_ol2 = 0.0, last_ol = 0.0;
        bool have_ol = false, dead = false;
        if (warp == 0) {
            zmu = S.zmu[w];
            zr = (double)zmu / (double)S.n_bonds;               // Zmu / Zmax, recomputed only when Zmu changes
            if (!REPLAY) {
                const unsigned long long *st = S.rng + (size_t)w * 4;
                g.s0 = st[0]; g.s1 = st[1]; g.s2 = st[2]; g.s3 = st[3];
            }
        }

        for (;;) {
            if (warp == 0) {
                int ev = 0;
                while (s_done < P.n_sweeps) {
                    const bool gate = pp == 0;                                  // :595 (pre-increment)
                    const size_t ro = (size_t)s_done * S.nw + w;
                    const double r = REPLAY ? P.rp_r[ro] : g.rand_f64();        // :546
                    bool accepted = false, reached = false;
                    if (!(r > zr)) {                                            // :547-550
                        long long b = REPLAY ? (long long)P.rp_bond[ro] : g.rand_index((unsigned long long)S.n_bonds);
                        if (b < 1) b = 1;
                        if (b > S.n_bonds) b = S.n_bonds;
                        const short2 bb = L.bond()[b - 1];
                        const int i = bb.x, site = bb.y;
                        const int ku_i = L.kap(0)[i], ku_s = L.kap(0)[site], kd_i = L.kap(1)[i], kd_s = L.kap(1)[site];
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
                            const int q_up = L.slot(0)[K_up], q_dn = L.slot(1)[K_dn];
                            const double wu = q_up >= 0 ? L.Wc(0)[(l_up - 1) * Mm[0] + q_up] : (L.kap(0)[K_up] == l_up ? 1.0 : 0.0);
                            const double wd = q_dn >= 0 ? L.Wc(1)[(l_dn - 1) * Mm[1] + q_dn] : (L.kap(1)[K_dn] == l_dn ? 1.0 : 0.0);
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
                                        double *temp = L.stg() + (sp ? S.n_up + Mm[0] : 0), *col = temp + N;
#pragma unroll
                                        for (int t = 0; t < NP; t++) {
                                            const int j = lane + 32 * t;
                                            if (j < N) {
                                                double v = L.Wc(sp)[j * M + q];
                                                if (j == l) v -= 1.0;
                                                temp[j] = alpha * v;
                                            }
                                            if (j < M) col[j] = L.Wc(sp)[l * M + j];
                                        }
                                        if (lane == 0) { L.ctrl()->l[sp] = l; L.ctrl()->q[sp] = q; }
                                    }
                                    ev |= RES_EV_UPDATE;
                                }
                                // Z_mu through the bonds incident to i or site (equals the full recount, :460-474)
                                const int ui_o = ku_i != 0, di_o = kd_i != 0, us_o = ku_s != 0, ds_o = kd_s != 0;
                                const int ui_n = flag == 1 ? 0 : 1, di_n = flag == 1 ? 1 : 0;
                                const int us_n = flag == 1 ? 1 : 0, ds_n = flag == 1 ? 0 : 1;
                                int delta = 0;
                                for (int qq = L.adj_off()[i] + lane; qq < L.adj_off()[i + 1]; qq += 32) {
                                    const int n = L.adj_nbr()[qq];
                                    if (n == site) continue;
                                    const int un = L.kap(0)[n] != 0, dn = L.kap(1)[n] != 0;
                                    delta += bond_is_anti(ui_n, di_n, un, dn) - bond_is_anti(ui_o, di_o, un, dn);
                                }
                                for (int qq = L.adj_off()[site] + lane; qq < L.adj_off()[site + 1]; qq += 32) {
                                    const int n = L.adj_nbr()[qq];
                                    if (n == i) continue;
                                    const int un = L.kap(0)[n] != 0, dn = L.kap(1)[n] != 0;
                                    delta += bond_is_anti(us_n, ds_n, un, dn) - bond_is_anti(us_o, ds_o, un, dn);
                                }
                                if (lane == 0)
                                    delta += bond_is_anti(ui_n, di_n, us_n, ds_n) - bond_is_anti(ui_o, di_o, us_o, ds_o);
                                zmu += warp_sum_int(delta);
                                zr = (double)zmu / (double)S.n_bonds;
                                __syncwarp();
                                if (lane == 0) {
                                    if (flag == 1) {                            // :502-503
                                        L.kap(0)[i] = 0; L.kap(0)[site] = (short)l_up;
                                        L.kap(1)[i] = (short)l_dn; L.kap(1)[site] = 0;
                                    } else {                                    // :508-509
                                        L.kap(0)[i] = (short)l_up; L.kap(0)[site] = 0;
                                        L.kap(1)[i] = 0; L.kap(1)[site] = (short)l_dn;
                                    }
                                    if (!gate) {
                                        // the vacated site takes over the slot of the newly occupied one
                                        const int R_up = flag == 1 ? i : site, R_dn = flag == 1 ? site : i;
                                        L.slot(0)[R_up] = (short)q_up; L.slot(0)[K_up] = -1; L.site(0)[q_up] = (short)R_up;
                                        L.slot(1)[R_dn] = (short)q_dn; L.slot(1)[K_dn] = -1; L.site(1)[q_dn] = (short)R_dn;
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
                    sw++;                                                       // Carlo: ctx.sweeps += 1
                    if (++pp == P.period) pp = 0;
                    if (++pm == S.n_occ) pm = 0;
                    if (P.therm >= 0 && sw > P.therm && pm == 0) ev |= RES_EV_MEASURE;   // :630 (post-increment)
                    if (ev) break;
                }
                if (s_done >= P.n_sweeps) ev |= RES_EV_DONE;
                if (lane == 0) L.ctrl()->ev = ev;
            }
            __syncthreads();
            const int ev = L.ctrl()->ev;
            if (ev & RES_EV_UPDATE) {
#pragma unroll
                for (int sp = 0; sp < 2; sp++) {
                    const int N = Nn[sp], M = Mm[sp];
                    const int l = L.ctrl()->l[sp], q = L.ctrl()->q[sp];
                    const double *temp = L.stg() + (sp ? S.n_up + Mm[0] : 0), *col = temp + N;
                    double *Wc = L.Wc(sp);
                    double cu[NP];                               // this lane's entries of the column W[:, l]
#pragma unroll
                    for (int t = 0; t < NP; t++) cu[t] = lane + 32 * t < M ? col[lane + 32 * t] : 0.0;
                    // one warp per label j (row of the compact storage), RU rows in flight.  Slot q now stands for the
                    // vacated site, whose old row was e_l: x = 1, A = delta_lj there.
                    bool isq[NP];
#pragma unroll
                    for (int t = 0; t < NP; t++) { isq[t] = lane + 32 * t == q; if (isq[t]) cu[t] = 1.0; }
                    constexpr int RU = 4;
                    for (int j0 = warp; j0 < N; j0 += RU * NWARP) {
                        double tj[RU], v[RU][NP];
#pragma unroll
                        for (int r = 0; r < RU; r++) {
                            const int j = j0 + r * NWARP;
                            tj[r] = j < N ? temp[j] : 0.0;
#pragma unroll
                            for (int t = 0; t < NP; t++) {
                                v[r][t] = (j < N && lane + 32 * t < M) ? Wc[j * M + lane + 32 * t] : 0.0;
                                if (isq[t]) v[r][t] = j == l ? 1.0 : 0.0;
                            }
                        }
#pragma unroll
                        for (int r = 0; r < RU; r++)
#pragma unroll
                            for (int t = 0; t < NP; t++) v[r][t] = fma(cu[t], tj[r], v[r][t]);
#pragma unroll
                        for (int r = 0; r < RU; r++) {
                            const int j = j0 + r * NWARP;
#pragma unroll
                            for (int t = 0; t < NP; t++)
                                if (j < N && lane + 32 * t < M) Wc[j * M + lane + 32 * t] = v[r][t];
                        }
                    }
                }
                __syncthreads();
            }
            if (ev & RES_EV_REFRESH) {
                bool ok = true;
#pragma unroll 1
                for (int sp = 0; sp < 2 && ok; sp++) {           // (one copy of the code for both species)
                    ok = res_reevaluate<T, NP>(S, L, sp, s_scan);
                    __syncthreads();
                }
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
                if (S.obs_on && warp == 0) {                     // structure factor / reweighting sums (kdsl_set_observables)
                    const double OLb = __shfl_sync(0xffffffffu, OL, 0);
                    const short *ku = L.kap(0);
                    extra_observables_warp(S, w, lane, zmu, OLb, [&](int i) { return ku[i] != 0; });
                }
            }
            if (ev & RES_EV_DONE) break;
        }

        // ---- write the walker back ----
        __syncthreads();
        for (int st = tid; st < ns; st += T) {
            S.kup[(size_t)w * ns + st] = L.kap(0)[st];
            S.kdn[(size_t)w * ns + st] = L.kap(1)[st];
        }
        if (!dead) {
#pragma unroll
            for (int sp = 0; sp < 2; sp++) {
                double *Wg = (sp ? S.W_dn : S.W_up) + (size_t)w * ns * Nn[sp];
                const int M = Mm[sp];
                for (int l = warp; l < Nn[sp]; l += NWARP)
                    for (int st = lane; st < ns; st += 32) {
                        const int u = L.slot(sp)[st];
                        Wg[(size_t)l * ns + st] = u >= 0 ? L.Wc(sp)[l * M + u] : (L.kap(sp)[st] == l + 1 ? 1.0 : 0.0);
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
