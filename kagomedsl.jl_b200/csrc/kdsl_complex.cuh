// kdsl_complex.cuh -- ComplexF64 mode of the VMC path: Hamiltonians with a Peierls flux B != 0
// (reference src/Hamiltonian.jl:201-204, 325-327; scripts/LL.jl) have complex Hermitian hopping matrices, so U and
// W = U inv(tilde_U) are genuinely complex, the acceptance uses abs2(ratio) (src/MonteCarlo.jl:581) and the local
// energy real(OL) (src/Hamiltonian.jl:777).
//
// First version: the reference's own algorithm, batched over walkers -- immediate rank-1 update per accepted
// move (update_W!, :279-292, unconjugated zgeru), unblocked Gauss-Jordan inverse with LAPACK's izamax pivot rule
// (|re| + |im|), FP64-FMA GEMM.  Complex arrays are interleaved (re, im) pairs in the same DevState buffers, which
// are allocated twice as large; every kernel here reinterprets them as double2.
#pragma once
#include "kdsl_common.cuh"
#include "kdsl_measure.cuh"
#include "kdsl_propose.cuh"
#include "kdsl_refresh.cuh"
#include "kdsl_update.cuh"

typedef double2 cplx;
__device__ __forceinline__ cplx c_make(double x, double y) { return make_double2(x, y); }
__device__ __forceinline__ cplx c_mul(cplx a, cplx b) {
    return make_double2(fma(a.x, b.x, -(a.y * b.y)), fma(a.x, b.y, a.y * b.x));
}
__device__ __forceinline__ cplx c_fma(cplx a, cplx b, cplx c) {      // a * b + c
    return make_double2(fma(-a.y, b.y, fma(a.x, b.x, c.x)), fma(a.y, b.x, fma(a.x, b.y, c.y)));
}
__device__ __forceinline__ cplx c_neg(cplx a) { return make_double2(-a.x, -a.y); }
__device__ __forceinline__ double c_abs2(cplx a) { return fma(a.x, a.x, a.y * a.y); }
__device__ __forceinline__ double c_abs1(cplx a) { return fabs(a.x) + fabs(a.y); }
__device__ __forceinline__ cplx c_inv(cplx a) {                       // Smith's algorithm (as robust as C's 1.0 / z)
    if (fabs(a.x) >= fabs(a.y)) {
        const double r = a.y / a.x, d = fma(a.y, r, a.x);
        return make_double2(1.0 / d, -r / d);
    }
    const double r = a.x / a.y, d = fma(a.x, r, a.y);
    return make_double2(r / d, -1.0 / d);
}
__device__ __forceinline__ const cplx *cW(const double *base, size_t off) { return reinterpret_cast<const cplx *>(base) + off; }
__device__ __forceinline__ cplx *cW(double *base, size_t off) { return reinterpret_cast<cplx *>(base) + off; }

// Carlo.sweep! (reference src/MonteCarlo.jl:538-607), complex W; same structure as k_propose.
template <bool REPLAY>
__global__ void __launch_bounds__(256)
k_propose_c(DevState S, int parity, int gate_refresh, const double *__restrict__ rp_r,
            const int *__restrict__ rp_bond, const int *__restrict__ rp_pick) {
    const int w = (int)((blockIdx.x * (unsigned)blockDim.x + threadIdx.x) >> 5);
    const int lane = threadIdx.x & 31;
    if (w >= S.nw) return;
    if (S.flags[w] & 1) return;      // frozen after a singular re-evaluation (see k_decide_wb; KDSL_FLAG_SINGULAR)
    const int ns = S.ns;
    int *kup = S.kup + (size_t)w * ns;
    int *kdn = S.kdn + (size_t)w * ns;
    const int zmu = S.zmu[w];
    Xoshiro g;
    if (!REPLAY) {
        const unsigned long long *st = S.rng + (size_t)w * 4;
        g.s0 = st[0]; g.s1 = st[1]; g.s2 = st[2]; g.s3 = st[3];
    }
    const double r = REPLAY ? rp_r[w] : g.rand_f64();           // :546
    const double zr = (double)zmu / (double)S.n_bonds;
    bool accepted = false, reached = false;
    int i = 0, site = 0, flag = 0, l_up = 0, l_dn = 0, K_up = 0, K_dn = 0;
    int ku_i = 0, ku_s = 0, kd_i = 0, kd_s = 0;
    cplx wu = c_make(0.0, 0.0), wd = c_make(0.0, 0.0);
    const cplx *Wu = cW(S.W_up, (size_t)w * ns * S.n_up);
    const cplx *Wd = cW(S.W_dn, (size_t)w * ns * S.n_dn);
    if (!(r > zr)) {                                            // :547-550
        long long b = REPLAY ? (long long)rp_bond[w] : g.rand_index((unsigned long long)S.n_bonds);
        if (b < 1) b = 1;
        if (b > S.n_bonds) b = S.n_bonds;
        i = S.bi[b - 1];
        site = S.bj[b - 1];
        ku_i = kup[i]; ku_s = kup[site]; kd_i = kdn[i]; kd_s = kdn[site];
        const bool f1 = ku_i != 0 && kd_s != 0;
        const bool f2 = ku_s != 0 && kd_i != 0;
        if (f1 || f2) {
            const int nm = (int)f1 + (int)f2;
            long long pick;
            if (REPLAY) pick = rp_pick ? (long long)rp_pick[w] : 1;
            else pick = g.rand_index((unsigned long long)nm);
            flag = (f1 && f2) ? (pick == 1 ? 1 : 2) : (f1 ? 1 : 2);
            l_up = flag == 1 ? ku_i : ku_s;
            l_dn = flag == 1 ? kd_s : kd_i;
            K_up = flag == 1 ? site : i;
            K_dn = flag == 1 ? i : site;
            wu = Wu[(size_t)(l_up - 1) * ns + K_up];
            wd = Wd[(size_t)(l_dn - 1) * ns + K_dn];
            const double p = c_abs2(c_mul(wu, wd));             // abs2(ratio), :581
            if (p >= 1.0 && r < zr) accepted = true;
            else if (p < 1.0 && r < zr * p) accepted = true;
            if (!(p == p) || p > 1.79e308) {
                if (lane == 0) atomicOr(&S.flags[w], KDSL_FLAG_NONFINITE_DEV);
            }
            reached = true;
        }
    }
    if (accepted) {
        if (!gate_refresh) {
            const cplx au = c_neg(c_inv(wu)), ad = c_neg(c_inv(wd));       // :289
            cplx *cu = cW(S.col_up, (size_t)w * ns), *cd = cW(S.col_dn, (size_t)w * ns);
            const cplx *srcu = Wu + (size_t)(l_up - 1) * ns, *srcd = Wd + (size_t)(l_dn - 1) * ns;
            for (int t = lane; t < ns; t += 32) {               // :288 col_cache = W[:, l]
                cu[t] = srcu[t];
                cd[t] = srcd[t];
            }
            cplx *tu = cW(S.trow_up, (size_t)w * S.n_up), *td = cW(S.trow_dn, (size_t)w * S.n_dn);
            for (int j = lane; j < S.n_up; j += 32) {           // :286-287 row_cache = W[K, :] - e_l, times alpha
                cplx v = Wu[(size_t)j * ns + K_up];
                if (j == l_up - 1) v.x -= 1.0;
                tu[j] = c_mul(au, v);
            }
            for (int j = lane; j < S.n_dn; j += 32) {
                cplx v = Wd[(size_t)j * ns + K_dn];
                if (j == l_dn - 1) v.x -= 1.0;
                td[j] = c_mul(ad, v);
            }
            if (lane == 0) {
                const int slot = atomicAdd(&S.cnt[parity], 1);
                S.acc_list[(size_t)parity * S.nw + slot] = w;
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
            if (flag == 1) {
                kup[i] = 0; kup[site] = l_up;
                kdn[i] = l_dn; kdn[site] = 0;
            } else {
                kup[i] = l_up; kup[site] = 0;
                kdn[i] = 0; kdn[site] = l_dn;
            }
            S.n_acc[w] += 1ull;
        }
    }
    if (lane == 0) {
        if (reached) {
            S.n_reach[w] += 1ull;
            if (gate_refresh) {
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

// stage (col, alpha*row) for explicit moves (kdsl_update_W): one warp per move
__global__ void __launch_bounds__(256)
k_stage_moves_c(DevState S, int parity, int n_moves, const int *__restrict__ mv) {
    const int m = (int)((blockIdx.x * (unsigned)blockDim.x + threadIdx.x) >> 5);
    const int lane = threadIdx.x & 31;
    if (m >= n_moves) return;
    const int w = mv[m], l_up = mv[n_moves + m], K_up = mv[2 * n_moves + m] - 1, l_dn = mv[3 * n_moves + m], K_dn = mv[4 * n_moves + m] - 1;
    const int ns = S.ns;
    const cplx *Wu = cW(S.W_up, (size_t)w * ns * S.n_up), *Wd = cW(S.W_dn, (size_t)w * ns * S.n_dn);
    const cplx au = c_neg(c_inv(Wu[(size_t)(l_up - 1) * ns + K_up])), ad = c_neg(c_inv(Wd[(size_t)(l_dn - 1) * ns + K_dn]));
    cplx *cu = cW(S.col_up, (size_t)w * ns), *cd = cW(S.col_dn, (size_t)w * ns);
    for (int t = lane; t < ns; t += 32) {
        cu[t] = Wu[(size_t)(l_up - 1) * ns + t];
        cd[t] = Wd[(size_t)(l_dn - 1) * ns + t];
    }
    cplx *tu = cW(S.trow_up, (size_t)w * S.n_up), *td = cW(S.trow_dn, (size_t)w * S.n_dn);
    for (int j = lane; j < S.n_up; j += 32) {
        cplx v = Wu[(size_t)j * ns + K_up];
        if (j == l_up - 1) v.x -= 1.0;
        tu[j] = c_mul(au, v);
    }
    for (int j = lane; j < S.n_dn; j += 32) {
        cplx v = Wd[(size_t)j * ns + K_dn];
        if (j == l_dn - 1) v.x -= 1.0;
        td[j] = c_mul(ad, v);
    }
    if (lane == 0) S.acc_list[(size_t)parity * S.nw + m] = w;
    if (m == 0 && lane == 0) S.cnt[parity] = n_moves;
}

// W += col (x) trow for the accepted walkers (unconjugated rank-1 update, zgeru): HBM-bound, 32 ns^2 bytes per
// accepted move.  Work item = (accepted walker, species, slab of CH columns); 128-bit streaming accesses.
template <int THREADS>
__global__ void __launch_bounds__(THREADS)
k_update_c(DevState S, int parity, int tiles_up, int tiles_dn, int CH) {
    extern __shared__ double smem_c[];
    cplx *s_col = reinterpret_cast<cplx *>(smem_c);            // [ns] the staged column, then [CH] the scaled row entries
    const int tid = threadIdx.x;
    const int ns = S.ns;
    cplx *s_trow = s_col + ns;
    const int n_acc = S.cnt[parity];
    const int tpw = tiles_up + tiles_dn;
    const long long total = (long long)n_acc * tpw;
    if (blockIdx.x == 0 && tid == 0) {
        S.cnt[parity ^ 1] = 0;
        *S.upd_moves += (unsigned long long)n_acc;
    }
    for (long long t = blockIdx.x; t < total; t += gridDim.x) {
        const int a = (int)(t / tpw);
        const int rem = (int)(t - (long long)a * tpw);
        const int spin = rem >= tiles_up;
        const int tile = spin ? rem - tiles_up : rem;
        const int w = S.acc_list[(size_t)parity * S.nw + a];
        const int N = spin ? S.n_dn : S.n_up;
        cplx *W = cW(spin ? S.W_dn : S.W_up, (size_t)w * ns * N);
        const cplx *col = cW(spin ? S.col_dn : S.col_up, (size_t)w * ns);
        const cplx *trow = cW(spin ? S.trow_dn : S.trow_up, (size_t)w * N);
        const int j0 = tile * CH, jn = min(CH, N - j0);
        __syncthreads();
        for (int x = tid; x < ns; x += THREADS) s_col[x] = col[x];
        for (int x = tid; x < jn; x += THREADS) s_trow[x] = trow[j0 + x];
        __syncthreads();
        // the slab of jn columns is contiguous: four 16-byte elements per thread in flight (one element per thread and
        // column left the loads of a column waiting on the stores of the previous one: 4.4 TB/s)
        cplx *base = W + (size_t)j0 * ns;
        const int n_el = jn * ns;
        for (int q = tid; q < n_el; q += THREADS * 4) {
            cplx v[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int idx = q + u * THREADS;
                if (idx < n_el) v[u] = ldg_stream(base + idx);
            }
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int idx = q + u * THREADS;
                if (idx < n_el) {
                    const int jq = idx / ns, xq = idx - jq * ns;
                    v[u] = c_fma(s_col[xq], s_trow[jq], v[u]);
                    stg_stream(base + idx, v[u]);
                }
            }
        }
    }
}

// tilde_U (src/MonteCarlo.jl:92-115), complex; grid (nw, 2), dynamic smem N ints
__global__ void __launch_bounds__(256)
k_gather_tilde_c(DevState S, const int *__restrict__ list, double *__restrict__ A_up,
                 double *__restrict__ A_dn, int *__restrict__ status) {
    extern __shared__ int s_site_c[];
    const int b = blockIdx.x, spin = blockIdx.y;
    if (b >= batch_count(S, list)) return;
    const int w = list ? list[b] : b;
    const int ns = S.ns, N = spin ? S.n_dn : S.n_up;
    const int *kap = (spin ? S.kdn : S.kup) + (size_t)w * ns;
    const cplx *U = cW(spin ? S.U_dn : S.U_up, 0);
    cplx *A = cW(spin ? A_dn : A_up, (size_t)b * N * N);
    for (int R = threadIdx.x; R < ns; R += blockDim.x) {
        const int l = kap[R];
        if (l != 0) s_site_c[l - 1] = R;
    }
    if (threadIdx.x == 0) status[2 * b + spin] = 0;
    __syncthreads();
    for (int e = threadIdx.x; e < N * N; e += blockDim.x) {
        const int c = e / N, l = e - c * N;
        A[e] = U[(size_t)c * ns + s_site_c[l]];                // tilde_U[l, c] = U[R_l, c]
    }
}

// in-place inverse, unblocked Gauss-Jordan with partial pivoting (LAPACK zgetrf pivot rule: first row of maximal
// |re| + |im|; an exactly zero or non-finite pivot => singular).  One CTA per (batch entry, species).
// dynamic smem: 2 N complex + N ints.
__global__ void __launch_bounds__(256)
k_inverse_gj_c(DevState S, const int *__restrict__ list, double *__restrict__ A_base, int spin,
               int *__restrict__ status) {
    extern __shared__ double sm_dc[];
    const int b = blockIdx.x;
    if (b >= batch_count(S, list)) return;
    const int N = spin ? S.n_dn : S.n_up;
    cplx *prow = reinterpret_cast<cplx *>(sm_dc), *colk = prow + N;
    int *piv = reinterpret_cast<int *>(colk + N);
    __shared__ double r_val[8];
    __shared__ int r_idx[8];
    __shared__ int s_p;
    cplx *A = cW(A_base, (size_t)b * N * N);
    const int tid = threadIdx.x, T = blockDim.x, lane = tid & 31, warp = tid >> 5;
    for (int k = 0; k < N; k++) {
        double best = -1.0;
        int bi = k;
        for (int i = k + tid; i < N; i += T) {
            const double v = c_abs1(A[(size_t)k * N + i]);
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
            if (!(bv > 0.0) || bv == INFINITY) bx = -1;
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
                const cplx t0 = A[(size_t)j * N + k];
                A[(size_t)j * N + k] = A[(size_t)j * N + p];
                A[(size_t)j * N + p] = t0;
            }
        __syncthreads();
        for (int i = tid; i < N; i += T) colk[i] = A[(size_t)k * N + i];
        __syncthreads();
        const cplx d = c_inv(colk[k]);
        for (int j = tid; j < N; j += T) prow[j] = (j == k) ? d : c_mul(A[(size_t)j * N + k], d);
        __syncthreads();
        for (int e = tid; e < N * N; e += T) {
            const int j = e / N, i = e - j * N;
            cplx v;
            if (i == k) v = prow[j];
            else if (j == k) v = c_neg(c_mul(colk[i], d));
            else v = c_fma(c_neg(colk[i]), prow[j], A[e]);
            A[e] = v;
        }
        __syncthreads();
    }
    for (int k = N - 1; k >= 0; k--) {                        // undo the row interchanges on the columns
        const int p = piv[k];
        if (p != k)
            for (int i = tid; i < N; i += T) {
                const cplx t0 = A[(size_t)k * N + i];
                A[(size_t)k * N + i] = A[(size_t)p * N + i];
                A[(size_t)p * N + i] = t0;
            }
        __syncthreads();
    }
}

// ---- inverse of the complex tilde_U through its real embedding (inverse_variant 0 of the ComplexF64 engine) ----
// A = X + iY (N x N complex) is invertible iff E = [[X, -Y], [Y, X]] (2N x 2N real) is, and inv(E) = [[P, -Q], [Q, P]]
// with inv(A) = P + iQ.  E goes through the production real kernels (blocked implicit-pivoting Gauss-Jordan on the
// FP64 tensor pipe, k_inverse_v4 / k_inverse_v5): twice the flops of a complex elimination, but DMMA instead of a
// latency-bound unblocked loop in global memory (measured 3.4x faster at 432 sites).
// k_gather_tilde_emb_c: E padded to Ne = roundup(2N, 8) with an identity block; grid (nw, 2), dynamic smem N ints.
__global__ void __launch_bounds__(256)
k_gather_tilde_emb_c(DevState S, const int *__restrict__ list, double *__restrict__ E_up, double *__restrict__ E_dn,
                     int *__restrict__ status, int Ne_up, int Ne_dn) {
    extern __shared__ int s_site_e[];
    const int b = blockIdx.x, spin = blockIdx.y;
    if (b >= batch_count(S, list)) return;
    const int w = list ? list[b] : b;
    const int ns = S.ns, N = spin ? S.n_dn : S.n_up, Ne = spin ? Ne_dn : Ne_up;
    const int *kap = (spin ? S.kdn : S.kup) + (size_t)w * ns;
    const cplx *U = cW(spin ? S.U_dn : S.U_up, 0);
    double *E = (spin ? E_dn : E_up) + (size_t)b * Ne * Ne;
    for (int R = threadIdx.x; R < ns; R += blockDim.x) {
        const int l = kap[R];
        if (l != 0) s_site_e[l - 1] = R;
    }
    if (threadIdx.x == 0) status[2 * b + spin] = 0;
    __syncthreads();
    for (int e = threadIdx.x; e < Ne * Ne; e += blockDim.x) {
        const int c = e / Ne, r = e - c * Ne;
        double v;
        if (r < 2 * N && c < 2 * N) {
            const int l = r < N ? r : r - N, cc = c < N ? c : c - N;
            const cplx u = U[(size_t)cc * ns + s_site_e[l]];        // tilde_U[l, cc] = U[R_l, cc]
            v = (r < N) == (c < N) ? u.x : (r < N ? -u.y : u.y);
        } else {
            v = r == c ? 1.0 : 0.0;
        }
        E[e] = v;
    }
}

// k_unembed_c: the stored result of the implicit-pivoting inverse is S[p_k, c] = inv(E)[k, p_c] (colsrc[i] = step at
// which row i was the pivot); X[k, j] = (inv(E)[k, j], inv(E)[N + k, j]) as a plain column-major complex N x N matrix
// for k_gemm_W_c.  grid (nw, 2), dynamic smem Ne ints.
__global__ void __launch_bounds__(256)
k_unembed_c(DevState S, const int *__restrict__ list, const double *__restrict__ E_up, const double *__restrict__ E_dn,
            double *__restrict__ X_up, double *__restrict__ X_dn, const int *__restrict__ status,
            const int *__restrict__ colsrc_base, int Ne_up, int Ne_dn, int cs_stride,
            int *__restrict__ urow_base, int urow_stride) {
    extern __shared__ int s_row_of_step[];
    const int b = blockIdx.x, spin = blockIdx.y;
    if (b >= batch_count(S, list)) return;
    if (status[2 * b] | status[2 * b + 1]) return;
    const int N = spin ? S.n_dn : S.n_up, Ne = spin ? Ne_dn : Ne_up;
    const double *E = (spin ? E_dn : E_up) + (size_t)b * Ne * Ne;
    cplx *X = cW(spin ? X_dn : X_up, (size_t)b * N * N);
    const int *colsrc = colsrc_base + ((size_t)2 * b + spin) * cs_stride;
    for (int i = threadIdx.x; i < Ne; i += blockDim.x) s_row_of_step[colsrc[i]] = i;
    __syncthreads();
    for (int e = threadIdx.x; e < N * N; e += blockDim.x) {
        const int j = e / N, k = e - j * N;
        const double *col = E + (size_t)colsrc[j] * Ne;
        X[e] = c_make(col[s_row_of_step[k]], col[s_row_of_step[N + k]]);
    }
    // ordered list of the sites NOT occupied by this species (the non-trivial rows of W, computed by k_gemm_W_c) and
    // the unit rows W[R_l, :] = e_l of the occupied ones (written here, never computed)
    __shared__ int s_wsum_u[8];
    __shared__ int s_base_u;
    const int w = list ? list[b] : b;
    const int ns = S.ns;
    const int *kap = (spin ? S.kdn : S.kup) + (size_t)w * ns;
    int *urow = urow_base + ((size_t)2 * b + spin) * urow_stride;
    cplx *W = cW(spin ? S.W_dn : S.W_up, (size_t)w * ns * N);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_base_u = 0;
    __syncthreads();
    for (int s0 = 0; s0 < ns; s0 += 256) {
        const int site = s0 + threadIdx.x;
        const bool un = site < ns && kap[site] == 0;
        const unsigned m = __ballot_sync(0xffffffffu, un);
        if (lane == 0) s_wsum_u[warp] = __popc(m);
        __syncthreads();
        int off = s_base_u;
        for (int q = 0; q < warp; q++) off += s_wsum_u[q];
        if (un) urow[off + __popc(m & ((1u << lane) - 1u))] = site;
        __syncthreads();
        if (threadIdx.x == 0) {
            int t = 0;
            for (int q = 0; q < 8; q++) t += s_wsum_u[q];
            s_base_u += t;
        }
        __syncthreads();
    }
    for (int site = threadIdx.x; site < ns; site += blockDim.x) {
        const int l = kap[site];
        if (l != 0)
            for (int n = 0; n < N; n++) W[(size_t)n * ns + site] = c_make(l - 1 == n ? 1.0 : 0.0, 0.0);
    }
}

// W[w] (ns x N) = U (ns x N) * X_b (N x N), complex, shared-memory tiled FP64 FMA.  grid (tiles_m * tiles_n, nw, 2).
template <int BM, int BN, int BK>
__global__ void __launch_bounds__(256)
k_gemm_W_c(DevState S, const int *__restrict__ list, const double *__restrict__ X_up,
           const double *__restrict__ X_dn, const int *__restrict__ status,
           const int *__restrict__ urow_base = nullptr, int urow_stride = 0) {
    __shared__ cplx As[BK][BM];
    __shared__ cplx Bs[BK][BN + 1];
    const int b = blockIdx.y, spin = blockIdx.z;
    if (b >= batch_count(S, list)) return;
    if (status[2 * b] | status[2 * b + 1]) return;
    const int w = list ? list[b] : b;
    const int ns = S.ns, N = spin ? S.n_dn : S.n_up;
    // with urow: only the M = ns - N rows on unoccupied sites are computed (row m of the product is site urow[m])
    const int *urow = urow_base ? urow_base + ((size_t)2 * b + spin) * urow_stride : nullptr;
    const int Mrows = urow ? ns - N : ns;
    const int tiles_m = (ns + BM - 1) / BM;
    const int tm = blockIdx.x % tiles_m, tn = blockIdx.x / tiles_m;
    if (tn * BN >= N || tm * BM >= Mrows) return;
    const cplx *U = cW(spin ? S.U_dn : S.U_up, 0);
    const cplx *X = cW(spin ? X_dn : X_up, (size_t)b * N * N);
    cplx *W = cW(spin ? S.W_dn : S.W_up, (size_t)w * ns * N);
    const int m0 = tm * BM, n0 = tn * BN;
    const int tid = threadIdx.x;
    constexpr int TM = BM / 16, TN = BN / 16;
    const int tx = tid & 15, ty = tid >> 4;
    cplx acc[TM][TN];
#pragma unroll
    for (int a = 0; a < TM; a++)
#pragma unroll
        for (int c = 0; c < TN; c++) acc[a][c] = c_make(0.0, 0.0);
    for (int k0 = 0; k0 < N; k0 += BK) {
        for (int e = tid; e < BK * BM; e += 256) {
            const int kk = e / BM, mm = e - kk * BM;
            const int m = m0 + mm, k = k0 + kk;
            As[kk][mm] = (m < Mrows && k < N) ? U[(size_t)k * ns + (urow ? urow[m] : m)] : c_make(0.0, 0.0);
        }
        for (int e = tid; e < BK * BN; e += 256) {
            const int nn = e / BK, kk = e - nn * BK;
            const int n = n0 + nn, k = k0 + kk;
            Bs[kk][nn] = (n < N && k < N) ? X[(size_t)n * N + k] : c_make(0.0, 0.0);
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < BK; kk++) {
            cplx av[TM], bv[TN];
#pragma unroll
            for (int a = 0; a < TM; a++) av[a] = As[kk][tx + 16 * a];
#pragma unroll
            for (int c = 0; c < TN; c++) bv[c] = Bs[kk][ty + 16 * c];
#pragma unroll
            for (int a = 0; a < TM; a++)
#pragma unroll
                for (int c = 0; c < TN; c++) acc[a][c] = c_fma(av[a], bv[c], acc[a][c]);
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
            if (m < Mrows) W[(size_t)n * ns + (urow ? urow[m] : m)] = acc[a][c];
        }
    }
}

// O_L = real( sum_bonds Sz_i Sz_j + sum_flips (-1/2) W_up[K_up, l_up] W_dn[K_dn, l_dn] ), one warp per walker
__global__ void __launch_bounds__(256)
k_measure_c(DevState S, double *__restrict__ ol_out, int accumulate) {
    const int w = (int)((blockIdx.x * (unsigned)blockDim.x + threadIdx.x) >> 5);
    const int lane = threadIdx.x & 31;
    if (w >= S.nw) return;
    const int ns = S.ns;
    const int *kup = S.kup + (size_t)w * ns;
    const int *kdn = S.kdn + (size_t)w * ns;
    const cplx *Wu = cW(S.W_up, (size_t)w * ns * S.n_up);
    const cplx *Wd = cW(S.W_dn, (size_t)w * ns * S.n_dn);
    double flips = 0.0;
    int diag4 = 0, bad = 0;
    for (int b = lane; b < S.n_bonds; b += 32) {
        const int i = S.bi[b], j = S.bj[b];
        const int iu = kup[i], ju = kup[j], id = kdn[i], jd = kdn[j];
        if (ju != 0 && id != 0) flips += -0.5 * c_mul(Wu[(size_t)(ju - 1) * ns + i], Wd[(size_t)(id - 1) * ns + j]).x;
        if (iu != 0 && jd != 0) flips += -0.5 * c_mul(Wu[(size_t)(iu - 1) * ns + j], Wd[(size_t)(jd - 1) * ns + i]).x;
        const int oi = (iu != 0) + (id != 0), oj = (ju != 0) + (jd != 0);
        if (oi != 1 || oj != 1) bad = 1;
        diag4 += (iu != 0 ? 1 : -1) * (ju != 0 ? 1 : -1);
    }
    flips = warp_sum_f64(flips);
    diag4 = warp_sum_int(diag4);
    bad = warp_sum_int(bad);
    if (lane == 0) {
        const double OL = flips + 0.25 * (double)diag4;
        if (bad) atomicOr(&S.flags[w], KDSL_FLAG_BAD_SITE_DEV);
        if (ol_out) ol_out[w] = OL;
        if (accumulate && !(S.flags[w] & 1)) {
            S.ol_last[w] = OL;
            S.ol_sum[w] += OL;
            S.ol_sq[w] += OL * OL;
            S.ol_n[w] += 1ull;
        }
    }
}
