// kdsl_inverse_cl_c.cuh -- ComplexF64 engine: batched in-place inversion of the COMPLEX tilde_U by thread-block clusters,
// one matrix per cluster (reference work: reevaluateW!, src/MonteCarlo.jl:55-66, `tilde_U \ I` with ComplexF64 U for a
// Peierls flux B != 0, src/Hamiltonian.jl:201-204).
//
// Round 1 / 2 inverted the real 2N x 2N embedding [[X, -Y], [Y, X]] on the real kernels: 2 (2N)^3 = 16 N^3 real flop and
// 2N pivots per matrix.  A complex elimination needs N pivots and 8 N^3 real flop.  This kernel is k_inverse_cl's scheme
// (kdsl_inverse_cl.cuh) in complex arithmetic on SPLIT planes: the matrix is kept as two column-major Np x Np real arrays
// (re, im), so every DMMA operand fragment is a plain real fragment and a complex block product is four real ones,
//     Cr += Xr Mr - Xi Mi,   Ci += Xr Mi + Xi Mr      (the minus sign is carried by a negated copy of the Mi fragment).
//
//   CTA 0 of the cluster ("P", thread = matrix row): factors the panel of NB = 8 complex columns in registers (implicit row
//       pivoting: largest |z|^2 by its top 28 bits over the rows that were never a pivot, ties to the lowest row; zero or
//       non-finite => singular) while the other CTAs apply the previous block step; before that it brings the next panel's
//       columns up to date (DMMA, all 16 warps).
//   CTAs 1 .. CL-1 ("G"): trailing update A[:, J] += (R_s - E) A_old[(p_q), J] for the column-tile groups dealt round-robin
//       to them; items (group, row slice) over the 16 warps, the next row tile's loads in flight behind the DMMAs.
//   Exchange per block step through a double-buffered per-cluster scratch in global memory and ONE barrier.cluster.
// Stored layout as in k_inverse_v4: with p_k the pivot row of elimination step k, S[p_k, c] = inv(A)[k, p_c];
// colsrc[i] = step at which row i was the pivot (k_unsplit_c reads both maps).
#pragma once
#include "kdsl_common.cuh"
#include "kdsl_refresh.cuh"
#include "kdsl_refresh_fast.cuh"
#include "kdsl_inverse_cl.cuh"
#include "kdsl_complex.cuh"

// doubles of per-cluster scratch: two buffers of (re, im) operands + two pivot-row buffers
__host__ __device__ inline size_t inverse_clc_scratch_doubles(int NB, int NpMax) {
    return (size_t)4 * NB * NpMax + (2 * KDSL_CL_PGS * sizeof(int) + 7) / 8;
}
inline size_t inverse_clc_smem(int NB, int NpMax) {
    return ((size_t)4 * NB * NpMax + 2 * NB + 2) * sizeof(double) + ((size_t)16 + 4 + NB) * sizeof(int);
}

// tilde_U (src/MonteCarlo.jl:92-115) of the listed walkers as split planes, padded to Np = roundup(N, 8) with an identity
// block; entry b of the batch lives at A + b * stride: re plane [Np x Np] column-major, then the im plane.
// grid (nw, 2), dynamic smem N ints.
__global__ void __launch_bounds__(256)
k_gather_tilde_split_c(DevState S, const int *__restrict__ list, double *__restrict__ A_up, double *__restrict__ A_dn,
                       int *__restrict__ status, int Np_up, int Np_dn, size_t str_up, size_t str_dn) {
    extern __shared__ int s_site_sp[];
    const int b = blockIdx.x, spin = blockIdx.y;
    if (b >= batch_count(S, list)) return;
    const int w = list ? list[b] : b;
    const int ns = S.ns, N = spin ? S.n_dn : S.n_up, Np = spin ? Np_dn : Np_up;
    const int *kap = (spin ? S.kdn : S.kup) + (size_t)w * ns;
    const cplx *U = cW(spin ? S.U_dn : S.U_up, 0);
    double *Ar = (spin ? A_dn : A_up) + (size_t)b * (spin ? str_dn : str_up);
    double *Ai = Ar + (size_t)Np * Np;
    for (int R = threadIdx.x; R < ns; R += blockDim.x) {
        const int l = kap[R];
        if (l != 0) s_site_sp[l - 1] = R;
    }
    if (threadIdx.x == 0) status[2 * b + spin] = 0;
    __syncthreads();
    for (int e = threadIdx.x; e < Np * Np; e += blockDim.x) {
        const int c = e / Np, r = e - c * Np;
        cplx v = c_make(r == c ? 1.0 : 0.0, 0.0);
        if (r < N && c < N) v = U[(size_t)c * ns + s_site_sp[r]];       // tilde_U[l, c] = U[R_l, c]
        Ar[e] = v.x;
        Ai[e] = v.y;
    }
}

template <int NB, int CT, int T, int MINB>
__global__ void __launch_bounds__(T, MINB)
k_inverse_cl_c(DevState S, const int *__restrict__ list, double *__restrict__ A_up, double *__restrict__ A_dn,
               int *__restrict__ status, int *__restrict__ colsrc_base, int Np_up, int Np_dn, size_t str_up, size_t str_dn,
               int cs_stride, double *__restrict__ scratch, int RS) {
    constexpr int NWARPS = T / 32, KS = NB / 4;
    static_assert(NB % 8 == 0 && NB < KDSL_CL_PGS && NWARPS <= 16, "panel width / key encoding");
    extern __shared__ double sm[];
    const int NpMax = max(Np_up, Np_dn);
    double *sMr = sm;                                   // [Np x NB] frag-major (r = row, k = q): Re(R - E) of the current step
    double *sMi = sMr + (size_t)NB * NpMax;             //                                       Im(R - E)
    double *sXr = sMi + (size_t)NB * NpMax;             // [Np x NB] frag-major (r = column j, k = q): Re A[p_q, j]
    double *sXi = sXr + (size_t)NB * NpMax;
    double *sRow = sXi + (size_t)NB * NpMax;            // [NB] (re, im): the scaled pivot row of the current step
    double *sRinv = sRow + 2 * NB;                      // (re, im) of 1 / pivot
    unsigned *sKey = reinterpret_cast<unsigned *>(sRinv + 2);   // [16] per-warp candidate keys
    int *sIdx = reinterpret_cast<int *>(sKey + 16);     // [4]: [0] pivot row
    int *sPivRow = sIdx + 4;                            // [NB] pivot rows of the panel (P: being factored; G: of this step)

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int gr = lane >> 2, tg = lane & 3;
    const int rank = (int)cl_ctarank(), NG = (int)cl_nctarank() - 1, grank = rank - 1;
    const bool isP = rank == 0;
    double *Mg = scratch + (size_t)cl_clusterid() * inverse_clc_scratch_doubles(NB, NpMax);
    int *Pg = reinterpret_cast<int *>(Mg + (size_t)4 * NB * NpMax);
    const int count = batch_count(S, list);
    if (tid < 16) sKey[tid] = 0u;                       // (keys of warps that do not exist never win)
    __syncthreads();

    for (int item = (int)cl_clusterid(); item < 2 * count; item += (int)cl_nclusterid()) {
        const int b = item >> 1, spin = item & 1;
        const int Np = spin ? Np_dn : Np_up;
        double *Ar = (spin ? A_dn : A_up) + (size_t)b * (spin ? str_dn : str_up);
        double *Ai = Ar + (size_t)Np * Np;
        const int nrt = Np >> 3;
        const bool has_row = isP && tid < Np;           // P: this thread owns matrix row `tid`
        bool pivoted = false;                           // my row has been a pivot
        int gstep = 0;                                  // ... at this elimination step

        // ---- P: factor the panel [k0, k0 + kw); operands to shared memory and to the scratch buffer `buf` ----
        auto factor_panel = [&](int k0, int kw, int buf) {
            double *Mgr = Mg + (size_t)buf * 2 * NB * NpMax, *Mgi = Mgr + (size_t)NB * NpMax;
            int *Pgb = Pg + buf * KDSL_CL_PGS;
            double ar[NB], ai[NB];
            int mypiv = -1;
#pragma unroll
            for (int c = 0; c < NB; c++) {
                ar[c] = (has_row && c < kw) ? __ldcg(Ar + (size_t)(k0 + c) * Np + tid) : 0.0;
                ai[c] = (has_row && c < kw) ? __ldcg(Ai + (size_t)(k0 + c) * Np + tid) : 0.0;
            }
            bool singular = false;
#pragma unroll
            for (int k = 0; k < NB; k++) {
                if (k < kw && !singular) {              // (uniform over the CTA)
                    const bool valid = has_row && !pivoted;
                    const double n2 = fma(ar[k], ar[k], ai[k] * ai[k]);
                    const unsigned hi = valid ? ((unsigned)__double2hiint(n2) & 0x7fffffffu) : 0u;
                    const unsigned mhi = __reduce_max_sync(0xffffffffu, hi);
                    const unsigned win = __ballot_sync(0xffffffffu, valid && hi == mhi);
                    const bool leader = win != 0u && lane == __ffs(win) - 1;
                    if (lane == 0) sKey[warp] = (win != 0u) ? ((mhi & 0xfffffff0u) | (unsigned)(15 - warp)) : 0u;
                    __syncthreads();
                    unsigned bk;
                    {
                        const uint4 v0 = *reinterpret_cast<const uint4 *>(sKey);
                        const uint4 v1 = *reinterpret_cast<const uint4 *>(sKey + 4);
                        const uint4 v2 = *reinterpret_cast<const uint4 *>(sKey + 8);
                        const uint4 v3 = *reinterpret_cast<const uint4 *>(sKey + 12);
                        const unsigned m0 = max(max(v0.x, v0.y), max(v0.z, v0.w)), m1 = max(max(v1.x, v1.y), max(v1.z, v1.w));
                        const unsigned m2 = max(max(v2.x, v2.y), max(v2.z, v2.w)), m3 = max(max(v3.x, v3.y), max(v3.z, v3.w));
                        bk = max(max(m0, m1), max(m2, m3));
                    }
                    if ((bk >> 4) == 0u || bk >= 0x7ff00000u) {      // zero / non-finite pivot: singular
                        singular = true;
                    } else {
                        const int wq = 15 - (int)(bk & 15u);
                        if (warp == wq && leader) {
                            // the leader publishes its RAW row and 1 / pivot; the scaling by 1 / pivot is done by every
                            // thread for itself (one complex product more per thread, sixteen fewer on the serial path)
                            const double rinv = rcp_fast(n2);        // 1 / z = conj(z) / |z|^2
                            sIdx[0] = tid;
                            sPivRow[k] = tid;
                            *reinterpret_cast<double2 *>(sRinv) = c_make(ar[k] * rinv, -(ai[k] * rinv));
                            double2 *dst = reinterpret_cast<double2 *>(sRow);
#pragma unroll
                            for (int j = 0; j < NB; j++) dst[j] = c_make(ar[j], ai[j]);
                        }
                        __syncthreads();
                        const int p = sIdx[0];
                        const cplx pinv = *reinterpret_cast<const double2 *>(sRinv);
                        const double2 *prow = reinterpret_cast<const double2 *>(sRow);
                        if (has_row) {
                            if (tid == p) {
                                pivoted = true;
                                mypiv = k;
                                gstep = k0 + k;
#pragma unroll
                                for (int j = 0; j < NB; j++) {
                                    const cplx v = j == k ? pinv : c_mul(c_make(ar[j], ai[j]), pinv);
                                    ar[j] = v.x;
                                    ai[j] = v.y;
                                }
                            } else {
                                const cplx f = c_mul(c_make(-ar[k], -ai[k]), pinv);   // -a[k] / pivot
#pragma unroll
                                for (int j = 0; j < NB; j++) {
                                    const cplx v = j == k ? f : c_fma(f, prow[j], c_make(ar[j], ai[j]));
                                    ar[j] = v.x;
                                    ai[j] = v.y;
                                }
                            }
                        }
                    }
                }
            }
            if (singular) {
                if (tid == 0) {
                    status[2 * b + spin] = 1;
                    __stcg(Pgb + NB, 1);
                }
                return;
            }
            // publish: final panel columns to the matrix, R - E to shared memory AND to the scratch (fragment order)
            if (has_row) {
#pragma unroll
                for (int c = 0; c < NB; c += 2) {
                    if (c < kw) {
#pragma unroll
                        for (int e = 0; e < 2; e++) {
                            __stcg(Ar + (size_t)(k0 + c + e) * Np + tid, ar[c + e]);
                            __stcg(Ai + (size_t)(k0 + c + e) * Np + tid, ai[c + e]);
                        }
                    }
                    const double2 vr = make_double2(ar[c] - (c == mypiv ? 1.0 : 0.0), ar[c + 1] - (c + 1 == mypiv ? 1.0 : 0.0));
                    const double2 vi = make_double2(ai[c], ai[c + 1]);
                    const int fi = frag_idx(tid, c, NB);
                    *reinterpret_cast<double2 *>(sMr + fi) = vr;
                    *reinterpret_cast<double2 *>(sMi + fi) = vi;
                    __stcg(reinterpret_cast<double2 *>(Mgr + fi), vr);
                    __stcg(reinterpret_cast<double2 *>(Mgi + fi), vi);
                }
            }
            if (tid < kw) __stcg(Pgb + tid, sPivRow[tid]);       // (written before the loop's last barrier)
            if (tid == NB) __stcg(Pgb + NB, 0);
        };
        // ---- raw pivot rows X[q, j] = A[p_q, j] of the columns j the predicate selects ----
        auto gather_cols = [&](int kw, auto pred) {
            for (int j = tid; j < Np; j += T) {
                if (pred(j)) {
                    const double *cr = Ar + (size_t)j * Np, *ci = Ai + (size_t)j * Np;
#pragma unroll
                    for (int q = 0; q < NB; q += 4) {
                        double vr[4], vi[4];
#pragma unroll
                        for (int e = 0; e < 4; e++) {
                            vr[e] = (q + e < kw) ? __ldcg(cr + sPivRow[q + e]) : 0.0;
                            vi[e] = (q + e < kw) ? __ldcg(ci + sPivRow[q + e]) : 0.0;
                        }
                        const int fi = frag_idx(j, q, NB);
                        double2 *dr = reinterpret_cast<double2 *>(sXr + fi), *di = reinterpret_cast<double2 *>(sXi + fi);
                        dr[0] = make_double2(vr[0], vr[1]);
                        dr[1] = make_double2(vr[2], vr[3]);
                        di[0] = make_double2(vi[0], vi[1]);
                        di[1] = make_double2(vi[2], vi[3]);
                    }
                }
            }
        };
        // one complex 8 x 8 tile step: (cr, ci) += X (8 columns j x 4 q) * M (4 q x 8 rows), transposed form
        auto cmma = [&](double2 &cr, double2 &ci, double xr, double xi, double mr, double mi, double mni) {
            dmma_8x8x4(cr.x, cr.y, xr, mr);
            dmma_8x8x4(ci.x, ci.y, xr, mi);
            dmma_8x8x4(cr.x, cr.y, xi, mni);
            dmma_8x8x4(ci.x, ci.y, xi, mr);
        };

        // ---- panel 0 ----
        long long t_phase = PHASE_CLOCK();                  // (developer phase clocks: make TICKS=1, tools/clc_phases.py)
        if (isP) factor_panel(0, min(NB, Np), 0);
        CL_TICK(0, 1);
        cl_sync();
        CL_TICK(0, 2);
        CL_TICK(1, 6);
        bool sing = false;
        for (int k0 = 0, s = 0; k0 < Np; k0 += NB, s++) {
            const int *Pgs = Pg + (s & 1) * KDSL_CL_PGS;
            if (__ldcg(Pgs + NB) != 0) { sing = true; break; }   // uniform over the cluster
            const int kw = min(NB, Np - k0);            // multiple of 8
            const int k1 = k0 + kw, kn = min(NB, Np - k1);   // next panel (kn <= 0: none)
            const int ex0 = k0 >> 3, exn = (kw + max(kn, 0)) >> 3;
            if (isP) {
                if (kn > 0) {
                    // columns of panel s+1: A[:, J] += (R_s - E) X_s[:, J]  (shared memory still holds step s)
                    gather_cols(kw, [&](int j) { return j >= k1 && j < k1 + kn; });
                    __syncthreads();
                    {
                        const int ktn = kn >> 3, nchunks = max(NWARPS / ktn, 1);
                        const int c = warp % ktn, chunk = warp / ktn;
                        if (chunk < nchunks) {
                            const int ct = (k1 >> 3) + c;
                            double xr[KS], xi[KS];
#pragma unroll
                            for (int q = 0; q < KS; q++) {
                                xr[q] = sXr[(((ct * KS) + q) << 5) + lane];
                                xi[q] = sXi[(((ct * KS) + q) << 5) + lane];
                            }
                            const size_t off = (size_t)((ct << 3) + gr) * Np + 2 * tg;
                            double2 *cpr = reinterpret_cast<double2 *>(Ar + off), *cpi = reinterpret_cast<double2 *>(Ai + off);
                            const int r_lo = chunk * nrt / nchunks, r_hi = (chunk + 1) * nrt / nchunks;
                            for (int rt = r_lo; rt < r_hi; rt++) {
                                double2 dr = __ldcg(cpr + (rt << 2)), di = __ldcg(cpi + (rt << 2));
#pragma unroll
                                for (int q = 0; q < KS; q++) {
                                    const double mr = sMr[(((rt * KS) + q) << 5) + lane], mi = sMi[(((rt * KS) + q) << 5) + lane];
                                    cmma(dr, di, xr[q], xi[q], mr, mi, -mi);
                                }
                                __stcg(cpr + (rt << 2), dr);
                                __stcg(cpi + (rt << 2), di);
                            }
                        }
                    }
                    __syncthreads();
                    CL_TICK(0, 0);
                    factor_panel(k1, kn, (s + 1) & 1);
                    CL_TICK(0, 1);
                }
            } else {
                // operands and pivot rows of step s: scratch -> shared memory (re and im planes are adjacent in both)
                {
                    const double2 *src = reinterpret_cast<const double2 *>(Mg + (size_t)(s & 1) * 2 * NB * NpMax);
                    double2 *dst = reinterpret_cast<double2 *>(sMr);
                    const int n2 = NB * NpMax;              // doubles of both planes / 2
                    for (int i = tid; i < n2; i += T) dst[i] = __ldcg(src + i);
                    if (tid < NB) sPivRow[tid] = __ldcg(Pgs + tid);
                }
                __syncthreads();
                CL_TICK(1, 3);
                // my column-tile groups: g = grank, grank + NG, ... over the tiles outside [ex0, ex0 + exn)
                const int nct = nrt - exn;
                const int groups = (nct + CT - 1) / CT;
                gather_cols(kw, [&](int j) {
                    const int ct = j >> 3;
                    if (ct >= ex0 && ct < ex0 + exn) return false;
                    const int t = ct < ex0 ? ct : ct - exn;
                    return (t / CT) % NG == grank;
                });
                __syncthreads();
                CL_TICK(1, 4);
                const int n_my = groups > grank ? (groups - grank + NG - 1) / NG : 0;
                const int items = n_my * RS;
                for (int it = warp; it < items; it += NWARPS) {
                    const int lg = it / RS, rs = it - lg * RS;
                    const int g = grank + lg * NG;
                    const int r_lo = rs * nrt / RS, r_hi = (rs + 1) * nrt / RS;
                    if (r_lo >= r_hi) continue;
                    double xr[CT][KS], xi[CT][KS];
                    double2 *cpr[CT], *cpi[CT];
                    bool cv[CT];
#pragma unroll
                    for (int c = 0; c < CT; c++) {
                        const int t = g * CT + c;
                        cv[c] = t < nct;
                        const int ct = cv[c] ? (t < ex0 ? t : t + exn) : 0;
#pragma unroll
                        for (int q = 0; q < KS; q++) {
                            xr[c][q] = cv[c] ? sXr[(((ct * KS) + q) << 5) + lane] : 0.0;
                            xi[c][q] = cv[c] ? sXi[(((ct * KS) + q) << 5) + lane] : 0.0;
                        }
                        const size_t off = (size_t)((ct << 3) + gr) * Np + 2 * tg;
                        cpr[c] = reinterpret_cast<double2 *>(Ar + off);
                        cpi[c] = reinterpret_cast<double2 *>(Ai + off);
                    }
                    double2 cur_r[CT], cur_i[CT], nxt_r[CT], nxt_i[CT];
#pragma unroll
                    for (int c = 0; c < CT; c++) {
                        cur_r[c] = cv[c] ? __ldcg(cpr[c] + (r_lo << 2)) : make_double2(0.0, 0.0);
                        cur_i[c] = cv[c] ? __ldcg(cpi[c] + (r_lo << 2)) : make_double2(0.0, 0.0);
                    }
                    for (int rt = r_lo; rt < r_hi; rt++) {
                        const int rn = min(rt + 1, r_hi - 1);
#pragma unroll
                        for (int c = 0; c < CT; c++) {
                            nxt_r[c] = cv[c] ? __ldcg(cpr[c] + (rn << 2)) : make_double2(0.0, 0.0);
                            nxt_i[c] = cv[c] ? __ldcg(cpi[c] + (rn << 2)) : make_double2(0.0, 0.0);
                        }
#pragma unroll
                        for (int q = 0; q < KS; q++) {
                            const double mr = sMr[(((rt * KS) + q) << 5) + lane], mi = sMi[(((rt * KS) + q) << 5) + lane];
                            const double mni = -mi;
#pragma unroll
                            for (int c = 0; c < CT; c++) cmma(cur_r[c], cur_i[c], xr[c][q], xi[c][q], mr, mi, mni);
                        }
#pragma unroll
                        for (int c = 0; c < CT; c++) {
                            if (cv[c]) {
                                __stcg(cpr[c] + (rt << 2), cur_r[c]);
                                __stcg(cpi[c] + (rt << 2), cur_i[c]);
                            }
                            cur_r[c] = nxt_r[c];
                            cur_i[c] = nxt_i[c];
                        }
                    }
                }
            }
            CL_TICK(1, 5);
            cl_sync();
            CL_TICK(0, 2);
            CL_TICK(1, 6);
        }
#ifdef KDSL_PHASE_TICKS
        if (blockIdx.x == 0 && tid == 0) g_inv_phase_cycles[15] += 1;
#endif
        if (sing) cl_sync();                            // nobody re-reads the flag after P has moved on to the next item
        // ---- index map for the consumer: colsrc[i] = elimination step at which row i was the pivot ----
        if (has_row && !sing) colsrc_base[((size_t)2 * b + spin) * cs_stride + tid] = gstep;
    }
}

// k_unsplit_c: the stored result of the implicit-pivoting inverse is S[p_k, c] = inv(A)[k, p_c] (colsrc[i] = step at which
// row i was the pivot); X[k, j] = inv(A)[k, j] as a plain column-major interleaved complex N x N matrix for k_gemm_W_c, plus
// the ordered list of the sites NOT occupied by the species (the non-trivial rows of W) and the unit rows W[R_l, :] = e_l
// of the occupied ones (written here, never computed).  grid (nw, 2), dynamic smem Np ints.
__global__ void __launch_bounds__(256)
k_unsplit_c(DevState S, const int *__restrict__ list, const double *__restrict__ A_up, const double *__restrict__ A_dn,
            double *__restrict__ X_up, double *__restrict__ X_dn, const int *__restrict__ status,
            const int *__restrict__ colsrc_base, int Np_up, int Np_dn, size_t str_up, size_t str_dn, int cs_stride,
            int *__restrict__ urow_base, int urow_stride, int write_X) {
    extern __shared__ int s_row_of_step_sp[];
    const int b = blockIdx.x, spin = blockIdx.y;
    if (b >= batch_count(S, list)) return;
    if (status[2 * b] | status[2 * b + 1]) return;
    const int N = spin ? S.n_dn : S.n_up, Np = spin ? Np_dn : Np_up;
    const double *Ar = (spin ? A_dn : A_up) + (size_t)b * (spin ? str_dn : str_up);
    const double *Ai = Ar + (size_t)Np * Np;
    cplx *X = cW(spin ? X_dn : X_up, (size_t)b * N * N);
    const int *colsrc = colsrc_base + ((size_t)2 * b + spin) * cs_stride;
    for (int i = threadIdx.x; i < Np; i += blockDim.x) s_row_of_step_sp[colsrc[i]] = i;
    __syncthreads();
    for (int e = threadIdx.x; e < (write_X ? N * N : 0); e += blockDim.x) {   // (k_gemm_W_dmma_c reads the planes itself)
        const int j = e / N, k = e - j * N;
        const size_t o = (size_t)colsrc[j] * Np + s_row_of_step_sp[k];
        X[e] = c_make(Ar[o], Ai[o]);
    }
    __shared__ int s_wsum_sp[8];
    __shared__ int s_base_sp;
    const int w = list ? list[b] : b;
    const int ns = S.ns;
    const int *kap = (spin ? S.kdn : S.kup) + (size_t)w * ns;
    int *urow = urow_base + ((size_t)2 * b + spin) * urow_stride;
    cplx *W = cW(spin ? S.W_dn : S.W_up, (size_t)w * ns * N);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_base_sp = 0;
    __syncthreads();
    for (int s0 = 0; s0 < ns; s0 += 256) {
        const int site = s0 + threadIdx.x;
        const bool un = site < ns && kap[site] == 0;
        const unsigned m = __ballot_sync(0xffffffffu, un);
        if (lane == 0) s_wsum_sp[warp] = __popc(m);
        __syncthreads();
        int off = s_base_sp;
        for (int q = 0; q < warp; q++) off += s_wsum_sp[q];
        if (un) urow[off + __popc(m & ((1u << lane) - 1u))] = site;
        __syncthreads();
        if (threadIdx.x == 0) {
            int t = 0;
            for (int q = 0; q < 8; q++) t += s_wsum_sp[q];
            s_base_sp += t;
        }
        __syncthreads();
    }
    if (!write_X) return;                                 // (k_gemm_W_dmma_c writes the unit rows with its own tiles)
    for (int site = threadIdx.x; site < ns; site += blockDim.x) {
        const int l = kap[site];
        if (l != 0)
            for (int n = 0; n < N; n++) W[(size_t)n * ns + site] = c_make(l - 1 == n ? 1.0 : 0.0, 0.0);
    }
}

// W[w][unoccupied sites, :] = U[unoccupied sites, :] * inv(tilde_U), complex, on the FP64 tensor pipe, straight from the split
// planes k_inverse_cl_c left behind: with rho the stored row index, W[m, j] = sum_rho U[m, colsrc[rho]] S[rho, colsrc[j]]
// (both permutations of the implicit pivoting folded into the operand loads; rows rho >= N are the identity padding).
// A complex block product is four real DMMAs (the minus sign rides on a negated copy of the Im(U) fragment).
// CTA = 3 warps, tile 72 sites x 24 columns (warp tile 24 x 24), four CTAs per SM (12 warps = 3 per sub-partition: a 9-warp
// CTA loses 10-25 % to the 4-way split); register-staged double buffering, fragment-major shared layout.
// grid (tiles_m * tiles_n, nw, 2).  The urow lists come from k_unsplit_c; the unit rows of the occupied sites are written
// here, by the CTA whose row tile spans them.
template <int KT>
__global__ void __launch_bounds__(96, 4)
k_gemm_W_dmma_c(DevState S, const int *__restrict__ list, const double *__restrict__ A_up, const double *__restrict__ A_dn,
                const int *__restrict__ status, const int *__restrict__ colsrc_base, int Np_up, int Np_dn, size_t str_up,
                size_t str_dn, int cs_stride, const int *__restrict__ urow_base, int urow_stride) {
    constexpr int TM = 72, TN = 24, NT = 96, KS = KT / 4;
    constexpr int PA = TM * KT / NT, PB = TN * KT / NT;
    static_assert(TM * KT % NT == 0 && TN * KT % NT == 0 && KT % 4 == 0, "stage must divide evenly");
    __shared__ double sAr[2][TM * KT], sAi[2][TM * KT], sBr[2][TN * KT], sBi[2][TN * KT];
    __shared__ int sSrc[TN];
    __shared__ int sRowSite[TM + 1];
    __shared__ int sKp[512 + KT];
    const int b = blockIdx.y, spin = blockIdx.z;
    if (b >= batch_count(S, list)) return;
    if (status[2 * b] | status[2 * b + 1]) return;
    const int w = list ? list[b] : b;
    const int ns = S.ns, N = spin ? S.n_dn : S.n_up, Np = spin ? Np_dn : Np_up;
    const int M = ns - N;
    const int tiles_m = (M + TM - 1) / TM;
    const int tm = blockIdx.x % tiles_m, tn = blockIdx.x / tiles_m;
    const int m0 = tm * TM, n0 = tn * TN;
    if (n0 >= N) return;
    const cplx *U = cW(spin ? S.U_dn : S.U_up, 0);
    const double *Sr = (spin ? A_dn : A_up) + (size_t)b * (spin ? str_dn : str_up);
    const double *Si = Sr + (size_t)Np * Np;
    const int *colsrc = colsrc_base + ((size_t)2 * b + spin) * cs_stride;
    const int *urow = urow_base + ((size_t)2 * b + spin) * urow_stride;
    cplx *W = cW(spin ? S.W_dn : S.W_up, (size_t)w * ns * N);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid < TN) sSrc[tid] = (n0 + tid < N) ? colsrc[n0 + tid] : -1;
    if (tid <= TM) sRowSite[tid] = (m0 + tid < M) ? urow[m0 + tid] : -1;
    for (int k = tid; k < 512 + KT; k += NT) sKp[k] = k < N ? colsrc[k] : 0;
    __syncthreads();

    cplx ra[PA];
    double rbr[PB], rbi[PB];
    auto load_stage = [&](int kk) {
#pragma unroll
        for (int q = 0; q < PA; q++) {
            const int e = tid + q * NT;
            const int r = e % TM, k = e / TM;                      // A: gathered rows of U, permuted columns
            const int site = sRowSite[r];
            ra[q] = (site >= 0 && kk + k < N) ? U[(size_t)sKp[kk + k] * ns + site] : c_make(0.0, 0.0);
        }
#pragma unroll
        for (int q = 0; q < PB; q++) {
            const int e = tid + q * NT;
            const int kb = e % KT, n = e / KT;                     // B: contiguous along the stored row index
            const int src = sSrc[n];
            const bool ok = src >= 0 && kk + kb < N;
            rbr[q] = ok ? Sr[(size_t)src * Np + kk + kb] : 0.0;
            rbi[q] = ok ? Si[(size_t)src * Np + kk + kb] : 0.0;
        }
    };
    auto store_stage = [&](int buf) {
#pragma unroll
        for (int q = 0; q < PA; q++) {
            const int e = tid + q * NT;
            const int fi = frag_idx(e % TM, e / TM, KT);
            sAr[buf][fi] = ra[q].x;
            sAi[buf][fi] = ra[q].y;
        }
#pragma unroll
        for (int q = 0; q < PB; q++) {
            const int e = tid + q * NT;
            const int fi = frag_idx(e / KT, e % KT, KT);
            sBr[buf][fi] = rbr[q];
            sBi[buf][fi] = rbi[q];
        }
    };
    double cr[3][3][2], ci[3][3][2];
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) cr[i][j][0] = cr[i][j][1] = ci[i][j][0] = ci[i][j][1] = 0.0;
    load_stage(0);
    store_stage(0);
    __syncthreads();
    int buf = 0;
    for (int kk = 0; kk < N; kk += KT) {
        const bool more = kk + KT < N;
        if (more) load_stage(kk + KT);
#pragma unroll
        for (int s = 0; s < KS; s++) {
            double ar[3], ai[3], an[3];
#pragma unroll
            for (int i = 0; i < 3; i++) {
                ar[i] = sAr[buf][((((warp * 3 + i) * KS) + s) << 5) + lane];
                ai[i] = sAi[buf][((((warp * 3 + i) * KS) + s) << 5) + lane];
                an[i] = -ai[i];
            }
#pragma unroll
            for (int j = 0; j < 3; j++) {
                const double br = sBr[buf][(((j * KS) + s) << 5) + lane], bi = sBi[buf][(((j * KS) + s) << 5) + lane];
#pragma unroll
                for (int i = 0; i < 3; i++) {
                    dmma_8x8x4(cr[i][j][0], cr[i][j][1], ar[i], br);
                    dmma_8x8x4(ci[i][j][0], ci[i][j][1], ar[i], bi);
                    dmma_8x8x4(cr[i][j][0], cr[i][j][1], an[i], bi);
                    dmma_8x8x4(ci[i][j][0], ci[i][j][1], ai[i], br);
                }
            }
        }
        if (more) store_stage(buf ^ 1);
        __syncthreads();
        buf ^= 1;
    }
    const int gr = lane >> 2, tg = lane & 3;
#pragma unroll
    for (int j = 0; j < 3; j++) {
#pragma unroll
        for (int e = 0; e < 2; e++) {
            const int n = n0 + 8 * j + 2 * tg + e;
            if (n >= N) continue;
#pragma unroll
            for (int i = 0; i < 3; i++) {
                const int site = sRowSite[24 * warp + 8 * i + gr];
                if (site >= 0) W[(size_t)n * ns + site] = c_make(cr[i][j][e], ci[i][j][e]);
            }
        }
    }
    // unit rows of the occupied sites inside this tile's site range [s_lo, s_hi), for this CTA's 24 columns: together with
    // the stores above every 32-byte sector of the range is completed by this CTA
    const int *kap = (spin ? S.kdn : S.kup) + (size_t)w * ns;
    const int s_lo = (tm == 0) ? 0 : sRowSite[0];
    const int s_hi = (tm == tiles_m - 1 || sRowSite[TM] < 0) ? ns : sRowSite[TM];
    const int ncols = min(TN, N - n0);
    for (int site = s_lo + tid; site < s_hi; site += NT) {
        const int l = kap[site];
        if (l != 0) {
            cplx *dst = W + (size_t)n0 * ns + site;
            for (int cc = 0; cc < ncols; cc++) dst[(size_t)cc * ns] = c_make((l - 1 == n0 + cc) ? 1.0 : 0.0, 0.0);
        }
    }
}
