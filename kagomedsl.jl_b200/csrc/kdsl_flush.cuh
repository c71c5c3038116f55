// kdsl_flush.cuh -- the HBM-bound pass of the delayed update through a bulk-async shared-memory ring (flush_variant 3):
//     W0 += C G,   C = W0[:, (l_m)],   G = -T Rt          (Woodbury form, see kdsl_woodbury.cuh)
//   k_flush_G    one CTA per (listed walker, species): G (KPAD x N) by DMMA from the row copies, written ONCE in DMMA
//                fragment order to a global buffer with a bulk-async store (shared -> global);
//   k_flush_tma  persistent CTAs; a producer lane requests [8 columns x <= 216 rows] stages with cp.async.bulk (SASS
//                UBLKCP) + mbarrier, nine consumer warps pull fragments with LDS.128, hand the stage back and run the
//                DMMAs and 128-bit stores; G arrives with one bulk copy per matrix visit.
// MEASURED (round 2, 432 sites, 4096 walkers, ~215 walkers = 642 MB per launch, every 8 sweeps): this pipeline reaches
// 3.8-4.0 TB/s including k_flush_G against 4.6 TB/s of the register-pipelined k_flush_wb (kdsl_woodbury.cuh), which
// therefore stays the default.  What the experiments showed (profiles/r2_flush_experiments.txt):
//   * the launch is short (~120 us) and quantised: 692 items of 216 rows over 296 CTAs = 2.3 rounds; a whole-matrix
//     variant (RT = 6, six 28 KB stages, one CTA per SM) was slower still (3.0-3.4 TB/s), so contiguity of the stream is
//     not what limits it; self-scheduling with shrinking items or a static even split of the tile range lose more to
//     partially filled warps / extra G prologues than they gain in balance (3.7-3.8 TB/s);
//   * with the DMMAs switched off the ring moves 4.1 TB/s (incl. k_flush_G): the pipeline, not the arithmetic, is the
//     limit at this launch size; with them on the tensor pipe is 56 % busy and DRAM 51 % (ncu), i.e. the two overlap
//     poorly inside one CTA and both run at half speed;
//   * k_flush_G costs 17 us per launch (13 %), a price the in-kernel prologue of k_flush_wb does not pay twice over.
// The kernels stay selectable for further work; the production path does not launch them.
#pragma once
#include "kdsl_common.cuh"
#include "kdsl_woodbury.cuh"

// ---- mbarrier / bulk-async helpers (sm_90+ PTX) ----
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait_parity(unsigned bar, unsigned parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "MBAR_WAITP_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@!p bra MBAR_WAITP_%=;\n\t"
        "}" ::"r"(bar), "r"(parity) : "memory");
}
// global -> shared, completion counted in bytes on the mbarrier (size and both addresses multiples of 16)
__device__ __forceinline__ void bulk_g2s(unsigned dst_smem, const void *src, unsigned bytes, unsigned bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst_smem), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
// shared -> global
__device__ __forceinline__ void bulk_s2g(void *dst, unsigned src_smem, unsigned bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src_smem), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit_wait_read() {
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
__device__ __forceinline__ void bulk_commit_wait_all() {
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// G of the listed walkers, fragment-major (r = column j, k = m), Np8 * KPAD doubles per (entry, species).
template <int KPAD>
__global__ void __launch_bounds__(288, 2)
k_flush_G(DevState S, const int *__restrict__ list, const int *__restrict__ count_ptr, int count_fixed,
          double *__restrict__ Gbuf, size_t Gstride) {
    constexpr int KS = KPAD / 4, MT = (KPAD + 7) / 8;
    extern __shared__ __align__(128) double gsm[];       // G, frag-major
    __shared__ double sT[MT * 8 * KPAD];                 // T, frag-major (r = m, k = n), zero padded
    __shared__ int sL[KPAD];
    const int count = count_ptr ? *count_ptr : count_fixed;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int gr = lane >> 2, tg = lane & 3;
    for (int item = blockIdx.x; item < 2 * count; item += gridDim.x) {
        const int e = item >> 1, spin = item & 1;
        const int w = list ? list[e] : e;
        const int cnt = S.fcnt[2 * w + spin];
        if (cnt == 0) continue;                           // (uniform: nothing pending for this species)
        const int N = spin ? S.n_dn : S.n_up;
        const int ctiles = (N + 7) >> 3;
        const WbView v = wb_view(S, w, spin);
        __syncthreads();                                  // (the previous item's bulk store has been waited for by thread 0)
        for (int x = tid; x < MT * 8 * KPAD; x += 288) {
            const int m = x / KPAD, n = x - m * KPAD;
            sT[frag_idx(m, n, KPAD)] = (m < cnt && n < cnt) ? v.T[m * S.kmax + n] : 0.0;
        }
        if (tid < KPAD) sL[tid] = tid < cnt ? v.Ls[tid] : -1;
        __syncthreads();
        // Rt[n][j] = W0[K_n, j] - delta(l_n, j) from the row copies made when the moves were accepted; G = -T Rt
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
                    double *dst = gsm + ((((size_t)ct * KS) + (m >> 2)) << 5) + (m & 3);
                    dst[(2 * tg) << 2] = -d0;
                    dst[(2 * tg + 1) << 2] = -d1;
                }
            }
        }
        fence_async_smem();                               // generic-proxy writes -> visible to the bulk copy engine
        __syncthreads();
        if (tid == 0) {
            bulk_s2g(Gbuf + (size_t)item * Gstride, smem_u32(gsm), (unsigned)(ctiles * 8 * KPAD * sizeof(double)));
            bulk_commit_wait_read();                      // the shared-memory source may be overwritten after this
        }
    }
    if (tid == 0) bulk_commit_wait_all();                 // writes complete before the kernel ends (belt and braces)
}

// ---- third generation: the W0 tiles travel through a shared-memory ring filled by bulk-async copies -------------------
// ncu on k_flush_stream (profiles/r2k_flushstream_*): 70 % of the warp samples sit on the first DMMA of a tile waiting for
// its loads (long scoreboard), DRAM at 56 % of peak: with two tiles per warp in flight in REGISTERS (96 registers, two
// CTAs per SM) there are ~55 KB per SM on the way, not enough for the latency of a loaded HBM.  Here the in-flight data
// lives in shared memory: a producer warp (one elected lane) keeps NST stages of [8 columns x <= 216 rows] (13.8 KB each)
// requested with cp.async.bulk (one copy per column: 1728 contiguous bytes), completion counted by an mbarrier per stage;
// the nine consumer warps wait for a stage, pull their fragments with LDS.128, hand the stage back (mbarrier arrive) and
// only then run the DMMAs and the 128-bit stores, so the refill overlaps the arithmetic.  The ring does NOT drain between
// work pieces: the producer owns the scheduling (the two-level self-scheduling of k_flush_stream), publishes piece
// descriptors in shared memory and already requests the first tiles of the next piece while the consumers finish the
// current one; only the G operand (one buffer) is handed over at a piece boundary when the matrix changes.
// Column stride in a stage = (rows | 8) * 8 bytes == 64 (mod 128): the two columns a quarter warp touches per LDS.128
// phase fall on disjoint banks.
struct FlushPiece {
    int w, spin, t0, nt, cnt, m, need_g, pad;
};
struct FlushStageMeta {
    int piece, ct, flags, pad;
};
#define KDSL_FL_LAST 1       /* last column tile of its piece */
#define KDSL_FL_RELEASE_G 2  /* ... and the next piece brings another G: the consumers hand the G buffer back */
#define KDSL_FL_END 4        /* no more work */

#ifdef KDSL_SPIN_GUARD
#define KDSL_SPIN_DECL unsigned long long spins_ = 0
#define KDSL_SPIN_CHECK() do { if (++spins_ > (1ull << 26)) __trap(); } while (0)
#else
#define KDSL_SPIN_DECL
#define KDSL_SPIN_CHECK() do { } while (0)
#endif
__device__ __forceinline__ void mbar_wait_guarded(unsigned bar, unsigned parity) {
    KDSL_SPIN_DECL;
    unsigned done = 0;
    do {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        KDSL_SPIN_CHECK();
    } while (!done);
}
__device__ __forceinline__ void mbar_arrive_plain(unsigned bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

// RT = row tiles per consumer warp (a piece has up to 9 RT tiles = 72 RT rows).  RT = 6 makes a piece a WHOLE 432-row
// matrix: its columns are adjacent in memory, so the CTA walks one contiguous 746 KB range front to back like the flat
// rank-1 kernel does (k_update_ldg: 6.7 TB/s), instead of 1728-byte pieces at a 3456-byte stride; one CTA per SM then,
// with NST = 6 stages of 28 KB in flight.
template <int KPAD, int NST, int RT>
__global__ void __launch_bounds__(320, RT <= 3 ? 2 : 1)
k_flush_tma(DevState S, const int *__restrict__ list, const int *__restrict__ count_ptr, int count_fixed,
            int *__restrict__ work_counter, const double *__restrict__ Gbuf, size_t Gstride, int dbg) {
    constexpr int KS = KPAD / 4, NCW = 9, MAXT = NCW * RT;
    constexpr int CST_MAX = ((MAXT * 8) | 8) * 8;         // bytes per column of a stage
    constexpr int STAGE_BYTES = 8 * CST_MAX;
    extern __shared__ __align__(128) unsigned char dyn[]; // G (frag-major) | NST stages
    __shared__ __align__(8) unsigned long long s_full[NST], s_empty[NST], s_gfull, s_gempty;
    __shared__ FlushStageMeta s_meta[NST];
    __shared__ FlushPiece s_piece[4];
    const int count = count_ptr ? *count_ptr : count_fixed;
    const int ns = S.ns;
    const int nrt8 = (ns + 7) >> 3;                       // 8-row tiles per matrix
    const int total = count * 2 * nrt8;
    const int Nmax8 = (max(S.n_up, S.n_dn) + 7) & ~7;
    const int gbytes_max = Nmax8 * KPAD * (int)sizeof(double);
    const double *fsm = reinterpret_cast<const double *>(dyn);
    unsigned char *ring = dyn + ((gbytes_max + 127) & ~127);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        for (int s = 0; s < NST; s++) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&s_full[s])) : "memory");
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&s_empty[s])), "r"(NCW) : "memory");
        }
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&s_gfull)) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&s_gempty)), "r"(NCW) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp == NCW) {
        // ================= producer: scheduling + bulk copies (one elected lane) =================
        if (lane != 0) return;
        unsigned it = 0, pc = 0, gloads = 0;
        // STATIC even split of the flattened (matrix, tile) range over the CTAs (see k_flush_stream)
        int lo = (int)((long long)total * blockIdx.x / gridDim.x);
        const int hi = (int)((long long)total * (blockIdx.x + 1) / gridDim.x);
        int cur_m = -1;
        (void)work_counter;
        auto next_piece = [&](FlushPiece &P) -> bool {
            for (;;) {
                if (lo >= hi) return false;
                const int m = lo / nrt8;
                const int t0 = lo - m * nrt8;
                const int nt = min(min(hi - lo, nrt8 - t0), MAXT);
                lo += nt;
                const int e = m >> 1, spin = m & 1;
                const int w = list ? list[e] : e;
                const int cnt = S.fcnt[2 * w + spin];
                if (cnt == 0) continue;                   // nothing pending for this species
                P.w = w; P.spin = spin; P.t0 = t0; P.nt = nt; P.cnt = cnt; P.m = m; P.need_g = 0; P.pad = 0;
                return true;
            }
        };
        FlushPiece cur, nxt;
        bool have = next_piece(cur);
        while (have) {
            const bool have_n = next_piece(nxt);
            const bool release = have_n && nxt.m != cur.m;
            cur.need_g = cur.m != cur_m;
            const int N = cur.spin ? S.n_dn : S.n_up;
            const int ctiles = (N + 7) >> 3;
            s_piece[pc & 3] = cur;                        // (slot of piece pc - 4: its stages were handed back long ago)
            if (cur.need_g) {
                if (gloads > 0) mbar_wait_guarded(smem_u32(&s_gempty), (gloads - 1) & 1u);
                const unsigned gb = (unsigned)(ctiles * 8 * KPAD * sizeof(double));
                mbar_expect_tx(smem_u32(&s_gfull), gb);
                bulk_g2s(smem_u32(dyn), Gbuf + (size_t)cur.m * Gstride, gb, smem_u32(&s_gfull));
                gloads++;
                cur_m = cur.m;
            }
            const double *W0 = (cur.spin ? S.W_dn : S.W_up) + (size_t)cur.w * ns * N;
            const int r0 = cur.t0 << 3;
            const unsigned colbytes = (unsigned)((min(r0 + cur.nt * 8, ns) - r0) * sizeof(double));
            const unsigned cst = (unsigned)(((cur.nt * 8) | 8) * sizeof(double));
            for (int ct = 0; ct < ctiles; ct++, it++) {
                const unsigned s = it % NST;
                mbar_wait_guarded(smem_u32(&s_empty[s]), ((it / NST) & 1u) ^ 1u);
                FlushStageMeta mt;
                mt.piece = (int)pc; mt.ct = ct; mt.pad = 0;
                mt.flags = ct == ctiles - 1 ? (KDSL_FL_LAST | (release ? KDSL_FL_RELEASE_G : 0)) : 0;
                s_meta[s] = mt;
                const int ncols = min(8, N - (ct << 3));
                const unsigned bar = smem_u32(&s_full[s]);
                mbar_expect_tx(bar, colbytes * ncols);
                const unsigned dst = smem_u32(ring + (size_t)s * STAGE_BYTES);
                for (int c = 0; c < ncols; c++)
                    bulk_g2s(dst + c * cst, W0 + (size_t)((ct << 3) + c) * ns + r0, colbytes, bar);
            }
            pc++;
            cur = nxt;
            have = have_n;
        }
        {   // end marker
            const unsigned s = it % NST;
            mbar_wait_guarded(smem_u32(&s_empty[s]), ((it / NST) & 1u) ^ 1u);
            FlushStageMeta mt;
            mt.piece = -1; mt.ct = 0; mt.flags = KDSL_FL_END; mt.pad = 0;
            s_meta[s] = mt;
            mbar_arrive_plain(smem_u32(&s_full[s]));
        }
        return;
    }

    // ================= consumers: nine warps, each a strip of up to three 8-row tiles =================
    const int gr = lane >> 2, tg = lane & 3;
    unsigned gwaits = 0;
    double af[RT][KS];                                    // C^T fragments (k = m, n = row): W0[row, l_m]
    bool rv[RT];
    double2 *gbase = nullptr;                             // W0 + gr * ns + first row of this warp + 2 tg
    size_t cstride = 0;
    int N = 0, lrow = 0;                                  // local row offset of this warp's strip inside the stage
    unsigned cst = 0;
    bool active = false;                                  // this warp has rows in the current piece
#pragma unroll
    for (int t = 0; t < RT; t++) rv[t] = false;
    for (unsigned it = 0;; it++) {
        const unsigned s = it % NST;
        mbar_wait_guarded(smem_u32(&s_full[s]), (it / NST) & 1u);
        const FlushStageMeta mt = s_meta[s];
        if (mt.flags & KDSL_FL_END) break;
        if (mt.ct == 0) {
            // ---- a new piece: this warp's rows, the C fragments (before any of these rows is overwritten), G ----
            const FlushPiece P = s_piece[mt.piece & 3];
            N = P.spin ? S.n_dn : S.n_up;
            double *W0 = (P.spin ? S.W_dn : S.W_up) + (size_t)P.w * ns * N;
            const int *Ls = S.wbL + ((size_t)P.w * 2 + P.spin) * S.kmax;
            const int rtw = (P.nt + NCW - 1) / NCW;       // <= RT
            const int tw0 = P.t0 + warp * rtw;
            const int tw1 = min(P.t0 + P.nt, tw0 + rtw);
            const int r0 = tw0 << 3;
            lrow = (tw0 - P.t0) << 3;
            cst = (unsigned)(((P.nt * 8) | 8) * sizeof(double));
            active = tw0 < tw1;
#pragma unroll
            for (int q = 0; q < KS; q++) {
                const int mm = 4 * q + tg;
                const int lm = mm < P.cnt ? Ls[mm] : 0;
#pragma unroll
                for (int t = 0; t < RT; t++) {
                    const int row = r0 + 8 * t + gr;
                    af[t][q] = (tw0 + t < tw1 && row < ns && mm < P.cnt) ? W0[(size_t)lm * ns + row] : 0.0;
                }
            }
#pragma unroll
            for (int t = 0; t < RT; t++) rv[t] = tw0 + t < tw1 && r0 + 8 * t + 2 * tg < ns;   // ns is even: both rows or none
            gbase = reinterpret_cast<double2 *>(W0 + (size_t)gr * ns + r0 + 2 * tg);
            cstride = (size_t)4 * ns;                     // 8 columns, in double2 units
            if (P.need_g) {
                mbar_wait_guarded(smem_u32(&s_gfull), gwaits & 1u);
                gwaits++;
            }
        }
        // ---- this warp's fragments of the stage, then the stage goes back to the producer ----
        double2 cc[RT];
        const bool cok = (mt.ct << 3) + gr < N;
        {
            const unsigned char *st = ring + (size_t)s * STAGE_BYTES + gr * cst + (lrow + 2 * tg) * sizeof(double);
#pragma unroll
            for (int t = 0; t < RT; t++)
                cc[t] = (rv[t] && cok) ? *reinterpret_cast<const double2 *>(st + t * 64) : make_double2(0.0, 0.0);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive_plain(smem_u32(&s_empty[s]));
        if (active) {
#pragma unroll
            for (int q = 0; q < KS; q++) {
                if (dbg & 1) break;                       // (developer probe "flush_dbg": the memory-only rate of this pipeline)
                const double gf = fsm[(((mt.ct * KS) + q) << 5) + lane];
#pragma unroll
                for (int t = 0; t < RT; t++) dmma_8x8x4(cc[t].x, cc[t].y, gf, af[t][q]);
            }
#pragma unroll
            for (int t = 0; t < RT; t++)
                if (rv[t] && cok) gbase[(size_t)mt.ct * cstride + 4 * t] = cc[t];
        }
        if (mt.flags & KDSL_FL_RELEASE_G) {
            __syncwarp();
            if (lane == 0) mbar_arrive_plain(smem_u32(&s_gempty));
        }
    }
}
