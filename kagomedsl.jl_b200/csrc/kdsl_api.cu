// kdsl_api.cu -- host side of libkdsl.so: handle, device memory, launch sequencing and the
// extern "C" entry points declared in include/kdsl.h.  sm_100a only; there is no CPU path:
// every entry point fails with KDSL_ERR_CUDA when no CUDA device is usable.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>              // types and prototypes only: libnccl.so.2 is opened with dlopen at kdsl_comm_* time

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/kdsl.h"
#include "kdsl_common.cuh"
#include "kdsl_measure.cuh"
#include "kdsl_propose.cuh"
#include "kdsl_refresh.cuh"
#include "kdsl_refresh_fast.cuh"
#include "kdsl_inverse_v4.cuh"
#include "kdsl_inverse_v5.cuh"
#include "kdsl_inverse_cl.cuh"
#ifdef KDSL_DEV_VARIANTS
#include "kdsl_reeval_cl.cuh"   // cluster re-evaluation: measured slower than the product paths everywhere
#endif
#include "kdsl_inverse_cl_c.cuh"
#include "kdsl_reeval_fused.cuh"
#ifdef KDSL_DEV_VARIANTS   // superseded kernels kept for A/B measurements: `make DEV=1` (not in the product library)
#include "kdsl_inverse_v3.cuh"
#include "kdsl_delayed.cuh"
#endif
#include "kdsl_woodbury.cuh"
#include "kdsl_flush.cuh"
#include "kdsl_resident.cuh"
#include "kdsl_update.cuh"
#include "kdsl_complex.cuh"
#include "kdsl_woodbury_c.cuh"

#define KDSL_VERSION_NUM 110
#define KDSL_KTH 16          /* pending factors that trigger a flush */
#define KDSL_FLUSH_EVERY 8   /* sweeps between flush launches    */
#define KDSL_KMAX (KDSL_KTH + KDSL_FLUSH_EVERY)

static thread_local std::string g_err;

static int fail(int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define CK(call)                                                                                 \
    do {                                                                                         \
        cudaError_t e_ = (call);                                                                 \
        if (e_ != cudaSuccess)                                                                   \
            return fail(KDSL_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_),   \
                        __FILE__, __LINE__);                                                     \
    } while (0)

struct TimedSpan {
    int cls;
    cudaEvent_t a, b;
};

struct kdsl_handle_s {
    int device = 0;
    int num_sms = 148;
    DevState S{};
    cudaStream_t stream = nullptr;
    std::vector<void *> allocs;
    // refresh workspace (tilde_U / inverse), sized for all walkers
    double *A_up = nullptr, *A_dn = nullptr;
    int *status = nullptr;
    int *colsrc = nullptr;        // [nw][2][Np] column map of the pivoted inverse
    int *urow = nullptr;          // [nw][2][ns] sites not occupied by the species (non-trivial rows of W)
    int Np_up = 0, Np_dn = 0;     // tilde_U dimensions padded to a multiple of 8
    // fused re-evaluation (k_reeval_fused): transposed U with padding rows, one workspace per resident CTA
    double *UT_up = nullptr, *UT_dn = nullptr;
    double *ws_fused = nullptr;
    size_t ws_stride = 0;
    size_t fused_smem = 0;        // dynamic shared memory of k_reeval_fused (0: the fused kernel does not apply)
    int fused_NpMax = 0, fused_CpMax = 0, fused_NB = 24, fused_stage = 0;
    size_t fused_smem24 = 0, fused_smem16 = 0;   // the same for panel widths 24 (default) and 16
    int fused_stage24 = 0, fused_stage16 = 0;
    int fused_ctas = 0;           // resident CTAs of k_reeval_fused (0: one per SM)
    bool fused_small = false;     // N <= 128, M <= 128, ns <= 256: the 256-thread instantiation, two CTAs per SM
    double *Gbuf = nullptr;       // [nw][2][Gstride] flush operands G = -T Rt in DMMA fragment order (k_flush_G -> k_flush_tma; flush_variant 3 only)
    size_t Gstride = 0;
    size_t res_smem = 0;          // dynamic shared memory of k_resident (0: a walker does not fit one CTA)
    int res_ctas_per_sm = 0;      // resident CTAs of k_resident per SM
    int res_np = 0;               // its template parameter: column passes of 32 (ceil(max(N_up, N_dn) / 32), 1..4)
    size_t smem_optin = 0;        // largest dynamic shared memory k_measure_wb may use on this handle's device
    int *d_tmp_i = nullptr;       // [nw] scratch
    double *d_tmp_d = nullptr;    // [nw] scratch
    double *d_acc8 = nullptr;     // [8]
    double *d_qcos = nullptr, *d_qsin = nullptr, *d_obs = nullptr;   // extra observables (kdsl_set_observables)
    // replay staging (device)
    double *rp_r = nullptr;
    int *rp_bond = nullptr, *rp_pick = nullptr;
    size_t rp_cap = 0;            // capacity in sweeps*walkers entries
    // move staging for kdsl_update_W
    int64_t sweeps = 0;
    int parity = 0;
    bool have_config = false, W_valid = false;
    double *X_up = nullptr, *X_dn = nullptr;   // ComplexF64 mode: un-embedded complex inverses [nw][N*N] (re, im)
    ncclComm_t comm = nullptr;    // NCCL communicator of this handle's rank (kdsl_comm_init_rank / kdsl_comm_init_all)
    int comm_rank = -1, comm_size = 0;
    bool cplx = false;            // ComplexF64 mode (kdsl_create_c128): W, U, staging and workspace hold (re, im) pairs
    int64_t walker_sweeps = 0;
    // options
    int64_t refresh_every = 0;
    int update_variant = 2, update_ctas_per_sm = 0, inverse_variant = 0, gemm_variant = 0;
    int since_flush = 0;
    int flush_dbg = 0;            // developer probe bits of k_flush_tma
    int flush_variant = 0;        // 0: k_flush_wb (persistent, 128-bit), 1: k_flush
    int flush_every = KDSL_FLUSH_EVERY;   // sweeps between flush launches (kmax = kth + flush_every <= 32)
    int fuse_sweeps = 1;          // fuse consecutive proposals into one launch where the loop allows it
    int inverse_tuning = 0;
    int inverse_cluster = 4;      // CTAs per matrix of k_inverse_cl (256 < Np <= 512): 1 P + (n - 1) G; 0 = one CTA per matrix (k_inverse_v4)
    int inverse_rs = 8;           // row slices per column-tile group of its trailing update
    double *cl_scratch = nullptr; // per-cluster exchange buffers of k_inverse_cl
    int cl_scratch_clusters = 0;
    // cluster re-evaluation k_reeval_cl (256 < Np <= 512, real engine): one kernel, one matrix per cluster
    bool rcl_ok = false;          // UT_up / UT_dn are built and the sizes fit
    int rcl_NpMax = 0, rcl_CpMax = 0, rcl_nvtMax = 0, rcl_ntcMax = 0;
    double *rcl_scratch = nullptr;   // per-cluster row-major workspace + exchange buffers
    int rcl_scratch_clusters = 0;
    int reeval_cluster = 4;       // CTAs per matrix of k_reeval_cl: 1 P + (n - 1) G
    // ComplexF64 engine: complex cluster inverse on split (re, im) planes (k_inverse_cl_c)
    int Npc_up = 0, Npc_dn = 0;   // complex tilde_U dimensions padded to a multiple of 8
    double *clc_scratch = nullptr;
    int clc_scratch_clusters = 0;
    int fdc_RB = 0;               // rows per item of k_flush_dmma_c (0: it does not fit, k_flush_c runs)
    int fd2_RB = 0;               // rows per item of k_flush_dmma2_c (two CTAs per SM; 0: it does not fit)
    int reeval_rs = 4;            // row slices per column-tile pair of its trailing update
    int update_ch = 8;
    // profiling
    bool profiling = false;
    std::vector<TimedSpan> spans;
    std::vector<cudaEvent_t> ev_pool;
    double t_ms[KDSL_N_TIMERS] = {0};
    int64_t t_launch[KDSL_N_TIMERS] = {0};
    cudaEvent_t user_ev[16] = {nullptr};
};

namespace {

template <typename T>
int dev_alloc(kdsl_handle h, T **p, size_t n) {
    void *q = nullptr;
    cudaError_t e = cudaMalloc(&q, std::max<size_t>(n, 1) * sizeof(T));
    if (e != cudaSuccess)
        return fail(KDSL_ERR_CUDA, "cudaMalloc of %zu bytes failed: %s", n * sizeof(T), cudaGetErrorString(e));
    // zero-fill ON THE ENGINE'S STREAM: that stream is non-blocking, so a memset on the legacy default stream would not be
    // ordered with the kernels that use the buffer
    e = h->stream ? cudaMemsetAsync(q, 0, std::max<size_t>(n, 1) * sizeof(T), h->stream) : cudaMemset(q, 0, std::max<size_t>(n, 1) * sizeof(T));
    if (e == cudaSuccess && !h->stream) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) return fail(KDSL_ERR_CUDA, "cudaMemset failed: %s", cudaGetErrorString(e));
    h->allocs.push_back(q);
    *p = static_cast<T *>(q);
    return KDSL_OK;
}

// Host <-> device copies go through the ENGINE'S stream and are complete when this returns.  A plain cudaMemcpy runs on the
// legacy default stream, which a cudaStreamNonBlocking stream does not synchronise with -- and a pageable host-to-device
// cudaMemcpy may return once its last chunk sits in the staging buffer, before the DMA has landed: a kernel launched on the
// engine's stream right afterwards could read the old contents of the buffer's tail (seen as run-to-run differences of the
// Z_mu of the last walkers of a 512-walker handle).
cudaError_t copy_sync(kdsl_handle h, void *dst, const void *src, size_t n, cudaMemcpyKind kind) {
    cudaError_t e = cudaMemcpyAsync(dst, src, n, kind, h->stream);
    if (e != cudaSuccess) return e;
    return cudaStreamSynchronize(h->stream);
}
cudaError_t memset_sync(kdsl_handle h, void *dst, int v, size_t n) {
    cudaError_t e = cudaMemsetAsync(dst, v, n, h->stream);
    if (e != cudaSuccess) return e;
    return cudaStreamSynchronize(h->stream);
}

int use_device(kdsl_handle h) {
    if (!h) return fail(KDSL_ERR_INVALID_ARGUMENT, "null handle");
    CK(cudaSetDevice(h->device));
    return KDSL_OK;
}

cudaEvent_t get_event(kdsl_handle h) {
    if (!h->ev_pool.empty()) {
        cudaEvent_t e = h->ev_pool.back();
        h->ev_pool.pop_back();
        return e;
    }
    cudaEvent_t e;
    cudaEventCreate(&e);
    return e;
}

// fold finished spans into the per-class timers (synchronises the stream)
int flush_spans(kdsl_handle h) {
    if (h->spans.empty()) return KDSL_OK;
    CK(cudaStreamSynchronize(h->stream));
    for (auto &s : h->spans) {
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, s.a, s.b));
        h->t_ms[s.cls] += ms;
        h->ev_pool.push_back(s.a);
        h->ev_pool.push_back(s.b);
    }
    h->spans.clear();
    return KDSL_OK;
}

struct Span {
    kdsl_handle h;
    int cls;
    cudaEvent_t a = nullptr;
    Span(kdsl_handle h_, int cls_) : h(h_), cls(cls_) {
        h->t_launch[cls] += 1;
        if (h->profiling) {
            a = get_event(h);
            cudaEventRecord(a, h->stream);
        }
    }
    ~Span() {
        if (h->profiling) {
            cudaEvent_t b = get_event(h);
            cudaEventRecord(b, h->stream);
            h->spans.push_back({cls, a, b});
        }
    }
};

int grid_for_warps(int nw) { return (nw * 32 + 255) / 256; }

size_t measure_wb_smem_c(const DevState &S) {
    return ((size_t)S.kmax * (S.n_up + S.n_dn) + 2 * (size_t)S.kmax * S.kmax) * 2 * sizeof(double) + 2 * S.kmax * sizeof(int);
}

size_t measure_wb_smem(const DevState &S) {
    return ((size_t)S.kmax * (S.n_up + S.n_dn) + 2 * (size_t)S.kmax * S.kmax) * sizeof(double) + 4 * S.kmax * sizeof(int);
}

int launch_update(kdsl_handle h, int parity) {
    const DevState &S = h->S;
    const int CH = h->update_ch;
    const int tiles_up = (S.n_up + CH - 1) / CH, tiles_dn = (S.n_dn + CH - 1) / CH;
    const size_t smem = (size_t)(S.ns + CH) * sizeof(double);
    int per_sm = h->update_ctas_per_sm > 0 ? h->update_ctas_per_sm : 4;
    Span sp(h, KDSL_T_UPDATE);
    if (h->cplx) {
        k_update_c<256><<<h->num_sms * per_sm, 256, (size_t)(S.ns + CH) * 2 * sizeof(double), h->stream>>>(S, parity, tiles_up, tiles_dn, CH);
        CK(cudaGetLastError());
        return KDSL_OK;
    }
    k_update_ldg<256, 8><<<h->num_sms * per_sm, 256, smem, h->stream>>>(S, parity, tiles_up, tiles_dn, CH);
    CK(cudaGetLastError());
    return KDSL_OK;
}

#ifdef KDSL_DEV_VARIANTS
template <int NB, int RPT, int T, int MINB>
int launch_inverse_blocked(kdsl_handle h, const int *list, double *A, int spin, int Np) {
    const size_t smem = ((size_t)2 * Np * NB + NB * NB + 2 * NB + T / 32) * sizeof(double) +
                        ((size_t)T / 32 + Np + 5 * NB) * sizeof(int);
    CK(cudaFuncSetAttribute(k_inverse_blocked<NB, RPT, T, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_inverse_blocked<NB, RPT, T, MINB><<<h->S.nw, T, smem, h->stream>>>(h->S, list, A, spin, h->status, h->colsrc, Np, std::max(h->Np_up, h->Np_dn));
    CK(cudaGetLastError());
    return KDSL_OK;
}

template <int NB, int RPT, int T>
int launch_inverse_v3(kdsl_handle h, const int *list, double *A, int spin, int Np) {
    constexpr int NWARP = T / 32;
    const size_t smem = ((size_t)3 * NB * Np + 5 * NB * NB + 2 * NWARP + 2 * NWARP) * sizeof(double) +
                        ((size_t)2 * NWARP + 3 * Np + 5 * NB) * sizeof(int);
    CK(cudaFuncSetAttribute(k_inverse_v3<NB, RPT, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_inverse_v3<NB, RPT, T><<<h->S.nw, T, smem, h->stream>>>(h->S, list, A, spin, h->status, h->colsrc, Np, std::max(h->Np_up, h->Np_dn));
    CK(cudaGetLastError());
    return KDSL_OK;
}

#endif  // KDSL_DEV_VARIANTS

template <int NB, int RPT, int T, int TP, int MINB = 1, int CT = 3>
int launch_inverse_v4(kdsl_handle h, const int *list, double *A, int spin, int Np) {
    static_assert(RPT * TP >= 8 && TP <= 256, "rows per CTA");
    const size_t smem = ((size_t)2 * NB * Np + NB + 2) * sizeof(double) + ((size_t)12 + NB) * sizeof(int);
    CK(cudaFuncSetAttribute(k_inverse_v4<NB, RPT, T, TP, MINB, CT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_inverse_v4<NB, RPT, T, TP, MINB, CT><<<h->S.nw, T, smem, h->stream>>>(h->S, list, A, spin, h->status, h->colsrc, Np, std::max(h->Np_up, h->Np_dn));
    CK(cudaGetLastError());
    return KDSL_OK;
}

template <int NB, int CT, int GW = 8>
int launch_inverse_v5(kdsl_handle h, const int *list, double *A, int spin, int Np) {
    const size_t smem = ((size_t)3 * NB * Np + NB + 2) * sizeof(double) + ((size_t)12 + NB) * sizeof(int);
    CK(cudaFuncSetAttribute(k_inverse_v5<NB, CT, GW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_inverse_v5<NB, CT, GW><<<h->S.nw, 256 + 32 * GW, smem, h->stream>>>(h->S, list, A, spin, h->status, h->colsrc, Np, std::max(h->Np_up, h->Np_dn));
    CK(cudaGetLastError());
    return KDSL_OK;
}

// Both species in one launch of the cluster kernel (kdsl_inverse_cl.cuh); returns -1 when it does not apply.
int launch_inverse_cl(kdsl_handle h, const int *list) {
    const int NpMax = std::max(h->Np_up, h->Np_dn), NpMin = std::min(h->Np_up, h->Np_dn);
    const int cl = h->inverse_cluster;
    if (!(h->inverse_variant == 0 || h->inverse_variant == 7) || cl < 2 || cl > 8 || NpMax > 512 || NpMin < 8) return -1;
    if (h->inverse_variant == 0 && NpMax <= 256) return -1;
    constexpr int NB = 24, CT = 2;
    const size_t smem = inverse_cl_smem(NB, NpMax);
    if (smem > (size_t)227 * 1024) return -1;
    auto kern = k_inverse_cl<NB, CT>;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = cl; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.blockDim = dim3(512); cfg.dynamicSmemBytes = smem; cfg.stream = h->stream; cfg.attrs = at; cfg.numAttrs = 1;
    cfg.gridDim = dim3(cl * (h->num_sms / cl));
    int ncl = 0;
    CK(cudaOccupancyMaxActiveClusters(&ncl, kern, &cfg));
    if (ncl < 1) return -1;
    ncl = std::min(ncl, h->num_sms / cl);
    if (getenv("KDSL_DEBUG_OCC")) fprintf(stderr, "k_inverse_cl: cluster size %d, %d active clusters, smem %zu\n", cl, ncl, smem);
    if (!h->cl_scratch || h->cl_scratch_clusters < ncl) {
        int rc = dev_alloc(h, &h->cl_scratch, (size_t)ncl * inverse_cl_scratch_doubles(NB, NpMax));
        if (rc) return rc;
        h->cl_scratch_clusters = ncl;
    }
    cfg.gridDim = dim3(cl * ncl);
    const int cs = NpMax, rs = std::max(1, h->inverse_rs);
    CK(cudaLaunchKernelEx(&cfg, kern, h->S, list, h->A_up, h->A_dn, h->status, h->colsrc, h->Np_up, h->Np_dn, cs, h->cl_scratch, rs));
    CK(cudaGetLastError());
    return KDSL_OK;
}

// ComplexF64 engine: W = U X on the unoccupied rows (FP64 FMA tiles; gemm_variant picks the tile shape)
int launch_gemm_W_c(kdsl_handle h, const int *list) {
    const DevState &S = h->S;
    const int Nmax = std::max(S.n_up, S.n_dn);
#define KDSL_GEMM_C(BM, BN, BK)                                                                                     \
    do {                                                                                                            \
        const int tiles = ((S.ns + BM - 1) / BM) * ((Nmax + BN - 1) / BN);                                          \
        k_gemm_W_c<BM, BN, BK><<<dim3(tiles, S.nw, 2), 256, 0, h->stream>>>(S, list, h->X_up, h->X_dn, h->status, h->urow, S.ns); \
    } while (0)
    if (h->gemm_variant == 4) KDSL_GEMM_C(64, 64, 8);
    else KDSL_GEMM_C(64, 32, 8);                          // (64 x 64 and deeper k tiles measured 5 % slower, 128 x 64 50 % slower)
#undef KDSL_GEMM_C
    CK(cudaGetLastError());
    return KDSL_OK;
}

// ComplexF64 engine: both species in one launch of the complex cluster inverse (kdsl_inverse_cl_c.cuh) on the split planes
// that k_gather_tilde_split_c left in A_up / A_dn; returns -1 when it does not apply.
bool inverse_clc_applies(kdsl_handle h) {
    if (!h->cplx || !(h->inverse_variant == 0 || h->inverse_variant == 9)) return false;
    const int cl = h->inverse_cluster;
    if (cl < 2 || cl > 8) return false;
    const int NpMax = std::max(h->Npc_up, h->Npc_dn);
    if (NpMax > 512 || std::min(h->Npc_up, h->Npc_dn) < 8) return false;
    // the split planes live in the embedding's workspace: 2 Npc^2 <= Np^2 doubles per entry
    if ((size_t)2 * h->Npc_up * h->Npc_up > (size_t)h->Np_up * h->Np_up || (size_t)2 * h->Npc_dn * h->Npc_dn > (size_t)h->Np_dn * h->Np_dn) return false;
    return inverse_clc_smem(8, NpMax) <= (size_t)227 * 1024;     // (panels of 16 columns are used when two CTAs per SM fit)
}
template <int NB, int CT, int T, int MINB>
int launch_inverse_clc_t(kdsl_handle h, const int *list) {
    const int NpMax = std::max(h->Npc_up, h->Npc_dn);
    if (NpMax > T) return -1;                             // thread = matrix row in the pivot CTA
    const int cl = h->inverse_cluster;
    const size_t smem = inverse_clc_smem(NB, NpMax);
    auto kern = k_inverse_cl_c<NB, CT, T, MINB>;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = cl; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.blockDim = dim3(T); cfg.dynamicSmemBytes = smem; cfg.stream = h->stream; cfg.attrs = at; cfg.numAttrs = 1;
    cfg.gridDim = dim3(cl * (MINB * h->num_sms / cl));
    int ncl = 0;
    CK(cudaOccupancyMaxActiveClusters(&ncl, kern, &cfg));
    if (ncl < 1) return -1;
    ncl = std::min(ncl, MINB * h->num_sms / cl);
    if (getenv("KDSL_DEBUG_OCC")) fprintf(stderr, "k_inverse_cl_c<%d,%d,%d,%d>: cluster size %d, %d active clusters, smem %zu\n", NB, CT, T, MINB, cl, ncl, smem);
    if (!h->clc_scratch || h->clc_scratch_clusters < ncl) {
        int rc = dev_alloc(h, &h->clc_scratch, (size_t)ncl * inverse_clc_scratch_doubles(16, NpMax));
        if (rc) return rc;
        h->clc_scratch_clusters = ncl;
    }
    cfg.gridDim = dim3(cl * ncl);
    const int cs = std::max(h->Np_up, h->Np_dn), rs = std::max(1, h->inverse_rs);
    const size_t su = (size_t)h->Np_up * h->Np_up, sd = (size_t)h->Np_dn * h->Np_dn;
    CK(cudaLaunchKernelEx(&cfg, kern, h->S, list, h->A_up, h->A_dn, h->status, h->colsrc, h->Npc_up, h->Npc_dn, su, sd, cs, h->clc_scratch, rs));
    CK(cudaGetLastError());
    return KDSL_OK;
}
int launch_inverse_clc(kdsl_handle h, const int *list) {
    const int NpMax = std::max(h->Npc_up, h->Npc_dn);
    const int v = h->inverse_tuning & 15;
    // two CTAs per SM double the matrices in flight (26.6 ms against 35.1 per 4096-walker bin at 432 sites); panels of 16
    // complex columns halve the block steps (21.8 ms)
    // (measured slower: three 224-thread CTAs per SM with panels of 8: 31.8 ms, 224-thread CTAs with panels of 16: 23.9 ms)
    if (NpMax <= 256 && v == 2) return launch_inverse_clc_t<8, 2, 256, 2>(h, list);
    if (NpMax <= 256 && v != 1 && 2 * (inverse_clc_smem(16, NpMax) + 1024) <= (size_t)227 * 1024) return launch_inverse_clc_t<16, 2, 256, 2>(h, list);
    if (NpMax <= 256 && v != 1) return launch_inverse_clc_t<8, 2, 256, 2>(h, list);
    return launch_inverse_clc_t<8, 2, 512, 1>(h, list);
}

#ifdef KDSL_DEV_VARIANTS
// reevaluateW! as ONE cluster kernel (kdsl_reeval_cl.cuh); returns -1 when it does not apply.
template <int T, int MINB>
int launch_reeval_cl_t(kdsl_handle h, const int *list) {
    const int cl = h->reeval_cluster;
    constexpr int NB = 24, DG = 4, NWARPS = T / 32;
    const int NG = cl - 1;
    if (h->rcl_NpMax > 4 * NWARPS * 8 || h->rcl_NpMax > T) return -1;   // row tiles of the next-panel update; thread = row
    // X slots (column tiles) a CTA needs: its pairs of the first block step, its V tiles of the last step, one panel
    const int pairs0 = (h->rcl_ntcMax + 1) / 2;
    const int XT = std::max(std::max(2 * ((pairs0 + NG - 1) / NG), (h->rcl_nvtMax + cl - 1) / cl + 1), NB / 8);
    const int nv = (h->rcl_nvtMax + cl - 1) / cl + 1;
    if (nv > 2 * NWARPS) return -1;                       // two column tiles per warp in the last step
    const int SW = 8 * nv;
    const size_t smem = reeval_cl_smem(NB, h->rcl_NpMax, h->rcl_CpMax, h->S.ns, XT, SW);
    if (MINB * (smem + 1024) > (size_t)227 * 1024) return -1;
    auto kern = k_reeval_cl<NB, DG, T, MINB>;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = cl; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.blockDim = dim3(T); cfg.dynamicSmemBytes = smem; cfg.stream = h->stream; cfg.attrs = at; cfg.numAttrs = 1;
    cfg.gridDim = dim3(cl * (MINB * h->num_sms / cl));
    int ncl = 0;
    CK(cudaOccupancyMaxActiveClusters(&ncl, kern, &cfg));
    if (ncl < 1) return -1;
    ncl = std::min(ncl, MINB * h->num_sms / cl);
    if (getenv("KDSL_DEBUG_OCC")) fprintf(stderr, "k_reeval_cl<%d,%d>: cluster size %d, %d active clusters, smem %zu, XT %d\n", T, MINB, cl, ncl, smem, XT);
    if (!h->rcl_scratch || h->rcl_scratch_clusters < ncl) {
        int rc = dev_alloc(h, &h->rcl_scratch, (size_t)ncl * reeval_cl_scratch_doubles(NB, h->rcl_NpMax, h->rcl_CpMax));
        if (rc) return rc;
        h->rcl_scratch_clusters = ncl;
    }
    cfg.gridDim = dim3(cl * ncl);
    const int rs = std::max(1, h->reeval_rs);
    CK(cudaLaunchKernelEx(&cfg, kern, h->S, list, h->rcl_scratch, (const double *)h->UT_up, (const double *)h->UT_dn, h->status,
                          h->Np_up, h->Np_dn, h->rcl_NpMax, h->rcl_CpMax, XT, SW, rs));
    CK(cudaGetLastError());
    return KDSL_OK;
}
int launch_reeval_cl(kdsl_handle h, const int *list) {
    if (!h->rcl_ok || h->cplx) return -1;
    if (h->reeval_cluster < 2 || h->reeval_cluster > 8) return -1;
    if (h->rcl_NpMax <= 256 && (h->inverse_tuning & 15) != 1) {          // two 256-thread CTAs per SM: twice the matrices in flight
        const int rc = launch_reeval_cl_t<256, 2>(h, list);
        if (rc >= 0) return rc;
    }
    return launch_reeval_cl_t<512, 1>(h, list);
}

#endif  // KDSL_DEV_VARIANTS

int launch_inverse(kdsl_handle h, const int *list, double *A, int spin, int Np) {
    if ((h->inverse_variant == 0 || h->inverse_variant == 5) && Np <= 256) {
        // look-ahead version: pivot loop of panel s+1 concurrent with the DMMA update of step s
        if ((h->inverse_tuning & 15) == 1) return launch_inverse_v5<24, 3>(h, list, A, spin, Np);
        if ((h->inverse_tuning & 15) == 2) return launch_inverse_v5<32, 2>(h, list, A, spin, Np);
        if ((h->inverse_tuning & 15) == 3) return launch_inverse_v5<16, 2>(h, list, A, spin, Np);
        return launch_inverse_v5<24, 2>(h, list, A, spin, Np);
    }
    if (h->inverse_variant == 0 || h->inverse_variant >= 4) {
        // implicit-pivoting blocked Gauss-Jordan, one CTA per matrix and ONE CTA per SM (matrices stay L2 resident)
        if (Np <= 256) {
            if (h->inverse_tuning == 1) return launch_inverse_v4<32, 1, 256, 256>(h, list, A, spin, Np);
            if (h->inverse_tuning == 2) return launch_inverse_v4<24, 1, 256, 256, 2, 2>(h, list, A, spin, Np);
            if (h->inverse_tuning == 3) return launch_inverse_v4<24, 1, 256, 256, 2, 3>(h, list, A, spin, Np);
            if (h->inverse_tuning == 4) return launch_inverse_v4<16, 1, 256, 256, 2, 3>(h, list, A, spin, Np);
            if (h->inverse_tuning == 5) return launch_inverse_v4<24, 1, 256, 256, 1, 2>(h, list, A, spin, Np);
            return launch_inverse_v4<24, 1, 256, 256>(h, list, A, spin, Np);
        }
        if (Np <= 512) return launch_inverse_v4<24, 2, 256, 256>(h, list, A, spin, Np);
        if (Np <= 1024) return launch_inverse_v4<8, 4, 256, 256>(h, list, A, spin, Np);
        return fail(KDSL_ERR_INVALID_ARGUMENT, "N = %d exceeds the supported maximum of 1024 orbitals per species", Np);
    }
#ifdef KDSL_DEV_VARIANTS
    if (h->inverse_variant == 3) {
        // one CTA per matrix and ONE CTA per SM: 148 x N^2 x 8 B of live matrices stay L2 resident
        if (Np <= 256) return h->inverse_tuning == 1 ? launch_inverse_v3<24, 1, 512>(h, list, A, spin, Np)
                                                     : launch_inverse_v3<32, 1, 512>(h, list, A, spin, Np);
        if (Np <= 512) return launch_inverse_v3<16, 1, 512>(h, list, A, spin, Np);
        if (Np <= 1024) return launch_inverse_v3<8, 2, 512>(h, list, A, spin, Np);
        return fail(KDSL_ERR_INVALID_ARGUMENT, "N = %d exceeds the supported maximum of 1024 orbitals per species", Np);
    }
    const int v = h->inverse_tuning;
    if (Np <= 256) {
        if (v == 1) return launch_inverse_blocked<24, 1, 256, 2>(h, list, A, spin, Np);
        if (v == 2) return launch_inverse_blocked<16, 2, 128, 4>(h, list, A, spin, Np);
        if (v == 3) return launch_inverse_blocked<8, 2, 128, 4>(h, list, A, spin, Np);
        return launch_inverse_blocked<16, 1, 256, 2>(h, list, A, spin, Np);
    }
    if (Np <= 512) return launch_inverse_blocked<16, 2, 256, 1>(h, list, A, spin, Np);
    if (Np <= 1024) return launch_inverse_blocked<8, 4, 256, 1>(h, list, A, spin, Np);
    return fail(KDSL_ERR_INVALID_ARGUMENT, "N = %d exceeds the supported maximum of 1024 orbitals per species", Np);
#else
    return fail(KDSL_ERR_INVALID_ARGUMENT, "inverse_variant %d is a developer variant (build with make DEV=1)", h->inverse_variant);
#endif
}

// reevaluateW! for the walkers in `list` (device list with device count cnt[2]) or all (list = null)
int launch_refresh(kdsl_handle h, const int *list) {
    const DevState &S = h->S;
    const int Nmax = std::max(S.n_up, S.n_dn);
    if (inverse_clc_applies(h)) {
        // ComplexF64: complex blocked Gauss-Jordan by clusters on split (re, im) planes, then W = U X on the unoccupied rows
        const int cs = std::max(h->Np_up, h->Np_dn);
        const size_t su = (size_t)h->Np_up * h->Np_up, sd = (size_t)h->Np_dn * h->Np_dn;
        int rc;
        {
            Span sp(h, KDSL_T_REFRESH_GATHER);
            k_gather_tilde_split_c<<<dim3(S.nw, 2), 256, Nmax * sizeof(int), h->stream>>>(S, list, h->A_up, h->A_dn, h->status, h->Npc_up, h->Npc_dn, su, sd);
            CK(cudaGetLastError());
        }
        {
            Span sp(h, KDSL_T_REFRESH_INVERSE);
            rc = launch_inverse_clc(h, list);
            if (rc > 0) return rc;
            if (rc == 0) {
                k_refresh_status<<<(S.nw + 255) / 256, 256, 0, h->stream>>>(S, list, h->status);
                CK(cudaGetLastError());
                h->t_launch[KDSL_T_REFRESH_INVERSE] += 1;
            }
        }
        if (rc == 0) {
            Span sp(h, KDSL_T_REFRESH_GEMM);
            const bool dm = h->gemm_variant != 4 && Nmax <= 512;      // tensor-pipe product straight from the split planes
            k_unsplit_c<<<dim3(S.nw, 2), 256, std::max(h->Npc_up, h->Npc_dn) * sizeof(int), h->stream>>>(S, list, h->A_up, h->A_dn, h->X_up, h->X_dn, h->status, h->colsrc, h->Npc_up, h->Npc_dn, su, sd, cs, h->urow, S.ns, dm ? 0 : 1);
            CK(cudaGetLastError());
            if (dm) {
                const int Mmax = S.ns - std::min(S.n_up, S.n_dn);
                const int tiles = ((Mmax + 71) / 72) * ((Nmax + 23) / 24);
                k_gemm_W_dmma_c<8><<<dim3(tiles, S.nw, 2), 96, 0, h->stream>>>(S, list, h->A_up, h->A_dn, h->status, h->colsrc, h->Npc_up, h->Npc_dn, su, sd, cs, h->urow, S.ns);
                CK(cudaGetLastError());
            } else {
                int rg = launch_gemm_W_c(h, list);
                if (rg) return rg;
            }
            h->t_launch[KDSL_T_REFRESH_GEMM] += 1;
            return KDSL_OK;
        }
        // (no cluster launch possible: fall through to the embedding, which rebuilds the workspace)
    }
    if (h->cplx && h->inverse_variant != 1) {
        // ComplexF64: inverse of tilde_U through its real 2N x 2N embedding on the production real kernels
        if (std::max(h->Np_up, h->Np_dn) > 1024)
            return fail(KDSL_ERR_INVALID_ARGUMENT, "ComplexF64 mode supports at most 512 orbitals per species");
        const int cs = std::max(h->Np_up, h->Np_dn);
        {
            Span sp(h, KDSL_T_REFRESH_GATHER);
            k_gather_tilde_emb_c<<<dim3(S.nw, 2), 256, Nmax * sizeof(int), h->stream>>>(S, list, h->A_up, h->A_dn, h->status, h->Np_up, h->Np_dn);
            CK(cudaGetLastError());
        }
        {
            Span sp(h, KDSL_T_REFRESH_INVERSE);
            int rc = launch_inverse_cl(h, list);
            if (rc < 0) {
                rc = launch_inverse(h, list, h->A_up, 0, h->Np_up);
                if (rc) return rc;
                rc = launch_inverse(h, list, h->A_dn, 1, h->Np_dn);
            }
            if (rc) return rc;
            k_refresh_status<<<(S.nw + 255) / 256, 256, 0, h->stream>>>(S, list, h->status);
            CK(cudaGetLastError());
            h->t_launch[KDSL_T_REFRESH_INVERSE] += 2;
        }
        {
            Span sp(h, KDSL_T_REFRESH_GEMM);
            k_unembed_c<<<dim3(S.nw, 2), 256, cs * sizeof(int), h->stream>>>(S, list, h->A_up, h->A_dn, h->X_up, h->X_dn, h->status, h->colsrc, h->Np_up, h->Np_dn, cs, h->urow, S.ns);
            CK(cudaGetLastError());
            int rg = launch_gemm_W_c(h, list);
            if (rg) return rg;
            h->t_launch[KDSL_T_REFRESH_GEMM] += 1;
        }
        return KDSL_OK;
    }
    if (h->cplx) {
        {
            Span sp(h, KDSL_T_REFRESH_GATHER);
            k_gather_tilde_c<<<dim3(S.nw, 2), 256, Nmax * sizeof(int), h->stream>>>(S, list, h->A_up, h->A_dn, h->status);
            CK(cudaGetLastError());
        }
        {
            Span sp(h, KDSL_T_REFRESH_INVERSE);
            const size_t smem = (size_t)Nmax * (4 * sizeof(double) + sizeof(int));
            CK(cudaFuncSetAttribute(k_inverse_gj_c, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(smem, 1024)));
            k_inverse_gj_c<<<S.nw, 256, smem, h->stream>>>(S, list, h->A_up, 0, h->status);
            CK(cudaGetLastError());
            k_inverse_gj_c<<<S.nw, 256, smem, h->stream>>>(S, list, h->A_dn, 1, h->status);
            CK(cudaGetLastError());
            k_refresh_status<<<(S.nw + 255) / 256, 256, 0, h->stream>>>(S, list, h->status);
            CK(cudaGetLastError());
            h->t_launch[KDSL_T_REFRESH_INVERSE] += 2;
        }
        {
            Span sp(h, KDSL_T_REFRESH_GEMM);
            constexpr int BM = 64, BN = 32;
            const int tiles = ((S.ns + BM - 1) / BM) * ((Nmax + BN - 1) / BN);
            k_gemm_W_c<BM, BN, 8><<<dim3(tiles, S.nw, 2), 256, 0, h->stream>>>(S, list, h->A_up, h->A_dn, h->status);
            CK(cudaGetLastError());
        }
        return KDSL_OK;
    }
    if ((h->inverse_variant == 0 || h->inverse_variant == 6) && h->fused_smem > 0) {
        // one kernel: Gauss-Jordan on [tilde_U^T | V^T] carries V^T to the non-trivial rows of W (kdsl_reeval_fused.cuh)
        Span sp(h, KDSL_T_REFRESH_INVERSE);
        const int v = h->inverse_tuning & 15;
        const int fgrid = h->fused_ctas > 0 ? std::min(h->fused_ctas, h->num_sms) : h->num_sms;
        const bool nb32 = h->fused_NB == 32 && v == 6;   // wider panels: fewer passes, but a longer pivot chain (measured slower)
        const size_t fsm = nb32 ? h->fused_smem : v == 5 ? h->fused_smem16 : h->fused_smem24;
        const int fst = nb32 ? h->fused_stage : v == 5 ? h->fused_stage16 : h->fused_stage24;
#define KDSL_FUSED_LAUNCH(...) k_reeval_fused<__VA_ARGS__><<<fgrid, 512, fsm, h->stream>>>(S, list, h->ws_fused, h->ws_stride, h->UT_up, h->UT_dn, h->status, h->Np_up, h->Np_dn, h->fused_NpMax, h->fused_CpMax, fst)
        if (h->fused_small && v == 0) {
            // small lattices: 256-thread CTAs, two per SM (twice the pivot chains in flight)
            const int sgrid = h->fused_ctas > 0 ? std::min(h->fused_ctas, 2 * h->num_sms) : 2 * h->num_sms;
            k_reeval_fused<24, 2, 6, 0, 256><<<sgrid, 256, fsm, h->stream>>>(S, list, h->ws_fused, h->ws_stride, h->UT_up, h->UT_dn, h->status, h->Np_up, h->Np_dn, h->fused_NpMax, h->fused_CpMax, fst);
        } else if (nb32) {
            KDSL_FUSED_LAUNCH(32, 2, 6);
        } else if (v == 1) KDSL_FUSED_LAUNCH(24, 1, 8);
        else if (v == 2) KDSL_FUSED_LAUNCH(24, 2, 4);
        else if (v == 5) KDSL_FUSED_LAUNCH(16, 2, 6);
        else if (v == 8) KDSL_FUSED_LAUNCH(24, 2, 6, 1);
        else if (v == 9) KDSL_FUSED_LAUNCH(24, 2, 6, 2);
        else KDSL_FUSED_LAUNCH(24, 2, 6);
#undef KDSL_FUSED_LAUNCH
        CK(cudaGetLastError());
        k_refresh_status_fused<<<(S.nw + 255) / 256, 256, 0, h->stream>>>(S, list, h->status);
        CK(cudaGetLastError());
        h->t_launch[KDSL_T_REFRESH_INVERSE] += 1;
        return KDSL_OK;
    }
#ifdef KDSL_DEV_VARIANTS
    if (h->inverse_variant == 8 && h->rcl_ok) {
        // one cluster kernel for 256 < Np <= 512 (kdsl_reeval_cl.cuh)
        Span sp(h, KDSL_T_REFRESH_INVERSE);
        int rc = launch_reeval_cl(h, list);
        if (rc > 0) return rc;
        if (rc < 0) return fail(KDSL_ERR_INVALID_ARGUMENT, "inverse_variant 8 (cluster re-evaluation) does not fit this problem / reeval_cluster");
        k_refresh_status_fused<<<(S.nw + 255) / 256, 256, 0, h->stream>>>(S, list, h->status);
        CK(cudaGetLastError());
        h->t_launch[KDSL_T_REFRESH_INVERSE] += 1;
        return KDSL_OK;
    }
#endif
    const bool fast = h->inverse_variant != 1;
    {
        Span sp(h, KDSL_T_REFRESH_GATHER);
        if (fast)
            k_gather_tilde_padded<<<dim3(S.nw, 2), 256, Nmax * sizeof(int), h->stream>>>(S, list, h->A_up, h->A_dn, h->status, h->Np_up, h->Np_dn, h->urow, S.ns);
        else
            k_gather_tilde<<<dim3(S.nw, 2), 256, Nmax * sizeof(int), h->stream>>>(S, list, h->A_up, h->A_dn, h->status);
        CK(cudaGetLastError());
    }
    {
        Span sp(h, KDSL_T_REFRESH_INVERSE);
        if (fast) {
            int rc = launch_inverse_cl(h, list);
            if (rc < 0) {
                rc = launch_inverse(h, list, h->A_up, 0, h->Np_up);
                if (rc) return rc;
                rc = launch_inverse(h, list, h->A_dn, 1, h->Np_dn);
            }
            if (rc) return rc;
        } else {
            const size_t smem = (size_t)Nmax * (2 * sizeof(double) + sizeof(int));
            k_inverse_gj<<<S.nw, 256, smem, h->stream>>>(S, list, h->A_up, 0, h->status);
            CK(cudaGetLastError());
            k_inverse_gj<<<S.nw, 256, smem, h->stream>>>(S, list, h->A_dn, 1, h->status);
            CK(cudaGetLastError());
        }
        k_refresh_status<<<(S.nw + 255) / 256, 256, 0, h->stream>>>(S, list, h->status);
        CK(cudaGetLastError());
        h->t_launch[KDSL_T_REFRESH_INVERSE] += 2;          // (the span counted one of the three launches)
    }
    {
        Span sp(h, KDSL_T_REFRESH_GEMM);
        if (fast) {
            constexpr int KT = 24;
            const int Mmax = S.ns - std::min(S.n_up, S.n_dn);
            const int tiles = ((Mmax + 71) / 72) * ((Nmax + 71) / 72);
            const size_t smem = (size_t)4 * 72 * KT * sizeof(double);
            CK(cudaFuncSetAttribute(k_gemm_W_dmma<KT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            CK(cudaFuncSetAttribute(k_gemm_W_dmma<KT>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
            if (getenv("KDSL_DEBUG_OCC")) {
                int nb = 0;
                cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_gemm_W_dmma<KT>, 288, smem);
                fprintf(stderr, "k_gemm_W_dmma: %d CTAs/SM (dynamic smem %zu)\n", nb, smem);
            }
            const int perm_k = (h->inverse_variant == 0 || h->inverse_variant >= 4) ? 1 : 0;
#ifdef KDSL_DEV_VARIANTS
            const int cs = std::max(h->Np_up, h->Np_dn);
            if (h->gemm_variant == 2 || h->gemm_variant == 3) {   // cp.async pipeline (measured slightly slower than the register-staged kernel)
                constexpr int ST = 3;
                const size_t sm3 = (size_t)ST * 2 * 72 * KT * sizeof(double);
                if (h->gemm_variant == 3) {
                    CK(cudaFuncSetAttribute(k_gemm_W_cpasync<KT, ST, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm3));
                    k_gemm_W_cpasync<KT, ST, 3><<<dim3(tiles, S.nw, 2), 288, sm3, h->stream>>>(S, list, h->A_up, h->A_dn, h->status, h->colsrc, h->Np_up, h->Np_dn, cs, h->urow, S.ns, perm_k);
                } else {
                    CK(cudaFuncSetAttribute(k_gemm_W_cpasync<KT, ST, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm3));
                    k_gemm_W_cpasync<KT, ST, 2><<<dim3(tiles, S.nw, 2), 288, sm3, h->stream>>>(S, list, h->A_up, h->A_dn, h->status, h->colsrc, h->Np_up, h->Np_dn, cs, h->urow, S.ns, perm_k);
                }
            } else
#endif
            k_gemm_W_dmma<KT><<<dim3(tiles, S.nw, 2), 288, smem, h->stream>>>(S, list, h->A_up, h->A_dn, h->status, h->colsrc, h->Np_up, h->Np_dn, std::max(h->Np_up, h->Np_dn), h->urow, S.ns, perm_k);
        } else {
            constexpr int BM = 64, BN = 64;
            const int tiles = ((S.ns + BM - 1) / BM) * ((Nmax + BN - 1) / BN);
            k_gemm_W_simt<BM, BN, 16><<<dim3(tiles, S.nw, 2), 256, 0, h->stream>>>(S, list, h->A_up, h->A_dn, h->status);
        }
        CK(cudaGetLastError());
    }
    return KDSL_OK;
}

template <int KPAD>
int launch_flush_wb_kernel(kdsl_handle h, const int *list, int *cptr) {
    const DevState &S = h->S;
    const int Nmax = std::max(S.n_up, S.n_dn);
    const size_t smem = (size_t)((Nmax + 7) / 8) * 8 * KPAD * sizeof(double);
    {
        Span sp(h, KDSL_T_UPDATE);
        if (h->flush_variant == 0) {                      // register-pipelined kernel, G built per 216-row item (fastest measured)
            k_flush_wb<KPAD><<<h->num_sms * 2, 288, smem, h->stream>>>(S, list, cptr, S.nw, S.cnt + 5);
        } else {                                          // 3: bulk-async shared-memory ring (k_flush_G + k_flush_tma), see kdsl_flush.cuh
            k_flush_G<KPAD><<<h->num_sms * 2, 288, smem, h->stream>>>(S, list, cptr, S.nw, h->Gbuf, h->Gstride);
            CK(cudaGetLastError());
            const size_t ring3 = ((smem + 127) & ~(size_t)127) + (size_t)4 * 8 * 1728;
            const int per_sm = 2 * (ring3 + 2048) <= (size_t)227 * 1024 ? 2 : 1;
            k_flush_tma<KPAD, 4, 3><<<h->num_sms * per_sm, 320, ring3, h->stream>>>(S, list, cptr, S.nw, S.cnt + 5, h->Gbuf, h->Gstride, h->flush_dbg);
            h->t_launch[KDSL_T_UPDATE] += 1;
        }
        CK(cudaGetLastError());
    }
    k_flush_finish_wb<<<1, 1024, 0, h->stream>>>(S, list, cptr, S.nw, S.cnt + 5);
    CK(cudaGetLastError());
    h->t_launch[KDSL_T_UPDATE] += 1;
    return KDSL_OK;
}

// W0 += pending factors for the listed walkers (list = device list with count cnt[4]) or for all walkers
int launch_flush(kdsl_handle h, bool all) {
    const DevState &S = h->S;
    const int *list = all ? nullptr : S.flush_list;
    int *cptr = all ? nullptr : S.cnt + 4;
    if (h->cplx) {
        // ComplexF64: FMA-pipe flush, one thread per row (kdsl_woodbury_c.cuh); kmax <= 24 (checked in kdsl_set_option)
        const int Np = std::max(S.n_up, S.n_dn);
        const size_t smem = (size_t)24 * Np * 2 * sizeof(double);
        const int per_sm = 2 * (smem + 12 * 1024) <= (size_t)227 * 1024 ? 2 : 1;
        {
            Span sp(h, KDSL_T_UPDATE);
            if (h->flush_variant == 0 && h->fd2_RB > 0) {     // tensor-pipe flush, two persistent 256-thread CTAs per SM
                k_flush_dmma2_c<24, 4><<<2 * h->num_sms, 256, flush_dmma2_c_smem(24, h->fd2_RB), h->stream>>>(S, list, cptr, S.nw, S.cnt + 5, h->fd2_RB);
            } else if ((h->flush_variant == 0 || h->flush_variant == 7) && h->fdc_RB > 0) {     // k_flush_dmma_c, one persistent CTA per SM
                const int Npad = (Np + 7) / 8 * 8;
                k_flush_dmma_c<24, 4><<<h->num_sms, 512, flush_dmma_c_smem(24, Npad, h->fdc_RB), h->stream>>>(S, list, cptr, S.nw, S.cnt + 5, Npad, h->fdc_RB);
            } else {
                k_flush_c<24, 216><<<h->num_sms * per_sm, 224, smem, h->stream>>>(S, list, cptr, S.nw, S.cnt + 5, Np);
            }
            CK(cudaGetLastError());
        }
        k_flush_finish_wb<<<1, 1024, 0, h->stream>>>(S, list, cptr, S.nw, S.cnt + 5);
        CK(cudaGetLastError());
        h->t_launch[KDSL_T_UPDATE] += 1;
        h->since_flush = 0;
        return KDSL_OK;
    }
    if (h->update_variant == 2 && (h->flush_variant == 0 || h->flush_variant == 3)) {
        int rc = S.kmax <= 20 ? launch_flush_wb_kernel<20>(h, list, cptr)
               : S.kmax <= 24 ? launch_flush_wb_kernel<24>(h, list, cptr)
                              : launch_flush_wb_kernel<32>(h, list, cptr);
        if (rc) return rc;
        h->since_flush = 0;
        return KDSL_OK;
    }
#ifdef KDSL_DEV_VARIANTS
    const int Nmax = std::max(S.n_up, S.n_dn);
    const size_t smem = (size_t)((Nmax + 7) / 8) * 8 * KDSL_KMAX * sizeof(double);
    CK(cudaFuncSetAttribute(k_flush<KDSL_KMAX, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CK(cudaFuncSetAttribute(k_flush<KDSL_KMAX, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    {
        Span sp(h, KDSL_T_UPDATE);
        if (h->update_variant == 2) {
            if (S.kmax != KDSL_KMAX) return fail(KDSL_ERR_STATE, "flush_variant 1 needs kmax = %d", KDSL_KMAX);
            const size_t psm = (size_t)S.kmax * S.kmax * sizeof(double) + 2 * S.kmax * sizeof(int);
            k_flush_prepare<<<h->num_sms * 4, 256, psm, h->stream>>>(S, list, cptr, S.nw);
            CK(cudaGetLastError());
            k_flush<KDSL_KMAX, true><<<h->num_sms * 4, 288, smem, h->stream>>>(S, list, cptr, S.nw);
        } else {
            k_flush<KDSL_KMAX, false><<<h->num_sms * 4, 288, smem, h->stream>>>(S, list, cptr, S.nw);
        }
        CK(cudaGetLastError());
        k_flush_done<<<8, 256, 0, h->stream>>>(S, list, cptr, S.nw);
        CK(cudaGetLastError());
        k_zero_int<<<1, 1, 0, h->stream>>>(S.cnt + 4);
        CK(cudaGetLastError());
    }
#endif
    h->since_flush = 0;
    return KDSL_OK;
}

int ensure_replay_capacity(kdsl_handle h, size_t n) {
    if (n <= h->rp_cap) return KDSL_OK;
    if (h->rp_r) { cudaFree(h->rp_r); cudaFree(h->rp_bond); cudaFree(h->rp_pick); }
    h->rp_r = nullptr; h->rp_bond = nullptr; h->rp_pick = nullptr; h->rp_cap = 0;
    CK(cudaMalloc(&h->rp_r, n * sizeof(double)));
    CK(cudaMalloc(&h->rp_bond, n * sizeof(int)));
    CK(cudaMalloc(&h->rp_pick, n * sizeof(int)));
    h->rp_cap = n;
    return KDSL_OK;
}

// the lock-step Carlo loop: n x { sweep!; ctx.sweeps += 1; [measure!] }
// In Woodbury mode consecutive sweeps without a gate (refresh), flush or measurement in between are fused
// into one launch of k_decide_wb (the walkers are independent; only those events need all walkers in step).
int run_sweeps(kdsl_handle h, int64_t n, int64_t therm, bool replay, bool have_pick) {
    const DevState &S = h->S;
    const int64_t period = h->refresh_every > 0 ? h->refresh_every : S.n_occ;
    const int pgrid = grid_for_warps(S.nw);
    const bool delayed = h->update_variant == 1 || h->update_variant == 2;
    const bool woodbury = h->update_variant == 2;
    if (h->update_variant == 3) {
        // small lattices: the whole call in ONE persistent kernel, every walker resident in shared memory (kdsl_resident.cuh)
        for (int64_t s = 0; s < n;) {
            const int64_t g = std::min<int64_t>(n - s, 1 << 30);
            ResParams P;
            P.n_sweeps = (int)g; P.sweep0 = h->sweeps; P.period = period; P.therm = therm;
            P.phase_p = h->sweeps % period; P.phase_m = h->sweeps % S.n_occ;
            P.rp_r = replay ? h->rp_r + (size_t)s * S.nw : nullptr;
            P.rp_bond = replay ? h->rp_bond + (size_t)s * S.nw : nullptr;
            P.rp_pick = (replay && have_pick) ? h->rp_pick + (size_t)s * S.nw : nullptr;
            P.work_counter = S.cnt + 6;
            CK(cudaMemsetAsync(S.cnt + 6, 0, sizeof(int), h->stream));
            const int grid = std::min(S.nw, h->num_sms * h->res_ctas_per_sm);
            {
                Span sp(h, KDSL_T_PROPOSE);
#define KDSL_RES_LAUNCH(NP_)                                                                         \
    do {                                                                                             \
        if (replay) k_resident<true, NP_><<<grid, KDSL_RES_THREADS, h->res_smem, h->stream>>>(S, P); \
        else k_resident<false, NP_><<<grid, KDSL_RES_THREADS, h->res_smem, h->stream>>>(S, P);       \
    } while (0)
                if (h->res_np <= 1) KDSL_RES_LAUNCH(1);
                else if (h->res_np == 2) KDSL_RES_LAUNCH(2);
                else if (h->res_np == 3) KDSL_RES_LAUNCH(3);
                else KDSL_RES_LAUNCH(4);
#undef KDSL_RES_LAUNCH
                CK(cudaGetLastError());
            }
            h->sweeps += g;
            h->walker_sweeps += g * S.nw;
            s += g;
        }
        return KDSL_OK;
    }
    for (int64_t s = 0; s < n;) {
        const bool gate = (h->sweeps % period) == 0;            // src/MonteCarlo.jl:595 (pre-increment)
        int64_t g = 1;
        if (woodbury && !gate && h->fuse_sweeps) {
            g = std::min<int64_t>(n - s, h->flush_every - h->since_flush);      // up to the next flush
            g = std::min<int64_t>(g, period - (h->sweeps % period));               // ... the next gate sweep
            g = std::min<int64_t>(g, S.n_occ - (h->sweeps % S.n_occ));             // ... the next measurement
            if (g < 1) g = 1;
        }
        {
            Span sp(h, KDSL_T_PROPOSE);
            const size_t off = (size_t)s * S.nw;
            if (woodbury && h->cplx) {
                if (replay)
                    k_decide_wb_c<true><<<pgrid, 256, 0, h->stream>>>(S, gate ? 1 : 0, (int)g, h->rp_r + off, h->rp_bond + off,
                                                                      have_pick ? h->rp_pick + off : nullptr);
                else
                    k_decide_wb_c<false><<<pgrid, 256, 0, h->stream>>>(S, gate ? 1 : 0, (int)g, nullptr, nullptr, nullptr);
                CK(cudaGetLastError());
            } else if (woodbury) {
                if (replay)
                    k_decide_wb<true><<<pgrid, 256, 0, h->stream>>>(S, gate ? 1 : 0, (int)g, h->rp_r + off, h->rp_bond + off,
                                                                    have_pick ? h->rp_pick + off : nullptr);
                else
                    k_decide_wb<false><<<pgrid, 256, 0, h->stream>>>(S, gate ? 1 : 0, (int)g, nullptr, nullptr, nullptr);
                CK(cudaGetLastError());
#ifdef KDSL_DEV_VARIANTS
            } else if (delayed) {
                if (replay)
                    k_decide<true><<<pgrid, 256, 0, h->stream>>>(S, h->parity, gate ? 1 : 0, h->rp_r + off, h->rp_bond + off,
                                                                 have_pick ? h->rp_pick + off : nullptr);
                else
                    k_decide<false><<<pgrid, 256, 0, h->stream>>>(S, h->parity, gate ? 1 : 0, nullptr, nullptr, nullptr);
                CK(cudaGetLastError());
                if (!gate) {                    // (at a gate sweep nobody is listed: accepted walkers are re-evaluated)
                    k_build_factors<KDSL_KMAX><<<h->num_sms * 8, 128, 0, h->stream>>>(S, h->parity);
                    h->parity ^= 1;
                }
#endif
            } else if (h->cplx) {
                if (replay)
                    k_propose_c<true><<<pgrid, 256, 0, h->stream>>>(S, h->parity, gate ? 1 : 0, h->rp_r + off,
                                                                    h->rp_bond + off, have_pick ? h->rp_pick + off : nullptr);
                else
                    k_propose_c<false><<<pgrid, 256, 0, h->stream>>>(S, h->parity, gate ? 1 : 0, nullptr, nullptr, nullptr);
            } else if (replay) {
                k_propose<true><<<pgrid, 256, 0, h->stream>>>(S, h->parity, gate ? 1 : 0, h->rp_r + off,
                                                              h->rp_bond + off, have_pick ? h->rp_pick + off : nullptr);
            } else {
                k_propose<false><<<pgrid, 256, 0, h->stream>>>(S, h->parity, gate ? 1 : 0, nullptr, nullptr, nullptr);
            }
            CK(cudaGetLastError());
        }
        if (gate) {
            int rc = launch_refresh(h, S.ref_list);
            if (rc) return rc;
            CK(cudaMemsetAsync(S.cnt + 2, 0, sizeof(int), h->stream));
        } else if (!delayed) {
            int rc = launch_update(h, h->parity);
            if (rc) return rc;
            h->parity ^= 1;
        }
        h->since_flush += (int)g;
        if (delayed && h->since_flush >= h->flush_every) {
            int rc = launch_flush(h, false);
            if (rc) return rc;
        }
        h->sweeps += g;                                          // Carlo: ctx.sweeps += 1
        h->walker_sweeps += g * S.nw;
        s += g;
        if (therm >= 0 && h->sweeps > therm && (h->sweeps % S.n_occ) == 0) {   // :630 (post-increment)
            Span sp(h, KDSL_T_MEASURE);
            if (woodbury && h->cplx) k_measure_wb_c<<<S.nw, 256, measure_wb_smem_c(S), h->stream>>>(S, nullptr, 1);
            else if (woodbury) k_measure_wb<<<S.nw, 256, measure_wb_smem(S), h->stream>>>(S, nullptr, 1);
#ifdef KDSL_DEV_VARIANTS
            else if (delayed) k_measure_delayed<<<pgrid, 256, 0, h->stream>>>(S, nullptr, 1);
#endif
            else if (h->cplx) k_measure_c<<<pgrid, 256, 0, h->stream>>>(S, nullptr, 1);
            else k_measure<<<pgrid, 256, 0, h->stream>>>(S, nullptr, 1);
            CK(cudaGetLastError());
            if (S.obs_on) {
                k_measure_extra<<<pgrid, 256, 0, h->stream>>>(S);
                CK(cudaGetLastError());
            }
        }
        if (h->profiling && h->spans.size() > 16384) {
            int rc = flush_spans(h);
            if (rc) return rc;
        }
    }
    return KDSL_OK;
}

int check_ready(kdsl_handle h) {
    if (!h->have_config) return fail(KDSL_ERR_STATE, "no configuration set: call kdsl_set_config first");
    if (!h->W_valid) return fail(KDSL_ERR_STATE, "W is stale: call kdsl_refresh after kdsl_set_config");
    return KDSL_OK;
}

// stage (col, alpha*row) for explicit moves: one warp per move
__global__ void __launch_bounds__(256)
k_stage_moves(DevState S, int parity, int n_moves, const int *__restrict__ mv) {
    const int m = (int)((blockIdx.x * (unsigned)blockDim.x + threadIdx.x) >> 5);
    const int lane = threadIdx.x & 31;
    if (m >= n_moves) return;
    const int w = mv[m], l_up = mv[n_moves + m], K_up = mv[2 * n_moves + m] - 1, l_dn = mv[3 * n_moves + m], K_dn = mv[4 * n_moves + m] - 1;
    const int ns = S.ns;
    const double *Wu = S.W_up + (size_t)w * ns * S.n_up, *Wd = S.W_dn + (size_t)w * ns * S.n_dn;
    const double au = -1.0 / Wu[(size_t)(l_up - 1) * ns + K_up], ad = -1.0 / Wd[(size_t)(l_dn - 1) * ns + K_dn];
    for (int t = lane; t < ns; t += 32) {
        S.col_up[(size_t)w * ns + t] = Wu[(size_t)(l_up - 1) * ns + t];
        S.col_dn[(size_t)w * ns + t] = Wd[(size_t)(l_dn - 1) * ns + t];
    }
    for (int j = lane; j < S.n_up; j += 32) {
        double v = Wu[(size_t)j * ns + K_up];
        if (j == l_up - 1) v -= 1.0;
        S.trow_up[(size_t)w * S.n_up + j] = au * v;
    }
    for (int j = lane; j < S.n_dn; j += 32) {
        double v = Wd[(size_t)j * ns + K_dn];
        if (j == l_dn - 1) v -= 1.0;
        S.trow_dn[(size_t)w * S.n_dn + j] = ad * v;
    }
    if (lane == 0) S.acc_list[(size_t)parity * S.nw + m] = w;
    if (m == 0 && lane == 0) S.cnt[parity] = n_moves;
}

// FP64 tensor-pipe peak probe: 8 independent DMMA accumulation chains per warp
__global__ void __launch_bounds__(256) k_dmma_peak(double *sink, int iters) {
    double c[8][2];
#pragma unroll
    for (int q = 0; q < 8; q++) c[q][0] = c[q][1] = 0.0;
    const double a = 1.0 + 1e-9 * threadIdx.x, b = 1.0 - 1e-9 * threadIdx.x;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int q = 0; q < 8; q++) dmma_8x8x4(c[q][0], c[q][1], a, b);
    }
    double s = 0.0;
#pragma unroll
    for (int q = 0; q < 8; q++) s += c[q][0] + c[q][1];
    if (s == 123.456) sink[0] = s;
}
}  // namespace

extern "C" {

int kdsl_version(void) { return KDSL_VERSION_NUM; }

const char *kdsl_last_error(void) { return g_err.c_str(); }

int kdsl_device_count(int *n) {
    int c = 0;
    cudaError_t e = cudaGetDeviceCount(&c);
    if (n) *n = (e == cudaSuccess) ? c : 0;
    if (e != cudaSuccess) return fail(KDSL_ERR_CUDA, "cudaGetDeviceCount failed: %s", cudaGetErrorString(e));
    if (c == 0) return fail(KDSL_ERR_CUDA, "no CUDA device visible (libkdsl has no CPU path)");
    return KDSL_OK;
}

static int create_impl(kdsl_handle *out, int device, int ns, int n_up, int n_dn, int n_bonds,
                       const int32_t *bonds, const double *U_up, const double *U_dn, int n_walkers, bool cplx) {
    const size_t cz = cplx ? 2 : 1;                               // doubles per matrix element
    if (!out) return fail(KDSL_ERR_INVALID_ARGUMENT, "out is null");
    *out = nullptr;
    if (ns <= 0 || (ns & 1)) return fail(KDSL_ERR_INVALID_ARGUMENT, "ns must be positive and even, got %d", ns);
    if (n_up <= 0 || n_dn <= 0 || n_up > ns || n_dn > ns)
        return fail(KDSL_ERR_INVALID_ARGUMENT, "need 0 < N_up, N_down <= ns (got %d, %d, ns=%d)", n_up, n_dn, ns);
    if (n_up + n_dn != ns)
        return fail(KDSL_ERR_INVALID_ARGUMENT, "Mott constraint: N_up + N_down must equal ns (got %d + %d != %d)", n_up, n_dn, ns);
    if (n_bonds <= 0 || !bonds || !U_up || !U_dn) return fail(KDSL_ERR_INVALID_ARGUMENT, "bonds / U_up / U_dn missing");
    if (n_walkers <= 0) return fail(KDSL_ERR_INVALID_ARGUMENT, "n_walkers must be positive");
    for (int b = 0; b < n_bonds; b++) {
        const int i = bonds[2 * b], j = bonds[2 * b + 1];
        if (i < 1 || j < 1 || i > ns || j > ns || i == j)
            return fail(KDSL_ERR_INVALID_ARGUMENT, "bond %d = (%d, %d) out of range", b + 1, i, j);
    }
    int ndev = 0;
    int rc = kdsl_device_count(&ndev);
    if (rc) return rc;
    if (device < 0 || device >= ndev) return fail(KDSL_ERR_INVALID_ARGUMENT, "device %d not in [0, %d)", device, ndev);
    CK(cudaSetDevice(device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10)
        return fail(KDSL_ERR_CUDA, "device %d is sm_%d%d; libkdsl is built for sm_100a only", device, prop.major, prop.minor);

    kdsl_handle h = new kdsl_handle_s();
    h->cplx = cplx;
    if (cplx) { h->update_variant = 2; h->inverse_variant = 0; }  // complex: Woodbury delayed updates too; inverse through the real embedding
    h->device = device;
    h->num_sms = prop.multiProcessorCount;
    DevState &S = h->S;
    S.ns = ns; S.n_up = n_up; S.n_dn = n_dn; S.n_bonds = n_bonds; S.nw = n_walkers;
    S.n_occ = std::min(n_up, n_dn);                               // src/MonteCarlo.jl:594
    const size_t nw = n_walkers;

#define ALLOC(ptr, n)                                        \
    do {                                                     \
        rc = dev_alloc(h, &(ptr), (n));                      \
        if (rc) { kdsl_destroy(h); return rc; }              \
    } while (0)
#define CKD(call)                                                                          \
    do {                                                                                   \
        cudaError_t e_ = (call);                                                           \
        if (e_ != cudaSuccess) {                                                           \
            kdsl_destroy(h);                                                               \
            return fail(KDSL_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(e_));    \
        }                                                                                  \
    } while (0)

    CKD(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    int *bi, *bj, *adj_off, *adj_nbr;
    double *dUu, *dUd;
    ALLOC(bi, n_bonds); ALLOC(bj, n_bonds); ALLOC(adj_off, ns + 1); ALLOC(adj_nbr, 2 * (size_t)n_bonds);
    ALLOC(dUu, cz * ns * n_up); ALLOC(dUd, cz * ns * n_dn);
    {
        std::vector<int> hbi(n_bonds), hbj(n_bonds), off(ns + 1, 0), nbr(2 * (size_t)n_bonds);
        for (int b = 0; b < n_bonds; b++) {
            hbi[b] = bonds[2 * b] - 1; hbj[b] = bonds[2 * b + 1] - 1;
            off[hbi[b] + 1]++; off[hbj[b] + 1]++;
        }
        for (int s = 0; s < ns; s++) off[s + 1] += off[s];
        std::vector<int> fill(off.begin(), off.end() - 1);
        for (int b = 0; b < n_bonds; b++) {
            nbr[fill[hbi[b]]++] = hbj[b];
            nbr[fill[hbj[b]]++] = hbi[b];
        }
        CKD(copy_sync(h, bi, hbi.data(), n_bonds * sizeof(int), cudaMemcpyHostToDevice));
        CKD(copy_sync(h, bj, hbj.data(), n_bonds * sizeof(int), cudaMemcpyHostToDevice));
        CKD(copy_sync(h, adj_off, off.data(), (ns + 1) * sizeof(int), cudaMemcpyHostToDevice));
        CKD(copy_sync(h, adj_nbr, nbr.data(), 2 * (size_t)n_bonds * sizeof(int), cudaMemcpyHostToDevice));
        CKD(copy_sync(h, dUu, U_up, cz * ns * n_up * sizeof(double), cudaMemcpyHostToDevice));
        CKD(copy_sync(h, dUd, U_dn, cz * ns * n_dn * sizeof(double), cudaMemcpyHostToDevice));
    }
    S.bi = bi; S.bj = bj; S.adj_off = adj_off; S.adj_nbr = adj_nbr; S.U_up = dUu; S.U_dn = dUd;
    ALLOC(S.kup, nw * ns); ALLOC(S.kdn, nw * ns);
    ALLOC(S.rng, nw * 4); ALLOC(S.zmu, nw);
    ALLOC(S.W_up, cz * nw * ns * n_up); ALLOC(S.W_dn, cz * nw * ns * n_dn);
    ALLOC(S.col_up, cz * nw * ns); ALLOC(S.col_dn, cz * nw * ns);
    ALLOC(S.trow_up, cz * nw * n_up); ALLOC(S.trow_dn, cz * nw * n_dn);
    ALLOC(S.acc_list, 12 * nw); ALLOC(S.cnt, 8); ALLOC(S.ref_list, nw); ALLOC(S.flags, nw);
    ALLOC(S.n_acc, nw); ALLOC(S.n_reach, nw); ALLOC(S.n_refresh, nw);
    ALLOC(S.ol_sum, nw); ALLOC(S.ol_sq, nw); ALLOC(S.ol_last, nw); ALLOC(S.ol_n, nw); ALLOC(S.upd_moves, 4);
    S.kmax = KDSL_KMAX; S.kth = KDSL_KTH;                  // (the buffers are sized for the largest kmax = 32)
    const size_t kal = KDSL_KALLOC;
    ALLOC(S.facA_up, cz * nw * kal * ns); ALLOC(S.facA_dn, cz * nw * kal * ns);
#ifdef KDSL_DEV_VARIANTS
    ALLOC(S.facB_up, nw * kal * ((n_up + 7) / 8 * 8)); ALLOC(S.facB_dn, nw * kal * ((n_dn + 7) / 8 * 8));
#endif
    ALLOC(S.fcnt, 2 * nw); ALLOC(S.flush_list, nw); ALLOC(S.listed, nw);
    ALLOC(S.wbT, cz * 2 * nw * KDSL_KALLOC * KDSL_KALLOC); ALLOC(S.wbK, 2 * nw * KDSL_KALLOC); ALLOC(S.wbL, 2 * nw * KDSL_KALLOC);
    h->Gstride = (size_t)((std::max(n_up, n_dn) + 7) / 8 * 8) * KDSL_KALLOC;   // (Gbuf itself is allocated when flush_variant 3 is chosen)
    if (cplx) {
        // the refresh workspace holds the real embedding [[X, -Y], [Y, X]] of tilde_U, padded: Np = roundup(2 N, 8)
        h->Np_up = (2 * n_up + 7) / 8 * 8; h->Np_dn = (2 * n_dn + 7) / 8 * 8;
        h->Npc_up = (n_up + 7) / 8 * 8; h->Npc_dn = (n_dn + 7) / 8 * 8;
        ALLOC(h->A_up, nw * h->Np_up * h->Np_up); ALLOC(h->A_dn, nw * h->Np_dn * h->Np_dn);
        ALLOC(h->X_up, 2 * nw * n_up * n_up); ALLOC(h->X_dn, 2 * nw * n_dn * n_dn);
    } else {
        h->Np_up = (n_up + 7) / 8 * 8; h->Np_dn = (n_dn + 7) / 8 * 8;
        ALLOC(h->A_up, nw * h->Np_up * h->Np_up); ALLOC(h->A_dn, nw * h->Np_dn * h->Np_dn);
    }
    ALLOC(h->colsrc, 2 * nw * std::max(h->Np_up, h->Np_dn));
    ALLOC(h->urow, 2 * nw * ns);
    ALLOC(h->status, 2 * nw); ALLOC(h->d_tmp_i, 2 * nw); ALLOC(h->d_tmp_d, nw); ALLOC(h->d_acc8, 8);
    if (!cplx && h->Np_up <= 256 && h->Np_dn <= 256) {
        // fused re-evaluation: shared-memory budget (operands R - E twice, pivot rows, output staging, tables) for panel
        // width NB; the staging area is dropped when the idle R - E buffer can hold 8 columns of W for both species
        const int Mp_up = (ns - n_up + 7) / 8 * 8, Mp_dn = (ns - n_dn + 7) / 8 * 8;
        const int NpMax = std::max(h->Np_up, h->Np_dn), NpMin = std::min(h->Np_up, h->Np_dn);
        const int CpMax = std::max(h->Np_up + Mp_up, h->Np_dn + Mp_dn);
        auto budget = [&](int NB, int *stage) {
            *stage = (NB * NpMin >= 8 * ns && !getenv("KDSL_NO_ALIAS")) ? 0 : 8 * ns;
            return ((size_t)2 * NB * NpMax + (size_t)NB * CpMax + (size_t)*stage + 2 * NB + 2) * sizeof(double) +
                   ((size_t)12 + 8 + 4 + NB + 32 + NpMax + CpMax + ns + 1) * sizeof(int);
        };
        const size_t smem24 = budget(24, &h->fused_stage24), smem32 = budget(32, &h->fused_stage);
        if (smem24 <= (size_t)prop.sharedMemPerBlockOptin && Mp_up <= 256 && Mp_dn <= 256 && ns <= 512) {
            h->fused_smem24 = smem24;
            h->fused_smem16 = budget(16, &h->fused_stage16);
            if (smem32 <= (size_t)prop.sharedMemPerBlockOptin) { h->fused_smem = smem32; h->fused_NB = 32; }
            else { h->fused_smem = smem24; h->fused_stage = h->fused_stage24; h->fused_NB = 24; }
            h->fused_NpMax = NpMax;
            h->fused_CpMax = CpMax;
            h->ws_stride = std::max((size_t)h->Np_up * (h->Np_up + Mp_up), (size_t)h->Np_dn * (h->Np_dn + Mp_dn));
            ALLOC(h->UT_up, (size_t)(ns + 9) * h->Np_up); ALLOC(h->UT_dn, (size_t)(ns + 9) * h->Np_dn);
            h->fused_small = NpMax <= 128 && Mp_up <= 128 && Mp_dn <= 128 && ns <= 256 && 2 * (smem24 + 1024) <= (size_t)prop.sharedMemPerMultiprocessor;
            ALLOC(h->ws_fused, (size_t)(h->fused_small ? 2 : 1) * h->num_sms * h->ws_stride);
            k_build_UT<<<64, 256, 0, h->stream>>>(dUu, h->UT_up, ns, n_up, h->Np_up);
            k_build_UT<<<64, 256, 0, h->stream>>>(dUd, h->UT_dn, ns, n_dn, h->Np_dn);
            CKD(cudaGetLastError());
            const int optin = (int)smem24;                 // exact sizes: a larger limit changes the L1 carve-out the driver picks
            const int optin32 = (int)std::min(smem32, (size_t)prop.sharedMemPerBlockOptin);
            CKD(cudaFuncSetAttribute(k_reeval_fused<32, 2, 6>, cudaFuncAttributeMaxDynamicSharedMemorySize, optin32));
            CKD(cudaFuncSetAttribute(k_reeval_fused<24, 2, 6>, cudaFuncAttributeMaxDynamicSharedMemorySize, optin));
            CKD(cudaFuncSetAttribute(k_reeval_fused<16, 2, 6>, cudaFuncAttributeMaxDynamicSharedMemorySize, optin));
            CKD(cudaFuncSetAttribute(k_reeval_fused<24, 2, 6, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, optin));
            CKD(cudaFuncSetAttribute(k_reeval_fused<24, 2, 6, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, optin));
            CKD(cudaFuncSetAttribute(k_reeval_fused<24, 1, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, optin));
            CKD(cudaFuncSetAttribute(k_reeval_fused<24, 2, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, optin));
            if (h->fused_small) CKD(cudaFuncSetAttribute(k_reeval_fused<24, 2, 6, 0, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, optin));
            CKD(cudaStreamSynchronize(h->stream));
        }
    }
#ifdef KDSL_DEV_VARIANTS
    if (!cplx && std::max(h->Np_up, h->Np_dn) <= 512 && std::min(h->Np_up, h->Np_dn) >= 8) {
        // cluster re-evaluation (k_reeval_cl; the default for 256 < Np <= 512, inverse_variant 8 forces it on smaller lattices):
        // transposed U with padding rows (shared with k_reeval_fused); the workspaces are allocated at the first launch
        const int Mp_up = (ns - n_up + 7) / 8 * 8, Mp_dn = (ns - n_dn + 7) / 8 * 8;
        h->rcl_NpMax = std::max(h->Np_up, h->Np_dn);
        h->rcl_CpMax = std::max(h->Np_up + Mp_up, h->Np_dn + Mp_dn);
        h->rcl_nvtMax = std::max(Mp_up, Mp_dn) / 8;
        h->rcl_ntcMax = h->rcl_CpMax / 8;
        if (!h->UT_up) {
            ALLOC(h->UT_up, (size_t)(ns + 9) * h->Np_up); ALLOC(h->UT_dn, (size_t)(ns + 9) * h->Np_dn);
            k_build_UT<<<64, 256, 0, h->stream>>>(dUu, h->UT_up, ns, n_up, h->Np_up);
            k_build_UT<<<64, 256, 0, h->stream>>>(dUd, h->UT_dn, ns, n_dn, h->Np_dn);
            CKD(cudaGetLastError());
            CKD(cudaStreamSynchronize(h->stream));
        }
        h->rcl_ok = true;
    }
#endif
#undef ALLOC
    // default xoshiro states must not be all-zero: seed walker w with a fixed SplitMix64 stream
    {
        std::vector<unsigned long long> st(nw * 4);
        unsigned long long x = 0x243F6A8885A308D3ull;
        for (auto &v : st) {
            unsigned long long z = (x += 0x9E3779B97F4A7C15ull);
            z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
            z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
            v = z ^ (z >> 31);
        }
        CKD(copy_sync(h, S.rng, st.data(), st.size() * sizeof(unsigned long long), cudaMemcpyHostToDevice));
    }
    CKD(cudaFuncSetAttribute(k_inverse_gj, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    // Opt-in dynamic shared memory of the Woodbury kernels.  cudaFuncSetAttribute is per device (context), so this is
    // done for every handle, not once per process; the sizes are validated against the device limit here and in
    // kdsl_set_option ("flush_every" / "flush_threshold" change kmax).
    if (cplx) {
        cudaFuncAttributes fa;
        CKD(cudaFuncGetAttributes(&fa, (const void *)k_measure_wb_c));
        h->smem_optin = (size_t)prop.sharedMemPerBlockOptin - fa.sharedSizeBytes;
        CKD(cudaFuncSetAttribute(k_measure_wb_c, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_optin));
        CKD(cudaFuncGetAttributes(&fa, (const void *)k_flush_c<24, 216>));
        CKD(cudaFuncSetAttribute(k_flush_c<24, 216>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)((size_t)prop.sharedMemPerBlockOptin - fa.sharedSizeBytes)));
        {
            // k_flush_dmma_c: the fewest row blocks per species whose operands fit the shared memory
            CKD(cudaFuncGetAttributes(&fa, (const void *)k_flush_dmma_c<24, 4>));
            const size_t lim = (size_t)prop.sharedMemPerBlockOptin - fa.sharedSizeBytes;
            const int Npad = (std::max(n_up, n_dn) + 7) / 8 * 8;
            for (int nrb = 1; nrb <= 64 && !h->fdc_RB; nrb++) {
                const int RB = (((int)ns + nrb - 1) / nrb + 7) / 8 * 8;
                if (flush_dmma_c_smem(24, Npad, RB) <= lim) h->fdc_RB = RB;
            }
            if (h->fdc_RB) CKD(cudaFuncSetAttribute(k_flush_dmma_c<24, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)flush_dmma_c_smem(24, Npad, h->fdc_RB)));
            // k_flush_dmma2_c: two CTAs per SM -- the C planes of a row block and T must fit half an SM
            for (int nrb = 1; nrb <= 64 && !h->fd2_RB; nrb++) {
                const int RB = (((int)ns + nrb - 1) / nrb + 7) / 8 * 8;
                if (2 * (flush_dmma2_c_smem(24, RB) + 2048) <= (size_t)prop.sharedMemPerMultiprocessor) h->fd2_RB = RB;
            }
            if (h->fd2_RB) CKD(cudaFuncSetAttribute(k_flush_dmma2_c<24, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)flush_dmma2_c_smem(24, h->fd2_RB)));
        }
        const size_t need_f = (size_t)24 * std::max(n_up, n_dn) * 2 * sizeof(double);
        if (measure_wb_smem_c(S) > h->smem_optin || need_f + fa.sharedSizeBytes > (size_t)prop.sharedMemPerBlockOptin) {
            // too large for the Woodbury kernels' shared memory: the reference's immediate rank-1 update (any size)
            h->update_variant = 0;
        }
    }
    if (!cplx) {
        // the largest dynamic allocation a kernel may ask for = the device's opt-in limit minus its static shared memory
        auto optin_dynamic = [&](const void *func, size_t *out) -> cudaError_t {
            cudaFuncAttributes fa;
            cudaError_t e = cudaFuncGetAttributes(&fa, func);
            if (e != cudaSuccess) return e;
            const size_t lim = (size_t)prop.sharedMemPerBlockOptin - fa.sharedSizeBytes;
            if (out) *out = lim;
            return cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lim);
        };
        CKD(optin_dynamic((const void *)k_measure_wb, &h->smem_optin));
        CKD(optin_dynamic((const void *)k_flush_wb<20>, nullptr));
        CKD(optin_dynamic((const void *)k_flush_wb<24>, nullptr));
        CKD(optin_dynamic((const void *)k_flush_wb<32>, nullptr));
        CKD(optin_dynamic((const void *)k_flush_G<20>, nullptr));
        CKD(optin_dynamic((const void *)k_flush_G<24>, nullptr));
        CKD(optin_dynamic((const void *)k_flush_G<32>, nullptr));
        CKD(optin_dynamic((const void *)k_flush_tma<20, 4, 3>, nullptr));
        CKD(optin_dynamic((const void *)k_flush_tma<24, 4, 3>, nullptr));
        CKD(optin_dynamic((const void *)k_flush_tma<32, 4, 3>, nullptr));
        // small lattices: the walker-resident kernel when at least two CTAs (walkers) share an SM
        {
            const size_t need = resident_smem_bytes(ns, n_up, n_dn, n_bonds);
            const int np = (std::max(n_up, n_dn) + 31) / 32;     // (M_up = N_dn, M_dn = N_up)
            const void *fn[2] = {nullptr, nullptr};
            if (np <= 1) { fn[0] = (const void *)k_resident<false, 1>; fn[1] = (const void *)k_resident<true, 1>; }
            else if (np == 2) { fn[0] = (const void *)k_resident<false, 2>; fn[1] = (const void *)k_resident<true, 2>; }
            else if (np == 3) { fn[0] = (const void *)k_resident<false, 3>; fn[1] = (const void *)k_resident<true, 3>; }
            else if (np == 4) { fn[0] = (const void *)k_resident<false, 4>; fn[1] = (const void *)k_resident<true, 4>; }
            if (fn[0]) {
                cudaFuncAttributes fa;
                CKD(cudaFuncGetAttributes(&fa, fn[0]));
                if (need + fa.sharedSizeBytes <= (size_t)prop.sharedMemPerBlockOptin) {
                    CKD(cudaFuncSetAttribute(fn[0], cudaFuncAttributeMaxDynamicSharedMemorySize, (int)need));
                    CKD(cudaFuncSetAttribute(fn[1], cudaFuncAttributeMaxDynamicSharedMemorySize, (int)need));
                    int nb = 0;
                    CKD(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, fn[0], KDSL_RES_THREADS, need));
                    if (nb >= 1) {
                        h->res_smem = need;
                        h->res_ctas_per_sm = nb;
                        h->res_np = np;
                        if (nb >= 2 && !getenv("KDSL_NO_RESIDENT")) h->update_variant = 3;
                    }
                }
            }
        }
        if (measure_wb_smem(S) > h->smem_optin) {
            const size_t need = measure_wb_smem(S), lim = h->smem_optin;
            kdsl_destroy(h);
            return fail(KDSL_ERR_INVALID_ARGUMENT, "k_measure_wb needs %zu bytes of shared memory at ns = %d, kmax = %d (device limit %zu)",
                        need, ns, KDSL_KMAX, lim);
        }
    }
#undef CKD
    *out = h;
    return KDSL_OK;
}

int kdsl_create(kdsl_handle *out, int device, int ns, int n_up, int n_dn, int n_bonds,
                const int32_t *bonds, const double *U_up, const double *U_dn, int n_walkers) {
    return create_impl(out, device, ns, n_up, n_dn, n_bonds, bonds, U_up, U_dn, n_walkers, false);
}

int kdsl_create_c128(kdsl_handle *out, int device, int ns, int n_up, int n_dn, int n_bonds,
                     const int32_t *bonds, const double *U_up, const double *U_dn, int n_walkers) {
    return create_impl(out, device, ns, n_up, n_dn, n_bonds, bonds, U_up, U_dn, n_walkers, true);
}

int kdsl_is_complex(kdsl_handle h, int *is_complex) {
    if (!h || !is_complex) return fail(KDSL_ERR_INVALID_ARGUMENT, "null handle / output");
    *is_complex = h->cplx ? 1 : 0;
    return KDSL_OK;
}

int kdsl_destroy(kdsl_handle h) {
    if (!h) return KDSL_OK;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    if (h->comm) kdsl_comm_destroy(h);
    for (auto &s : h->spans) { cudaEventDestroy(s.a); cudaEventDestroy(s.b); }
    for (auto e : h->ev_pool) cudaEventDestroy(e);
    for (auto e : h->user_ev) if (e) cudaEventDestroy(e);
    for (void *p : h->allocs) cudaFree(p);
    if (h->d_qcos) cudaFree(h->d_qcos);
    if (h->d_qsin) cudaFree(h->d_qsin);
    if (h->d_obs) cudaFree(h->d_obs);
    if (h->S.obs_w) cudaFree(h->S.obs_w);
    if (h->rp_r) { cudaFree(h->rp_r); cudaFree(h->rp_bond); cudaFree(h->rp_pick); }
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
    return KDSL_OK;
}

int kdsl_info(kdsl_handle h, int64_t *out) {
    if (!h || !out) return fail(KDSL_ERR_INVALID_ARGUMENT, "null argument");
    out[0] = h->S.ns; out[1] = h->S.n_up; out[2] = h->S.n_dn; out[3] = h->S.n_bonds; out[4] = h->S.nw; out[5] = h->S.n_occ;
    return KDSL_OK;
}

int kdsl_set_config(kdsl_handle h, const int32_t *kup, const int32_t *kdn) {
    int rc = use_device(h);
    if (rc) return rc;
    if (!kup || !kdn) return fail(KDSL_ERR_INVALID_ARGUMENT, "kappa_up / kappa_dn is null");
    const DevState &S = h->S;
    std::vector<char> seen(std::max(S.n_up, S.n_dn) + 1);
    for (int w = 0; w < S.nw; w++) {
        const int32_t *ku = kup + (size_t)w * S.ns, *kd = kdn + (size_t)w * S.ns;
        for (int spin = 0; spin < 2; spin++) {
            const int32_t *k = spin ? kd : ku;
            const int N = spin ? S.n_dn : S.n_up;
            std::fill(seen.begin(), seen.end(), 0);
            int cnt = 0;
            for (int R = 0; R < S.ns; R++) {
                const int l = k[R];
                if (l == 0) continue;
                if (l < 1 || l > N)
                    return fail(KDSL_ERR_INVALID_ARGUMENT, "walker %d: kappa_%s[%d] = %d is not a label in 1..%d (BoundsError in tilde_U)",
                                w, spin ? "down" : "up", R + 1, l, N);
                if (seen[l])
                    return fail(KDSL_ERR_INVALID_ARGUMENT, "walker %d: label %d appears twice in kappa_%s", w, l, spin ? "down" : "up");
                seen[l] = 1;
                cnt++;
            }
            if (cnt != N)
                return fail(KDSL_ERR_INVALID_ARGUMENT, "walker %d: kappa_%s holds %d particles, expected %d (kappa is not valid)",
                            w, spin ? "down" : "up", cnt, N);
        }
        for (int R = 0; R < S.ns; R++)
            if ((ku[R] != 0) == (kd[R] != 0))
                return fail(KDSL_ERR_INVALID_ARGUMENT, "walker %d: site %d is %s (not a Mott state)", w, R + 1,
                            ku[R] != 0 ? "doubly occupied" : "unoccupied");
    }
    CK(cudaStreamSynchronize(h->stream));
    CK(copy_sync(h, S.kup, kup, (size_t)S.nw * S.ns * sizeof(int), cudaMemcpyHostToDevice));
    CK(copy_sync(h, S.kdn, kdn, (size_t)S.nw * S.ns * sizeof(int), cudaMemcpyHostToDevice));
    k_count_Z<<<grid_for_warps(S.nw), 256, 0, h->stream>>>(S, nullptr, 1);
    CK(cudaGetLastError());
    CK(cudaMemsetAsync(S.cnt, 0, 8 * sizeof(int), h->stream));
    CK(cudaMemsetAsync(S.flags, 0, (size_t)S.nw * sizeof(int), h->stream));
    CK(cudaMemsetAsync(S.fcnt, 0, (size_t)2 * S.nw * sizeof(int), h->stream));
    CK(cudaMemsetAsync(S.listed, 0, (size_t)S.nw * sizeof(int), h->stream));
    CK(cudaStreamSynchronize(h->stream));
    h->parity = 0;
    h->since_flush = 0;
    h->have_config = true;
    h->W_valid = false;
    return KDSL_OK;
}

int kdsl_get_config(kdsl_handle h, int32_t *kup, int32_t *kdn) {
    int rc = use_device(h);
    if (rc) return rc;
    if (!kup || !kdn) return fail(KDSL_ERR_INVALID_ARGUMENT, "null output");
    CK(cudaStreamSynchronize(h->stream));
    CK(copy_sync(h, kup, h->S.kup, (size_t)h->S.nw * h->S.ns * sizeof(int), cudaMemcpyDeviceToHost));
    CK(copy_sync(h, kdn, h->S.kdn, (size_t)h->S.nw * h->S.ns * sizeof(int), cudaMemcpyDeviceToHost));
    return KDSL_OK;
}

int kdsl_set_rng(kdsl_handle h, const uint64_t *states) {
    int rc = use_device(h);
    if (rc) return rc;
    if (!states) return fail(KDSL_ERR_INVALID_ARGUMENT, "states is null");
    for (int w = 0; w < h->S.nw; w++)
        if (!(states[4 * w] | states[4 * w + 1] | states[4 * w + 2] | states[4 * w + 3]))
            return fail(KDSL_ERR_INVALID_ARGUMENT, "walker %d: all-zero Xoshiro state", w);
    CK(cudaStreamSynchronize(h->stream));
    CK(copy_sync(h, h->S.rng, states, (size_t)h->S.nw * 4 * sizeof(uint64_t), cudaMemcpyHostToDevice));
    return KDSL_OK;
}

int kdsl_get_rng(kdsl_handle h, uint64_t *states) {
    int rc = use_device(h);
    if (rc) return rc;
    if (!states) return fail(KDSL_ERR_INVALID_ARGUMENT, "states is null");
    CK(cudaStreamSynchronize(h->stream));
    CK(copy_sync(h, states, h->S.rng, (size_t)h->S.nw * 4 * sizeof(uint64_t), cudaMemcpyDeviceToHost));
    return KDSL_OK;
}

int kdsl_set_sweeps(kdsl_handle h, int64_t sweeps) {
    if (!h) return fail(KDSL_ERR_INVALID_ARGUMENT, "null handle");
    if (sweeps < 0) return fail(KDSL_ERR_INVALID_ARGUMENT, "sweeps must be >= 0");
    h->sweeps = sweeps;
    return KDSL_OK;
}

int kdsl_get_sweeps(kdsl_handle h, int64_t *sweeps) {
    if (!h || !sweeps) return fail(KDSL_ERR_INVALID_ARGUMENT, "null argument");
    *sweeps = h->sweeps;
    return KDSL_OK;
}

int kdsl_refresh(kdsl_handle h, int *n_singular) {
    int rc = use_device(h);
    if (rc) return rc;
    if (!h->have_config) return fail(KDSL_ERR_STATE, "no configuration set: call kdsl_set_config first");
    CK(cudaMemsetAsync(h->S.cnt + 3, 0, sizeof(int), h->stream));
    rc = launch_refresh(h, nullptr);
    if (rc) return rc;
    int ns_ = 0;
    CK(cudaMemcpyAsync(&ns_, h->S.cnt + 3, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    CK(cudaGetLastError());
    if (n_singular) *n_singular = ns_;
    h->W_valid = true;
    if (ns_ > 0) return fail(KDSL_ERR_SINGULAR, "SingularException: tilde_U is singular for %d walker(s)", ns_);
    return KDSL_OK;
}

int kdsl_sweep(kdsl_handle h, int64_t n_sweeps, int64_t thermalization) {
    int rc = use_device(h);
    if (rc) return rc;
    if ((rc = check_ready(h))) return rc;
    if (n_sweeps < 0) return fail(KDSL_ERR_INVALID_ARGUMENT, "n_sweeps must be >= 0");
    return run_sweeps(h, n_sweeps, thermalization, false, false);
}

int kdsl_replay(kdsl_handle h, int64_t n_sweeps, int64_t thermalization, const double *r,
                const int32_t *bond_idx, const int32_t *pick) {
    int rc = use_device(h);
    if (rc) return rc;
    if ((rc = check_ready(h))) return rc;
    if (n_sweeps < 0) return fail(KDSL_ERR_INVALID_ARGUMENT, "n_sweeps must be >= 0");
    if (n_sweeps == 0) return KDSL_OK;
    if (!r || !bond_idx) return fail(KDSL_ERR_INVALID_ARGUMENT, "r / bond_idx is null");
    const size_t n = (size_t)n_sweeps * h->S.nw;
    if ((rc = ensure_replay_capacity(h, n))) return rc;
    CK(cudaMemcpyAsync(h->rp_r, r, n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(h->rp_bond, bond_idx, n * sizeof(int), cudaMemcpyHostToDevice, h->stream));
    if (pick) CK(cudaMemcpyAsync(h->rp_pick, pick, n * sizeof(int), cudaMemcpyHostToDevice, h->stream));
    return run_sweeps(h, n_sweeps, thermalization, true, pick != nullptr);
}

int kdsl_measure(kdsl_handle h, double *ol) {
    int rc = use_device(h);
    if (rc) return rc;
    if ((rc = check_ready(h))) return rc;
    if (!ol) return fail(KDSL_ERR_INVALID_ARGUMENT, "ol is null");
    const DevState &S = h->S;
    {
        Span sp(h, KDSL_T_MEASURE);
        if (h->cplx && h->update_variant == 2) k_measure_wb_c<<<S.nw, 256, measure_wb_smem_c(S), h->stream>>>(S, h->d_tmp_d, 0);
        else if (h->cplx) k_measure_c<<<grid_for_warps(S.nw), 256, 0, h->stream>>>(S, h->d_tmp_d, 0);
        else if (h->update_variant == 2) k_measure_wb<<<S.nw, 256, measure_wb_smem(S), h->stream>>>(S, h->d_tmp_d, 0);
        // (update_variant 0 and 3 keep W itself up to date: the plain kernel below)
#ifdef KDSL_DEV_VARIANTS
        else if (h->update_variant == 1) k_measure_delayed<<<grid_for_warps(S.nw), 256, 0, h->stream>>>(S, h->d_tmp_d, 0);
#endif
        else k_measure<<<grid_for_warps(S.nw), 256, 0, h->stream>>>(S, h->d_tmp_d, 0);
        CK(cudaGetLastError());
    }
    CK(cudaMemcpyAsync(ol, h->d_tmp_d, (size_t)S.nw * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    std::vector<int> fl(S.nw);
    CK(cudaMemcpyAsync(fl.data(), S.flags, (size_t)S.nw * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    for (int w = 0; w < S.nw; w++)
        if (fl[w] & KDSL_FLAG_BAD_SITE)
            return fail(KDSL_ERR_INVALID_ARGUMENT, "walker %d: a site is unoccupied or doubly occupied (ArgumentError in Sz)", w);
    return KDSL_OK;
}

int kdsl_last_OL(kdsl_handle h, double *ol, int64_t *n_samples) {
    int rc = use_device(h);
    if (rc) return rc;
    if (!ol) return fail(KDSL_ERR_INVALID_ARGUMENT, "ol is null");
    const DevState &S = h->S;
    CK(cudaMemcpyAsync(ol, S.ol_last, (size_t)S.nw * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    if (n_samples)
        CK(cudaMemcpyAsync(n_samples, S.ol_n, (size_t)S.nw * sizeof(int64_t), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return KDSL_OK;
}

int kdsl_accumulators(kdsl_handle h, double *out, int64_t *acc_per_walker, double *ol_sum_per_walker) {
    int rc = use_device(h);
    if (rc) return rc;
    if (!out) return fail(KDSL_ERR_INVALID_ARGUMENT, "out is null");
    const DevState &S = h->S;
    k_reduce_acc<<<1, 1024, 0, h->stream>>>(S, h->d_acc8, (double)h->walker_sweeps);
    CK(cudaGetLastError());
    double tmp[8];
    CK(cudaMemcpyAsync(tmp, h->d_acc8, 8 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    if (acc_per_walker)
        CK(cudaMemcpyAsync(acc_per_walker, S.n_acc, (size_t)S.nw * sizeof(int64_t), cudaMemcpyDeviceToHost, h->stream));
    if (ol_sum_per_walker)
        CK(cudaMemcpyAsync(ol_sum_per_walker, S.ol_sum, (size_t)S.nw * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    memcpy(out, tmp, sizeof tmp);                       // (k_reduce_acc writes the KDSL_ACC_* order)
    return KDSL_OK;
}

int kdsl_reset_accumulators(kdsl_handle h) {
    int rc = use_device(h);
    if (rc) return rc;
    const DevState &S = h->S;
    const size_t nw = S.nw;
    CK(cudaMemsetAsync(S.n_acc, 0, nw * 8, h->stream));
    CK(cudaMemsetAsync(S.n_reach, 0, nw * 8, h->stream));
    CK(cudaMemsetAsync(S.n_refresh, 0, nw * 8, h->stream));
    CK(cudaMemsetAsync(S.ol_sum, 0, nw * 8, h->stream));
    CK(cudaMemsetAsync(S.ol_sq, 0, nw * 8, h->stream));
    CK(cudaMemsetAsync(S.ol_n, 0, nw * 8, h->stream));
    CK(cudaMemsetAsync(S.cnt + 3, 0, sizeof(int), h->stream));
    if (S.obs_w) CK(cudaMemsetAsync(S.obs_w, 0, nw * (4 + 2 * (size_t)S.nq) * sizeof(double), h->stream));
    h->walker_sweeps = 0;
    return KDSL_OK;
}

int kdsl_set_observables(kdsl_handle h, int nq, const double *cos_qr, const double *sin_qr) {
    int rc = use_device(h);
    if (rc) return rc;
    if (nq < 0 || nq > 4096 || (nq > 0 && (!cos_qr || !sin_qr)))
        return fail(KDSL_ERR_INVALID_ARGUMENT, "need 0 <= nq <= 4096 and both phase tables");
    DevState &S = h->S;
    CK(cudaStreamSynchronize(h->stream));
    const size_t nw = S.nw, ns = S.ns;
    auto renew = [&](double **p, size_t n) -> cudaError_t {
        if (*p) cudaFree(*p);
        *p = nullptr;
        return cudaMalloc((void **)p, std::max<size_t>(n, 1) * sizeof(double));
    };
    CK(renew(&h->d_qcos, (size_t)nq * ns));
    CK(renew(&h->d_qsin, (size_t)nq * ns));
    CK(renew(&h->d_obs, 4 + 2 * (size_t)nq));
    CK(renew(&S.obs_w, nw * (4 + 2 * (size_t)nq)));
    if (nq > 0) {
        CK(copy_sync(h, h->d_qcos, cos_qr, (size_t)nq * ns * sizeof(double), cudaMemcpyHostToDevice));
        CK(copy_sync(h, h->d_qsin, sin_qr, (size_t)nq * ns * sizeof(double), cudaMemcpyHostToDevice));
    }
    CK(memset_sync(h, S.obs_w, 0, nw * (4 + 2 * (size_t)nq) * sizeof(double)));
    S.q_cos = h->d_qcos; S.q_sin = h->d_qsin; S.nq = nq; S.obs_on = 1;
    return KDSL_OK;
}


int kdsl_get_W(kdsl_handle h, int walker, int spin, double *out) {
    int rc = use_device(h);
    if (rc) return rc;
    const DevState &S = h->S;
    if (!out || walker < 0 || walker >= S.nw || spin < 0 || spin > 1) return fail(KDSL_ERR_INVALID_ARGUMENT, "bad walker / spin / out");
    const int N = spin ? S.n_dn : S.n_up;
    const double *src = (spin ? S.W_dn : S.W_up) + (size_t)walker * S.ns * N * (h->cplx ? 2 : 1);
    if ((h->update_variant == 1 || h->update_variant == 2) && (rc = launch_flush(h, true))) return rc;   // fold pending delayed factors into W0
    CK(cudaStreamSynchronize(h->stream));
    CK(copy_sync(h, out, src, (size_t)S.ns * N * sizeof(double) * (h->cplx ? 2 : 1), cudaMemcpyDeviceToHost));
    return KDSL_OK;
}

int kdsl_set_W(kdsl_handle h, int walker, int spin, const double *in) {
    int rc = use_device(h);
    if (rc) return rc;
    const DevState &S = h->S;
    if (!in || walker < 0 || walker >= S.nw || spin < 0 || spin > 1) return fail(KDSL_ERR_INVALID_ARGUMENT, "bad walker / spin / in");
    const int N = spin ? S.n_dn : S.n_up;
    double *dst = (spin ? S.W_dn : S.W_up) + (size_t)walker * S.ns * N * (h->cplx ? 2 : 1);
    if ((h->update_variant == 1 || h->update_variant == 2) && (rc = launch_flush(h, true))) return rc;
    CK(cudaStreamSynchronize(h->stream));
    CK(copy_sync(h, dst, in, (size_t)S.ns * N * sizeof(double) * (h->cplx ? 2 : 1), cudaMemcpyHostToDevice));
    return KDSL_OK;
}

int kdsl_update_W(kdsl_handle h, int n_moves, const int32_t *walker, const int32_t *l_up,
                  const int32_t *K_up, const int32_t *l_dn, const int32_t *K_dn) {
    int rc = use_device(h);
    if (rc) return rc;
    const DevState &S = h->S;
    if (n_moves < 0 || n_moves > S.nw) return fail(KDSL_ERR_INVALID_ARGUMENT, "n_moves must be in [0, n_walkers]");
    if (n_moves == 0) return KDSL_OK;
    if (!walker || !l_up || !K_up || !l_dn || !K_dn) return fail(KDSL_ERR_INVALID_ARGUMENT, "null move array");
    std::vector<char> used(S.nw, 0);
    std::vector<int> mv(5 * (size_t)n_moves);
    for (int m = 0; m < n_moves; m++) {
        if (walker[m] < 0 || walker[m] >= S.nw || used[walker[m]]) return fail(KDSL_ERR_INVALID_ARGUMENT, "move %d: walker id invalid or repeated", m);
        used[walker[m]] = 1;
        if (l_up[m] < 1 || l_up[m] > S.n_up || l_dn[m] < 1 || l_dn[m] > S.n_dn || K_up[m] < 1 || K_up[m] > S.ns || K_dn[m] < 1 || K_dn[m] > S.ns)
            return fail(KDSL_ERR_INVALID_ARGUMENT, "move %d: label or site out of range (BoundsError)", m);
        mv[m] = walker[m]; mv[n_moves + m] = l_up[m]; mv[2 * (size_t)n_moves + m] = K_up[m];
        mv[3 * (size_t)n_moves + m] = l_dn[m]; mv[4 * (size_t)n_moves + m] = K_dn[m];
    }
    if ((h->update_variant == 1 || h->update_variant == 2) && (rc = launch_flush(h, true))) return rc;
    int *d_mv = nullptr;
    CK(cudaMalloc(&d_mv, mv.size() * sizeof(int)));
    CK(cudaMemcpyAsync(d_mv, mv.data(), mv.size() * sizeof(int), cudaMemcpyHostToDevice, h->stream));
    if (h->cplx) k_stage_moves_c<<<grid_for_warps(n_moves), 256, 0, h->stream>>>(S, h->parity, n_moves, d_mv);
    else k_stage_moves<<<grid_for_warps(n_moves), 256, 0, h->stream>>>(S, h->parity, n_moves, d_mv);
    CK(cudaGetLastError());
    rc = launch_update(h, h->parity);
    h->parity ^= 1;
    cudaStreamSynchronize(h->stream);
    cudaFree(d_mv);
    if (rc) return rc;
    CK(cudaGetLastError());
    return KDSL_OK;
}

int kdsl_get_Z(kdsl_handle h, int32_t *zmu, int32_t *zmu_recount) {
    int rc = use_device(h);
    if (rc) return rc;
    const DevState &S = h->S;
    if (!zmu) return fail(KDSL_ERR_INVALID_ARGUMENT, "zmu is null");
    if (zmu_recount) {
        k_count_Z<<<grid_for_warps(S.nw), 256, 0, h->stream>>>(S, h->d_tmp_i, 0);
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(zmu_recount, h->d_tmp_i, (size_t)S.nw * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    }
    CK(cudaMemcpyAsync(zmu, S.zmu, (size_t)S.nw * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return KDSL_OK;
}

int kdsl_get_flags(kdsl_handle h, int32_t *flags) {
    int rc = use_device(h);
    if (rc) return rc;
    if (!flags) return fail(KDSL_ERR_INVALID_ARGUMENT, "flags is null");
    CK(cudaMemcpyAsync(flags, h->S.flags, (size_t)h->S.nw * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return KDSL_OK;
}

int kdsl_set_profiling(kdsl_handle h, int enabled) {
    int rc = use_device(h);
    if (rc) return rc;
    if (!enabled && h->profiling) {
        rc = flush_spans(h);
        if (rc) return rc;
    }
    h->profiling = enabled != 0;
    return KDSL_OK;
}

int kdsl_timers(kdsl_handle h, double *ms, int64_t *launches, int64_t *update_moves) {
    int rc = use_device(h);
    if (rc) return rc;
    if ((rc = flush_spans(h))) return rc;
    if (ms) memcpy(ms, h->t_ms, sizeof h->t_ms);
    if (launches) memcpy(launches, h->t_launch, sizeof h->t_launch);
    if (update_moves) {
        unsigned long long v[2] = {0, 0};
        CK(copy_sync(h, v, h->S.upd_moves, sizeof v, cudaMemcpyDeviceToHost));
        update_moves[0] = (int64_t)v[0];
        update_moves[1] = (int64_t)v[1];
    }
    return KDSL_OK;
}

int kdsl_reset_timers(kdsl_handle h) {
    int rc = use_device(h);
    if (rc) return rc;
    if ((rc = flush_spans(h))) return rc;
    memset(h->t_ms, 0, sizeof h->t_ms);
    memset(h->t_launch, 0, sizeof h->t_launch);
    CK(memset_sync(h, h->S.upd_moves, 0, 4 * sizeof(unsigned long long)));
    return KDSL_OK;
}

int kdsl_set_option(kdsl_handle h, const char *name, int64_t value) {
    if (h && h->cplx && name) {
        const std::string nm(name);
        if ((nm == "update_variant" && value != 0 && value != 2) || (nm == "inverse_variant" && value != 0 && value != 1 && value != 4 && value != 5 && value != 7 && value != 9) ||
            (nm == "flush_variant" && value != 0 && value != 4 && value != 7))
            return fail(KDSL_ERR_STATE, "option %s = %lld is not available in ComplexF64 mode (update_variant 0 / 2; flush_variant 0 = tensor-pipe flush, 4 = FMA flush; inverse_variant 0 / 9 = complex cluster inverse, 4 / 5 / 7 = blocked inverse of the real embedding, 1 = unblocked complex)", name, (long long)value);
        if (nm == "update_variant" && value == 2 && measure_wb_smem_c(h->S) > h->smem_optin)
            return fail(KDSL_ERR_INVALID_ARGUMENT, "the ComplexF64 Woodbury kernels need %zu bytes of shared memory at ns = %d (device limit %zu)",
                        measure_wb_smem_c(h->S), h->S.ns, h->smem_optin);
    }
    if (!h || !name) return fail(KDSL_ERR_INVALID_ARGUMENT, "null argument");
    const std::string n(name);
    if (n == "refresh_every") {
        if (value < 0) return fail(KDSL_ERR_INVALID_ARGUMENT, "refresh_every must be >= 0");
        h->refresh_every = value;
    } else if (n == "update_variant") {
        if (value < 0 || value > 3) return fail(KDSL_ERR_INVALID_ARGUMENT, "update_variant must be 0 (rank-1 streaming), 2 (delayed, Woodbury form) or 3 (walker resident in shared memory)");
        if (value == 3 && h->res_smem == 0)
            return fail(KDSL_ERR_INVALID_ARGUMENT, "update_variant 3 needs a walker's compact state (%zu bytes at ns = %d) in one CTA's shared memory",
                        resident_smem_bytes(h->S.ns, h->S.n_up, h->S.n_dn, h->S.n_bonds), h->S.ns);
#ifndef KDSL_DEV_VARIANTS
        if (value == 1) return fail(KDSL_ERR_INVALID_ARGUMENT, "update_variant 1 (delayed factor lists) is a developer variant (build with make DEV=1)");
#endif
        if (h->have_config && h->W_valid && (h->update_variant == 1 || h->update_variant == 2)) {
            int rc = use_device(h);
            if (rc) return rc;
            if ((rc = launch_flush(h, true))) return rc;
        }
        if (h->have_config) {
            int rc = use_device(h);
            if (rc) return rc;
            CK(cudaMemsetAsync(h->S.cnt, 0, 2 * sizeof(int), h->stream));
        }
        h->parity = 0;
        h->update_variant = (int)value;
        if (h->have_config) CK(cudaStreamSynchronize(h->stream));
        h->flush_every = KDSL_FLUSH_EVERY;                // the factor-list kernels are compiled for kmax = 20
        h->S.kth = KDSL_KTH;
        h->S.kmax = KDSL_KMAX;
    }
    else if (n == "update_ctas_per_sm") h->update_ctas_per_sm = (int)value;
    else if (n == "update_cols_per_item") {
        if (value < 1 || value > 4096) return fail(KDSL_ERR_INVALID_ARGUMENT, "update_cols_per_item out of range");
        h->update_ch = (int)value;
    } else if (n == "inverse_variant") {
#ifndef KDSL_DEV_VARIANTS
        if (value == 2 || value == 3) return fail(KDSL_ERR_INVALID_ARGUMENT, "inverse_variant %lld is a developer variant (build with make DEV=1)", (long long)value);
#endif
#ifndef KDSL_DEV_VARIANTS
        if (value == 8) return fail(KDSL_ERR_INVALID_ARGUMENT, "inverse_variant 8 (cluster re-evaluation) is a developer variant (build with make DEV=1)");
#endif
        if (value < 0 || value > 9 || (value == 9 && !h->cplx)) return fail(KDSL_ERR_INVALID_ARGUMENT, "inverse_variant must be 0, 1, 4, 5, 6, 7 or 8 (9: ComplexF64 engine only)");
        h->inverse_variant = (int)value;
    }
    else if (n == "fuse_sweeps") h->fuse_sweeps = (int)value;
    else if (n == "flush_variant") {
#ifndef KDSL_DEV_VARIANTS
        if (value == 3 && !h->Gbuf) {                    // the bulk-async path keeps G = -T Rt of every listed walker in global memory
            int rc = use_device(h);
            if (rc) return rc;
            if ((rc = dev_alloc(h, &h->Gbuf, (size_t)2 * h->S.nw * h->Gstride))) return rc;
        }
        if (value != 0 && value != 3 && !(h->cplx && (value == 4 || value == 7))) return fail(KDSL_ERR_INVALID_ARGUMENT, "flush_variant %lld is a developer variant (build with make DEV=1)", (long long)value);
#endif
        h->flush_variant = (int)value;
    }
    else if (n == "flush_every" || n == "flush_threshold") {
        // Woodbury mode only: a walker is flushed at the first flush launch after it reached `flush_threshold`
        // pending updates; launches come every `flush_every` sweeps, so at most threshold + every are ever pending
        if (h->update_variant != 2) return fail(KDSL_ERR_STATE, "%s applies to update_variant 2 only", name);
        const int fe = n == "flush_every" ? (int)value : h->flush_every;
        const int kth = n == "flush_threshold" ? (int)value : h->S.kth;
        if (fe < 1 || kth < 1 || fe + kth > (h->cplx ? 24 : KDSL_KALLOC))
            return fail(KDSL_ERR_INVALID_ARGUMENT, "need flush_every >= 1, flush_threshold >= 1 and their sum <= %d", h->cplx ? 24 : KDSL_KALLOC);
        {
            DevState T = h->S;
            T.kmax = fe + kth;
            if ((h->cplx ? measure_wb_smem_c(T) : measure_wb_smem(T)) > h->smem_optin)
                return fail(KDSL_ERR_INVALID_ARGUMENT, "flush_every + flush_threshold = %d needs %zu bytes of shared memory in k_measure_wb at ns = %d (device limit %zu)",
                            fe + kth, measure_wb_smem(T), T.ns, h->smem_optin);
        }
        int rc = use_device(h);
        if (rc) return rc;
        if (h->have_config && h->W_valid) {               // strides change: no update may be pending
            rc = launch_flush(h, true);
            if (rc) return rc;
            CK(cudaStreamSynchronize(h->stream));
        }
        h->flush_every = fe;
        h->S.kth = kth;
        h->S.kmax = fe + kth;
    }
    else if (n == "gemm_variant") {
#ifndef KDSL_DEV_VARIANTS
        if (value == 2 || value == 3) return fail(KDSL_ERR_INVALID_ARGUMENT, "gemm_variant %lld is a developer variant (build with make DEV=1)", (long long)value);
#endif
        h->gemm_variant = (int)value;
    }
    else if (n == "fused_ctas") h->fused_ctas = (int)value;
    else if (n == "flush_dbg") h->flush_dbg = (int)value;
    else if (n == "inverse_tuning") h->inverse_tuning = (int)value;
    else if (n == "reeval_cluster") {
        if (value < 2 || value > 8) return fail(KDSL_ERR_INVALID_ARGUMENT, "reeval_cluster must be 2..8 CTAs per matrix");
        h->reeval_cluster = (int)value;
    }
    else if (n == "reeval_rs") {
        if (value < 1 || value > 16) return fail(KDSL_ERR_INVALID_ARGUMENT, "reeval_rs must be 1..16");
        h->reeval_rs = (int)value;
    }
    else if (n == "inverse_cluster") {
        if (value != 0 && (value < 2 || value > 8)) return fail(KDSL_ERR_INVALID_ARGUMENT, "inverse_cluster must be 0 (one CTA per matrix) or 2..8 CTAs per matrix");
        h->inverse_cluster = (int)value;
    }
    else if (n == "inverse_row_slices") {
        if (value < 1 || value > 64) return fail(KDSL_ERR_INVALID_ARGUMENT, "inverse_row_slices must be 1..64");
        h->inverse_rs = (int)value;
    }
    else return fail(KDSL_ERR_INVALID_ARGUMENT, "unknown option '%s'", name);
    return KDSL_OK;
}

int kdsl_event_record(kdsl_handle h, int slot) {
    int rc = use_device(h);
    if (rc) return rc;
    if (slot < 0 || slot >= 16) return fail(KDSL_ERR_INVALID_ARGUMENT, "event slot must be in 0..15");
    if (!h->user_ev[slot]) CK(cudaEventCreate(&h->user_ev[slot]));
    CK(cudaEventRecord(h->user_ev[slot], h->stream));
    return KDSL_OK;
}

int kdsl_event_elapsed(kdsl_handle h, int a, int b, double *ms) {
    int rc = use_device(h);
    if (rc) return rc;
    if (!ms || a < 0 || a >= 16 || b < 0 || b >= 16 || !h->user_ev[a] || !h->user_ev[b])
        return fail(KDSL_ERR_INVALID_ARGUMENT, "event slots not recorded");
    CK(cudaEventSynchronize(h->user_ev[b]));
    float f = 0.f;
    CK(cudaEventElapsedTime(&f, h->user_ev[a], h->user_ev[b]));
    *ms = f;
    return KDSL_OK;
}

#ifdef KDSL_PHASE_TICKS
/* developer hook (not in kdsl.h): cycles of CTA 0 per phase of the re-evaluation kernels since the last call */
int kdsl_debug_inverse_phases(kdsl_handle h, long long *out) {
    int rc = use_device(h);
    if (rc) return rc;
    CK(cudaStreamSynchronize(h->stream));
    CK(cudaMemcpyFromSymbol(out, g_inv_phase_cycles, 16 * sizeof(long long)));   // out: 16 entries
    long long z[16] = {0};
    CK(cudaMemcpyToSymbol(g_inv_phase_cycles, z, sizeof z));
    return KDSL_OK;
}
#endif

int kdsl_bench_fp64_dmma_sustained(kdsl_handle h, double seconds, double *sustained, double *burst) {
    int rc = use_device(h);
    if (rc) return rc;
    if (!sustained && !burst) return fail(KDSL_ERR_INVALID_ARGUMENT, "no output");
    if (!(seconds >= 0.0) || seconds > 30.0) return fail(KDSL_ERR_INVALID_ARGUMENT, "seconds must be in [0, 30]");
    const int iters = 4096, grid = h->num_sms * 8;
    const double flops = (double)grid * 8 /*warps*/ * iters * 8 /*chains*/ * 512.0;   // per launch (~10 ms)
    cudaEvent_t a, b, c;
    CK(cudaEventCreate(&a));
    CK(cudaEventCreate(&b));
    CK(cudaEventCreate(&c));
    double best = 0.0, total_ms = 0.0;
    int launches = 0;
    CK(cudaEventRecord(c, h->stream));
    // back-to-back launches until `seconds` of device time have passed (at least four): the sustained figure is the
    // total flop over the total time, the burst figure the best single launch
    while (launches < 4 || total_ms < seconds * 1e3) {
        CK(cudaEventRecord(a, h->stream));
        k_dmma_peak<<<grid, 256, 0, h->stream>>>(h->d_acc8, iters);
        CK(cudaEventRecord(b, h->stream));
        CK(cudaEventSynchronize(b));
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, a, b));
        best = std::max(best, flops / (ms * 1e-3) / 1e12);
        CK(cudaEventElapsedTime(&ms, c, b));
        total_ms = ms;
        launches++;
    }
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    cudaEventDestroy(c);
    if (sustained) *sustained = flops * launches / (total_ms * 1e-3) / 1e12;
    if (burst) *burst = best;
    return KDSL_OK;
}

int kdsl_bench_fp64_dmma(kdsl_handle h, double *tflops) {
    if (!tflops) return fail(KDSL_ERR_INVALID_ARGUMENT, "tflops is null");
    return kdsl_bench_fp64_dmma_sustained(h, 0.0, nullptr, tflops);
}

/* ---- multi-GPU: NCCL sum of the observable accumulators --------------------------------------------
 * libnccl.so.2 is not a link-time dependency: it is opened on the first kdsl_comm_* call (KDSL_NCCL_LIB overrides
 * the name), so a host process that already carries an NCCL (e.g. PyTorch's) shares that copy. */
namespace {
struct NcclApi {
    void *dl = nullptr;
    decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
    decltype(&ncclCommInitRank) CommInitRank = nullptr;
    decltype(&ncclCommInitAll) CommInitAll = nullptr;
    decltype(&ncclCommDestroy) CommDestroy = nullptr;
    decltype(&ncclAllReduce) AllReduce = nullptr;
    decltype(&ncclGroupStart) GroupStart = nullptr;
    decltype(&ncclGroupEnd) GroupEnd = nullptr;
    decltype(&ncclGetErrorString) GetErrorString = nullptr;
    decltype(&ncclGetVersion) GetVersion = nullptr;
};
NcclApi g_nccl;

int nccl_load() {
    if (g_nccl.dl) return KDSL_OK;
    const char *name = getenv("KDSL_NCCL_LIB");
    void *dl = dlopen(name && *name ? name : "libnccl.so.2", RTLD_NOW | RTLD_LOCAL);
    if (!dl) return fail(KDSL_ERR_CUDA, "cannot open NCCL (%s): %s", name && *name ? name : "libnccl.so.2", dlerror());
#define KDSL_NCCL_SYM(f)                                                                  \
    g_nccl.f = reinterpret_cast<decltype(g_nccl.f)>(dlsym(dl, "nccl" #f));                \
    if (!g_nccl.f) { dlclose(dl); return fail(KDSL_ERR_CUDA, "NCCL symbol nccl" #f " missing"); }
    KDSL_NCCL_SYM(GetUniqueId) KDSL_NCCL_SYM(CommInitRank) KDSL_NCCL_SYM(CommInitAll) KDSL_NCCL_SYM(CommDestroy)
    KDSL_NCCL_SYM(AllReduce) KDSL_NCCL_SYM(GroupStart) KDSL_NCCL_SYM(GroupEnd) KDSL_NCCL_SYM(GetErrorString)
    KDSL_NCCL_SYM(GetVersion)
#undef KDSL_NCCL_SYM
    g_nccl.dl = dl;
    return KDSL_OK;
}
}  // namespace
#define NCK(call)                                                                                          \
    do {                                                                                                   \
        ncclResult_t r_ = (call);                                                                          \
        if (r_ != ncclSuccess)                                                                             \
            return fail(KDSL_ERR_CUDA, "%s failed: %s", #call, g_nccl.GetErrorString(r_));                 \
    } while (0)

int kdsl_comm_version(int *version) {
    int rc = nccl_load();
    if (rc) return rc;
    if (!version) return fail(KDSL_ERR_INVALID_ARGUMENT, "version is null");
    NCK(g_nccl.GetVersion(version));
    return KDSL_OK;
}

int kdsl_comm_unique_id(uint8_t *id) {
    int rc = nccl_load();
    if (rc) return rc;
    if (!id) return fail(KDSL_ERR_INVALID_ARGUMENT, "id is null");
    ncclUniqueId u;
    NCK(g_nccl.GetUniqueId(&u));
    static_assert(sizeof u == KDSL_COMM_ID_BYTES, "ncclUniqueId size");
    memcpy(id, &u, sizeof u);
    return KDSL_OK;
}

int kdsl_comm_init_rank(kdsl_handle h, int n_ranks, int rank, const uint8_t *id) {
    int rc = use_device(h);
    if (rc) return rc;
    if ((rc = nccl_load())) return rc;
    if (!id || n_ranks < 1 || rank < 0 || rank >= n_ranks) return fail(KDSL_ERR_INVALID_ARGUMENT, "bad rank / n_ranks / id");
    if (h->comm) return fail(KDSL_ERR_STATE, "handle already has a communicator");
    ncclUniqueId u;
    memcpy(&u, id, sizeof u);
    NCK(g_nccl.CommInitRank(&h->comm, n_ranks, u, rank));
    h->comm_rank = rank;
    h->comm_size = n_ranks;
    return KDSL_OK;
}

int kdsl_comm_init_all(int n, kdsl_handle *handles) {
    if (n < 1 || !handles) return fail(KDSL_ERR_INVALID_ARGUMENT, "need n >= 1 handles");
    int rc = nccl_load();
    if (rc) return rc;
    std::vector<int> devs(n);
    for (int i = 0; i < n; i++) {
        if (!handles[i]) return fail(KDSL_ERR_INVALID_ARGUMENT, "handle %d is null", i);
        if (handles[i]->comm) return fail(KDSL_ERR_STATE, "handle %d already has a communicator", i);
        devs[i] = handles[i]->device;
        for (int j = 0; j < i; j++)
            if (devs[j] == devs[i]) return fail(KDSL_ERR_INVALID_ARGUMENT, "handles %d and %d share device %d (one handle per GPU)", j, i, devs[i]);
    }
    std::vector<ncclComm_t> comms(n);
    NCK(g_nccl.CommInitAll(comms.data(), n, devs.data()));
    for (int i = 0; i < n; i++) {
        handles[i]->comm = comms[i];
        handles[i]->comm_rank = i;
        handles[i]->comm_size = n;
    }
    return KDSL_OK;
}

int kdsl_comm_destroy(kdsl_handle h) {
    if (!h) return fail(KDSL_ERR_INVALID_ARGUMENT, "null handle");
    if (!h->comm) return KDSL_OK;
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    g_nccl.CommDestroy(h->comm);
    h->comm = nullptr;
    h->comm_rank = -1;
    h->comm_size = 0;
    return KDSL_OK;
}

int kdsl_comm_info(kdsl_handle h, int *rank, int *n_ranks) {
    if (!h) return fail(KDSL_ERR_INVALID_ARGUMENT, "null handle");
    if (rank) *rank = h->comm_rank;
    if (n_ranks) *n_ranks = h->comm_size;
    return KDSL_OK;
}

/* one rank's share of the reduction (enqueue only) */
static int enqueue_acc_allreduce(kdsl_handle h) {
    CK(cudaSetDevice(h->device));
    k_reduce_acc<<<1, 1024, 0, h->stream>>>(h->S, h->d_acc8, (double)h->walker_sweeps);
    CK(cudaGetLastError());
    NCK(g_nccl.AllReduce(h->d_acc8, h->d_acc8, KDSL_N_ACC, ncclDouble, ncclSum, h->comm, h->stream));
    return KDSL_OK;
}

int kdsl_accumulators_allreduce(kdsl_handle h, double *out) {
    int rc = use_device(h);
    if (rc) return rc;
    if (!out) return fail(KDSL_ERR_INVALID_ARGUMENT, "out is null");
    if (!h->comm) return kdsl_accumulators(h, out, nullptr, nullptr);      // a job of one rank
    if ((rc = enqueue_acc_allreduce(h))) return rc;
    CK(cudaMemcpyAsync(out, h->d_acc8, KDSL_N_ACC * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return KDSL_OK;
}

int kdsl_group_accumulators_allreduce(int n, kdsl_handle *handles, double *out) {
    if (n < 1 || !handles || !out) return fail(KDSL_ERR_INVALID_ARGUMENT, "need n >= 1 handles and an output vector");
    for (int i = 0; i < n; i++)
        if (!handles[i] || !handles[i]->comm || handles[i]->comm_size != n)
            return fail(KDSL_ERR_STATE, "handle %d is not part of a %d-rank communicator (call kdsl_comm_init_all first)", i, n);
    NCK(g_nccl.GroupStart());
    for (int i = 0; i < n; i++) {
        int rc = enqueue_acc_allreduce(handles[i]);
        if (rc) { g_nccl.GroupEnd(); return rc; }
    }
    NCK(g_nccl.GroupEnd());
    CK(cudaSetDevice(handles[0]->device));
    CK(cudaMemcpyAsync(out, handles[0]->d_acc8, KDSL_N_ACC * sizeof(double), cudaMemcpyDeviceToHost, handles[0]->stream));
    for (int i = 0; i < n; i++) {
        CK(cudaSetDevice(handles[i]->device));
        CK(cudaStreamSynchronize(handles[i]->stream));
    }
    return KDSL_OK;
}

int kdsl_get_observables(kdsl_handle h, double *out, int allreduce) {
    int rc = use_device(h);
    if (rc) return rc;
    const DevState &S = h->S;
    if (!S.obs_on) return fail(KDSL_ERR_STATE, "no extra observables: call kdsl_set_observables first");
    if (!out) return fail(KDSL_ERR_INVALID_ARGUMENT, "out is null");
    const int n = 4 + 2 * S.nq;
    k_reduce_obs<<<(n + 127) / 128, 128, 0, h->stream>>>(S, h->d_obs);
    CK(cudaGetLastError());
    if (allreduce && h->comm) NCK(g_nccl.AllReduce(h->d_obs, h->d_obs, n, ncclDouble, ncclSum, h->comm, h->stream));
    CK(cudaMemcpyAsync(out, h->d_obs, n * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return KDSL_OK;
}

int kdsl_synchronize(kdsl_handle h) {
    int rc = use_device(h);
    if (rc) return rc;
    int n_sing = 0;
    CK(cudaMemcpyAsync(&n_sing, h->S.cnt + 3, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    CK(cudaGetLastError());
    if (n_sing > 0)
        return fail(KDSL_ERR_SINGULAR, "SingularException: tilde_U was singular in %d walker re-evaluation(s) since the last "
                    "kdsl_reset_accumulators (those walkers are frozen and carry KDSL_FLAG_SINGULAR)", n_sing);
    return KDSL_OK;
}

}  // extern "C"
