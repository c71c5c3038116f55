/*
 * kdsl.h -- C ABI of libkdsl.so: the B200-native (sm_100a) walker-batched VMC
 * sampling engine that replaces the hot path of hz-xiaxz/KagomeDSL.jl.
 *
 * One handle = one GPU = `n_walkers` independent Markov chains advanced in
 * lock step.  All pointers are HOST pointers unless stated otherwise; every
 * entry point returns an int status (0 = KDSL_OK, < 0 = error; the message is
 * available from kdsl_last_error()) and never throws.  The entry points are
 * the ones the reference's Julia methods would bind with `ccall`
 * (INTEGRATION.md shows the stubs); each cites the reference interface it
 * replaces, paths relative to the reference repository root.
 *
 * Conventions shared with the reference:
 *   - sites and orbital labels are 1-based; kappa[R] = 0 means "no particle of
 *     this species on site R", otherwise the label l of the particle
 *     (src/MonteCarlo.jl:17-28);
 *   - matrices are column-major (Julia layout): U is ns x N, W is ns x N with
 *     column l contiguous;
 *   - the bond list is Ham.nn in the reference's order
 *     (src/Hamiltonian.jl:447-451), pairs (i < j);
 *   - one "sweep" is ONE proposed spin exchange per walker
 *     (src/MonteCarlo.jl:538-607).
 * W is real FP64 (valid for B = 0, where the mean-field Hamiltonian is real
 * symmetric).
 */
#ifndef KDSL_H
#define KDSL_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct kdsl_handle_s *kdsl_handle;

/* status codes */
#define KDSL_OK 0
#define KDSL_ERR_INVALID_ARGUMENT (-1) /* ArgumentError / DimensionMismatch / BoundsError analogue */
#define KDSL_ERR_CUDA (-2)             /* CUDA runtime failure (no device, OOM, launch error) */
#define KDSL_ERR_SINGULAR (-3)         /* LinearAlgebra.SingularException analogue */
#define KDSL_ERR_STATE (-4)            /* call sequence error (e.g. sweep before a configuration is set) */

/* per-walker flag bits (kdsl_get_flags) */
#define KDSL_FLAG_SINGULAR 1   /* exact-zero / non-finite pivot in the last W re-evaluation */
#define KDSL_FLAG_NONFINITE 2  /* non-finite determinant ratio met in a proposal */
#define KDSL_FLAG_BAD_SITE 4   /* Sz ArgumentError analogue (site empty or doubly occupied) in measure */

/* indices into the double[KDSL_N_ACC] vector of kdsl_accumulators */
#define KDSL_ACC_WALKER_SWEEPS 0 /* number of Carlo.sweep! call-equivalents (walkers x sweeps)      */
#define KDSL_ACC_SUM_ACC 1       /* sum of the :acc observable (src/MonteCarlo.jl:548-589)          */
#define KDSL_ACC_SUM_OL 2        /* sum of the :OL observable (src/MonteCarlo.jl:632)               */
#define KDSL_ACC_SUM_OL2 3       /* sum of OL^2                                                      */
#define KDSL_ACC_N_OL 4          /* number of :OL samples                                            */
#define KDSL_ACC_N_REACH 5       /* sweeps that reached the refresh block (src/MonteCarlo.jl:594)   */
#define KDSL_ACC_N_REFRESH 6     /* walker re-evaluations of W executed                              */
#define KDSL_ACC_N_SINGULAR 7    /* re-evaluations that hit a singular tilde_U                       */
#define KDSL_N_ACC 8

/* indices into kdsl_timers (milliseconds of device time / launch counts per kernel class) */
#define KDSL_T_PROPOSE 0
#define KDSL_T_UPDATE 1
#define KDSL_T_REFRESH_GATHER 2
#define KDSL_T_REFRESH_INVERSE 3
#define KDSL_T_REFRESH_GEMM 4
#define KDSL_T_MEASURE 5
#define KDSL_N_TIMERS 6

int kdsl_version(void);
/* thread-local message of the last failing call on this thread */
const char *kdsl_last_error(void);
/* number of visible CUDA devices (0 and KDSL_ERR_CUDA when there is none: there is no CPU path) */
int kdsl_device_count(int *n);

/*
 * Create an engine on CUDA device `device`.
 * Replaces the state built by MC(params) / MC(Ham, kappa_up, kappa_down, W_up, W_down)
 * (src/MonteCarlo.jl:165-185, 209-234) from a Hamiltonian (src/Hamiltonian.jl:420-427):
 *   bonds  int32 [n_bonds][2]  Ham.nn, 1-based, reference order
 *   U_up   double [ns x n_up]  column-major Ham.U_up (real)
 *   U_dn   double [ns x n_dn]  column-major Ham.U_down
 * All walkers start with kappa = 0 and W = 0 like the reference (:179-182); a
 * configuration must be supplied with kdsl_set_config before sweeping.
 * Requires ns even (true for every DoubleKagome: ns = 3*n1*n2 with n1 even).
 */
int kdsl_create(kdsl_handle *out, int device, int ns, int n_up, int n_dn, int n_bonds,
                const int32_t *bonds, const double *U_up, const double *U_dn, int n_walkers);
/*
 * ComplexF64 engine for Hamiltonians with a Peierls flux B != 0 (complex Hermitian hopping matrix,
 * src/Hamiltonian.jl:201-204, 325-327; scripts/LL.jl): U_up / U_dn are column-major Matrix{ComplexF64}, i.e.
 * interleaved (re, im) doubles [ns x N x 2].  Same entry points as the real engine; kdsl_get_W / kdsl_set_W then
 * move interleaved complex matrices (2 * ns * N doubles), the acceptance uses abs2(ratio) (src/MonteCarlo.jl:581)
 * and O_L is real(OL) (src/Hamiltonian.jl:777).  The accepted-move updates are delayed in Woodbury form like in the
 * real engine ("update_variant" 2, default: kdsl_woodbury_c.cuh; "flush_every" + "flush_threshold" <= 24) or applied
 * immediately as the reference does ("update_variant" 0: rank-1 zgeru per accepted move); reevaluateW! inverts tilde_U
 * through its real 2N x 2N embedding on the blocked DMMA kernels ("inverse_variant" 0 / 4 / 5; 1 = unblocked complex
 * Gauss-Jordan).  At most 512 orbitals per species.  kdsl_is_complex reports the kind of a handle.
 */
int kdsl_create_c128(kdsl_handle *out, int device, int ns, int n_up, int n_dn, int n_bonds,
                     const int32_t *bonds, const double *U_up, const double *U_dn, int n_walkers);
int kdsl_is_complex(kdsl_handle h, int *is_complex);
int kdsl_destroy(kdsl_handle h);

/*
 * Set / get the configurations of all walkers: int32 [n_walkers][ns] each.
 * Replaces the assignments mc.kappa_up = ..., mc.kappa_down = ... of init_conf_qr!
 * (src/MonteCarlo.jl:326-357) and read_checkpoint! (:751-755), and the reads of
 * write_checkpoint (:715-719).  Every walker must be a Mott state (each site holds exactly one
 * particle) whose labels are a permutation of 1..N per species, else KDSL_ERR_INVALID_ARGUMENT
 * (the analogue of tilde_U's ArgumentError / BoundsError, src/MonteCarlo.jl:96-110).
 * Setting a configuration invalidates W; call kdsl_refresh next (deviation from the reference,
 * whose W stays zero after read_checkpoint!: SURVEY section 9 item 14).
 */
int kdsl_set_config(kdsl_handle h, const int32_t *kappa_up, const int32_t *kappa_dn);
int kdsl_get_config(kdsl_handle h, int32_t *kappa_up, int32_t *kappa_dn);

/*
 * Per-walker Xoshiro256++ states, uint64 [n_walkers][4] (Julia's Random.Xoshiro layout s0..s3).
 * Replaces ctx.rng (src/MonteCarlo.jl:546,552,569): the host owns seeding, the device owns the
 * streams afterwards.
 */
int kdsl_set_rng(kdsl_handle h, const uint64_t *states);
int kdsl_get_rng(kdsl_handle h, uint64_t *states);

/* ctx.sweeps (src/MonteCarlo.jl:595,630): the lock-step sweep counter shared by all walkers */
int kdsl_set_sweeps(kdsl_handle h, int64_t sweeps);
int kdsl_get_sweeps(kdsl_handle h, int64_t *sweeps);

/*
 * reevaluateW!(mc) for every walker (src/MonteCarlo.jl:55-66): W = U * inv(tilde_U(U, kappa)).
 * On return *n_singular (may be NULL) holds the number of walkers whose tilde_U was singular
 * (their KDSL_FLAG_SINGULAR is set and their W is left unchanged); the status is
 * KDSL_ERR_SINGULAR when that number is non-zero -- the SingularException of :59-60 / the
 * "QR-based configuration is singular" error of :401-405.
 */
int kdsl_refresh(kdsl_handle h, int *n_singular);

/*
 * n_sweeps x { Carlo.sweep!(mc, ctx); ctx.sweeps += 1 } for every walker, random numbers drawn
 * on the device from the walker's Xoshiro stream exactly in the reference's order
 * (src/MonteCarlo.jl:538-607).  If thermalization >= 0 the Carlo step loop is completed:
 * after each increment, if ctx.sweeps > thermalization, Carlo.measure! (:628-634) runs and
 * feeds the device-side accumulators.  Pass thermalization < 0 to never measure.
 * Asynchronous; errors of the device work surface at the next synchronising call.
 */
int kdsl_sweep(kdsl_handle h, int64_t n_sweeps, int64_t thermalization);

/*
 * Same as kdsl_sweep but the three random draws of each sweep are REPLAYED from host buffers
 * laid out [n_sweeps][n_walkers]:  r (the Float64 of :546), bond_idx (1-based index into nn
 * drawn at :552; read only if the gate :547 passes), pick (1-based index into maybe_update,
 * :569; may be NULL = all ones).  The walkers' RNG states are not touched.  This is the parity
 * interface, and the path a host that owns ctx.rng uses.
 */
int kdsl_replay(kdsl_handle h, int64_t n_sweeps, int64_t thermalization, const double *r,
                const int32_t *bond_idx, const int32_t *pick);

/*
 * getOL(mc, kappa_up, kappa_down) for every walker, now, regardless of the cadence test
 * (src/Hamiltonian.jl:762-778): ol double [n_walkers].  Does not touch the accumulators.
 * KDSL_ERR_INVALID_ARGUMENT if a walker holds an empty / doubly occupied site (Sz's ArgumentError).
 */
int kdsl_measure(kdsl_handle h, double *ol);

/*
 * The latest :OL sample of every walker taken by the cadence inside kdsl_sweep / kdsl_replay
 * (double [n_walkers]) and how many samples each walker has taken so far (int64 [n_walkers],
 * may be NULL).  This is what Carlo.measure!(mc, ctx) hands to measure!(ctx, :OL, .).
 */
int kdsl_last_OL(kdsl_handle h, double *ol, int64_t *n_samples);

/*
 * Sums over this handle's walkers since the last reset: double [KDSL_N_ACC] (indices above).
 * Optionally per-walker: acc_per_walker int64 [n_walkers] (accepted moves), ol_sum_per_walker
 * double [n_walkers]; either may be NULL.  Multi-GPU jobs add these vectors across ranks.
 */
int kdsl_accumulators(kdsl_handle h, double *out, int64_t *acc_per_walker,
                      double *ol_sum_per_walker);
int kdsl_reset_accumulators(kdsl_handle h);

/*
 * Extra observables taken with every :OL sample (not in the reference; the hook is register_evaluables,
 * src/MonteCarlo.jl:675-687).  After kdsl_set_observables(h, nq, cos_qr, sin_qr) with
 * cos_qr / sin_qr = double [nq][ns] holding cos(q . r_i) / sin(q . r_i) for the wave vectors of interest (nq may be 0),
 * every cadence measurement also accumulates, per walker,
 *   S(q) = |sum_i exp(i q r_i) Sz_i|^2 / ns      the longitudinal spin structure factor (from kappa alone), and
 *   Z_mu, OL * Z_mu, S(q) * Z_mu                 the weights that turn chain averages into |psi|^2 averages: the
 *                                                chain of src/MonteCarlo.jl:538-607 samples |psi|^2 / Z_mu, so
 *                                                <O>_psi = <O Z_mu> / <Z_mu>.
 * kdsl_get_observables returns the sums over this handle's walkers (over all ranks if allreduce != 0 and a communicator
 * exists): out[0] = samples, [1] = sum Z_mu, [2] = sum OL Z_mu, [3] = 0, [4 .. 4+nq) = sum S(q), [4+nq .. 4+2nq) =
 * sum S(q) Z_mu.  kdsl_reset_accumulators clears them.
 */
int kdsl_set_observables(kdsl_handle h, int nq, const double *cos_qr, const double *sin_qr);
int kdsl_get_observables(kdsl_handle h, double *out, int allreduce);

/* Copy one walker's W matrix to the host: spin 0 = up (ns x n_up), 1 = down (ns x n_dn) */
int kdsl_get_W(kdsl_handle h, int walker, int spin, double *out);
/* Overwrite one walker's W matrix (test hook for the update_W! known answers) */
int kdsl_set_W(kdsl_handle h, int walker, int spin, const double *in);

/*
 * update_W!(W, l, K, col_cache, row_cache) (src/MonteCarlo.jl:279-292) applied to the listed
 * walkers through the production rank-1 kernel: for m in 0..n_moves-1, walker[m]'s W_up is
 * updated with (l_up[m], K_up[m]) and its W_down with (l_dn[m], K_dn[m]) (1-based).  Walker ids
 * must be distinct.  kappa is not touched (test / benchmarking hook for the W-update kernel).
 */
int kdsl_update_W(kdsl_handle h, int n_moves, const int32_t *walker, const int32_t *l_up,
                  const int32_t *K_up, const int32_t *l_dn, const int32_t *K_dn);

/* Z(nn, kappa_up, kappa_down) (src/MonteCarlo.jl:460-474) of every walker:
 * zmu = the incrementally maintained value used by the proposals, zmu_recount (may be NULL) =
 * a full recount over the bond list; the two must agree. int32 [n_walkers] each */
int kdsl_get_Z(kdsl_handle h, int32_t *zmu, int32_t *zmu_recount);

/* per-walker KDSL_FLAG_* bits, int32 [n_walkers] */
int kdsl_get_flags(kdsl_handle h, int32_t *flags);

/*
 * Device-time profile.  kdsl_set_profiling(h, 1) brackets every kernel launch with CUDA events on
 * the engine's stream; kdsl_timers then returns, per kernel class, the summed device
 * milliseconds (ms[KDSL_N_TIMERS]) and launch counts (launches[KDSL_N_TIMERS]) since the last
 * kdsl_reset_timers, plus (any may be NULL) update_counts[2]: [0] accepted moves streamed by the
 * rank-1 W-update launches (update_variant 0), [1] walker flushes W0 += A B^T executed by the delayed
 * update launches (update_variant 1).  Launch counts are maintained even when profiling is off.
 */
int kdsl_set_profiling(kdsl_handle h, int enabled);
int kdsl_timers(kdsl_handle h, double *ms, int64_t *launches, int64_t *update_counts);
int kdsl_reset_timers(kdsl_handle h);

/* Tunables: "refresh_every" (0 = reference cadence n_occ); "update_variant" (2 = delayed rank-k updates in
 * Woodbury form, default; 0 = the reference's immediate rank-1 update streamed per accepted move);
 * "flush_every" / "flush_threshold" (Woodbury mode: sweeps between flush launches / pending updates that make a
 * walker due; sum <= 32 and small enough for the shared memory of the measurement kernel); "fuse_sweeps";
 * "update_ctas_per_sm", "update_cols_per_item" (rank-1 kernel tiling);
 * "inverse_variant": how reevaluateW! (src/MonteCarlo.jl:55-66) is computed -- 0 (default) / 6 = the one-kernel
 * re-evaluation k_reeval_fused when it applies (N <= 256 per species, ns <= 512), else gather + cluster inverse
 * k_inverse_cl (one matrix per thread-block cluster, 256 < N <= 512; 7 forces it) + DMMA product; 5 / 4 = gather +
 * one-CTA blocked implicit-pivoting inverse (with / without look-ahead) + DMMA product; 1 = simple cross-check kernels;
 * (8 = the cluster re-evaluation k_reeval_cl, one kernel for N <= 512: `make DEV=1` builds only, measured slower).  ComplexF64 engine: 0 (default) / 9 = complex
 * cluster inverse on split (re, im) planes + tensor-pipe product, 4 / 5 / 7 = blocked inverse of the real 2N x 2N
 * embedding, 1 = unblocked complex elimination; "flush_variant" 0 = tensor-pipe flush, 4 = FMA flush.
 * "inverse_cluster" (CTAs per matrix of the cluster kernels, 2..8), "inverse_row_slices", "inverse_tuning",
 * "gemm_variant", "fused_ctas": developer knobs.  The superseded variants (update_variant 1,
 * flush_variant 1, inverse_variant 2 / 3, gemm_variant 2 / 3) exist only in a `make DEV=1` build.
 * KDSL_ERR_INVALID_ARGUMENT if the name is unknown or the value is not available in this build. */
int kdsl_set_option(kdsl_handle h, const char *name, int64_t value);

/*
 * Device-side stopwatch on the engine's stream: kdsl_event_record(h, slot) enqueues a CUDA event
 * (slot in 0..15); kdsl_event_elapsed returns the milliseconds between two recorded slots after
 * synchronising on the later one.  Used by bench.py to time whole steps on the device.
 */
int kdsl_event_record(kdsl_handle h, int slot);
int kdsl_event_elapsed(kdsl_handle h, int slot_start, int slot_stop, double *ms);

/* Measured FP64 tensor-pipe (DMMA m8n8k4) peak of this device in TFLOP/s: the roofline denominator of the
 * W re-evaluation kernels (MEASURED_PEAKS.json holds no FP64 figure). Runs a ~10 ms register-only probe. */
int kdsl_bench_fp64_dmma(kdsl_handle h, double *tflops);
/* The same probe launched back to back for `seconds` (0..30) of device time: *sustained = total flop / total time
 * (the figure to use for a kernel timed inside a long step), *burst = the best single ~10 ms launch; either may be NULL. */
int kdsl_bench_fp64_dmma_sustained(kdsl_handle h, double seconds, double *sustained, double *burst);

/* Block until all device work of this handle is complete; returns the first deferred error: a CUDA error, or
 * KDSL_ERR_SINGULAR when a re-evaluation inside kdsl_sweep / kdsl_replay met a singular tilde_U since the last
 * kdsl_reset_accumulators (the SingularException that the reference re-throws out of sweep!, src/MonteCarlo.jl:596-603).
 * Such a walker is frozen: it carries KDSL_FLAG_SINGULAR, proposes no further moves and feeds no :OL samples; the other
 * walkers of the batch are not affected.  kdsl_sweep / kdsl_replay are asynchronous and return KDSL_OK themselves. */
int kdsl_synchronize(kdsl_handle h);

/*
 * Multi-GPU: walkers are independent Markov chains, so a job shards them over one handle per GPU and
 * the ONLY exchange of the path is the sum of the observable accumulators per bin -- what Carlo.jl does across MPI
 * ranks when it merges the runs' :acc / :OL observables (src/MonteCarlo.jl:548-589, 632).  No W or kappa ever leaves
 * its GPU.  The reduction is one ncclAllReduce(sum) of the KDSL_N_ACC doubles, enqueued on the engine's stream behind
 * the device-side reduction over the handle's own walkers.  libnccl.so.2 is opened with dlopen on first use
 * (environment KDSL_NCCL_LIB overrides the file name); it is not needed for single-GPU use.
 *
 *   one process, several GPUs (the Julia host: one handle per device, one thread):
 *       kdsl_comm_init_all(n, handles)                      ncclCommInitAll over the handles' devices
 *       kdsl_group_accumulators_allreduce(n, handles, out)  grouped all-reduce; out = global sums
 *   one process per GPU (MPI-style launchers):
 *       rank 0: kdsl_comm_unique_id(id); ship the KDSL_COMM_ID_BYTES bytes to the other ranks by any means
 *       every rank: kdsl_comm_init_rank(h, n_ranks, rank, id)   (collective)
 *       every rank: kdsl_accumulators_allreduce(h, out)         (collective; without a communicator = kdsl_accumulators)
 * kdsl_comm_destroy is implied by kdsl_destroy.
 */
#define KDSL_COMM_ID_BYTES 128
int kdsl_comm_version(int *version);                      /* ncclGetVersion of the library that was opened */
int kdsl_comm_unique_id(uint8_t *id);
int kdsl_comm_init_rank(kdsl_handle h, int n_ranks, int rank, const uint8_t *id);
int kdsl_comm_init_all(int n, kdsl_handle *handles);
int kdsl_comm_info(kdsl_handle h, int *rank, int *n_ranks);  /* rank = -1, n_ranks = 0 without a communicator */
int kdsl_comm_destroy(kdsl_handle h);
int kdsl_accumulators_allreduce(kdsl_handle h, double *out);
int kdsl_group_accumulators_allreduce(int n, kdsl_handle *handles, double *out);

/* Handle geometry: out[0..5] = ns, n_up, n_dn, n_bonds, n_walkers, n_occ (= min(n_up, n_dn),
 * src/MonteCarlo.jl:594) */
int kdsl_info(kdsl_handle h, int64_t *out);

#ifdef __cplusplus
}
#endif
#endif /* KDSL_H */
