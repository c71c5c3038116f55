#!/usr/bin/env python
"""bench.py -- walker-sweeps/sec of the VMC sampling path at the 432-site pi-flux DSL
(BASELINE.json metric), one process per GPU.

  python bench.py --gpus N --steps K --warmup W            (N > 1: launched by torchrun)
  python bench.py --impl reference --gpus N --steps K --warmup W

A "step" is one bin of the hot path for every walker: n_occ = 216 lock-step Carlo sweeps (each ONE
proposal per walker, src/MonteCarlo.jl:538-607, including the accepted-move W updates and the periodic
re-evaluation of W) followed by one O_L measurement (src/MonteCarlo.jl:628-634).
Workload = BASELINE.json configs[2] restricted to what one GPU holds: 12x12 DoubleKagome (432 sites),
pi-flux, PBC, antiPBC=(true,false), N_up = N_down = 216, 4096 walkers per GPU (weak scaling:
32768 walkers on 8 GPUs).  Each walker's W is 1.49 MB, 6.1 GB per GPU >> the 126 MB L2, so every
timed iteration streams its inputs from HBM (no L2 flush needed).

The `--impl reference` arm times the CPU restatement of the reference's algorithm (oracle/, ComplexF64 W
like the reference) on all host cores: the reference itself is Julia and cannot run in this image.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

UNIT = "walker-sweeps/s"


def metric_name(args):
    """BASELINE.json's metric, named after the lattice actually run (432 sites = the headline)"""
    ns = 3 * args.lattice * args.lattice
    kind = ("zeroflux" if args.flux == "zero" else "piflux") + ("_peierlsB" if args.B != 0.0 else "")
    return f"walker_sweeps_per_sec_{ns}site_{kind}_dsl"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--lattice", type=int, default=12, help="n1 = n2 (12 -> 432 sites)")
    ap.add_argument("--flux", default="pi", choices=["pi", "zero"], help="mean-field ansatz: pi-flux DSL or the zero-flux tables of scripts/zero_flux.jl")
    ap.add_argument("--B", type=float, default=0.0, help="Peierls flux (B != 0 selects the ComplexF64 engine, scripts/LL.jl)")
    ap.add_argument("--options", default="", help="engine options name=value,... (developer knob)")
    ap.add_argument("--no-carlo", action="store_true", help="skip the call-per-sweep e2e_carlo legs")
    ap.add_argument("--walkers-per-gpu", type=int, default=4096)
    ap.add_argument("--thermalization", type=int, default=-1, help="untimed sweeps before warm-up (default 10*ns)")
    ap.add_argument("--seed", type=int, default=1234)
    ap.add_argument("--cpu-baseline-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


def make_config(args, world):
    """the `config` object of the JSON line: identical for the GPU arm and the --impl reference arm of the same
    command line (how many walkers each arm actually advanced is in `walkers_total` / `cpu_baseline.sample`)"""
    n, nw = args.lattice, args.walkers_per_gpu
    ns = 3 * n * n
    flux = "zero-flux" if args.flux == "zero" else "pi-flux DSL"
    return {"workload": f"{n}x{n} DoubleKagome ({ns} sites) {flux}, PBC, antiPBC=(true,false), half filling, B = {args.B:g}; "
                        f"GPU arm: {nw} walkers per GPU; CPU arm: bounded sample, one walker per host core",
            "sweeps_per_step": ns // 2, "walkers_per_gpu": nw, "n_gpus": world,
            "l2": "inputs_exceed_l2 (W working set %.1f GB per GPU)" % (nw * ns * ns * 8 * (2 if args.B != 0.0 else 1) / 1e9)
                  if nw * ns * ns * 8 > 126e6 else "W working set %.0f MB per GPU is L2 resident; L2 is not flushed (the chain's own state is the input)" % (nw * ns * ns * 8 / 1e6),
            "rng": "Xoshiro256++ per walker", "refresh": "reference cadence n_occ"}


# ----------------------------------------------------------------------------------------------
# CPU arm: the oracle (restated reference algorithm, ComplexF64 like the reference) on all host cores
# ----------------------------------------------------------------------------------------------
def cpu_chain_rate(ham, kup, kdn, seed, sweeps_per_step, steps, warmup, n_threads, dtype="c128"):
    """Every thread advances its OWN reference walker through `warmup` + `steps` steps of `sweeps_per_step` Carlo sweeps
    (thermalization 0, i.e. O_L measured every n_occ sweeps) and times its own `steps` steps: no join between steps,
    the threads only start together.  Returns (sum over threads of sweeps / own time, mean seconds per step, E/site)."""
    from oracle import oracle as O
    bonds = np.asarray(ham.nn, dtype=np.int32)
    mcs = []
    for t in range(n_threads):
        mc = O.MC(bonds, ham.U_up, ham.U_down, dtype)
        mc.set_kappa(kup, kdn)
        mc.reevaluateW()
        mcs.append((mc, O.Xoshiro.from_seed(seed + 7919 * t), np.zeros(4)))
    dts = [0.0] * n_threads
    go = threading.Barrier(n_threads)

    def work(t):
        mc, g, st = mcs[t]
        go.wait()
        for _ in range(warmup):
            mc.run(g, sweeps_per_step, 0)
        t0 = time.perf_counter()
        for _ in range(steps):
            mc.run(g, sweeps_per_step, 0, stats=st)
        dts[t] = time.perf_counter() - t0
    threads = [threading.Thread(target=work, args=(t,)) for t in range(n_threads)]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    rate = sum(sweeps_per_step * steps / dt for dt in dts)
    tot = sum(m[2] for m in mcs)
    e = tot[1] / tot[3] / len(kup) if tot[3] else float("nan")
    return rate, float(np.mean(dts)) / max(steps, 1), e


def cpu_baseline_sample(ham, kup, kdn, seed, n_occ, seconds, cores, dtype):
    """bounded sample (about `seconds` of wall time on all cores) of the CPU arm: sized from a two-bin probe"""
    probe, _, _ = cpu_chain_rate(ham, kup, kdn, seed, n_occ, 2, 0, cores, dtype)
    bins = int(max(4, round(seconds * probe / cores / n_occ)))
    rate, spb, e = cpu_chain_rate(ham, kup, kdn, seed, n_occ, bins, 1, cores, dtype)
    return rate, bins, e


def setup_problem(args):
    import kagomedsl.jl_b200 as kd
    n = args.lattice
    lat = kd.DoubleKagome(1.0, n, n, (True, True), (True, False))
    ns = kd.ns(lat)
    li, lx = (kd.zero_link_in, kd.zero_link_inter) if args.flux == "zero" else (kd.pi_link_in, kd.pi_link_inter)
    ham = kd.Hamiltonian(ns // 2, ns // 2, lat, link_in=li, link_inter=lx, B=args.B)
    kup, kdn = kd.init_conf_qr(ham, ns, ns // 2)
    return kd, lat, ham, ns, kup, kdn


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    kd, lat, ham, ns, kup, kdn = setup_problem(args)
    n_occ = ns // 2
    cores = os.cpu_count() or 1
    from oracle import oracle as O
    O.build()
    # one reference step = every core advances its own walker by `bins` bins of n_occ sweeps (sized for ~1 s per step at
    # 432 sites); the threads never wait for each other between steps
    bins = max(1, int(round(16 * (432.0 / ns) ** 3)))
    sweeps = bins * n_occ
    rate, sec_per_step, e = cpu_chain_rate(ham, kup, kdn, args.seed, sweeps, args.steps, args.warmup, cores, "c128")
    rate64, _, _ = cpu_chain_rate(ham, kup, kdn, args.seed, sweeps, max(1, args.steps // 4), 1, cores, "f64") if args.B == 0.0 else (None, 0, 0)
    line = {
        "impl": "reference", "metric": metric_name(args), "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * sec_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "c128", "data": "synthetic",
        "config": make_config(args, args.gpus),
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{cores} walkers (one per host core, each thread times its own steps: no join between steps) x "
                                   f"{sweeps} sweeps per step x {args.steps} steps, ComplexF64 W like the reference",
                         "value_f64": rate64,
                         "note": "CPU restatement (oracle/) of the Julia reference: the reference itself cannot run here (no Julia). "
                                 "No Dict / allocation / dispatch overhead of the Julia code, but no MKL either; value_f64 = the same "
                                 "chain with real FP64 W (what the GPU arm computes for B = 0)"},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "E_per_site": e,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi polled every 50 ms from before the warm-up; every line is stamped on arrival, and only the samples
    that fall inside the timed window are reported (all of them taken under load)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.p = None
        self.lines = []
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50"],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, bufsize=1)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.p = None

    def _read(self):
        for ln in self.p.stdout:
            self.lines.append((time.time(), ln))

    def stop(self, t0, t1):
        """samples with arrival time in [t0, t1] (the timed region); if the region was too short for three samples,
        the window is widened backwards over the warm-up steps (same load) and the JSON says so"""
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        return self.summarise(list(self.lines), t0, t1, getattr(self, "t_load", 0.0))

    @staticmethod
    def summarise(lines, t0, t1, t_load):
        """lines: (arrival time, csv line) pairs -> the `clocks` object of the JSON line"""
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

        def parse(lo, hi):
            sm, mx, pw, reasons = [], [], [], set()
            for ts, ln in lines:
                if ts < lo or ts > hi:
                    continue
                f = [x.strip() for x in ln.split(",")]
                if len(f) < 7:
                    continue
                try:
                    sm.append(float(f[0])); mx.append(float(f[1]))
                except ValueError:
                    continue
                try:
                    pw.append(float(f[2]))
                except ValueError:
                    pass
                for nme, v in zip(names, f[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nme)
            return sm, mx, pw, reasons
        window = "timed region"
        sm, mx, pw, reasons = parse(t0, t1 + 0.06)
        if len(sm) < 3:
            window = "warm-up + timed region (the timed region alone is shorter than three 50 ms samples)"
            sm, mx, pw, reasons = parse(t_load, t1 + 0.06)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w": float(np.median(pw)) if pw else None, "samples": len(sm), "window": window, "reasons": sorted(reasons)}


def run_ours(args):
    import torch
    import torch.distributed as dist
    rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        # NCCL prints its version banner on STDOUT when the communicator comes up: keep stdout for the ONE JSON line
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
            t0_ = torch.zeros(1, device=f"cuda:{local}")
            dist.all_reduce(t0_)                         # forces communicator creation now
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def allmax(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=f"cuda:{local}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    kd, lat, ham, ns, kup, kdn = setup_problem(args)
    from kagomedsl.jl_b200 import _lib
    n_occ = ns // 2
    nw = args.walkers_per_gpu
    K, W = args.steps, args.warmup
    cplx = args.B != 0.0
    li, lx = (kd.zero_link_in, kd.zero_link_inter) if args.flux == "zero" else (kd.pi_link_in, kd.pi_link_inter)
    mc = kd.MC({"n1": args.lattice, "n2": args.lattice, "PBC": (True, True), "antiPBC": (True, False), "N_up": ns // 2,
                "N_down": ns // 2, "n_walkers": nw, "device": local, "link_in": li, "link_inter": lx, "B": args.B})
    # public-API initialisation; walker streams are disjoint across ranks
    mc.load_configuration(kup, kdn, kd.walker_states(args.seed, nw, first_walker=rank * nw))
    eng = mc.engine
    for kv in filter(None, args.options.split(",")):
        name, val = kv.split("=")
        eng.set_option(name.strip(), int(val))
    if world > 1:
        # the per-bin reduction of the observable accumulators runs INSIDE libkdsl (ncclAllReduce on the engine's
        # stream, communicator made by kdsl_comm_init_rank); torch.distributed only carries the 128-byte unique id,
        # the barriers around the timed region and the max-over-ranks of the device time
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        try:
            kd.dist.init_comm(eng)
            eng.accumulators_allreduce()
        finally:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)
    ctx = kd.MCContext({"thermalization": 0, "seed": args.seed})
    therm = args.thermalization if args.thermalization >= 0 else 10 * ns
    therm = (therm // n_occ) * n_occ                    # keep the bin phase aligned
    sampler = ClockSampler(local) if rank == 0 else None
    eng.sweeps = 0
    eng.sweep(therm, -1)
    ctx.sweeps = therm
    eng.synchronize()
    if sampler:
        sampler.t_load = time.time()                    # from here on the GPU runs the timed workload (warm-up steps)
    for _ in range(W):                                  # warm-up steps through the public API
        kd.run_(mc, ctx, n_occ)
        eng.accumulators_allreduce()
    eng.synchronize()
    eng.reset_accumulators()
    eng.reset_timers()
    eng.set_profiling(True)                             # CUDA events around every launch on the engine's stream

    # ---- timed region: K steps, device-timed on the launching stream, max over ranks.  One step = one bin:
    #      n_occ lock-step sweeps + the O_L measurement + the per-bin sum of the accumulators over all ranks ----
    barrier()
    t_w0 = time.time()
    reduce_ms = 0.0
    eng.event_record(0)
    for _ in range(K):
        kd.run_(mc, ctx, n_occ)
        eng.event_record(2)
        acc_global = eng.accumulators_allreduce()        # device reduction + ncclAllReduce (N > 1) + D2H of 8 doubles
        eng.event_record(3)
        reduce_ms += eng.event_elapsed_ms(2, 3)
    eng.event_record(1)
    eng.synchronize()
    t_w1 = time.time()
    barrier()
    ms = allmax(eng.event_elapsed_ms(0, 1))
    clocks = sampler.stop(t_w0, t_w1) if sampler else None
    tm = eng.timers()
    eng.set_profiling(False)
    res = kd.accumulators(mc)                            # (global sums once more, as a dict)
    assert abs(res["sum_OL"] - acc_global[_lib.ACC_SUM_OL]) <= 1e-9 * abs(res["sum_OL"])
    total_walkers = nw * world
    value = total_walkers * n_occ * K / (ms * 1e-3)
    launches = sum(v["launches"] for v in tm.values()) + K      # + one k_reduce_acc per step
    B_acc = 16 * ns * ns * (2 if cplx else 1)            # bytes to read+write both W matrices of a walker once
    upd = tm["update"]
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 (of fallback)"
    traffic = {}
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
    except Exception:
        pass
    tkey = lambda k: traffic.get(f"{k}_{ns}", traffic.get(k) if ns == 432 else None)
    moves_total = res["sum_acc"] / world                 # accepted moves on this rank (all folded by flush or refresh)
    woodbury = upd.get("flushes", 0) > 0
    # Rooflines.  (1) The HBM-bound W update.  Production path: k_flush_wb folds the >= 16 pending rank-1 updates of a
    # walker into W with ONE read+write pass (delayed Sherman-Morrison, Woodbury form): algorithmic bytes per launch =
    # walkers flushed x 16 ns^2.  The reference's immediate update (and the ComplexF64 engine's) moves 16 ns^2 (32 ns^2)
    # per ACCEPTED move.
    if woodbury:
        n_units, unit_name = upd["flushes"], "walker flush"
        kname = ("k_flush_dmma2_c (delayed rank-k update of the ComplexF64 W: four real DMMAs per complex block product, two CTAs per SM, 128-bit streaming)" if cplx else
                 "k_flush_wb (delayed rank-k Sherman-Morrison update of W, DMMA, 128-bit streaming)")
        n_launch = max(upd["launches"] // 2, 1)
    else:
        n_units, unit_name = upd.get("moves", 0), "accepted move"
        kname = "k_update_c / k_update_ldg (immediate rank-1 update per accepted move, 128-bit streaming)"
        n_launch = max(upd["launches"], 1)
    achieved = n_units * B_acc / (upd["ms"] * 1e-3) / 1e9 if upd["ms"] > 0 else 0.0
    roofline_update = {"bound": "hbm", "kernel": kname, "achieved": achieved, "peak": peak, "peak_source": peak_src, "unit": "GB/s",
                       "frac": achieved / peak, "traffic": tkey("k_flush_wb_dram_bytes_per_walker_flush") if woodbury else None,
                       "units_per_launch": n_units / n_launch, "unit": "GB/s", "algorithmic_bytes_per_unit": B_acc, "unit_name": unit_name,
                       "avg_launch_us": 1e3 * upd["ms"] / n_launch, "kernel_share_of_step": upd["ms"] / ms if ms > 0 else None,
                       "rank1_equivalent_GBs": moves_total * B_acc / (upd["ms"] * 1e-3) / 1e9 if upd["ms"] > 0 else None,
                       "note": "traffic is per unit (ncu dram bytes of the captured launch / units it processed); rank1_equivalent = "
                               "accepted moves x 16 ns^2 / update time, the traffic the reference's per-move update would need"}
    # (2) FP64 tensor-pipe rooflines of the W re-evaluation.  Denominator: DMMA rate sustained for >= 1 s, clocks sampled
    dsampler = ClockSampler(local) if rank == 0 else None
    t_d0 = time.time()
    dmma_peak, dmma_burst = eng.fp64_dmma_peak_tflops(1.0)
    t_d1 = time.time()
    dclocks = dsampler.stop(t_d0, t_d1) if dsampler else None
    n_refresh = res["n_refresh"] / world
    Nh = ns // 2
    cmul = 4.0 if cplx else 1.0                          # real flop per complex multiply-add pair
    dsrc = ("FP64 DMMA rate sustained over 1 s by kdsl_bench_fp64_dmma_sustained in this run (MEASURED_PEAKS.json has no FP64 "
            "figure); clocks during the probe in dmma_probe")

    # algorithmic flop of one walker refresh (both species) = the reference's inverse + full product
    flop_survey = cmul * 2.0 * (2.0 * Nh ** 3 + 2.0 * ns * Nh ** 2)

    def tensor_roofline(kernel, flop_executed, t_ms, n_launch, traffic_key, flop_algorithmic):
        t = t_ms * 1e-3
        ach = n_refresh * flop_algorithmic / t / 1e12 if t > 0 else 0.0
        exe = n_refresh * flop_executed / t / 1e12 if t > 0 else 0.0
        return {"bound": "tensor", "kernel": kernel, "achieved": exe, "peak": dmma_peak, "peak_source": dsrc, "unit": "TFLOP/s",
                "frac": exe / dmma_peak if dmma_peak > 0 else None,
                "frac_meaning": "EXECUTED flop / sustained DMMA peak (how busy the tensor pipe is); the contract figure "
                                "(SURVEY 8(d) algorithmic flop of the reference's inverse + full product) is in achieved_contract / frac_contract",
                "achieved_contract": ach, "frac_contract": ach / dmma_peak if dmma_peak > 0 else None,
                "traffic": tkey(traffic_key), "algorithmic_flop_per_walker_refresh": flop_algorithmic,
                "executed_flop_per_walker_refresh": flop_executed, "walker_refreshes_per_step": n_refresh / max(K, 1),
                "avg_launch_us": 1e3 * t_ms / max(n_launch, 1), "kernel_share_of_step": t_ms / ms if ms > 0 else None}

    fused = tm["refresh_gemm"]["launches"] == 0 and tm["refresh_inverse"]["launches"] > 0
    cands = [roofline_update]
    roofline_inverse = roofline_gemm = None
    if fused:
        # k_reeval_fused: Gauss-Jordan on [tilde_U^T | V^T] (N x (N + M), M = ns - N unoccupied sites): step k touches the
        # N rows of the N - k - 1 + M unfinished columns -> executed flop per matrix = sum_k 2 N (N - k - 1 + M) ~ 3 N^3
        # (the unit rows and the explicit inverse are never computed)
        M_un = ns - Nh
        flop_exec = 2.0 * sum(2.0 * Nh * (Nh - k - 1 + M_un) for k in range(Nh))
        roofline_inverse = tensor_roofline(
            "k_reeval_fused (reevaluateW!: blocked Gauss-Jordan on [tilde_U^T | V^T] with look-ahead, FP64 DMMA, writes W directly)",
            flop_exec, tm["refresh_inverse"]["ms"], tm["refresh_inverse"]["launches"] // 2,
            "k_reeval_fused_dram_bytes_per_walker_refresh", flop_survey)
        cands.append(roofline_inverse)
    elif tm["refresh_inverse"]["launches"] > 0:
        # ComplexF64: complex elimination on split (re, im) planes (k_inverse_cl_c): N^3 complex multiply-adds per matrix
        iname = ("k_inverse_cl_c (ComplexF64: blocked Gauss-Jordan inverse of tilde_U, one matrix per thread-block cluster, four real DMMAs per complex block product)"
                 if cplx else
                 "k_inverse_cl / k_inverse_v5 / k_inverse_v4 (batched blocked Gauss-Jordan inverse of tilde_U, FP64 DMMA trailing update; one cluster per matrix for N > 256)")
        roofline_inverse = tensor_roofline(iname, cmul * 2.0 * 2.0 * Nh ** 3, tm["refresh_inverse"]["ms"], 2 * K,
                                           "k_inverse_cl_c_dram_bytes_per_walker_refresh" if cplx else "k_inverse_v4_dram_bytes_per_matrix",
                                           cmul * 2.0 * 2.0 * Nh ** 3)
        # (the rows of W on occupied sites are unit vectors and are written, not computed: the kernel executes
        #  2 (ns - N) N^2 flops per matrix where the reference's full product has 2 ns N^2)
        roofline_gemm = tensor_roofline("k_gemm_W_dmma_c (ComplexF64: W = U inv(tilde_U) on the unoccupied rows, straight from the split planes)" if cplx
                                        else "k_gemm_W_dmma (W = U inv(tilde_U) on the unoccupied rows)",
                                        cmul * 2.0 * 2.0 * (ns - Nh) * Nh ** 2, tm["refresh_gemm"]["ms"], 2 * K,
                                        "k_gemm_W_dmma_dram_bytes_per_matrix", cmul * 2.0 * 2.0 * ns * Nh ** 2)
        cands += [roofline_inverse, roofline_gemm]
    resident = tm["update"]["launches"] == 0 and tm["refresh_inverse"]["launches"] == 0 and tm["propose"]["launches"] > 0
    if resident:
        # k_resident (small lattices): the walker lives in shared memory for the whole call, so neither HBM nor the tensor
        # pipe bounds it ("for ns = 108 the update is SMEM-resident: report it as a fraction of SMEM throughput").
        # Shared-memory bytes of the arithmetic: an accepted move reads + writes the compact W of both species
        # (2 x 2 N_up N_dn doubles), a re-evaluation streams N x ns doubles per species and pivot step (read + write,
        # shrinking tilde_U^T block: ~ N (ns - N/2) on average).
        t = tm["propose"]["ms"] * 1e-3
        upd_b = moves_total * 2.0 * 2.0 * Nh * (ns - Nh) * 8.0
        ref_b = n_refresh * 2.0 * Nh * 2.0 * Nh * (ns - Nh / 2.0) * 8.0
        smem_peak = 148 * 128 * 1.965                     # GB/s: 128 B/clk/SM at the 1965 MHz the run holds
        ach = (upd_b + ref_b) / t / 1e9 if t > 0 else 0.0
        hbm_b = K * nw * 2.0 * (ns * ns * 8.0)            # W of both species loaded and stored once per walker and launch
        roofline_resident = {
            "bound": "smem", "kernel": "k_resident (whole call per walker in one persistent kernel: proposals, rank-1 updates, "
            "re-evaluation and O_L in shared memory)", "achieved": ach, "peak": smem_peak, "unit": "GB/s", "frac": ach / smem_peak,
            "peak_source": "148 SMs x 128 B/clk x 1.965 GHz shared-memory bandwidth (nominal)",
            "note": "latency bound: one walker per CTA, barriers at accepted moves / pivots; the figure counts only the shared-memory "
                    "bytes of the W update and re-evaluation arithmetic",
            "hbm_GBs": hbm_b / t / 1e9 if t > 0 else None, "hbm_frac": hbm_b / t / 1e9 / peak if t > 0 else None,
            "rank1_equivalent_GBs": moves_total * B_acc / t / 1e9 if t > 0 else None, "traffic": None,
            "avg_launch_us": 1e3 * tm["propose"]["ms"] / max(tm["propose"]["launches"], 1),
            "kernel_share_of_step": tm["propose"]["ms"] / ms if ms > 0 else None}
        cands = [roofline_resident]
    # `roofline` = the kernel with the largest share of the timed step
    roofline = max(cands, key=lambda r: r["kernel_share_of_step"] or 0.0)

    # ---- e2e: same work through the public API with HOST buffers (replayed proposal stream) ----
    e2e = None
    if not args.no_e2e:
        Ke = max(1, min(K, 20))
        rng = np.random.default_rng(args.seed + 17 * rank)
        r_host = torch.empty((Ke, n_occ, nw), dtype=torch.float64, pin_memory=True)
        b_host = torch.empty((Ke, n_occ, nw), dtype=torch.int32, pin_memory=True)
        r_host.copy_(torch.from_numpy(rng.random((Ke, n_occ, nw))))
        b_host.copy_(torch.from_numpy(rng.integers(1, len(ham.nn) + 1, size=(Ke, n_occ, nw)).astype(np.int32)))
        r_np, b_np = r_host.numpy(), b_host.numpy()
        eng.sweeps = ctx.sweeps
        eng.replay(r_np[0], b_np[0], thermalization=0)      # warm the path once (untimed)
        eng.last_OL()
        barrier()
        t0 = time.perf_counter()
        for k in range(Ke):
            eng.replay(r_np[k], b_np[k], thermalization=0)  # H2D of this step's proposal stream + n_occ sweeps + measure
            ol, n_ol = eng.last_OL()                        # D2H of the step's result
            acc_vec = eng.accumulators_allreduce()          # per-bin sum over all ranks
        barrier()
        dt = allmax(time.perf_counter() - t0)
        ctx.sweeps = eng.sweeps
        e2e = {"value": total_walkers * n_occ * Ke / dt, "unit": UNIT, "steps": Ke,
               "h2d_bytes_per_step": int(n_occ * nw * (8 + 4)), "d2h_bytes_per_step": int(nw * 16 + 64),
               "mode": "kdsl_replay: host supplies (r, bond) for every proposal from pinned memory; O_L, counters and the "
                       "all-rank accumulator sums copied back every step"}

    # ---- e2e_carlo: the call-per-sweep protocol a Carlo.jl run loop uses (sweep!; ctx.sweeps += 1; measure!), host-timed:
    #      sweeps_per_call = 1 is the reference's granularity (one kdsl_sweep + one :acc read-back per sweep),
    #      sweeps_per_call = n_occ lets one Carlo sweep stand for a whole bin ----
    e2e_carlo = None
    if not args.no_carlo and not args.no_e2e:
        e2e_carlo = {}
        for spc, bins in ((1, 2), (n_occ, 5)):
            mc.sweeps_per_call = spc
            cctx = kd.MCContext({"thermalization": 0, "seed": args.seed})
            cctx.sweeps = eng.sweeps // spc
            mc.sync_counters()
            calls = bins * n_occ // spc
            kd.step_(mc, cctx)                              # untimed first call
            barrier()
            t0 = time.perf_counter()
            for _ in range(calls):
                kd.step_(mc, cctx)
            eng.synchronize()
            barrier()
            dt = allmax(time.perf_counter() - t0)
            e2e_carlo[f"sweeps_per_call_{spc}"] = {"value": total_walkers * calls * spc / dt, "unit": UNIT, "calls": calls,
                                                   "us_per_call": 1e6 * dt / calls, "n_OL_recorded": len(cctx.measurements.get("OL", []))}
        mc.sweeps_per_call = 1
        ctx.sweeps = eng.sweeps

    # ---- the reference-style immediate rank-1 W update (update_W!, src/MonteCarlo.jl:279-292) timed alone:
    #      every walker applies one move; 16 ns^2 algorithmic bytes per move ----
    eng.set_option("update_variant", 0)
    eng.reset_timers()
    eng.set_profiling(True)
    wl = np.arange(nw, dtype=np.int32)
    one = np.ones(nw, dtype=np.int32)
    for rep in range(4):
        eng.update_W(wl, one * (1 + rep), one * (3 + rep), one * (2 + rep), one * (5 + rep))
    t1 = eng.timers()["update"]
    eng.set_profiling(False)
    r1 = t1["moves"] * B_acc / (t1["ms"] * 1e-3) / 1e9 if t1["ms"] > 0 else 0.0
    roofline_rank1 = {"bound": "hbm", "kernel": "k_update_ldg / k_update_c (immediate rank-1 update, one move per walker, timed alone)",
                      "achieved": r1, "peak": peak, "peak_source": peak_src, "unit": "GB/s", "frac": r1 / peak,
                      "moves_per_launch": t1["moves"] / max(t1["launches"], 1), "algorithmic_bytes_per_move": B_acc,
                      "note": "peak = the driver's measured COPY bandwidth (two address streams); this kernel is an in-place "
                              "read-modify-write over one stream and can slightly exceed it"}
    eng.set_option("update_variant", 2)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import oracle as O
        O.build()
        cores = os.cpu_count() or 1
        rate, bins, e_cpu = cpu_baseline_sample(ham, kup, kdn, args.seed, n_occ, args.cpu_baseline_seconds, cores, "c128")
        rate64 = None
        if not cplx:
            rate64, _, _ = cpu_baseline_sample(ham, kup, kdn, args.seed, n_occ, max(2.0, args.cpu_baseline_seconds / 3), cores, "f64")
        cpu = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"{cores} walkers (one per core, each thread times its own run: no join) x {bins} bins of {n_occ} sweeps, ComplexF64 oracle",
               "value_f64": rate64, "E_per_site": e_cpu}

    if rank == 0:
        cfg = make_config(args, world)
        cfg.update({"walkers_total": total_walkers, "thermalization_sweeps": therm,
                    "w_update": ("walker resident in shared memory (k_resident), rank-1 update per accepted move" if resident else
                                 "delayed rank-k, Woodbury form (flush launch every 8 sweeps for walkers with >= 16 pending updates)"
                                 if woodbury else "immediate rank-1 update per accepted move"),
                    "refresh_kernels": "inside k_resident" if resident else "k_reeval_fused (one kernel)" if fused else "gather + inverse + GEMM kernels",
                    "step": "n_occ sweeps + O_L measurement + per-bin all-rank sum of the accumulators (inside the timed region)"})
        line = {
            "metric": metric_name(args), "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / max(K, 1), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "c128" if cplx else "f64", "data": "synthetic", "config": cfg,
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "roofline_w_update": roofline_update,
            "roofline_refresh": roofline_inverse, "roofline_refresh_gemm": roofline_gemm,
            "roofline_rank1_update": roofline_rank1, "e2e": e2e, "e2e_carlo": e2e_carlo, "cpu_baseline": cpu,
            "reduce": {"us_per_step": 1e3 * reduce_ms / max(K, 1), "what": "k_reduce_acc + ncclAllReduce of 8 doubles (N > 1) + D2H, device-timed, "
                       "inside every timed step", "nccl_ranks": world},
            "dmma_probe": {"sustained_tflops": dmma_peak, "burst_tflops": dmma_burst, "seconds": 1.0, "clocks": dclocks},
            "observables": {"E_per_site": res["energy"], "acc": res["acc"], "n_OL": res["n_OL"], "n_singular": res["n_singular"]},
            "kernel_ms": {k: round(v["ms"], 3) for k, v in tm.items()},
        }
        print(json.dumps(line), flush=True)
    eng.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
