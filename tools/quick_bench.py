"""Developer timing probe (not the contract bench): per-kernel-class device times of the lock-step loop."""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kagomedsl.jl_b200 as kd


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=12)
    ap.add_argument("--walkers", type=int, default=4096)
    ap.add_argument("--sweeps", type=int, default=432)
    ap.add_argument("--therm", type=int, default=432)
    ap.add_argument("--opt", action="append", default=[], help="name=value engine options")
    ap.add_argument("--no-prof", action="store_true")
    ap.add_argument("--B", type=float, default=0.0, help="Peierls flux: B != 0 selects the ComplexF64 engine")
    args = ap.parse_args()
    lat = kd.DoubleKagome(1.0, args.n, args.n, (True, True), (True, False))
    ns = kd.ns(lat)
    t0 = time.time()
    ham = kd.Hamiltonian(ns // 2, ns // 2, lat, B=args.B)
    ku, kdn = kd.init_conf_qr(ham, ns, ns // 2)
    print(f"host setup {time.time()-t0:.2f}s ns={ns} bonds={len(ham.nn)}", flush=True)
    eng = kd.Engine(ham, args.walkers, 0)
    for o in args.opt:
        k, v = o.split("=")
        eng.set_option(k, int(v))
    eng.set_config(ku, kdn)
    eng.set_rng(kd.walker_states(1234, args.walkers))
    t0 = time.time()
    eng.refresh()
    print(f"initial refresh of {args.walkers} walkers: {time.time()-t0:.3f}s", flush=True)
    eng.sweep(args.therm, -1)
    eng.synchronize()
    eng.reset_accumulators()
    eng.reset_timers()
    eng.set_profiling(not args.no_prof)
    t0 = time.time()
    eng.sweep(args.sweeps, 0)
    eng.synchronize()
    dt = time.time() - t0
    tm = eng.timers()
    acc = eng.accumulators()
    ws = args.walkers * args.sweeps
    B_acc = 16 * ns * ns * (2 if eng.is_complex else 1)
    upd = tm["update"]
    print(json.dumps({
        "ns": ns, "walkers": args.walkers, "sweeps": args.sweeps, "wall_s": dt, "walker_sweeps_per_s": ws / dt,
        "acc": acc[1] / acc[0], "E_site": acc[2] / max(acc[4], 1) / ns, "n_OL": acc[4], "n_refresh": acc[6], "n_singular": acc[7],
        "timers": tm,
        "update_GBs": upd["moves"] * B_acc / (upd["ms"] * 1e-3) / 1e9 if upd["ms"] > 0 else None,
        "flush_GBs": upd["flushes"] * B_acc / (upd["ms"] * 1e-3) / 1e9 if upd["ms"] > 0 else None,
        "launches_total": sum(v["launches"] for v in tm.values()),
    }, indent=1))
    # pure W-update bandwidth: every walker gets one move
    eng.set_profiling(True)
    z = np.arange(args.walkers, dtype=np.int32)
    one = np.ones(args.walkers, dtype=np.int32)
    eng.reset_timers()
    for _ in range(5):
        eng.update_W(z, one, one * 3, one * 2, one * 5)
    tm = eng.timers()["update"]
    print("pure update: %.1f GB/s (%d moves, %.3f ms)" % (tm["moves"] * B_acc / (tm["ms"] * 1e-3) / 1e9, tm["moves"], tm["ms"]))
    eng.refresh()
    eng.close()


if __name__ == "__main__":
    main()
