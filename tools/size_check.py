"""Correctness + throughput probe at one lattice size: GPU vs oracle after a replayed trajectory, then timing."""
import os, sys, time, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kagomedsl.jl_b200 as kd
from oracle import oracle as O
n, nw, flux = int(sys.argv[1]), int(sys.argv[2]), (sys.argv[3] if len(sys.argv) > 3 else "pi")
lat = kd.DoubleKagome(1.0, n, n, (True, True), (True, False)); ns = kd.ns(lat)
li, lx = (kd.pi_link_in, kd.pi_link_inter) if flux == "pi" else (kd.zero_link_in, kd.zero_link_inter)
ham = kd.Hamiltonian(ns // 2, ns // 2, lat, link_in=li, link_inter=lx)
ku, kdn = kd.init_conf_qr(ham, ns, ns // 2)
print(f"ns={ns} flux={flux} gap={ham.gap():.4f} walkers={nw}", flush=True)
eng = kd.Engine(ham, nw, 0)
eng.set_config(ku, kdn); eng.set_rng(kd.walker_states(99, nw)); eng.refresh()
# parity on 3 walkers: replay through one refresh period + a bit
nchk, nsw = 3, ns // 2 + 37
rng = np.random.default_rng(5)
r = rng.random((nsw, nw)); bond = rng.integers(1, len(ham.nn) + 1, size=(nsw, nw)).astype(np.int32)
eng.replay(r, bond)
gku, gkd = eng.get_config()
worst = 0.0
for w in range(nchk):
    mc = O.MC(np.asarray(ham.nn, dtype=np.int32), ham.U_up, ham.U_down, "f64")
    mc.set_kappa(ku, kdn); mc.reevaluateW()
    for s in range(nsw):
        mc.sweep(replay=(r[s, w], int(bond[s, w]), 1)); mc.sweeps = mc.sweeps + 1
    oku, okd = mc.kappa()
    assert np.array_equal(gku[w], oku) and np.array_equal(gkd[w], okd), "kappa mismatch"
    Wu, Wd = mc.W()
    worst = max(worst, np.abs(eng.get_W(w, 0) - Wu).max() / max(1, np.abs(Wu).max()), np.abs(eng.get_W(w, 1) - Wd).max() / max(1, np.abs(Wd).max()))
    assert abs(eng.measure()[w] - mc.getOL()) < 1e-9 * max(1, abs(mc.getOL()))
print(f"parity ok: kappa bit-exact on {nchk} walkers after {nsw} replayed sweeps, max rel W err {worst:.2e}", flush=True)
assert worst < 1e-10
# timing
n_occ = ns // 2
eng.sweeps = 0
eng.sweep(4 * n_occ, -1); eng.synchronize(); eng.reset_accumulators(); eng.reset_timers(); eng.set_profiling(True)
t0 = time.time(); eng.sweep(4 * n_occ, 0); eng.synchronize(); dt = time.time() - t0
acc = eng.accumulators(); tm = eng.timers()
print(json.dumps({"ns": ns, "walkers": nw, "walker_sweeps_per_s": nw * 4 * n_occ / dt, "acc": acc[1] / acc[0], "E_site": acc[2] / acc[4] / ns,
                  "n_singular": acc[7], "ms": {k: round(v["ms"], 2) for k, v in tm.items()}}))
