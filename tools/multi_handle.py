"""Developer probe: the same 4096 walkers as ONE handle or split over g handles (one stream each) on the same GPU.
Kernels of different handles overlap wherever the block scheduler finds room (ramps, tails, latency-bound phases)."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kagomedsl.jl_b200 as kd

n = int(sys.argv[1]) if len(sys.argv) > 1 else 12
total = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
groups = int(sys.argv[3]) if len(sys.argv) > 3 else 2
bins = int(sys.argv[4]) if len(sys.argv) > 4 else 4
offset = int(sys.argv[5]) if len(sys.argv) > 5 else 0       # stagger: group i starts i * offset sweeps ahead
lat = kd.DoubleKagome(1.0, n, n, (True, True), (True, False)); ns = kd.ns(lat)
ham = kd.Hamiltonian(ns // 2, ns // 2, lat); ku, kdn = kd.init_conf_qr(ham, ns, ns // 2)
nw = total // groups
engs = []
for g in range(groups):
    e = kd.Engine(ham, nw, 0)
    e.set_config(ku, kdn); e.set_rng(kd.walker_states(1234, nw, first_walker=g * nw) if "first_walker" in kd.walker_states.__code__.co_varnames else kd.walker_states(1234 + g, nw))
    e.refresh()
    engs.append(e)
n_occ = ns // 2
for g, e in enumerate(engs):
    e.sweep(2 * n_occ + g * offset, -1)
for e in engs:
    e.synchronize()
t0 = time.time()
for b in range(bins):
    for e in engs:
        e.sweep(n_occ, 0)
for e in engs:
    e.synchronize()
dt = time.time() - t0
acc = [e.accumulators() for e in engs]
print(json.dumps({"ns": ns, "walkers": total, "handles": groups, "offset": offset, "bins": bins, "wall_s": dt,
                  "walker_sweeps_per_s": total * n_occ * bins / dt,
                  "E_site": float(sum(a[2] for a in acc) / max(sum(a[4] for a in acc), 1) / ns)}))
