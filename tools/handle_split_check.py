"""Do the chains depend on how the walkers are grouped into handles?  (They must not: walkers are independent.)"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kagomedsl.jl_b200 as kd
n = int(sys.argv[1]) if len(sys.argv) > 1 else 12
total = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
sweeps = int(sys.argv[3]) if len(sys.argv) > 3 else 432
opts = [o.split("=") for o in sys.argv[4:]]
lat = kd.DoubleKagome(1.0, n, n, (True, True), (True, False)); ns = kd.ns(lat)
ham = kd.Hamiltonian(ns // 2, ns // 2, lat); ku0, kd0 = kd.init_conf_qr(ham, ns, ns // 2)
def run(groups, chunk):
    nw = total // groups
    out_u, out_acc, out_rng = [], [], []
    for g in range(groups):
        e = kd.Engine(ham, nw, 0)
        for k, v in opts: e.set_option(k, int(v))
        e.set_config(ku0, kd0); e.set_rng(kd.walker_states(1234, nw, first_walker=g * nw)); e.refresh()
        done = 0
        while done < sweeps:
            e.sweep(min(chunk, sweeps - done), -1); done += chunk
        ku, kdn = e.get_config()
        acc, acc_w, _ = e.accumulators(per_walker=True)
        out_u.append(ku); out_acc.append(acc_w); out_rng.append(e.get_rng())
        e.close()
    return np.concatenate(out_u), np.concatenate(out_acc), np.concatenate(out_rng)
a = run(1, sweeps)
for g, chunk in ((8, sweeps), (1, 50), (2, sweeps)):
    b = run(g, chunk)
    bad = np.nonzero((a[0] != b[0]).any(axis=1))[0]
    print("handles", g, "chunk", chunk, ": walkers with different kappa:", len(bad), bad[:10], "acc differ:", int((a[1] != b[1]).sum()), "rng differ:", int((a[2] != b[2]).any(axis=1).sum()))
