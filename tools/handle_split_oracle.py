"""Which grouping matches the oracle?  (needs oracle/: developer / test tool only)"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kagomedsl.jl_b200 as kd
from oracle import oracle as O
n = int(sys.argv[1]) if len(sys.argv) > 1 else 12
total = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
sweeps = int(sys.argv[3]) if len(sys.argv) > 3 else 432
sub = int(sys.argv[4]) if len(sys.argv) > 4 else 512
opts = [o.split("=") for o in sys.argv[5:]]
lat = kd.DoubleKagome(1.0, n, n, (True, True), (True, False)); ns = kd.ns(lat)
ham = kd.Hamiltonian(ns // 2, ns // 2, lat); ku0, kd0 = kd.init_conf_qr(ham, ns, ns // 2)
states = kd.walker_states(1234, total)
def run(nw):
    e = kd.Engine(ham, nw, 0)
    for k, v in opts: e.set_option(k, int(v))
    e.set_config(ku0, kd0); e.set_rng(states[:nw]); e.refresh()
    e.sweep(sweeps, -1)
    ku, kdn = e.get_config(); acc, acc_w, _ = e.accumulators(per_walker=True); e.close()
    return ku, acc_w
big, acc_big = run(total)
small, acc_small = run(sub)
bad = np.nonzero((big[:sub] != small).any(axis=1))[0]
print("walkers 0..%d: %d differ between the %d-walker and the %d-walker handle" % (sub - 1, len(bad), total, sub), bad[:12])
for w in list(bad[:4]) + [0, 1]:
    mc = O.MC(np.asarray(ham.nn, dtype=np.int32), ham.U_up, ham.U_down, "f64")
    mc.set_kappa(ku0, kd0); mc.reevaluateW()
    st, _ = mc.run(O.Xoshiro(states[w]), sweeps, -1)
    oku, okd = mc.kappa()
    print("walker", w, "oracle == big:", bool(np.array_equal(oku, big[w])), " oracle == small:", bool(np.array_equal(oku, small[w])),
          " acc oracle/big/small:", st[0], acc_big[w], acc_small[w])
