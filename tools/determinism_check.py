"""Run the same walkers several times (same handle size): are the chains reproducible?"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kagomedsl.jl_b200 as kd
n = int(sys.argv[1]); nw = int(sys.argv[2]); sweeps = int(sys.argv[3]); reps = int(sys.argv[4])
opts = [o.split("=") for o in sys.argv[5:]]
lat = kd.DoubleKagome(1.0, n, n, (True, True), (True, False)); ns = kd.ns(lat)
ham = kd.Hamiltonian(ns // 2, ns // 2, lat, B=float(os.environ.get('KDSL_B', '0'))); ku0, kd0 = kd.init_conf_qr(ham, ns, ns // 2)
states = kd.walker_states(1234, nw)
def run():
    e = kd.Engine(ham, nw, 0)
    for k, v in opts: e.set_option(k, int(v))
    e.set_config(ku0, kd0); e.set_rng(states); e.refresh()
    e.sweep(sweeps, -1)
    ku, kdn = e.get_config(); z, zr = e.Z(); assert np.array_equal(z, zr); e.close()
    return ku
ref = run()
for r in range(reps):
    b = run()
    bad = np.nonzero((ref != b).any(axis=1))[0]
    print("nw", nw, opts, "rep", r, "walkers differing from the first run:", len(bad), bad[:8])
