"""Profiling driver: a few full refreshes (gather, inverse, GEMM) of nw walkers.  Usage: prof_refresh.py n nw reps"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kagomedsl.jl_b200 as kd
n = int(sys.argv[1]) if len(sys.argv) > 1 else 12
nw = int(sys.argv[2]) if len(sys.argv) > 2 else 296
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
lat = kd.DoubleKagome(1.0, n, n, (True, True), (True, False)); ns = kd.ns(lat)
ham = kd.Hamiltonian(ns // 2, ns // 2, lat); ku, kdn = kd.init_conf_qr(ham, ns, ns // 2)
eng = kd.Engine(ham, nw, 0)
for o in sys.argv[4:]:
    if not o.startswith("+"):
        k, v = o.split("="); eng.set_option(k, int(v))
eng.set_config(ku, kdn)
eng.set_rng(kd.walker_states(1234, nw))
eng.refresh()
eng.sweep(200, -1)          # decorrelate the walkers so that pivot orders differ
for o in sys.argv[4:]:      # "+name=value": options set after the first refresh (debug variants)
    if o.startswith("+"):
        k, v = o[1:].split("="); eng.set_option(k, int(v))
for _ in range(reps):
    eng.refresh()
eng.synchronize()
eng.close()
