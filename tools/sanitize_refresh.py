"""Small re-evaluation workload for compute-sanitizer (memcheck / racecheck): 108 sites, 8 walkers, two refreshes + a few sweeps."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import kagomedsl.jl_b200 as kd
n = int(sys.argv[1]) if len(sys.argv) > 1 else 6
lat = kd.DoubleKagome(1.0, n, n, (True, True), (True, False)); ns = kd.ns(lat)
ham = kd.Hamiltonian(ns // 2, ns // 2, lat); ku, kdn = kd.init_conf_qr(ham, ns, ns // 2)
eng = kd.Engine(ham, 8, 0)
eng.set_config(ku, kdn); eng.set_rng(kd.walker_states(7, 8))
eng.refresh()
eng.sweep(ns, 0)
eng.refresh()
W = eng.get_W(3, 0)
Wr = ham.U_up @ np.linalg.inv(kd.tilde_U(ham.U_up, eng.get_config()[0][3]))
print("max |dW| vs numpy:", float(np.abs(W - Wr).max()))
eng.close()
