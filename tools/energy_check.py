"""<E>/site with error bars at the 432-site pi-flux DSL: GPU engine (independent walkers -> standard error over the
per-walker means) against the CPU oracle chain (restated reference algorithm, one walker per host core; standard error
over walkers).  BASELINE.json correctness criterion (3): <E>/site within statistical error bars of the reference."""
import os, sys, threading, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kagomedsl.jl_b200 as kd
from oracle import oracle as O

n = int(sys.argv[1]) if len(sys.argv) > 1 else 12
nw = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
bins_gpu = int(sys.argv[3]) if len(sys.argv) > 3 else 60
bins_cpu = int(sys.argv[4]) if len(sys.argv) > 4 else 400
lat = kd.DoubleKagome(1.0, n, n, (True, True), (True, False)); ns = kd.ns(lat); n_occ = ns // 2
ham = kd.Hamiltonian(ns // 2, ns // 2, lat); ku, kdn = kd.init_conf_qr(ham, ns, ns // 2)
therm = (int(sys.argv[5]) if len(sys.argv) > 5 else 20 * ns) // n_occ * n_occ   # sweeps before measuring (all walkers start from the same QR state)
# ---- GPU ----
eng = kd.Engine(ham, nw, 0)
eng.set_config(ku, kdn); eng.set_rng(kd.walker_states(1234, nw)); eng.refresh()
eng.sweep(therm, -1)
eng.reset_accumulators()
t0 = time.time()
eng.sweep(bins_gpu * n_occ, 0)
acc, acc_w, ol_w = eng.accumulators(per_walker=True)
t_gpu = time.time() - t0
e_w = ol_w / bins_gpu / ns                       # per-walker mean of O_L / ns over its bins
e_gpu, s_gpu = e_w.mean(), e_w.std(ddof=1) / np.sqrt(nw)
print("thermalization: %d sweeps" % therm)
print("GPU : %d walkers x %d bins  E/site = %.6f +- %.6f   acc = %.5f   (%.1f s)" % (nw, bins_gpu, e_gpu, s_gpu, acc[1] / acc[0], t_gpu), flush=True)
eng.close()
# ---- CPU oracle (f64 instantiation; the chain's law does not depend on the storage type) ----
O.build()
cores = os.cpu_count() or 1
bonds = np.asarray(ham.nn, dtype=np.int32)
res = [None] * cores
def work(t):
    mc = O.MC(bonds, ham.U_up, ham.U_down, "f64")
    mc.set_kappa(ku, kdn); mc.reevaluateW()
    g = O.Xoshiro.from_seed(99 + 7919 * t)
    mc.run(g, therm, -1) if False else mc.run(g, therm, 10 ** 12)          # thermalise without measuring
    st = np.zeros(4)
    mc.run(g, bins_cpu * n_occ, 0, stats=st)
    res[t] = st
t0 = time.time()
ths = [threading.Thread(target=work, args=(t,)) for t in range(cores)]
[th.start() for th in ths]; [th.join() for th in ths]
t_cpu = time.time() - t0
e_c = np.array([r[1] / r[3] / ns for r in res])
e_cpu, s_cpu = e_c.mean(), e_c.std(ddof=1) / np.sqrt(cores)
print("CPU : %d walkers x %d bins  E/site = %.6f +- %.6f   acc = %.5f   (%.1f s, %d cores)" % (cores, bins_cpu, e_cpu, s_cpu, sum(r[0] for r in res) / (cores * bins_cpu * n_occ), t_cpu, cores))
z = (e_gpu - e_cpu) / np.hypot(s_gpu, s_cpu)
print("difference = %.6f = %.2f sigma" % (e_gpu - e_cpu, z))
