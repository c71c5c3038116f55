import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import kagomedsl.jl_b200 as kd
import _util as U
lat, ham = U.problem(2, 2, (False, False), (False, False))
ns, nw, n = 12, 8, 60
ku0, kd0 = kd.init_conf_qr(ham, ns, ham.N_up)
ku = np.tile(ku0, (nw, 1)); kdn = np.tile(kd0, (nw, 1))
rng = np.random.default_rng(2026)
nb = len(ham.nn)
r = rng.random((n, nw)); bond = rng.integers(1, nb + 1, size=(n, nw)).astype(np.int32)
eng = kd.Engine(ham, nw); eng.set_config(ku, kdn); eng.refresh()
orc = U.oracle_walkers(ham, ku, kdn)
for s in range(n):
    eng.replay(r[s:s+1], bond[s:s+1])
    for w, mc in enumerate(orc):
        rc = mc.sweep(replay=(r[s, w], int(bond[s, w]), 1)); mc.sweeps = mc.sweeps + 1
        gku, gkd = eng.get_config()
        oku, okd = mc.kappa()
        Wu, Wd = mc.W()
        eu, ed = U.relerr(eng.get_W(w, 0), Wu), U.relerr(eng.get_W(w, 1), Wd)
        keq = np.array_equal(gku[w], oku) and np.array_equal(gkd[w], okd)
        if eu > 1e-10 or ed > 1e-10 or not keq:
            print(f"sweep {s} walker {w} rc={rc} gate={s%6==0} kappa_eq={keq} err_up={eu:.3e} err_dn={ed:.3e}")
print("done")
