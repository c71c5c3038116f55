"""Regenerates tests/golden/energy_432.json: <E>/site of the reference's chain (law |psi|^2 / Z_mu) at the 432-site
pi-flux DSL from a long run of the CPU oracle (f64 instantiation), with the standard error over independent walkers.
The reference's own tests pin no energy, so this number is DERIVED with the oracle, whose sweep /
update / re-evaluation / getOL are pinned by the reference's known answers (tests/test_oracle_golden.py) and whose law
is pinned by exact enumeration on 12 sites (derived.json).
Run: python tests/golden/make_energy_432.py [walkers=16] [bins=8000] [thermalization=86400]   (~1-2 min on 8-16 cores)"""
import json
import os
import sys
import threading

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import kagomedsl.jl_b200 as kd          # host-side lattice / Hamiltonian / QR start state only (no GPU involved)
from oracle import oracle as O

walkers = int(sys.argv[1]) if len(sys.argv) > 1 else 16
bins = int(sys.argv[2]) if len(sys.argv) > 2 else 8000
therm = int(sys.argv[3]) if len(sys.argv) > 3 else 86400
lat = kd.DoubleKagome(1.0, 12, 12, (True, True), (True, False))
ns = kd.ns(lat)
n_occ = ns // 2
ham = kd.Hamiltonian(n_occ, n_occ, lat)
ku, kdn = kd.init_conf_qr(ham, ns, n_occ)
O.build()
bonds = np.asarray(ham.nn, dtype=np.int32)
res = [None] * walkers


def work(t):
    mc = O.MC(bonds, ham.U_up, ham.U_down, "f64")
    mc.set_kappa(ku, kdn)
    mc.reevaluateW()
    g = O.Xoshiro.from_seed(99 + 7919 * t)
    mc.run(g, therm, 10 ** 12)                       # thermalise without measuring
    st = np.zeros(4)
    mc.run(g, bins * n_occ, 0, stats=st)
    res[t] = st


cores = os.cpu_count() or 1
for lo in range(0, walkers, cores):
    ths = [threading.Thread(target=work, args=(t,)) for t in range(lo, min(walkers, lo + cores))]
    [th.start() for th in ths]
    [th.join() for th in ths]
e = np.array([r[1] / r[3] / ns for r in res])
out = {"lattice": "12x12 DoubleKagome (432 sites) pi-flux, PBC, antiPBC=(true,false), N_up = N_down = 216",
       "oracle_chain": {"E_per_site": float(e.mean()), "stderr": float(e.std(ddof=1) / np.sqrt(walkers)), "walkers": walkers,
                        "bins_per_walker": bins, "thermalization_sweeps": therm, "seeds": "Xoshiro.from_seed(99 + 7919 t)",
                        "acc": float(sum(r[0] for r in res) / (walkers * bins * n_occ))},
       "made_by": "tests/golden/make_energy_432.py"}
json.dump(out, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "energy_432.json"), "w"), indent=1)
print(out)
