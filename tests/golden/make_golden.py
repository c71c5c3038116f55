"""Regenerates tests/golden/derived.json: values DERIVED with the oracle (not published by the
reference): bond-table fingerprints and exact small-lattice energies by full enumeration.
The independent survey restatement obtained the same numbers; both are recorded.
Run: python tests/golden/make_golden.py"""
import hashlib
import itertools
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O


def fingerprint(n1, n2, pbc, anti, zero=False):
    L = O.Lattice(1.0, n1, n2, pbc, anti)
    H = L.hmat(O.ZERO_LINK_IN, O.ZERO_LINK_INTER) if zero else L.hmat()
    nn = O.get_nn(H)
    return {"n1": n1, "n2": n2, "PBC": list(pbc), "antiPBC": list(anti), "flux": "zero" if zero else "pi",
            "n_bonds": int(len(nn)), "sha256_16": hashlib.sha256(nn.astype("<i4").tobytes()).hexdigest()[:16],
            "first_bonds": nn[:6].tolist()}


def exact(n1, n2, pbc, anti, zero=False):
    L = O.Lattice(1.0, n1, n2, pbc, anti)
    H = L.hmat(O.ZERO_LINK_IN, O.ZERO_LINK_INTER) if zero else L.hmat()
    nn = O.get_nn(H)
    ns = L.ns
    N = ns // 2
    Uu, Ud, w = O.orbitals(H, N, N)
    mc = O.MC(nn, Uu, Ud, "c128")
    n1_ = d1 = n2_ = d2 = 0.0
    for up in itertools.combinations(range(ns), N):
        kup = np.zeros(ns, dtype=np.int64)
        kdn = np.zeros(ns, dtype=np.int64)
        for l, s in enumerate(up):
            kup[s] = l + 1
        for l, s in enumerate([s for s in range(ns) if s not in up]):
            kdn[s] = l + 1
        wgt = abs(np.linalg.det(O.tilde_U(Uu, kup))) ** 2 * abs(np.linalg.det(O.tilde_U(Ud, kdn))) ** 2
        if wgt < 1e-28:
            continue
        mc.set_kappa(kup, kdn)
        mc.reevaluateW()
        ol = mc.getOL()
        z = O.Z(nn, kup, kdn)
        n1_ += wgt * ol; d1 += wgt
        n2_ += wgt / z * ol; d2 += wgt / z
    return {"n1": n1, "n2": n2, "PBC": list(pbc), "antiPBC": list(anti), "flux": "zero" if zero else "pi",
            "gap": float(w[N] - w[N - 1]), "E_site_psi2": n1_ / d1 / ns, "E_site_chain_law": n2_ / d2 / ns}


if __name__ == "__main__":
    out = {
        "note": "derived with oracle/ (exact enumeration / sha256 of the int32 LE bond table); not reference-published",
        "fingerprints": [fingerprint(2, 2, (False, False), (False, False)), fingerprint(6, 6, (True, True), (True, False)),
                         fingerprint(12, 12, (True, True), (True, False)), fingerprint(12, 12, (True, True), (True, False), True),
                         fingerprint(18, 18, (True, True), (True, False))],
        "survey_fingerprints": {"2x2 OBC": "eb5c1f687a54036e", "6x6 PBC": "f76392c3315c43ca", "12x12 PBC": "2da61a622acfbc05",
                                "18x18 PBC": "97b99d953b7d07d8"},
        "exact": [exact(2, 2, (False, False), (False, False)), exact(2, 2, (True, True), (False, False)),
                  exact(2, 2, (True, True), (True, False)), exact(2, 2, (True, True), (True, False), True)],
        "survey_exact": {"pi 2x2 OBC": [-0.3720882491, -0.3714938624], "pi 2x2 PBC": [-0.4187331786, -0.4171722367],
                         "pi 2x2 PBC antiPBC(T,F)": [-0.4418508684, -0.4405702893],
                         "zero 2x2 PBC antiPBC(T,F)": [-0.4063342625, -0.4044681404]},
    }
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "derived.json"), "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out, indent=1)[:600])
