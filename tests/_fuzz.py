"""Randomised parity fuzz (test infrastructure: uses oracle/): random lattice shapes, boundary conditions, fluxes, fillings,
handle sizes and sweep counts; a few walkers of every case are compared bit for bit with the oracle's Carlo loop
(kappa, acceptance count, O_L sum) and every walker's incremental Z_mu with the recount."""
import time
import numpy as np

import kagomedsl.jl_b200 as kd
from oracle import oracle as O


def run_fuzz(seed, budget_s, big=False):
    rng = np.random.default_rng(seed)
    t0 = time.time()
    n_ok = n_skip = 0
    while time.time() - t0 < budget_s:
        n1 = int(rng.choice([2, 4, 6, 8, 10, 12] if big else [2, 4, 6, 8])); n2 = int(rng.integers(2, 13 if big else 9))
        PBC = (bool(rng.integers(0, 2)), bool(rng.integers(0, 2)))
        anti = (bool(rng.integers(0, 2)) and PBC[0], bool(rng.integers(0, 2)) and PBC[1])
        flux = str(rng.choice(["pi", "zero"]))
        B = float(rng.choice([0.0, 0.0, 0.03, 0.11]))
        lat = kd.DoubleKagome(1.0, n1, n2, PBC, anti)
        ns = kd.ns(lat)
        Nu = ns // 2 + int(rng.integers(-2, 3)) if ns >= 24 else ns // 2
        li, lx = (kd.pi_link_in, kd.pi_link_inter) if flux == "pi" else (kd.zero_link_in, kd.zero_link_inter)
        desc = f"n={n1}x{n2} PBC={PBC} anti={anti} flux={flux} B={B} ns={ns} N_up={Nu}"
        try:
            ham = kd.Hamiltonian(Nu, ns - Nu, lat, link_in=li, link_inter=lx, B=B)
            ku0, kd0 = kd.init_conf_qr(ham, ns, Nu)
            cu = np.linalg.cond(kd.tilde_U(np.asarray(ham.U_up), ku0))
            cd = np.linalg.cond(kd.tilde_U(np.asarray(ham.U_down), kd0))
            if not (cu < 1e6 and cd < 1e6):
                n_skip += 1
                continue
        except Exception:
            n_skip += 1
            continue
        nw = int(rng.choice([1, 7, 64, 300, 512, 777]))
        n = int(rng.integers(50, 700))
        therm = int(rng.integers(0, n))
        states = kd.walker_states(int(rng.integers(1, 10 ** 6)), nw)
        dtype = "c128" if B != 0.0 else "f64"
        eng = kd.Engine(ham, nw)
        # the default path of this size, or one of the other selectable paths (lock-step Woodbury / rank-1 updates, the
        # three-kernel re-evaluation; ComplexF64: embedding inverse, FMA flush and product)
        opts = [{}, {}, {"update_variant": 2}, {"update_variant": 0}, {"update_variant": 2, "inverse_variant": 5}]
        if B != 0.0:
            opts = [{}, {}, {"update_variant": 0}, {"inverse_variant": 7, "flush_variant": 4, "gemm_variant": 4}]
        opt = opts[int(rng.integers(0, len(opts)))]
        try:
            for k_, v_ in opt.items():
                eng.set_option(k_, v_)
        except kd.KdslError:
            pass
        desc += f" opts={opt}"
        eng.set_config(ku0, kd0)
        eng.set_rng(states)
        try:
            eng.refresh()
            cuts = sorted(set(int(c) for c in rng.integers(1, n, size=int(rng.integers(0, 3)))))   # 1..3 calls: the chain must not care
            done = 0
            for c in cuts + [n]:
                eng.sweep(c - done, thermalization=therm)
                done = c
            gku, gkd = eng.get_config()
            z, zr = eng.Z()
            acc, acc_w, ol_w = eng.accumulators(per_walker=True)
        except kd.KdslError:
            n_skip += 1
            eng.close()
            continue
        assert np.array_equal(z, zr), ("Z_mu != recount", desc)
        for w in sorted({0, nw - 1, int(rng.integers(0, nw))}):
            mc = O.MC(np.asarray(ham.nn, dtype=np.int32), ham.U_up, ham.U_down, dtype)
            mc.set_kappa(ku0, kd0)
            mc.reevaluateW()
            st, _ = mc.run(O.Xoshiro(states[w]), n, therm)
            oku, okd = mc.kappa()
            assert np.array_equal(gku[w], oku) and np.array_equal(gkd[w], okd), ("kappa", desc, nw, n, therm, w)
            assert acc_w[w] == st[0], ("acceptance count", desc, nw, n, therm, w, acc_w[w], st[0])
            assert abs(ol_w[w] - st[1]) <= 1e-8 * max(1.0, abs(st[1])), ("O_L sum", desc, nw, n, therm, w, ol_w[w], st[1])
        eng.close()
        n_ok += 1
    return n_ok, n_skip
