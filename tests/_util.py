"""shared helpers for the tests: problem setup through the PRODUCT host code, oracle walkers"""
import numpy as np

import kagomedsl.jl_b200 as kd
from oracle import oracle as O

_cache = {}


def problem(n1, n2, PBC=(True, True), antiPBC=(True, False), flux="pi", N_up=None, B=0.0):
    key = (n1, n2, PBC, antiPBC, flux, N_up, B)
    if key not in _cache:
        lat = kd.DoubleKagome(1.0, n1, n2, PBC, antiPBC)
        nsites = kd.ns(lat)
        Nu = nsites // 2 if N_up is None else N_up
        Nd = nsites - Nu
        li, lx = (kd.pi_link_in, kd.pi_link_inter) if flux == "pi" else (kd.zero_link_in, kd.zero_link_inter)
        ham = kd.Hamiltonian(Nu, Nd, lat, link_in=li, link_inter=lx, B=B)
        _cache[key] = (lat, ham)
    return _cache[key]


def random_mott(rng, ns, N_up, n_walkers):
    """random Mott configurations with random label permutations: int64 [nw, ns] x 2"""
    ku = np.zeros((n_walkers, ns), dtype=np.int64)
    kd_ = np.zeros((n_walkers, ns), dtype=np.int64)
    for w in range(n_walkers):
        sites = rng.permutation(ns)
        up, dn = sites[:N_up], sites[N_up:]
        ku[w, up] = rng.permutation(N_up) + 1
        kd_[w, dn] = rng.permutation(ns - N_up) + 1
    return ku, kd_


def oracle_walkers(ham, ku, kd_, dtype="f64", refresh=True):
    out = []
    for w in range(ku.shape[0]):
        mc = O.MC(np.asarray(ham.nn, dtype=np.int32), ham.U_up, ham.U_down, dtype)
        mc.set_kappa(ku[w], kd_[w])
        if refresh:
            mc.reevaluateW()
        out.append(mc)
    return out


def well_conditioned_mott(rng, ham, ns, N_up, n_walkers, tries=200, cond_max=1e4):
    """random Mott states whose tilde_U matrices are comfortably invertible"""
    ku = np.zeros((n_walkers, ns), dtype=np.int64)
    kd_ = np.zeros((n_walkers, ns), dtype=np.int64)
    for w in range(n_walkers):
        for _ in range(tries):
            a, b = random_mott(rng, ns, N_up, 1)
            cu = np.linalg.cond(kd.tilde_U(ham.U_up, a[0]))
            cd = np.linalg.cond(kd.tilde_U(ham.U_down, b[0]))
            if cu < cond_max and cd < cond_max:
                break
        ku[w], kd_[w] = a[0], b[0]
    return ku, kd_


def relerr(a, b):
    return float(np.max(np.abs(a - b)) / max(1.0, float(np.max(np.abs(b)))))
