"""ctypes front-end of the CPU oracle (TEST INFRASTRUCTURE, NOT PRODUCT CODE).

`libkdsl_oracle.so` restates the reference's VMC sampling path on the CPU
(oracle/kdsl_oracle.c, oracle/oracle_mc_core.inc; each function cites the
reference file:line it follows).  Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs may import this module.

The reference (hz-xiaxz/KagomeDSL.jl) is Julia and cannot run in this image,
so there is no oracle/_ref build; see DESIGN.md.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libkdsl_oracle.so")
_SRC = [os.path.join(_HERE, "kdsl_oracle.c"), os.path.join(_HERE, "oracle_mc_core.inc")]


def build(force: bool = False) -> str:
    """Compile the oracle with gcc if the shared object is missing or stale."""
    stale = (not os.path.exists(_SO)) or any(
        os.path.exists(s) and os.path.getmtime(s) > os.path.getmtime(_SO) for s in _SRC
    )
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-B", "libkdsl_oracle.so"], stdout=subprocess.DEVNULL)
    return _SO


_lib = None


class _Lattice(C.Structure):
    _fields_ = [
        ("n1", C.c_int), ("n2", C.c_int), ("t", C.c_double),
        ("a1", C.c_double * 2), ("a2", C.c_double * 2),
        ("r", (C.c_double * 2) * 6),
        ("pbc", C.c_int * 2), ("anti", C.c_int * 2),
    ]


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        L = _lib
        L.ko_xoshiro_next.restype = C.c_uint64
        L.ko_rand_f64.restype = C.c_double
        L.ko_rand_index.restype = C.c_int64
        L.ko_rand_index.argtypes = [C.c_void_p, C.c_uint64]
        L.ko_splitmix64.restype = C.c_uint64
        for sfx in ("_c128", "_f64"):
            getattr(L, "ko_mc_create" + sfx).restype = C.c_void_p
            getattr(L, "ko_mc_get_sweeps" + sfx).restype = C.c_int64
            getattr(L, "ko_mc_run" + sfx).restype = C.c_int64
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


# ----------------------------------------------------------------------------
# link tables (src/Hamiltonian.jl:219-261 and scripts/zero_flux.jl:15-42)
# ----------------------------------------------------------------------------
PI_LINK_IN = {(1, 2): 1, (1, 3): 1, (2, 3): 1, (2, 4): -1, (4, 6): 1, (4, 5): 1, (5, 6): 1,
              (2, 1): 1, (3, 1): 1, (3, 2): 1, (4, 2): -1, (6, 4): 1, (5, 4): 1, (6, 5): 1}
PI_LINK_INTER = {(3, 5, -1, 1): -1, (3, 1, 0, 1): -1, (6, 2, 0, 1): -1, (6, 4, 0, 1): 1,
                 (5, 1, 1, 0): 1, (1, 5, -1, 0): 1, (1, 3, 0, -1): -1, (2, 6, 0, -1): -1,
                 (4, 6, 0, -1): 1, (5, 3, 1, -1): -1}
ZERO_LINK_IN = {k: 1 for k in PI_LINK_IN}
ZERO_LINK_INTER = {k: 1 for k in PI_LINK_INTER}


def _flat_in(d):
    return np.ascontiguousarray([[a, b, v] for (a, b), v in d.items()], dtype=np.int32).reshape(-1, 3)


def _flat_inter(d):
    return np.ascontiguousarray([[a, b, dx, dy, v] for (a, b, dx, dy), v in d.items()], dtype=np.int32).reshape(-1, 5)


class Lattice:
    """DoubleKagome(t, n1, n2, PBC; antiPBC)  -- src/Lattice.jl:68-92"""

    def __init__(self, t, n1, n2, PBC, antiPBC=(False, False)):
        self.c = _Lattice()
        rc = lib().ko_lattice_init(C.byref(self.c), C.c_double(t), n1, n2, int(PBC[0]), int(PBC[1]),
                                   int(antiPBC[0]), int(antiPBC[1]))
        if rc == -1:
            raise ValueError("ArgumentError: antiperiodic without periodic boundary conditions")
        if rc == -2:
            raise AssertionError("n1 must be even in DoubleKagome")
        self.n1, self.n2 = n1, n2

    @property
    def ns(self):
        return lib().ko_ns(C.byref(self.c))

    def unitcell_coord(self, s):
        out = (C.c_double * 2)()
        if lib().ko_unitcell_coord(C.byref(self.c), s, out):
            raise AssertionError("s out of range")
        return np.array(out[:])

    def unitcell_diff(self, c1, c2):
        a = (C.c_double * 2)(*c1)
        b = (C.c_double * 2)(*c2)
        dx, dy = C.c_int(), C.c_int()
        lib().ko_unitcell_diff(C.byref(self.c), a, b, C.byref(dx), C.byref(dy))
        return dx.value, dy.value

    def get_site_coord(self, s):
        out = (C.c_double * 2)()
        if lib().ko_get_site_coord(C.byref(self.c), s, out):
            raise AssertionError("s out of range")
        return np.array(out[:])

    def get_boundary_shifts(self, s1, s2):
        dx = (C.c_int * 9)()
        dy = (C.c_int * 9)()
        sg = (C.c_double * 9)()
        n = lib().ko_get_boundary_shifts(C.byref(self.c), s1, s2, dx, dy, sg)
        if n < 0:
            raise AssertionError("bad site pair")
        return [(dx[q], dy[q], sg[q]) for q in range(n)]

    def apply_boundary_conditions(self, tunneling, s1, s2, link_inter, B=0.0):
        """tunneling: complex128 Fortran-ordered square matrix, modified in place"""
        assert tunneling.dtype == np.complex128 and tunneling.flags.f_contiguous
        li = _flat_inter(link_inter)
        rc = lib().ko_apply_boundary_conditions(_p(tunneling), tunneling.shape[0], C.byref(self.c), s1, s2,
                                                _p(li), li.shape[0], C.c_double(B))
        if rc:
            raise AssertionError("apply_boundary_conditions! assertion")

    def hmat(self, link_in=None, link_inter=None, B=0.0):
        """Hmat(lat; link_in, link_inter, B)  -- src/Hamiltonian.jl:308-354"""
        li = _flat_in(PI_LINK_IN if link_in is None else link_in)
        lx = _flat_inter(PI_LINK_INTER if link_inter is None else link_inter)
        ns = self.ns
        H = np.zeros((ns, ns), dtype=np.complex128, order="F")
        rc = lib().ko_hmat(C.byref(self.c), _p(li), li.shape[0], _p(lx), lx.shape[0], C.c_double(B), _p(H))
        if rc:
            raise RuntimeError("tunneling matrix must be upper triangular")
        return H


def get_nn(H):
    """get_nn(H_mat) -- src/Hamiltonian.jl:447-451; returns int32 [n_bonds, 2] 1-based pairs"""
    H = np.asfortranarray(H, dtype=np.complex128)
    ns = H.shape[0]
    n = lib().ko_get_nn(_p(H), ns, None, 0)
    bonds = np.zeros((max(n, 1), 2), dtype=np.int32)
    lib().ko_get_nn(_p(H), ns, _p(bonds), n)
    return bonds[:n]


def orbitals(H, N_up, N_down):
    """orbitals(H_mat, N_up, N_down) -- src/Hamiltonian.jl:382-393 (eigen(Hermitian) -> lowest-N
    eigenvectors).  LAPACK via numpy; the eigenvector gauge is unpinned (W is gauge invariant for
    closed shells).  Real H (B = 0) is diagonalised as a real symmetric matrix so U is real."""
    if np.abs(H.imag).max() == 0.0:
        w, v = np.linalg.eigh(np.ascontiguousarray(H.real))
    else:
        w, v = np.linalg.eigh(H)
    p = np.argsort(w, kind="stable")
    v = v[:, p]
    return np.asfortranarray(v[:, :N_up]), np.asfortranarray(v[:, :N_down]), w[p]


def Z(bonds, kup, kdn):
    b = np.ascontiguousarray(bonds, dtype=np.int32)
    ku = np.ascontiguousarray(kup, dtype=np.int64)
    kd = np.ascontiguousarray(kdn, dtype=np.int64)
    return lib().ko_Z(_p(b), b.shape[0], _p(ku), _p(kd))


def Sz(i, kup, kdn):
    ku = np.ascontiguousarray(kup, dtype=np.int64)
    kd = np.ascontiguousarray(kdn, dtype=np.int64)
    if len(ku) != len(kd):
        raise ValueError("DimensionMismatch")
    out = C.c_double()
    rc = lib().ko_Sz(i, _p(ku), _p(kd), len(ku), C.byref(out))
    if rc == -3:
        raise IndexError("BoundsError")
    if rc == -1:
        raise ValueError(f"ArgumentError: Site {i} is doubly occupied")
    if rc == -2:
        raise ValueError(f"ArgumentError: Site {i} is unoccupied")
    return out.value


def getxprime(bonds, kup, kdn):
    """returns dict {(K_up, l_up, K_down, l_down): coeff} with (-1,-1,-1,-1) the diagonal term"""
    b = np.ascontiguousarray(bonds, dtype=np.int32)
    ku = np.ascontiguousarray(kup, dtype=np.int64)
    kd = np.ascontiguousarray(kdn, dtype=np.int64)
    nk = 2 * b.shape[0] + 1
    keys = np.zeros((nk, 4), dtype=np.int64)
    coefs = np.zeros(nk)
    diag = C.c_double()
    n = lib().ko_getxprime(_p(b), b.shape[0], _p(ku), _p(kd), len(ku), _p(keys), _p(coefs), nk, C.byref(diag))
    if n < 0:
        raise ValueError("ArgumentError from Sz (code %d)" % n)
    out = {}
    for q in range(n):
        k = tuple(int(x) for x in keys[q])
        out[k] = out.get(k, 0.0) + coefs[q]
    out[(-1, -1, -1, -1)] = diag.value
    return out


def structure_factor(kup, cos_qr, sin_qr):
    """S(q) = |sum_i exp(i q r_i) Sz_i|^2 / ns of one configuration (Sz_i = +1/2 where kappa_up[i] != 0, else -1/2);
    checker of the product's extra observable (kdsl_set_observables), numpy"""
    sz = np.where(np.asarray(kup) != 0, 0.5, -0.5)
    re, im = np.asarray(cos_qr) @ sz, np.asarray(sin_qr) @ sz
    return (re * re + im * im) / len(sz)


class Xoshiro:
    """Julia's Random.Xoshiro stream (state supplied explicitly)."""

    def __init__(self, state):
        self.s = np.array(state, dtype=np.uint64)
        assert self.s.shape == (4,)

    @classmethod
    def from_seed(cls, seed):
        x = C.c_uint64(int(seed) & 0xFFFFFFFFFFFFFFFF)
        return cls([lib().ko_splitmix64(C.byref(x)) for _ in range(4)])

    def next_u64(self):
        return lib().ko_xoshiro_next(_p(self.s))

    def rand(self):
        return lib().ko_rand_f64(_p(self.s))

    def rand_index(self, n):
        return lib().ko_rand_index(_p(self.s), n)


def seed_states(seed, n_walkers):
    """per-walker Xoshiro256++ states: four SplitMix64 outputs of the stream started at mix(seed) ^ mix(~walker)
    (mix = one SplitMix64 output); the seeding policy of the product's rng.walker_states, restated"""
    M = 0xFFFFFFFFFFFFFFFF
    out = np.zeros((n_walkers, 4), dtype=np.uint64)
    key = lib().ko_splitmix64(C.byref(C.c_uint64(int(seed) & M)))
    for w in range(n_walkers):
        x = C.c_uint64(key ^ lib().ko_splitmix64(C.byref(C.c_uint64(~w & M))))
        for q in range(4):
            out[w, q] = lib().ko_splitmix64(C.byref(x))
    return out


class MC:
    """One reference walker: `mutable struct MC` + ctx.sweeps (src/MonteCarlo.jl:17-28)."""

    def __init__(self, bonds, U_up, U_dn, dtype="c128"):
        self.sfx = "_" + dtype
        self.dt = np.complex128 if dtype == "c128" else np.float64
        self.bonds = np.ascontiguousarray(bonds, dtype=np.int32).reshape(-1, 2)
        Uu = np.asfortranarray(U_up, dtype=self.dt)
        Ud = np.asfortranarray(U_dn, dtype=self.dt)
        self.ns, self.N_up = Uu.shape
        self.N_dn = Ud.shape[1]
        self.h = C.c_void_p(self._f("ko_mc_create")(self.ns, self.N_up, self.N_dn, self.bonds.shape[0],
                                                    _p(self.bonds), _p(Uu), _p(Ud)))

    def _f(self, name):
        return getattr(lib(), name + self.sfx)

    def __del__(self):
        try:
            self._f("ko_mc_destroy")(self.h)
        except Exception:
            pass

    def set_kappa(self, kup, kdn):
        ku = np.ascontiguousarray(kup, dtype=np.int64)
        kd = np.ascontiguousarray(kdn, dtype=np.int64)
        assert len(ku) == self.ns and len(kd) == self.ns
        self._f("ko_mc_set_kappa")(self.h, _p(ku), _p(kd))

    def kappa(self):
        ku = np.zeros(self.ns, dtype=np.int64)
        kd = np.zeros(self.ns, dtype=np.int64)
        self._f("ko_mc_get_kappa")(self.h, _p(ku), _p(kd))
        return ku, kd

    def W(self):
        Wu = np.zeros((self.ns, self.N_up), dtype=self.dt, order="F")
        Wd = np.zeros((self.ns, self.N_dn), dtype=self.dt, order="F")
        self._f("ko_mc_get_W")(self.h, _p(Wu), _p(Wd))
        return Wu, Wd

    def set_W(self, Wu, Wd):
        Wu = np.asfortranarray(Wu, dtype=self.dt)
        Wd = np.asfortranarray(Wd, dtype=self.dt)
        self._f("ko_mc_set_W")(self.h, _p(Wu), _p(Wd))

    @property
    def sweeps(self):
        return self._f("ko_mc_get_sweeps")(self.h)

    @sweeps.setter
    def sweeps(self, v):
        self._f("ko_mc_set_sweeps")(self.h, C.c_int64(v))

    def reevaluateW(self):
        rc = self._f("ko_mc_reevaluateW")(self.h)
        if rc == -10:
            raise np.linalg.LinAlgError("SingularException")
        if rc:
            raise ValueError("tilde_U error %d" % rc)

    def sweep(self, rng=None, replay=None):
        """Carlo.sweep!; returns flags (bit0 accepted, bit1 reached refresh block, bit2 refreshed)"""
        if replay is not None:
            r = C.c_double(replay[0])
            b = C.c_int32(replay[1])
            p = C.c_int32(replay[2] if len(replay) > 2 else 1)
            rc = self._f("ko_mc_sweep")(self.h, None, C.byref(r), C.byref(b), C.byref(p))
        else:
            rc = self._f("ko_mc_sweep")(self.h, _p(rng.s), None, None, None)
        if rc == -10:
            raise np.linalg.LinAlgError("SingularException")
        return rc

    def getOL(self, kup=None, kdn=None):
        if kup is None:
            kup, kdn = self.kappa()
        ku = np.ascontiguousarray(kup, dtype=np.int64)
        kd = np.ascontiguousarray(kdn, dtype=np.int64)
        out = C.c_double()
        rc = self._f("ko_getOL")(self.h, _p(ku), _p(kd), C.byref(out))
        if rc:
            raise ValueError("ArgumentError from Sz (code %d)" % rc)
        return out.value

    def measure(self):
        """Carlo.measure!: returns OL or None"""
        out = C.c_double()
        rc = self._f("ko_mc_measure")(self.h, C.byref(out))
        if rc < 0:
            raise ValueError("ArgumentError from Sz (code %d)" % rc)
        return out.value if rc == 1 else None

    def run(self, rng, n_steps, thermalization, stats=None, trace_len=0):
        """Carlo step loop; returns (stats[sum_acc, sum_OL, sum_OL2, n_OL], trace)"""
        if stats is None:
            stats = np.zeros(4)
        trace = np.zeros(max(trace_len, 1))
        rc = self._f("ko_mc_run")(self.h, _p(rng.s), C.c_int64(n_steps), C.c_int64(thermalization), _p(stats),
                                  _p(trace) if trace_len else None, C.c_int64(trace_len))
        if rc == -10:
            raise np.linalg.LinAlgError("SingularException")
        if rc < 0:
            raise ValueError("oracle run error %d" % rc)
        return stats, trace[: min(rc, trace_len)]

    def counters(self):
        out = np.zeros(3, dtype=np.int64)
        self._f("ko_mc_counters")(self.h, _p(out))
        return out


def tilde_U(U, kappa, dtype="c128"):
    dt = np.complex128 if dtype == "c128" else np.float64
    U = np.asfortranarray(U, dtype=dt)
    k = np.ascontiguousarray(kappa, dtype=np.int64)
    ns, N = U.shape
    if len(k) != ns:
        raise ValueError("DimensionMismatch")
    out = np.zeros((N, N), dtype=dt, order="F")
    rc = getattr(lib(), "ko_tilde_U_" + dtype)(_p(U), ns, N, _p(k), _p(out))
    if rc == -1:
        raise ValueError("ArgumentError: kappa is not valid")
    if rc == -2:
        raise IndexError("BoundsError")
    return out


def update_W(W, l, K, dtype="c128"):
    """update_W!(W, l, K, col_cache, row_cache) in place on a Fortran-ordered array"""
    dt = np.complex128 if dtype == "c128" else np.float64
    assert W.dtype == dt and W.flags.f_contiguous
    ns, N = W.shape
    col = np.zeros(ns, dtype=dt)
    row = np.zeros(N, dtype=dt)
    getattr(lib(), "ko_update_W_" + dtype)(_p(W), ns, N, l, K, _p(col), _p(row))
    return W


def init_conf_qr(U_up, U_dn, ns, N_up):
    """init_conf_qr! (src/MonteCarlo.jl:326-357): column-pivoted QR of U' picks N_up sites, then
    of U_down[available,:]' picks the ns-N_up down sites.  geqp3 pivot ties are implementation
    specific; this greedy max-residual-norm selection breaks ties by lowest index."""
    def pivots(A, k):      # A: m x n; returns first k column pivots of a column-pivoted QR
        A = np.array(A, dtype=np.complex128)
        m, n = A.shape
        norms = np.sum(np.abs(A) ** 2, axis=0)
        perm = np.arange(n)
        for step in range(min(k, m, n)):
            p = step + int(np.argmax(norms[step:]))
            if p != step:
                A[:, [step, p]] = A[:, [p, step]]
                perm[[step, p]] = perm[[p, step]]
                norms[[step, p]] = norms[[p, step]]
            x = A[step:, step]
            nx = np.linalg.norm(x)
            if nx == 0:
                break
            v = x.copy()
            v[0] += (x[0] / abs(x[0]) if x[0] != 0 else 1.0) * nx
            v /= np.linalg.norm(v)
            A[step:, step:] -= 2.0 * np.outer(v, v.conj() @ A[step:, step:])
            norms[step + 1:] = np.sum(np.abs(A[step + 1:, step + 1:]) ** 2, axis=0)
        return perm[:k]

    U_up = np.asarray(U_up)
    U_dn = np.asarray(U_dn)
    sites_up = pivots(U_up.conj().T, N_up)
    kup = np.zeros(ns, dtype=np.int64)
    for i, s in enumerate(sites_up):
        kup[s] = i + 1
    N_down = ns - N_up
    kdn = np.zeros(ns, dtype=np.int64)
    if N_down > 0:
        avail = np.setdiff1d(np.arange(ns), sites_up)
        sub = U_dn[avail, :]
        idx = pivots(sub.conj().T, N_down)
        for i, q in enumerate(idx):
            kdn[avail[q]] = i + 1
    return kup, kdn
