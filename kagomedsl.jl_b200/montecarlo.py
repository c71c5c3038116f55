"""Host-side mirror of the reference's Monte Carlo interface (`src/MonteCarlo.jl`) on top of the
libkdsl C ABI.

The reference drives ONE walker through Carlo.jl's `AbstractMC` protocol:
`MC(params)`, `Carlo.init!`, `Carlo.sweep!`, `Carlo.measure!`, `Carlo.register_evaluables`,
`Carlo.write_checkpoint`, `Carlo.read_checkpoint!`.  Here the same calls drive `n_walkers`
independent walkers resident on one B200 (Python has no `!`, so `f!` is spelled `f_`).  All
arithmetic of the path runs in the CUDA library; this module only marshals.  There is no CPU
fallback: without libkdsl.so or without a GPU the calls raise.

Julia is not available in this image, so this Python layer is the host side that is exercised by
the tests; `julia/KagomeDSLB200.jl` holds the equivalent `ccall` glue.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, List, Optional, Sequence

import numpy as np

from . import _lib
from ._lib import KdslError, SingularException, check
from .hamiltonian import Hamiltonian, pi_link_in, pi_link_inter
from .lattice import DoubleKagome
from .rng import Xoshiro, walker_states


class AbstractMC:
    """Carlo.AbstractMC stand-in"""


# ------------------------------------------------------------------------------------------
# small host utilities that the reference exports / tests (integer or O(N^2) bookkeeping)
# ------------------------------------------------------------------------------------------
def tilde_U(U: np.ndarray, kappa: Sequence[int]) -> np.ndarray:
    """src/MonteCarlo.jl:92-115: tilde_U[l, :] = U[R_l, :] (host utility for tests / inspection;
    the engine builds tilde_U on the device in k_gather_tilde)."""
    U = np.asarray(U)
    n, m = U.shape
    kappa = np.asarray(kappa)
    if len(kappa) != n:
        raise ValueError(f"DimensionMismatch: Length of kappa ({len(kappa)}) must match number of rows in U ({n})")
    if int(np.count_nonzero(kappa)) != m:
        raise ValueError(f"ArgumentError: kappa ({kappa.tolist()}) is not valid")
    out = np.zeros((m, m), dtype=U.dtype)
    for Rl, l in enumerate(kappa):
        if l != 0:
            if not 1 <= l <= m:
                raise IndexError(f"BoundsError: attempt to access {m}x{m} matrix at index [{l}, :]")
            out[l - 1, :] = U[Rl, :]
    return out


def Z(nn, kappa_up, kappa_down) -> int:
    """src/MonteCarlo.jl:460-474 (host utility; the engine maintains Z_mu on the device)"""
    count = 0
    for (s1, s2) in nn:
        if kappa_up[s1 - 1] != 0 and kappa_down[s2 - 1] != 0:
            count += 1
        elif kappa_up[s2 - 1] != 0 and kappa_down[s1 - 1] != 0:
            count += 1
    return count


def _pivoted_columns(A: np.ndarray, k: int) -> np.ndarray:
    """first k column pivots of a column-pivoted Householder QR of A (LAPACK geqp3's greedy
    largest-remaining-norm rule; exact ties go to the lowest index)"""
    A = np.array(A, dtype=np.complex128 if np.iscomplexobj(A) else np.float64)
    m, n = A.shape
    perm = np.arange(n)
    norms = np.sum(np.abs(A) ** 2, axis=0)
    for step in range(min(k, m, n)):
        p = step + int(np.argmax(norms[step:]))
        if p != step:
            A[:, [step, p]] = A[:, [p, step]]
            perm[[step, p]] = perm[[p, step]]
        x = A[step:, step]
        nx = np.linalg.norm(x)
        if nx == 0.0:
            break
        v = x.copy()
        v[0] += (x[0] / abs(x[0]) if x[0] != 0 else 1.0) * nx
        v /= np.linalg.norm(v)
        A[step:, step:] -= 2.0 * np.outer(v, v.conj() @ A[step:, step:])
        norms[step + 1:] = np.sum(np.abs(A[step + 1:, step + 1:]) ** 2, axis=0)
    return perm[:k]


def init_conf_qr(Ham: Hamiltonian, ns: int, N_up: int):
    """`init_conf_qr!` (src/MonteCarlo.jl:326-357): pick N_up sites by column-pivoted QR of U_up',
    then label all remaining ns - N_up sites (the reference ignores params[:N_down] here, :339)
    in the pivot order of U_down[available, :]'.  Returns (kappa_up, kappa_down) as int64 vectors."""
    sites_up = _pivoted_columns(np.asarray(Ham.U_up).conj().T, N_up)
    kappa_up = np.zeros(ns, dtype=np.int64)
    for i, site in enumerate(sites_up):
        kappa_up[site] = i + 1
    N_down = ns - N_up
    kappa_down = np.zeros(ns, dtype=np.int64)
    if N_down > 0:
        available = np.setdiff1d(np.arange(ns), sites_up)
        sub = np.asarray(Ham.U_down)[available, :]
        idx = _pivoted_columns(sub.conj().T, N_down)
        for i, q in enumerate(idx):
            kappa_down[available[q]] = i + 1
    return kappa_up, kappa_down


# ------------------------------------------------------------------------------------------
# Engine: object wrapper of one libkdsl handle
# ------------------------------------------------------------------------------------------
def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Engine:
    """All walkers of one GPU (one `kdsl_handle`)."""

    def __init__(self, Ham: Hamiltonian, n_walkers: int = 1, device: Optional[int] = None):
        if device is None:
            device = int(os.environ.get("LOCAL_RANK", "0"))
        U_up = np.asarray(Ham.U_up)
        U_dn = np.asarray(Ham.U_down)
        # complex orbitals (Peierls flux B != 0) select the ComplexF64 engine, real ones the FP64 engine
        self.is_complex = any(np.iscomplexobj(U) and np.abs(U.imag).max(initial=0.0) > 0.0 for U in (U_up, U_dn))
        self.dtype = np.complex128 if self.is_complex else np.float64
        if self.is_complex:
            self._Uu = np.asfortranarray(U_up, dtype=np.complex128)
            self._Ud = np.asfortranarray(U_dn, dtype=np.complex128)
        else:
            self._Uu = np.asfortranarray(U_up.real, dtype=np.float64)
            self._Ud = np.asfortranarray(U_dn.real, dtype=np.float64)
        self.ns, self.N_up = self._Uu.shape
        self.N_dn = self._Ud.shape[1]
        self.bonds = np.ascontiguousarray(np.asarray(Ham.nn, dtype=np.int32).reshape(-1, 2))
        self.n_bonds = self.bonds.shape[0]
        self.nw = int(n_walkers)
        self.device = int(device)
        self.n_occ = min(self.N_up, self.N_dn)
        self._h = C.c_void_p()
        L = _lib.lib()
        create = L.kdsl_create_c128 if self.is_complex else L.kdsl_create
        check(create(C.byref(self._h), self.device, self.ns, self.N_up, self.N_dn, self.n_bonds,
                     _ptr(self.bonds), _ptr(self._Uu), _ptr(self._Ud), self.nw))
        self._L = L

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self._L.kdsl_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- state ---------------------------------------------------------------------------
    def set_config(self, kappa_up, kappa_down):
        ku = np.asarray(kappa_up)
        kd = np.asarray(kappa_down)
        if ku.ndim == 1:
            ku = np.broadcast_to(ku, (self.nw, self.ns))
            kd = np.broadcast_to(kd, (self.nw, self.ns))
        if ku.shape != (self.nw, self.ns) or kd.shape != (self.nw, self.ns):
            raise ValueError(f"DimensionMismatch: kappa must be [{self.nw}, {self.ns}] or [{self.ns}]")
        ku = np.ascontiguousarray(ku, dtype=np.int32)
        kd = np.ascontiguousarray(kd, dtype=np.int32)
        check(self._L.kdsl_set_config(self._h, _ptr(ku), _ptr(kd)))

    def get_config(self):
        ku = np.zeros((self.nw, self.ns), dtype=np.int32)
        kd = np.zeros((self.nw, self.ns), dtype=np.int32)
        check(self._L.kdsl_get_config(self._h, _ptr(ku), _ptr(kd)))
        return ku, kd

    def set_rng(self, states):
        st = np.ascontiguousarray(states, dtype=np.uint64)
        if st.shape != (self.nw, 4):
            raise ValueError(f"rng states must be [{self.nw}, 4] uint64")
        check(self._L.kdsl_set_rng(self._h, _ptr(st)))

    def get_rng(self):
        st = np.zeros((self.nw, 4), dtype=np.uint64)
        check(self._L.kdsl_get_rng(self._h, _ptr(st)))
        return st

    @property
    def sweeps(self) -> int:
        v = C.c_int64()
        check(self._L.kdsl_get_sweeps(self._h, C.byref(v)))
        return v.value

    @sweeps.setter
    def sweeps(self, v: int):
        check(self._L.kdsl_set_sweeps(self._h, int(v)))

    # -- the hot path --------------------------------------------------------------------
    def refresh(self) -> None:
        """reevaluateW! for every walker; raises SingularException like the reference"""
        n = C.c_int(0)
        check(self._L.kdsl_refresh(self._h, C.byref(n)))

    def sweep(self, n_sweeps: int = 1, thermalization: int = -1) -> None:
        check(self._L.kdsl_sweep(self._h, int(n_sweeps), int(thermalization)))

    def replay(self, r, bond_idx, pick=None, thermalization: int = -1) -> None:
        r = np.ascontiguousarray(r, dtype=np.float64)
        b = np.ascontiguousarray(bond_idx, dtype=np.int32)
        if r.ndim == 1:
            r = r.reshape(-1, self.nw)
            b = b.reshape(-1, self.nw)
        n = r.shape[0]
        if r.shape != (n, self.nw) or b.shape != (n, self.nw):
            raise ValueError(f"replay buffers must be [n_sweeps, {self.nw}]")
        p = None
        if pick is not None:
            p = np.ascontiguousarray(pick, dtype=np.int32).reshape(n, self.nw)
        check(self._L.kdsl_replay(self._h, n, int(thermalization), _ptr(r), _ptr(b), _ptr(p)))
        check(self._L.kdsl_synchronize(self._h))      # host buffers may be released after return

    def measure(self) -> np.ndarray:
        ol = np.zeros(self.nw)
        check(self._L.kdsl_measure(self._h, _ptr(ol)))
        return ol

    def last_OL(self):
        ol = np.zeros(self.nw)
        n = np.zeros(self.nw, dtype=np.int64)
        check(self._L.kdsl_last_OL(self._h, _ptr(ol), _ptr(n)))
        return ol, n

    def accumulators(self, per_walker: bool = False):
        out = np.zeros(_lib.N_ACC)
        acc_w = np.zeros(self.nw, dtype=np.int64) if per_walker else None
        ol_w = np.zeros(self.nw) if per_walker else None
        check(self._L.kdsl_accumulators(self._h, _ptr(out), _ptr(acc_w), _ptr(ol_w)))
        return (out, acc_w, ol_w) if per_walker else out

    def reset_accumulators(self):
        check(self._L.kdsl_reset_accumulators(self._h))

    # -- extra observables ------------------------------------------------
    def set_observables(self, q_vectors=None, coords=None) -> None:
        """enable the structure factor S(q) at the given wave vectors ([nq, 2]; `coords` = site positions [ns, 2]) and the
        Z_mu reweighting sums; with no wave vectors only the reweighting sums are taken"""
        if q_vectors is None or len(q_vectors) == 0:
            self._nq = 0
            check(self._L.kdsl_set_observables(self._h, 0, None, None))
            return
        q = np.asarray(q_vectors, dtype=np.float64).reshape(-1, 2)
        r = np.asarray(coords, dtype=np.float64).reshape(self.ns, 2)
        ph = q @ r.T                                                    # [nq, ns]
        c, s = np.ascontiguousarray(np.cos(ph)), np.ascontiguousarray(np.sin(ph))
        self._nq = q.shape[0]
        check(self._L.kdsl_set_observables(self._h, self._nq, _ptr(c), _ptr(s)))

    def observables(self, allreduce: bool = False) -> Dict[str, np.ndarray]:
        """sums and the derived estimators: chain averages (law |psi|^2 / Z_mu) and |psi|^2 averages (reweighted)"""
        nq = getattr(self, "_nq", 0)
        out = np.zeros(4 + 2 * nq)
        check(self._L.kdsl_get_observables(self._h, _ptr(out), int(bool(allreduce))))
        n, z, olz = out[0], out[1], out[2]
        sq, sqz = out[4:4 + nq], out[4 + nq:]
        return {"n": n, "sum_Z": z, "sum_OL_Z": olz, "sum_Sq": sq, "sum_Sq_Z": sqz,
                "energy_psi2": olz / z / self.ns if z else float("nan"),   # <O_L>_{|psi|^2} / ns
                "Sq_chain": sq / n if n else sq, "Sq_psi2": sqz / z if z else sqz}

    # -- multi-GPU: NCCL sum of the accumulators inside the C ABI ----------------------------
    def comm_init_rank(self, n_ranks: int, rank: int, unique_id: bytes) -> None:
        """collective over the ranks of a one-process-per-GPU job (kdsl_comm_init_rank)"""
        buf = (C.c_uint8 * _lib.COMM_ID_BYTES).from_buffer_copy(bytes(unique_id))
        check(self._L.kdsl_comm_init_rank(self._h, int(n_ranks), int(rank), C.cast(buf, C.c_void_p)))

    def comm_info(self):
        r, n = C.c_int(-1), C.c_int(0)
        check(self._L.kdsl_comm_info(self._h, C.byref(r), C.byref(n)))
        return r.value, n.value

    def accumulators_allreduce(self) -> np.ndarray:
        """global sums over every rank's walkers (one ncclAllReduce of KDSL_N_ACC doubles on the engine's stream)"""
        out = np.zeros(_lib.N_ACC)
        check(self._L.kdsl_accumulators_allreduce(self._h, _ptr(out)))
        return out

    # -- inspection / test hooks ---------------------------------------------------------
    def get_W(self, walker: int, spin: int) -> np.ndarray:
        N = self.N_dn if spin else self.N_up
        out = np.zeros((self.ns, N), order="F", dtype=self.dtype)
        check(self._L.kdsl_get_W(self._h, int(walker), int(spin), _ptr(out)))
        return out

    def set_W(self, walker: int, spin: int, W) -> None:
        N = self.N_dn if spin else self.N_up
        W = np.asfortranarray(W, dtype=self.dtype)
        if W.shape != (self.ns, N):
            raise ValueError("DimensionMismatch")
        check(self._L.kdsl_set_W(self._h, int(walker), int(spin), _ptr(W)))

    def update_W(self, walker, l_up, K_up, l_dn, K_dn) -> None:
        arrs = [np.ascontiguousarray(np.atleast_1d(a), dtype=np.int32) for a in (walker, l_up, K_up, l_dn, K_dn)]
        n = len(arrs[0])
        check(self._L.kdsl_update_W(self._h, n, *[_ptr(a) for a in arrs]))

    def Z(self, recount: bool = True):
        z = np.zeros(self.nw, dtype=np.int32)
        zr = np.zeros(self.nw, dtype=np.int32) if recount else None
        check(self._L.kdsl_get_Z(self._h, _ptr(z), _ptr(zr)))
        return (z, zr) if recount else z

    def flags(self) -> np.ndarray:
        f = np.zeros(self.nw, dtype=np.int32)
        check(self._L.kdsl_get_flags(self._h, _ptr(f)))
        return f

    def set_profiling(self, on: bool):
        check(self._L.kdsl_set_profiling(self._h, int(bool(on))))

    def timers(self) -> Dict[str, Dict[str, float]]:
        ms = np.zeros(_lib.N_TIMERS)
        n = np.zeros(_lib.N_TIMERS, dtype=np.int64)
        mv = np.zeros(2, dtype=np.int64)
        check(self._L.kdsl_timers(self._h, _ptr(ms), _ptr(n), _ptr(mv)))
        out = {name: {"ms": float(ms[i]), "launches": int(n[i])} for i, name in enumerate(_lib.TIMER_NAMES)}
        out["update"]["moves"] = int(mv[0])
        out["update"]["flushes"] = int(mv[1])
        return out

    def reset_timers(self):
        check(self._L.kdsl_reset_timers(self._h))

    def set_option(self, name: str, value: int):
        check(self._L.kdsl_set_option(self._h, name.encode(), int(value)))

    def synchronize(self):
        check(self._L.kdsl_synchronize(self._h))

    def fp64_dmma_peak_tflops(self, seconds: float = 0.0):
        """(sustained, burst) FP64 DMMA TFLOP/s of this device: the probe kernel launched back to back for `seconds`"""
        s, b = C.c_double(0.0), C.c_double(0.0)
        check(self._L.kdsl_bench_fp64_dmma_sustained(self._h, float(seconds), C.byref(s), C.byref(b)))
        return s.value, b.value

    def event_record(self, slot: int):
        check(self._L.kdsl_event_record(self._h, int(slot)))

    def event_elapsed_ms(self, a: int, b: int) -> float:
        ms = C.c_double(0.0)
        check(self._L.kdsl_event_elapsed(self._h, int(a), int(b), C.byref(ms)))
        return ms.value


# ------------------------------------------------------------------------------------------
# Carlo.jl stand-ins (the real Carlo package is the reference's external scheduler)
# ------------------------------------------------------------------------------------------
class MCContext:
    """Carlo.MCContext{Random.Xoshiro}(params): sweeps, thermalization_sweeps, rng, measurements
    (SURVEY Appendix A.1)."""

    def __init__(self, params: Dict, seed_variation: int = 0):
        self.sweeps = 0
        self.thermalization_sweeps = int(params.get("thermalization", 0))
        self.binsize = int(params.get("binsize", 1))
        seed = params.get("seed", None)
        self.rng = Xoshiro(seed=None if seed is None else int(seed) * (1 + seed_variation))
        self.measurements: Dict[str, List] = {}

    def is_thermalized(self) -> bool:
        return self.sweeps > self.thermalization_sweeps

    def mean(self, name: str):
        v = np.asarray(self.measurements[name], dtype=np.float64)
        return v.mean(axis=0)


class Evaluator:
    """Carlo.Evaluator stand-in: records `evaluate!(eval, name, (obs...,)) do ... end`"""

    def __init__(self):
        self.evaluables = {}

    def evaluate_(self, name: str, observables: Sequence[str], func):
        self.evaluables[name] = (tuple(observables), func)

    def results(self, ctx: MCContext) -> Dict[str, float]:
        out = {}
        for name, (obs, func) in self.evaluables.items():
            out[name] = func(*[np.mean(np.asarray(ctx.measurements[o], dtype=np.float64)) for o in obs])
        return out


# ------------------------------------------------------------------------------------------
# MC: the reference's `mutable struct MC <: AbstractMC` for a batch of walkers
# ------------------------------------------------------------------------------------------
class MC(AbstractMC):
    """`MC(params)` (src/MonteCarlo.jl:165-185) or `MC(Ham, kappa_up, kappa_down, W_up, W_down)` (:209-234).

    params keys follow the reference (as strings): n1, n2, PBC, N_up, N_down, and optional antiPBC,
    lattice, B, link_in, link_inter; added: n_walkers (1), device (LOCAL_RANK or 0),
    sweeps_per_call (1: one Carlo sweep = one reference sweep of every walker).
    The GPU engine is created lazily (at init_/read_checkpoint_), so constructing an MC needs no GPU,
    exactly like the reference where MC(params) only builds the Hamiltonian and zero matrices.
    """

    def __init__(self, *args):
        self.engine: Optional[Engine] = None
        self._acc_w_seen = None          # per-walker accepted-move counts already reported as :acc (None: zeros)
        if len(args) == 1:
            params = args[0]
            n1, n2 = params["n1"], params["n2"]
            PBC = params["PBC"]
            antiPBC = params.get("antiPBC", (False, False))
            lat_type = params.get("lattice", DoubleKagome)
            B = params.get("B", 0.0)
            lat = lat_type(1.0, n1, n2, PBC, antiPBC)
            N_up, N_down = params["N_up"], params["N_down"]
            link_in = params.get("link_in", pi_link_in)
            link_inter = params.get("link_inter", pi_link_inter)
            self.Ham = Hamiltonian(N_up, N_down, lat, link_in=link_in, link_inter=link_inter, B=B)
            nsites = n1 * n2 * 3
            self.n_walkers = int(params.get("n_walkers", 1))
            self.device = params.get("device", None)
            self.sweeps_per_call = int(params.get("sweeps_per_call", 1))
            self._kappa_up = np.zeros(nsites, dtype=np.int64)        # :179-180
            self._kappa_down = np.zeros(nsites, dtype=np.int64)
            self._W0 = (np.zeros((nsites, N_up)), np.zeros((nsites, N_down)))   # :181-182
        elif len(args) == 5:
            Ham, kappa_up, kappa_down, W_up, W_down = args
            W_up, W_down = np.asarray(W_up), np.asarray(W_down)
            nsites, N_up = W_up.shape
            _, N_down = W_down.shape
            assert nsites != 0 and N_up != 0 and N_down != 0
            self.Ham = Ham
            self.n_walkers, self.device, self.sweeps_per_call = 1, None, 1
            self._kappa_up = np.asarray(kappa_up, dtype=np.int64).copy()
            self._kappa_down = np.asarray(kappa_down, dtype=np.int64).copy()
            self._W0 = (W_up.copy(), W_down.copy())
        else:
            raise TypeError("MC(params) or MC(Ham, kappa_up, kappa_down, W_up, W_down)")
        self.ns = len(self._kappa_up)

    # -- fields of the reference struct (walker 0 when batched; *_all for every walker) -------
    @property
    def kappa_up(self) -> np.ndarray:
        return self.kappa_all()[0][0] if self.engine is not None else self._kappa_up

    @property
    def kappa_down(self) -> np.ndarray:
        return self.kappa_all()[1][0] if self.engine is not None else self._kappa_down

    def kappa_all(self):
        ku, kd = self.engine.get_config()
        return ku.astype(np.int64), kd.astype(np.int64)

    @property
    def W_up(self) -> np.ndarray:
        return self.engine.get_W(0, 0) if self.engine is not None else self._W0[0]

    @property
    def W_down(self) -> np.ndarray:
        return self.engine.get_W(0, 1) if self.engine is not None else self._W0[1]

    @property
    def n_occupied(self) -> int:
        return min(self.Ham.N_up, self.Ham.N_down)

    # -- engine plumbing -----------------------------------------------------------------
    def _ensure_engine(self) -> Engine:
        if self.engine is None:
            self.engine = Engine(self.Ham, self.n_walkers, self.device)
        pending = getattr(self, "_pending", None)
        if pending is not None:
            self._pending = None
            self.load_configuration(*pending)
        return self.engine

    def sync_counters(self) -> None:
        """make the next sweep_ report :acc relative to the engine's CURRENT counters (after sweeps driven through
        run_ / the engine directly)"""
        self._acc_w_seen = self._ensure_engine().accumulators(per_walker=True)[1]

    def load_configuration(self, kappa_up, kappa_down, rng_states=None) -> None:
        """put configurations on the GPU and (re)compute W = U * inv(tilde_U) for every walker"""
        eng = self._ensure_engine()
        eng.set_config(kappa_up, kappa_down)
        if rng_states is not None:
            eng.set_rng(rng_states)
        eng.refresh()


def reevaluateW_(mc: MC) -> None:
    """`reevaluateW!(mc)` (src/MonteCarlo.jl:55-66) for every walker"""
    mc._ensure_engine().refresh()


def find_initial_configuration_(mc: MC, ns: int, N_up: int, rng_states=None) -> None:
    """`find_initial_configuration!` (src/MonteCarlo.jl:382-410): QR-selected configuration (the same
    for every walker, as every reference run starts from the same deterministic state), then the first W."""
    kappa_up, kappa_down = init_conf_qr(mc.Ham, ns, N_up)
    try:
        mc.load_configuration(kappa_up, kappa_down, rng_states)
    except SingularException:
        raise RuntimeError("QR-based configuration is singular. The Hamiltonian may be rank-deficient.") from None


def init_(mc: MC, ctx: MCContext, params: Dict) -> None:
    """`Carlo.init!(mc, ctx, params)` (src/MonteCarlo.jl:435-441).  The per-walker Xoshiro streams
    are split off ctx.rng (4 words per walker) so that ctx.rng stays the single source of seeding."""
    n1, n2 = params["n1"], params["n2"]
    nsites = n1 * n2 * 3
    N_up = params["N_up"]
    states = np.array([[ctx.rng.next_u64() for _ in range(4)] for _ in range(mc.n_walkers)], dtype=np.uint64)
    find_initial_configuration_(mc, nsites, N_up, states)
    mc.engine.reset_accumulators()
    mc._acc_w_seen = None


def sweep_(mc: MC, ctx: MCContext) -> None:
    """`Carlo.sweep!(mc, ctx)` (src/MonteCarlo.jl:538-607): one proposal for every walker (times sweeps_per_call), then
    the :acc observable (:548-589).  The reference records 0.0 / 1.0 for its one walker; a batch records the VECTOR of
    the walkers' acceptance fractions over this call (length n_walkers: a Carlo vector observable, one chain per
    component), a single walker the scalar."""
    eng = mc._ensure_engine()
    k = mc.sweeps_per_call
    eng.sweeps = ctx.sweeps * k
    eng.sweep(k, -1)
    acc, acc_w, _ = eng.accumulators(per_walker=True)
    seen = mc._acc_w_seen if mc._acc_w_seen is not None else np.zeros_like(acc_w)
    d_w = (acc_w - seen) / float(k)
    mc._acc_w_seen = acc_w
    if acc[_lib.ACC_N_SINGULAR] > 0:
        raise SingularException(_lib.KDSL_ERR_SINGULAR, "lu factorization failed in reevaluateW! (SingularException)")
    measure_(ctx, "acc", float(d_w[0]) if mc.n_walkers == 1 else d_w)


def measure_(*args):
    """`Carlo.measure!(mc, ctx)` (src/MonteCarlo.jl:628-634) or `measure!(ctx, name, value)`"""
    if len(args) == 3:
        ctx, name, value = args
        ctx.measurements.setdefault(name, []).append(value)
        return None
    mc, ctx = args
    n_occupied = mc.n_occupied
    if (ctx.sweeps * mc.sweeps_per_call) % n_occupied == 0:
        OL = getOL(mc)
        measure_(ctx, "OL", float(OL[0]) if mc.n_walkers == 1 else OL)
    return None


def getOL(mc: MC, kappa_up=None, kappa_down=None) -> np.ndarray:
    """`getOL(mc, kappa_up, kappa_down)` (src/Hamiltonian.jl:762-778) for every walker, evaluated on the
    GPU from the walkers' current configurations (explicit kappa arguments must equal them)."""
    if kappa_up is not None:
        ku, kd = mc.kappa_all()
        if not (np.array_equal(np.broadcast_to(kappa_up, ku.shape), ku) and np.array_equal(np.broadcast_to(kappa_down, kd.shape), kd)):
            raise ValueError("getOL evaluates the walkers' resident configurations; load_configuration first")
    return mc._ensure_engine().measure()


def step_(mc: MC, ctx: MCContext) -> None:
    """Carlo's per-run step (external package)"""
    sweep_(mc, ctx)
    ctx.sweeps += 1
    if ctx.is_thermalized():
        measure_(mc, ctx)


def run_(mc: MC, ctx: MCContext, n_steps: int) -> None:
    """n_steps of Carlo's step loop fused on the device (requires sweeps_per_call == 1): :acc and :OL
    go to the device accumulators instead of ctx; read them with `accumulators(mc)`."""
    assert mc.sweeps_per_call == 1, "run_ needs sweeps_per_call == 1"
    eng = mc._ensure_engine()
    eng.sweeps = ctx.sweeps
    eng.sweep(int(n_steps), ctx.thermalization_sweeps)
    ctx.sweeps += int(n_steps)


def accumulators(mc: MC, reduce_ranks: bool = True) -> Dict[str, float]:
    """Sums of :acc / :OL over all walkers (and, if torch.distributed is initialised, over all
    ranks: the only inter-GPU exchange of the path) plus the derived means."""
    eng = mc._ensure_engine()
    if reduce_ranks and eng.comm_info()[1] > 1:
        v = eng.accumulators_allreduce()                 # NCCL inside the C ABI (dist.init_comm set it up)
    else:
        v = eng.accumulators().copy()
        if reduce_ranks:
            from .dist import allreduce_sum
            v = allreduce_sum(v, device=eng.device)
    out = {
        "walker_sweeps": v[_lib.ACC_WALKER_SWEEPS], "sum_acc": v[_lib.ACC_SUM_ACC], "sum_OL": v[_lib.ACC_SUM_OL],
        "sum_OL2": v[_lib.ACC_SUM_OL2], "n_OL": v[_lib.ACC_N_OL], "n_reach": v[_lib.ACC_N_REACH],
        "n_refresh": v[_lib.ACC_N_REFRESH], "n_singular": v[_lib.ACC_N_SINGULAR],
    }
    out["acc"] = out["sum_acc"] / out["walker_sweeps"] if out["walker_sweeps"] else float("nan")
    out["OL"] = out["sum_OL"] / out["n_OL"] if out["n_OL"] else float("nan")
    out["energy"] = out["OL"] / mc.ns
    return out


def register_evaluables(mc_type, evaluator: Evaluator, params: Dict) -> None:
    """`Carlo.register_evaluables(::Type{MC}, eval, params)` (src/MonteCarlo.jl:675-687)"""
    nsites = params["n1"] * params["n2"] * 3
    evaluator.evaluate_("energy", ("OL",), lambda OL: OL / nsites)


def write_checkpoint(mc: MC, out) -> None:
    """`Carlo.write_checkpoint(mc, out::HDF5.Group)` (src/MonteCarlo.jl:715-719): datasets `kappa_up`,
    `kappa_down` (Int64; [ns] for one walker like the reference, [n_walkers, ns] when batched).  `out` is
    any mutable mapping (an h5py group works).  Added: `rng_state` [n_walkers, 4] since the device owns the streams."""
    if mc.engine is not None:
        ku, kd = mc.kappa_all()
        if mc.n_walkers == 1:
            ku, kd = ku[0], kd[0]
        out["kappa_up"] = ku.astype(np.int64)
        out["kappa_down"] = kd.astype(np.int64)
        out["rng_state"] = mc.engine.get_rng()
    else:
        out["kappa_up"] = np.asarray(mc._kappa_up, dtype=np.int64)
        out["kappa_down"] = np.asarray(mc._kappa_down, dtype=np.int64)


def save_checkpoint_npz(mc: MC, path: str, ctx: Optional[MCContext] = None) -> None:
    """The checkpoint group as a file: `.npz` with the reference's dataset names and dtypes (`kappa_up`, `kappa_down`
    Int64, src/MonteCarlo.jl:715-719) plus `rng_state` (uint64 [n_walkers, 4]) and, when a context is given, `sweeps`
    (ctx.sweeps, which Carlo keeps in its own part of the HDF5 file).  libhdf5 is not in this image; the Julia side
    (julia/KagomeDSLB200.jl) writes the same datasets into Carlo's HDF5 group."""
    group: Dict[str, np.ndarray] = {}
    write_checkpoint(mc, group)
    if ctx is not None:
        group["sweeps"] = np.asarray(ctx.sweeps, dtype=np.int64)
    np.savez(path, **group)


def load_checkpoint_npz(mc: MC, path: str, ctx: Optional[MCContext] = None, defer: bool = False) -> None:
    """inverse of save_checkpoint_npz (`read_checkpoint!`, src/MonteCarlo.jl:751-755)"""
    with np.load(path) as f:
        group = {k: f[k] for k in f.files}
    if ctx is not None and "sweeps" in group:
        ctx.sweeps = int(group["sweeps"])
    read_checkpoint_(mc, group, defer=defer)


def read_checkpoint_(mc: MC, inp, defer: bool = False) -> None:
    """`Carlo.read_checkpoint!(mc, in)` (src/MonteCarlo.jl:751-755).  Deviation: the
    reference leaves W = 0 after a resume; here W is recomputed from the restored configuration when it is
    put on the GPU (immediately, or at the first engine use if defer=True)."""
    ku = np.asarray(inp["kappa_up"], dtype=np.int64)
    kd = np.asarray(inp["kappa_down"], dtype=np.int64)
    mc._kappa_up, mc._kappa_down = (ku, kd) if ku.ndim == 1 else (ku[0], kd[0])
    states = np.asarray(inp["rng_state"], dtype=np.uint64) if "rng_state" in inp else None
    mc._pending = (ku, kd, states)
    if mc.engine is not None:
        mc.engine.close()
        mc.engine = None
    mc._acc_w_seen = None                               # the new engine's accumulators start from zero
    if not defer:
        mc._ensure_engine()
