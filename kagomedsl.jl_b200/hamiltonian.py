"""Mean-field spinon Hamiltonian, orbitals and bond table (host side, cold path), plus the
integer helpers of the local-energy expansion.

Mirrors the reference's `src/Hamiltonian.jl`.  The tight-binding matrix is assembled directly
(site x link-table entry, O(ns)) instead of the reference's all-pairs scan (:270-280), but
reproduces it entry for entry: same link keys looked up in the s1 < s2 direction, same
periodic images and antiperiodic signs (:77-123), same Peierls phase (:163-172), same `+=`
accumulation, and bonds whose images cancel vanish from `get_nn` (:370-373).
"""
from __future__ import annotations

import cmath
from typing import Dict, Iterable, List, Sequence, Tuple

import numpy as np

from .lattice import AbstractLattice, DoubleKagome, ns as _ns

# src/Hamiltonian.jl:219-240
pi_link_in: Dict[Tuple[int, int], int] = {
    (1, 2): 1, (1, 3): 1, (2, 3): 1, (2, 4): -1, (4, 6): 1, (4, 5): 1, (5, 6): 1,
    (2, 1): 1, (3, 1): 1, (3, 2): 1, (4, 2): -1, (6, 4): 1, (5, 4): 1, (6, 5): 1,
}
# src/Hamiltonian.jl:250-261
pi_link_inter: Dict[Tuple[int, int, int, int], int] = {
    (3, 5, -1, 1): -1, (3, 1, 0, 1): -1, (6, 2, 0, 1): -1, (6, 4, 0, 1): 1, (5, 1, 1, 0): 1,
    (1, 5, -1, 0): 1, (1, 3, 0, -1): -1, (2, 6, 0, -1): -1, (4, 6, 0, -1): 1, (5, 3, 1, -1): -1,
}
# scripts/zero_flux.jl:15-42 (uniform-RVB / zero-flux ansatz)
zero_link_in = {k: 1 for k in pi_link_in}
zero_link_inter = {k: 1 for k in pi_link_inter}


def unitcell_coord(lat: AbstractLattice, s: int) -> np.ndarray:
    """src/Hamiltonian.jl:28-38 (s is 1-based)"""
    n1 = lat.n1 // 2
    n2 = lat.n2
    nsites = n1 * n2 * 6
    assert 1 <= s <= nsites, f"s should be in the range of 1 to ns, got: {s}"
    uc = (s - 1) // 6
    return (uc % n1) * lat.a1 + (uc // n1) * lat.a2


def unitcell_diff(lat: AbstractLattice, c1: Sequence[float], c2: Sequence[float]) -> Tuple[int, int]:
    """src/Hamiltonian.jl:56-76"""
    d0, d1 = c1[0] - c2[0], c1[1] - c2[1]
    a1, a2 = lat.a1, lat.a2
    det = a1[0] * a2[1] - a1[1] * a2[0]
    dx = int(np.rint((a2[1] * d0 - a2[0] * d1) / det))
    dy = int(np.rint((-a1[1] * d0 + a1[0] * d1) / det))
    return dx, dy


def get_site_coord(lat: AbstractLattice, s: int) -> np.ndarray:
    """src/Hamiltonian.jl:279-283"""
    return unitcell_coord(lat, s) + lat.r[(s - 1) % 6]


def get_boundary_shifts(lat: AbstractLattice, s1: int, s2: int) -> List[Tuple[int, int, float]]:
    """src/Hamiltonian.jl:99-153: displacement of s2's cell from s1's plus its periodic images"""
    assert s1 != s2, f"s1 and s2 should not be the same, got: {s1} and {s2}"
    PBC1, PBC2 = lat.PBC
    anti1, anti2 = lat.antiPBC
    n1, n2 = lat.n1 // 2, lat.n2
    nsites = n1 * n2 * 6
    assert 1 <= s1 <= nsites, f"s1 should be in the range of 1 to ns, got: {s1} in {nsites}"
    assert 1 <= s2 <= nsites, f"s2 should be in the range of 1 to ns, got: {s2} in {nsites}"
    dx, dy = unitcell_diff(lat, unitcell_coord(lat, s2), unitcell_coord(lat, s1))
    shifts: List[Tuple[int, int, float]] = [(dx, dy, 1.0)]
    if not PBC1 and not PBC2:
        return shifts
    for sh1 in ((-n1, 0, n1) if PBC1 else (0,)):
        for sh2 in ((-n2, 0, n2) if PBC2 else (0,)):
            if sh1 == 0 and sh2 == 0:
                continue
            sign = 1.0
            if anti1 and sh1 != 0:
                sign *= -1.0
            if anti2 and sh2 != 0:
                sign *= -1.0
            item = (dx + sh1, dy + sh2, sign)
            if item not in shifts:
                shifts.append(item)
    return shifts


def apply_boundary_conditions_(tunneling: np.ndarray, lat: AbstractLattice, s1: int, s2: int,
                               link_inter: Dict, B: float) -> None:
    """`apply_boundary_conditions!` (src/Hamiltonian.jl:177-207); tunneling is modified in place"""
    nsites = (lat.n1 // 2) * lat.n2 * 6
    assert 1 <= s1 <= nsites and 1 <= s2 <= nsites, "site index out of range"
    cell1, cell2 = (s1 - 1) // 6 + 1, (s2 - 1) // 6 + 1
    assert cell1 != cell2, "s1 and s2 should not be in the same cell"
    label1, label2 = (s1 - 1) % 6 + 1, (s2 - 1) % 6 + 1
    r1 = get_site_coord(lat, s1)
    r_uc_1 = unitcell_coord(lat, s1)
    dr_2 = get_site_coord(lat, s2) - unitcell_coord(lat, s2)
    for dx, dy, sign in get_boundary_shifts(lat, s1, s2):
        key = (label1, label2, dx, dy)
        if key in link_inter:
            r2 = r_uc_1 + dx * lat.a1 + dy * lat.a2 + dr_2
            phase = (B / 2) * (r1[0] + r2[0]) * (r2[1] - r1[1])
            tunneling[s1 - 1, s2 - 1] += sign * link_inter[key] * cmath.exp(1j * phase)


def Hmat(lat: DoubleKagome, link_in: Dict = None, link_inter: Dict = None, B: float = 0.0) -> np.ndarray:
    """`Hmat(lat; link_in, link_inter, B)` (src/Hamiltonian.jl:308-354): ns x ns complex Hermitian
    hopping matrix, H = -(T + T') with T strictly upper triangular."""
    link_in = pi_link_in if link_in is None else link_in
    link_inter = pi_link_inter if link_inter is None else link_inter
    n1h, n2 = lat.n1 // 2, lat.n2
    ncell = n1h * n2
    nsites = ncell * 6
    PBC1, PBC2 = lat.PBC
    anti1, anti2 = lat.antiPBC
    T = np.zeros((nsites, nsites), dtype=np.complex128)
    site_xy = np.array([get_site_coord(lat, s + 1) for s in range(nsites)]) if nsites else np.zeros((0, 2))
    for c in range(ncell):                                   # in-cell links (:254-268)
        for l1 in range(1, 7):
            for l2 in range(l1 + 1, 7):
                if (l1, l2) in link_in:
                    s1, s2 = c * 6 + l1 - 1, c * 6 + l2 - 1
                    r1, r2 = site_xy[s1], site_xy[s2]
                    phase = (B / 2) * (r1[0] + r2[0]) * (r2[1] - r1[1])
                    T[s1, s2] = link_in[(l1, l2)] * cmath.exp(1j * phase)
    for c in range(ncell):                                   # inter-cell links (:270-280 via :145-174)
        cx, cy = c % n1h, c // n1h
        uc1 = cx * lat.a1 + cy * lat.a2
        for (l1, l2, ddx, ddy), val in link_inter.items():
            tx, ty = cx + ddx, cy + ddy
            # the target cell is reached directly (shift 0) or through ONE periodic image
            # (shift -+n); the image exists only in a periodic direction (:96-97)
            wrap_x = wrap_y = 0
            if tx < 0 or tx >= n1h:
                if not PBC1 or n1h == 0:
                    continue
                wrap_x = 1
                tx %= n1h
            if ty < 0 or ty >= n2:
                if not PBC2 or n2 == 0:
                    continue
                wrap_y = 1
                ty %= n2
            if (wrap_x or wrap_y) and not (PBC1 or PBC2):
                continue
            c2 = ty * n1h + tx
            if c2 == c:                                      # same-cell pairs are skipped (:274)
                continue
            s1, s2 = c * 6 + l1 - 1, c2 * 6 + l2 - 1
            if s1 >= s2:                                     # only the s1 < s2 key is consulted (:276)
                continue
            # a direct (unwrapped) displacement in a direction where |d| == period could also be
            # an image: handled because each (ddx, ddy) resolves to exactly one (cell, shift)
            sign = 1.0
            if anti1 and wrap_x:
                sign *= -1.0
            if anti2 and wrap_y:
                sign *= -1.0
            r1 = site_xy[s1]
            r2 = uc1 + ddx * lat.a1 + ddy * lat.a2 + (site_xy[s2] - (tx * lat.a1 + ty * lat.a2))
            phase = (B / 2) * (r1[0] + r2[0]) * (r2[1] - r1[1])
            T[s1, s2] += sign * val * cmath.exp(1j * phase)
    if np.any(np.tril(T, -1) != 0):
        raise RuntimeError("tunneling matrix must be upper triangular")
    return -(T + T.conj().T)


def orbitals(H_mat: np.ndarray, N_up: int, N_down: int):
    """src/Hamiltonian.jl:382-393: lowest-N eigenvectors of Hermitian(H_mat) as columns.
    A real H (B = 0) is diagonalised as real symmetric so that U, hence W, is real FP64."""
    if np.iscomplexobj(H_mat) and np.abs(H_mat.imag).max(initial=0.0) == 0.0:
        w, v = np.linalg.eigh(np.ascontiguousarray(H_mat.real))
    else:
        w, v = np.linalg.eigh(H_mat)
    p = np.argsort(w, kind="stable")
    v = v[:, p]
    return np.asfortranarray(v[:, :N_up]), np.asfortranarray(v[:, :N_down])


def get_nn(H_mat: np.ndarray) -> List[Tuple[int, int]]:
    """src/Hamiltonian.jl:447-451: `findall(!iszero, UpperTriangular(H))` -- 1-based (i, j) pairs in
    column-major order"""
    jj, ii = np.nonzero(np.triu(np.asarray(H_mat)).T)
    return [(int(i) + 1, int(j) + 1) for i, j in zip(ii, jj)]


class Hamiltonian:
    """`struct Hamiltonian` (src/Hamiltonian.jl:420-427) with both reference constructors:
    Hamiltonian(N_up, N_down, lat; link_in, link_inter, B)            (:399-411)
    Hamiltonian(N_up, N_down, U_up, U_down, H_mat, nn)                (field constructor)"""

    def __init__(self, N_up: int, N_down: int, *args, link_in=None, link_inter=None, B: float = 0.0):
        self.N_up, self.N_down = int(N_up), int(N_down)
        if len(args) == 1:
            lat = args[0]
            self.H_mat = Hmat(lat, link_in=link_in, link_inter=link_inter, B=B)
            self.U_up, self.U_down = orbitals(self.H_mat, self.N_up, self.N_down)
            self.nn = get_nn(self.H_mat)
        elif len(args) == 4:
            U_up, U_down, H_mat, nn = args
            self.U_up, self.U_down = np.asarray(U_up), np.asarray(U_down)
            self.H_mat = np.asarray(H_mat)
            self.nn = [tuple(int(x) for x in b) for b in nn]
        else:
            raise TypeError("Hamiltonian(N_up, N_down, lat; ...) or Hamiltonian(N_up, N_down, U_up, U_down, H_mat, nn)")

    def gap(self) -> float:
        """Fermi-level gap of the mean-field spectrum at filling N_up (closed shell <=> gap > 0)"""
        w = np.linalg.eigvalsh(self.H_mat)
        n = self.N_up
        return float(w[n] - w[n - 1]) if 0 < n < len(w) else float("inf")


def is_occupied(kappa: Sequence[int], l: int) -> bool:
    """src/MonteCarlo.jl:132-135 (l is a 1-based site index)"""
    if not 1 <= l <= len(kappa):
        raise IndexError(f"BoundsError: attempt to access {len(kappa)}-element vector at index [{l}]")
    return kappa[l - 1] != 0


def Sz(i: int, kappa_up: Sequence[int], kappa_down: Sequence[int]) -> float:
    """src/Hamiltonian.jl:531-565"""
    n = len(kappa_up)
    if not 1 <= i <= n:
        raise IndexError(f"BoundsError: attempt to access {n}-element vector at index [{i}]")
    if len(kappa_down) != n:
        raise ValueError(f"DimensionMismatch: kappa_up and kappa_down must have same length, got {n} and {len(kappa_down)}")
    up, down = kappa_up[i - 1] != 0, kappa_down[i - 1] != 0
    if up and not down:
        return 0.5
    if not up and down:
        return -0.5
    if up and down:
        raise ValueError(f"ArgumentError: Site {i} is doubly occupied, with kappa_up: {list(kappa_up)} and kappa_down: {list(kappa_down)}")
    raise ValueError(f"ArgumentError: Site {i} is unoccupied, with kappa_up: {list(kappa_up)} and kappa_down: {list(kappa_down)}")


def SzInteraction_(xprime: Dict, kappa_up, kappa_down, i: int, j: int) -> None:
    """`SzInteraction!` (src/Hamiltonian.jl:593-604)"""
    key = (-1, -1, -1, -1)
    xprime[key] = xprime.get(key, 0.0) + Sz(i, kappa_up, kappa_down) * Sz(j, kappa_up, kappa_down)


def spinInteraction_(xprime: Dict, kappa_up, kappa_down, i: int, j: int) -> None:
    """`spinInteraction!` (src/Hamiltonian.jl:636-672): keys (K_up, l_up, K_down, l_down) += -1/2"""
    i_up, j_up = kappa_up[i - 1], kappa_up[j - 1]
    i_down, j_down = kappa_down[i - 1], kappa_down[j - 1]
    if j_up != 0 and i_down != 0:
        key = (i, int(j_up), j, int(i_down))
        xprime[key] = xprime.get(key, 0.0) - 1.0 / 2.0
    if i_up != 0 and j_down != 0:
        key = (j, int(i_up), i, int(j_down))
        xprime[key] = xprime.get(key, 0.0) - 1.0 / 2.0


def getxprime(Ham: Hamiltonian, kappa_up, kappa_down) -> Dict[Tuple[int, int, int, int], float]:
    """src/Hamiltonian.jl:711-720 (host utility; on the GPU this expansion is fused into k_measure)"""
    xprime: Dict[Tuple[int, int, int, int], float] = {}
    for (i, j) in Ham.nn:
        spinInteraction_(xprime, kappa_up, kappa_down, i, j)
        SzInteraction_(xprime, kappa_up, kappa_down, i, j)
    return xprime
