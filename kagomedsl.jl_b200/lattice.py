"""DoubleKagome lattice geometry (host side, cold path).

Mirrors the reference's `src/Lattice.jl`: `DoubleKagome(t, n1, n2, PBC; antiPBC)` (:68-92),
`ns(lat)` (:97) and `validate_boundary_conditions` (:3-15).  Pure geometry; nothing here runs on
the GPU.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import List, Tuple

import numpy as np


class AbstractLattice:
    """`abstract type AbstractLattice` (src/Lattice.jl:1)"""


def validate_boundary_conditions(PBC: Tuple[bool, bool], antiPBC: Tuple[bool, bool]) -> None:
    """src/Lattice.jl:3-15 -- antiperiodic requires periodic (ArgumentError -> ValueError)."""
    for i, (pbc, apbc) in enumerate(zip(PBC, antiPBC)):
        if apbc and not pbc:
            direction = "first" if i == 0 else "second"
            raise ValueError(
                f"Invalid boundary conditions in {direction} direction: "
                "Cannot have antiperiodic boundary conditions without periodic boundary conditions"
            )


@dataclass
class DoubleKagome(AbstractLattice):
    """Double-unit-cell Kagome lattice, 6 sites per cell, `ns = 3*n1*n2` sites (n1 even).

    Positional signature follows the reference: DoubleKagome(t, n1, n2, PBC, antiPBC=(False, False)).
    """

    t: float
    n1: int
    n2: int
    PBC: Tuple[bool, bool]
    antiPBC: Tuple[bool, bool] = (False, False)
    trunc: float = math.inf
    a1: np.ndarray = field(init=False, repr=False)
    a2: np.ndarray = field(init=False, repr=False)
    r: List[np.ndarray] = field(init=False, repr=False)

    def __post_init__(self):
        self.PBC = (bool(self.PBC[0]), bool(self.PBC[1]))
        self.antiPBC = (bool(self.antiPBC[0]), bool(self.antiPBC[1]))
        validate_boundary_conditions(self.PBC, self.antiPBC)
        assert self.n1 % 2 == 0, "n1 must be even in DoubleKagome"
        t = float(self.t)
        a = 2.0 * t
        self.a1 = np.array([2.0 * a, 0.0])
        self.a2 = np.array([0.5 * a, 0.5 * math.sqrt(3.0) * a])
        self.r = [
            np.array([0.0, 0.0]),
            0.25 * self.a1,
            0.5 * self.a2,
            0.5 * self.a1,
            0.75 * self.a1,
            np.array([2.5 * t, 0.5 * math.sqrt(3.0) * t]),
        ]


def ns(lat: DoubleKagome) -> int:
    """src/Lattice.jl:97"""
    return lat.n1 * lat.n2 * 3
