"""kagomedsl.jl_b200 -- B200-native walker-batched VMC sampling path of hz-xiaxz/KagomeDSL.jl.

Exports follow the reference module (`src/KagomeDSL.jl:12,15,18`): DoubleKagome, Hamiltonian, Sz,
spinInteraction!, MC, MCContext, tilde_U (Julia's `f!` is spelled `f_`).
"""
from .lattice import AbstractLattice, DoubleKagome, ns, validate_boundary_conditions
from .hamiltonian import (Hamiltonian, Hmat, Sz, SzInteraction_, apply_boundary_conditions_, get_boundary_shifts,
                          get_nn, get_site_coord, getxprime, is_occupied, orbitals, pi_link_in, pi_link_inter,
                          spinInteraction_, unitcell_coord, unitcell_diff, zero_link_in, zero_link_inter)
from .montecarlo import (MC, AbstractMC, Engine, Evaluator, MCContext, Z, accumulators, find_initial_configuration_,
                         getOL, init_, init_conf_qr, measure_, read_checkpoint_, reevaluateW_, register_evaluables,
                         run_, step_, sweep_, tilde_U, write_checkpoint, save_checkpoint_npz, load_checkpoint_npz)
from ._lib import KdslError, SingularException
from .rng import Xoshiro, walker_states
from . import dist

__all__ = [n for n in dir() if not n.startswith("_")]
