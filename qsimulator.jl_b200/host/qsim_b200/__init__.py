"""qsim_b200: host-side mirror of QSimulator.jl's interface for the time-evolution hot path.

Same names and argument meaning as the reference (Julia's `!` suffix dropped):
systems (systems.jl), operators and drive factories (operators.jl), CompositeQSystem
assembly (composite_systems.jl) and the four solvers (time_evolution.jl:29-164), which call
the CUDA shared library through the C ABI of include/qsim_b200.h.  There is no CPU fallback.
"""
from .signals import Param, Signal, ComplexSignal, Const, Cos, Sin, Table, tabulated
from .systems import (HermitianSpec, ResonatorSpec, TransmonSpec, DuffingSpec, QSystem, LiteralHermitian,
                      Resonator, DuffingTransmon, PerturbativeTransmon, RotatingFrameSystem, label, dimension,
                      spec, hamiltonian, duffing_hamiltonian, fit_transmon, transmon_spec_from_duffing,
                      ChargeBasisTransmon, DiagonalChargeBasisTransmon, charge_basis_levels, charge_spectrum_tables)
from .transmon_series import (PERTURBATIVE_NUM_TERMS, xi_effective, perturbative_transmon_freq,
                              perturbative_transmon_anharm)
from .operators import (raising, lowering, number, X, Y, X_Y, XY, flip_flop, decay, dephasing, dipole_drive,
                        rwa_dipole, parametric_drive, xy_exchange_drive, scaled_operator, ParametricHamiltonian,
                        ParametricLindblad)
from .composite import (CompositeQSystem, add_hamiltonian, add_lindblad, embed, embed_indices, embed_add,
                        find_indices, add_parametric_hamiltonians, lower, ProblemSpec)
from . import _abi
from ._ffi import Context, Problem, QsimError, PinnedArray, default_context
from .time_evolution import unitary_propagator, unitary_state, me_propagator, me_state
from .sweep import shard_indices, shard_params, scatter_back, gather_results
from .postprocess import (populations, subspace_indices, subspace_block, average_gate_fidelity, superoperator,
                          kraus_mix)
from .floquet import (floquet_propagator, floquet_rise_fall_propagator, floquet_sweep, choose_times_floquet,
                      decompose_times_floquet, unique_tol)
