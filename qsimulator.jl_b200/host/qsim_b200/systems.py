"""Subsystem models: QSpecs, QSystems and their Hamiltonians.

Host-side mirror of /root/reference/src/systems.jl for the systems on the
time-evolution hot path:
  specs        systems.jl:30-62        DuffingSpec(TransmonSpec, phi) :79-83
  QSystems     systems.jl:91-166
  hamiltonian  Resonator :178-179, duffing_hamiltonian :189-194,
               DuffingTransmon :199-206, PerturbativeTransmon :213-216,
               RotatingFrameSystem :242-252, subtract_number! :259-265
  fit_transmon, TransmonSpec(DuffingSpec) :70-73, :288-323 (perturbative model)
ChargeBasisTransmon drives are out of scope: they need a
101x101 eigensolve per RHS call.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import numpy as np

from .transmon_series import (PERTURBATIVE_NUM_TERMS, perturbative_transmon_anharm,
                              perturbative_transmon_freq)


# ---- specs ---------------------------------------------------------------
@dataclass(frozen=True)
class HermitianSpec:
    matrix: np.ndarray

    def __post_init__(self):
        m = np.asarray(self.matrix)
        if m.ndim != 2 or m.shape[0] != m.shape[1] or not np.array_equal(m, m.conj().T):
            raise AssertionError("HermitianSpec needs a Hermitian matrix")

    def __eq__(self, other):
        return isinstance(other, HermitianSpec) and np.array_equal(self.matrix, other.matrix)

    def __hash__(self):
        return hash(np.asarray(self.matrix).tobytes())


@dataclass(frozen=True)
class ResonatorSpec:
    frequency: float


@dataclass(frozen=True)
class TransmonSpec:
    EC: float
    EJ1: float
    EJ2: float


@dataclass(frozen=True)
class DuffingSpec:
    frequency: float
    anharmonicity: float

    @staticmethod
    def from_transmon(t: TransmonSpec, phi: float = 0.0, num_terms: int = PERTURBATIVE_NUM_TERMS) -> "DuffingSpec":
        return DuffingSpec(perturbative_transmon_freq(t.EC, t.EJ1, t.EJ2, phi, num_terms),
                           perturbative_transmon_anharm(t.EC, t.EJ1, t.EJ2, phi, num_terms))


def fit_transmon(freq_max: float, freq_min: float, anharm_max: float,
                 num_terms: int = PERTURBATIVE_NUM_TERMS) -> TransmonSpec:
    """TransmonSpec whose perturbative model has 0-1 frequency `freq_max` and anharmonicity `anharm_max` at
    phi = 0 and 0-1 frequency `freq_min` at phi = 1/2; `freq_max == freq_min` enforces EJ2 = 0
    (systems.jl:288-323, PerturbativeTransmon model).  The reference minimises the L1 misfit with Optim's
    Nelder-Mead from the guesses below; the same conditions are solved here as equations (scipy least_squares
    from the same starting point), which meets the reference's test tolerances (test_systems.jl:49-92)."""
    from scipy.optimize import least_squares

    EC0 = -anharm_max
    EJ_max0 = -(freq_max - anharm_max) ** 2 / (8 * anharm_max)
    EJ_min0 = -(freq_min - anharm_max) ** 2 / (8 * anharm_max)
    tol = dict(xtol=1e-15, ftol=1e-15, gtol=1e-15)
    if freq_max == freq_min:
        def res(p):
            return [perturbative_transmon_freq(p[0], p[1], 0.0, 0.0, num_terms) - freq_max,
                    perturbative_transmon_anharm(p[0], p[1], 0.0, 0.0, num_terms) - anharm_max]
        sol = least_squares(res, [EC0, 0.5 * (EJ_max0 + EJ_min0)], **tol)
        return TransmonSpec(float(sol.x[0]), float(sol.x[1]), 0.0)

    def res(p):
        return [perturbative_transmon_freq(p[0], p[1], p[2], 0.0, num_terms) - freq_max,
                perturbative_transmon_anharm(p[0], p[1], p[2], 0.0, num_terms) - anharm_max,
                perturbative_transmon_freq(p[0], p[1], p[2], 0.5, num_terms) - freq_min]
    sol = least_squares(res, [EC0, 0.5 * (EJ_max0 + EJ_min0), 0.5 * abs(EJ_max0 - EJ_min0)], **tol)
    return TransmonSpec(*(float(x) for x in sol.x))


def transmon_spec_from_duffing(d: DuffingSpec, num_terms: int = PERTURBATIVE_NUM_TERMS) -> TransmonSpec:
    """TransmonSpec(::DuffingSpec) (systems.jl:70-73): a fixed-frequency transmon fitted to (frequency, anharmonicity)."""
    return fit_transmon(d.frequency, d.frequency, d.anharmonicity, num_terms)


# ---- systems -------------------------------------------------------------
class QSystem:
    label: str
    dimension: int


def label(q) -> str:
    return q.label


def dimension(q) -> int:
    return q.dimension


def spec(q):
    return q.spec


@dataclass(frozen=True)
class LiteralHermitian(QSystem):
    label: str
    spec: HermitianSpec

    @property
    def dimension(self) -> int:
        return np.asarray(self.spec.matrix).shape[0]


@dataclass(frozen=True)
class Resonator(QSystem):
    label: str
    dimension: int
    spec: ResonatorSpec


@dataclass(frozen=True)
class DuffingTransmon(QSystem):
    label: str
    dimension: int
    spec: DuffingSpec


@dataclass(frozen=True)
class PerturbativeTransmon(QSystem):
    label: str
    dimension: int
    spec: TransmonSpec
    num_terms: int = PERTURBATIVE_NUM_TERMS


class RotatingFrameSystem(QSystem):
    """A system viewed in a frame rotating at `rotation_rate` (systems.jl:155-165).  Constructed like the
    reference, `RotatingFrameSystem(system, rotation_rate)` (the label is the wrapped system's), or with an
    explicit label, `RotatingFrameSystem(label, system, rotation_rate)`."""
    __slots__ = ("label", "system", "rotation_rate")

    def __init__(self, *args):
        if len(args) == 2:
            system, rotation_rate = args
            lbl = system.label
        elif len(args) == 3:
            lbl, system, rotation_rate = args
        else:
            raise TypeError("RotatingFrameSystem(system, rotation_rate) or (label, system, rotation_rate)")
        object.__setattr__(self, "label", lbl)
        object.__setattr__(self, "system", system)
        object.__setattr__(self, "rotation_rate", float(rotation_rate))

    def __setattr__(self, name, value):
        raise AttributeError("RotatingFrameSystem is immutable")

    def __eq__(self, other):
        return (isinstance(other, RotatingFrameSystem) and self.label == other.label and self.system == other.system
                and self.rotation_rate == other.rotation_rate)

    def __hash__(self):
        return hash((self.label, self.system, self.rotation_rate))

    def __repr__(self):
        return f"RotatingFrameSystem({self.label!r}, {self.system!r}, {self.rotation_rate!r})"

    @staticmethod
    def wrap(q: QSystem, rotation_rate: float) -> "RotatingFrameSystem":
        return RotatingFrameSystem(q, rotation_rate)

    @property
    def dimension(self) -> int:
        return self.system.dimension

    @property
    def spec(self):
        return self.system.spec


# ---- hamiltonians --------------------------------------------------------
def duffing_hamiltonian(freq: float, anharm: float, dim: int) -> np.ndarray:
    """E_n = (freq - 0.5*anharm)*n + 0.5*anharm*n^2 on the diagonal (systems.jl:189-194)."""
    return np.diag([(freq - 0.5 * anharm) * n + 0.5 * anharm * n ** 2 for n in range(dim)])


def _subtract_number(h: np.ndarray, r: float) -> None:
    if r != 0:
        for i in range(h.shape[0]):
            h[i, i] -= i * r


def hamiltonian(q, arg: Optional[float] = None) -> np.ndarray:
    """hamiltonian(q) / hamiltonian(q, freq-or-flux); CompositeQSystem handled in composite.py."""
    from .composite import CompositeQSystem, composite_hamiltonian
    if isinstance(q, CompositeQSystem):
        return composite_hamiltonian(q)
    if isinstance(q, LiteralHermitian):
        if arg is not None:
            raise TypeError("LiteralHermitian has no parametric hamiltonian")
        return np.asarray(q.spec.matrix)
    if isinstance(q, Resonator):
        return np.diag([q.spec.frequency * n for n in range(q.dimension)]).astype(float)
    if isinstance(q, DuffingTransmon):
        f = q.spec.frequency if arg is None else arg
        return duffing_hamiltonian(f, q.spec.anharmonicity, q.dimension)
    if isinstance(q, PerturbativeTransmon):
        s = DuffingSpec.from_transmon(q.spec, 0.0 if arg is None else arg, q.num_terms)
        return duffing_hamiltonian(s.frequency, s.anharmonicity, q.dimension)
    if isinstance(q, RotatingFrameSystem):
        h = np.array(hamiltonian(q.system, arg), dtype=float)
        _subtract_number(h, q.rotation_rate)
        return h
    raise TypeError(f"no hamiltonian for {type(q).__name__}")
