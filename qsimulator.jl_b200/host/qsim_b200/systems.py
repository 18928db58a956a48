"""Subsystem models: QSpecs, QSystems and their Hamiltonians.

Host-side mirror of /root/reference/src/systems.jl for the systems on the
time-evolution hot path:
  specs        systems.jl:30-62        DuffingSpec(TransmonSpec, phi) :79-83
  QSystems     systems.jl:91-166
  hamiltonian  Resonator :178-179, duffing_hamiltonian :189-194,
               DuffingTransmon :199-206, PerturbativeTransmon :213-216,
               RotatingFrameSystem :242-252, subtract_number! :259-265
  fit_transmon, TransmonSpec(DuffingSpec) :70-73, :288-323 (both transmon models)
  ChargeBasisTransmon :131-139, :220-229; DiagonalChargeBasisTransmon :143-152, :231-237
A flux drive on a DiagonalChargeBasisTransmon costs the reference a 101x101 eigensolve per RHS call; here the
level energies are tabulated once per problem as functions of cos(pi*phi) (`charge_spectrum_tables`) and the
device evaluates the tables (QSIM_TERM_SPECTRUM in include/qsim_b200.h).
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


def fit_transmon(freq_max: float, freq_min: float, anharm_max: float, model=None,
                 num_terms: Optional[int] = None) -> TransmonSpec:
    """TransmonSpec whose model (`PerturbativeTransmon`, the default, or `DiagonalChargeBasisTransmon`) has 0-1
    frequency `freq_max` and anharmonicity `anharm_max` at phi = 0 and 0-1 frequency `freq_min` at phi = 1/2;
    `freq_max == freq_min` enforces EJ2 = 0 (systems.jl:288-323).  The reference minimises the L1 misfit with
    Optim's Nelder-Mead from the guesses below; the same conditions are solved here as equations (scipy
    least_squares from the same starting point), which meets the reference's test tolerances
    (test_systems.jl:49-92).  Also callable like the reference: fit_transmon(fmax, fmin, anh, Model, num_terms)."""
    from scipy.optimize import least_squares

    if isinstance(model, int):          # fit_transmon(fmax, fmin, anh, num_terms): the perturbative model
        model, num_terms = None, model
    charge = model is DiagonalChargeBasisTransmon
    if model is not None and not charge and model is not PerturbativeTransmon:
        raise TypeError("fit_transmon: model must be PerturbativeTransmon or DiagonalChargeBasisTransmon")
    if num_terms is None:
        num_terms = CHARGE_NUM_TERMS if charge else PERTURBATIVE_NUM_TERMS

    def f01_anh(EC, EJ1, EJ2, phi):
        if not charge:
            return (perturbative_transmon_freq(EC, EJ1, EJ2, phi, num_terms),
                    perturbative_transmon_anharm(EC, EJ1, EJ2, phi, num_terms))
        lev = charge_basis_levels(EC, EJ1, EJ2, phi, num_terms, 3)
        return lev[1] - lev[0], (lev[2] - lev[1]) - (lev[1] - lev[0])

    EC0 = -anharm_max
    EJ_max0 = -(freq_max - anharm_max) ** 2 / (8 * anharm_max)
    EJ_min0 = -(freq_min - anharm_max) ** 2 / (8 * anharm_max)
    tol = dict(xtol=1e-15, ftol=1e-15, gtol=1e-15)
    if freq_max == freq_min:
        def res(p):
            f, a = f01_anh(p[0], p[1], 0.0, 0.0)
            return [f - freq_max, a - anharm_max]
        sol = least_squares(res, [EC0, 0.5 * (EJ_max0 + EJ_min0)], **tol)
        return TransmonSpec(float(sol.x[0]), float(sol.x[1]), 0.0)

    def res(p):
        f0, a0 = f01_anh(p[0], p[1], p[2], 0.0)
        f1, _ = f01_anh(p[0], p[1], p[2], 0.5)
        return [f0 - freq_max, a0 - anharm_max, f1 - freq_min]
    sol = least_squares(res, [EC0, 0.5 * (EJ_max0 + EJ_min0), 0.5 * abs(EJ_max0 - EJ_min0)], **tol)
    return TransmonSpec(*(float(x) for x in sol.x))


def transmon_spec_from_duffing(d: DuffingSpec, num_terms: int = PERTURBATIVE_NUM_TERMS) -> TransmonSpec:
    """TransmonSpec(::DuffingSpec) (systems.jl:70-73): a fixed-frequency transmon fitted to (frequency, anharmonicity)."""
    return fit_transmon(d.frequency, d.frequency, d.anharmonicity, PerturbativeTransmon, num_terms)


# ---- charge-basis transmon spectra -------------------------------------------
CHARGE_NUM_TERMS = 101


def charge_basis_levels(EC: float, EJ1: float, EJ2: float, phi, num_terms: int = CHARGE_NUM_TERMS, n_levels: int = 3,
                        cos_half_flux=None) -> np.ndarray:
    """Lowest `n_levels` eigenvalues of the charge-basis Hamiltonian 4 EC n^2 - EJ(phi)/2 (|n><n+1| + h.c.),
    n = -N..N with 2N + 1 = num_terms, EJ(phi)^2 = EJ1^2 + EJ2^2 + 2 EJ1 EJ2 cos(2 pi phi) (systems.jl:220-237).
    `phi` (or `cos_half_flux` = cos(pi phi) directly) may be an array: the result is then [..., n_levels]."""
    from scipy.linalg import eigvalsh_tridiagonal
    if cos_half_flux is None:
        y = np.cos(2 * np.pi * np.asarray(phi, dtype=float))
    else:
        y = 2.0 * np.asarray(cos_half_flux, dtype=float) ** 2 - 1.0
    N = num_terms // 2
    diag = 4.0 * EC * (np.arange(-N, N + 1, dtype=float) ** 2)
    EJ = np.sqrt(np.maximum(EJ1 * EJ1 + EJ2 * EJ2 + 2.0 * EJ1 * EJ2 * y, 0.0))
    out = np.empty(y.shape + (n_levels,))
    for idx in np.ndindex(y.shape):
        off = np.full(num_terms - 1, -0.5 * EJ[idx])
        out[idx] = eigvalsh_tridiagonal(diag, off, select="i", select_range=(0, n_levels - 1))
    return out


def charge_spectrum_tables(EC: float, EJ1: float, EJ2: float, num_terms: int, n_levels: int, n_coef: int = 14):
    """One `signals.Table` per level: E_k as a function of z = cos(pi phi) on [-1, 1].  The spectrum depends on
    the flux only through EJ(phi)^2 = (EJ1 - EJ2)^2 + 4 EJ1 EJ2 z^2 and is a smooth even function of z (in
    cos(2 pi phi) it has a square-root-like edge at half flux for nearly symmetric junctions); sampled to the accuracy of the eigenvalue computation itself (1e-15 of the largest
    charging energy 4 EC N^2, a few 1e-12).  Built once per problem; the device evaluates E_k(cos(pi s(t))) from
    them instead of diagonalising per right-hand-side call."""
    from .signals import Table
    cache = {}

    def levels(ys):
        ys = np.asarray(ys, dtype=float)
        key = ys.tobytes()
        if key not in cache:
            cache.clear()
            cache[key] = charge_basis_levels(EC, EJ1, EJ2, None, num_terms, n_levels, cos_half_flux=ys)
        return cache[key]

    atol = 4e-15 * max(4.0 * abs(EC) * (num_terms // 2) ** 2, abs(EJ1) + abs(EJ2), 1.0)
    return [Table.from_function(lambda ys, k=k: levels(ys)[..., k], -1.0, 1.0, n_coef=n_coef, tol=1e-14, atol=atol,
                                max_seg=1 << 10)
            for k in range(n_levels)]


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


@dataclass(frozen=True)
class ChargeBasisTransmon(QSystem):
    """A transmon in the charge basis, n = -N..N (systems.jl:131-139); `dimension` (odd) is the basis size."""
    label: str
    dimension: int
    spec: TransmonSpec

    def __post_init__(self):
        if self.dimension % 2 != 1:
            raise AssertionError("ChargeBasisTransmon needs an odd dimension")


@dataclass(frozen=True)
class DiagonalChargeBasisTransmon(QSystem):
    """The lowest `dimension` levels of a charge-basis transmon of `num_terms` charge states, as a diagonal
    Hamiltonian (systems.jl:143-152)."""
    label: str
    dimension: int
    spec: TransmonSpec
    num_terms: int = CHARGE_NUM_TERMS

    def __post_init__(self):
        if self.num_terms % 2 != 1 or self.dimension > self.num_terms:
            raise AssertionError("DiagonalChargeBasisTransmon needs an odd num_terms >= dimension")


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
    if isinstance(q, ChargeBasisTransmon):
        d, sp = q.dimension, q.spec
        N = d // 2
        phi = 0.0 if arg is None else arg
        EJ = np.sqrt(sp.EJ1 ** 2 + sp.EJ2 ** 2 + 2 * sp.EJ1 * sp.EJ2 * np.cos(2 * np.pi * phi))
        return (4 * sp.EC * np.diag(np.arange(-N, N + 1, dtype=float) ** 2)
                - 0.5 * EJ * (np.diag(np.ones(d - 1), -1) + np.diag(np.ones(d - 1), 1)))
    if isinstance(q, DiagonalChargeBasisTransmon):
        sp = q.spec
        return np.diag(charge_basis_levels(sp.EC, sp.EJ1, sp.EJ2, 0.0 if arg is None else arg, q.num_terms,
                                           q.dimension))
    if isinstance(q, RotatingFrameSystem):
        h = np.array(hamiltonian(q.system, arg), dtype=float)
        _subtract_number(h, q.rotation_rate)
        return h
    raise TypeError(f"no hamiltonian for {type(q).__name__}")
