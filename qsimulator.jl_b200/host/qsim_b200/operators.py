"""Operators and drive factories.

Host-side mirror of /root/reference/src/operators.jl (SURVEY.md section 8 rows
a14, a15): ladder operators :32-86, X/Y/X_Y :94-139, flip_flop/XY :148-170,
decay :185-187, dephasing :201-203, and the drive factories dipole_drive
:220-226, rwa_dipole :242-250, parametric_drive :265-268.

The factories return `ParametricHamiltonian` objects instead of closures: they
are callable like the reference closures (`ham(t)` -> small matrix) so the CPU
oracle can use them as-is, and they carry the analytic descriptor the CUDA
library needs (closures cannot cross the C ABI).
"""
from __future__ import annotations

import cmath
import math
from dataclasses import dataclass
from functools import reduce
from typing import Callable, List, Optional, Sequence, Union

import numpy as np

from .signals import ComplexSignal, Const, Param, Scalar, Signal, max_param_index, resolve
from .systems import (DiagonalChargeBasisTransmon, DuffingTransmon, PerturbativeTransmon, QSystem, RotatingFrameSystem,
                      charge_spectrum_tables,
                      dimension, hamiltonian)

TERM_SCALAR, TERM_ROTATING, TERM_DUFFING, TERM_TRANSMON, TERM_SPECTRUM = 0, 1, 2, 3, 4


def _dim(q) -> int:
    return q if isinstance(q, (int, np.integer)) else dimension(q)


# ---- primitives ------------------------------------------------------------
def raising(q, phi: float = 0.0) -> np.ndarray:
    """m[i+1,i] = sqrt(i+1)*exp(2*pi*i*phi) (operators.jl:32-39)."""
    dim = _dim(q)
    m = np.zeros((dim, dim), dtype=complex)
    fac = cmath.exp(1j * 2 * math.pi * phi)
    for i in range(1, dim):
        m[i, i - 1] = math.sqrt(i) * fac
    return m


def lowering(q, phi: float = 0.0, factor: float = 1.0) -> np.ndarray:
    """m[i,i+1] = sqrt(i+1)*exp(-2*pi*i*phi)*factor (operators.jl:62-69)."""
    dim = _dim(q)
    m = np.zeros((dim, dim), dtype=complex)
    efac = cmath.exp(-1j * 2 * math.pi * phi) * factor
    for i in range(1, dim):
        m[i - 1, i] = math.sqrt(i) * efac
    return m


def number(q, factor: float = 1.0) -> np.ndarray:
    """diag(0, 1, ..) * factor (operators.jl:80-86)."""
    dim = _dim(q)
    m = np.zeros((dim, dim), dtype=complex)
    for i in range(1, dim):
        m[i, i] = i * factor
    return m


def _kron_all(ms: Sequence[np.ndarray]) -> np.ndarray:
    return reduce(np.kron, ms)


def X(q, phi=None) -> np.ndarray:
    """X = a + a^dagger; with phases, raising(q,phi)+lowering(q,phi); lists give tensor products
    (operators.jl:94-121)."""
    if isinstance(q, (list, tuple)):
        if phi is None:
            return _kron_all([X(qi) for qi in q])
        return _kron_all([X(qi, p) for qi, p in zip(q, phi)])
    if phi is None:
        dim = _dim(q)
        de = np.sqrt(np.arange(1, dim))
        return np.diag(de, -1) + np.diag(de, 1)
    return raising(q, phi) + lowering(q, phi)


def Y(q, phi=None) -> np.ndarray:
    """Y = -i a + i a^dagger (operators.jl:128-136)."""
    if isinstance(q, (list, tuple)):
        if phi is None:
            return _kron_all([Y(qi) for qi in q])
        return _kron_all([Y(qi, p) for qi, p in zip(q, phi)])
    if phi is None:
        dim = _dim(q)
        de = np.sqrt(np.arange(1, dim))
        return np.diag(1j * de, -1) + np.diag(-1j * de, 1)
    return 1j * (raising(q, phi) - lowering(q, phi))


def X_Y(qs, phis=None) -> np.ndarray:
    """X(qs[,phis]) + Y(qs[,phis]) (operators.jl:138-139)."""
    return X(qs, phis) + Y(qs, phis)


def flip_flop(a, b, phi: Optional[float] = None) -> np.ndarray:
    """a b^dagger + a^dagger b (operators.jl:148).  The phased method reproduces the
    reference as written, including its `lowering(b) (x) raising(b)` second factor
    (operators.jl:155)."""
    if phi is None:
        return np.kron(raising(a), lowering(b)) + np.kron(lowering(a), raising(b))
    return (cmath.exp(-1j * 2 * math.pi * phi) * np.kron(raising(a), lowering(b))
            + cmath.exp(1j * 2 * math.pi * phi) * np.kron(lowering(b), raising(b)))


def XY(a, b, phi: Optional[float] = None) -> np.ndarray:
    return 2.0 * flip_flop(a, b, phi)


def decay(q, gamma: float) -> np.ndarray:
    """sqrt(gamma)*a; T1 = 1/(2*pi*gamma) (operators.jl:185-187)."""
    return lowering(q, 0, math.sqrt(gamma))


def dephasing(q, gamma: float) -> np.ndarray:
    """sqrt(2*gamma)*n; Tphi = 1/(2*pi*gamma) (operators.jl:201-203)."""
    return number(q, math.sqrt(2 * gamma))


# ---- time-dependent terms ---------------------------------------------------
@dataclass
class LoweredTerm:
    """One qsim_term of the C ABI, before embedding."""
    kind: int
    signal: Signal
    atom_a: Optional[np.ndarray] = None   # small operator (SCALAR / ROTATING)
    atom_b: Optional[np.ndarray] = None   # small operator (ROTATING)
    rate: Scalar = 0.0
    anharm: Scalar = 0.0
    rot: Scalar = 0.0
    EC: Scalar = 0.0
    EJ1: Scalar = 0.0
    EJ2: Scalar = 0.0
    num_terms: int = 0
    dim: int = 0                           # level count (DUFFING / TRANSMON / SPECTRUM)
    level_tables: Optional[list] = None    # SPECTRUM: one signals.Table per level, E_k(cos(pi phi)) on [-1, 1]

    def max_param_index(self) -> int:
        return max_param_index(self.signal, self.rate, self.anharm, self.rot, self.EC, self.EJ1, self.EJ2)


class ParametricHamiltonian:
    """A time-dependent subsystem Hamiltonian: callable `ham(t[, params])` like the
    reference's closures, plus `lowered()` descriptors for the CUDA library."""

    def __init__(self, fn: Callable, terms: List[LoweredTerm]):
        self._fn = fn
        self._terms = terms

    def __call__(self, t, params=None) -> np.ndarray:
        return self._fn(t, params)

    def lowered(self) -> List[LoweredTerm]:
        return self._terms

    def max_param_index(self) -> int:
        return max([t.max_param_index() for t in self._terms], default=-1)


def _as_signal(drive) -> Signal:
    if isinstance(drive, Signal):
        return drive
    if isinstance(drive, (int, float, Param)):
        return Const(drive)
    raise TypeError(
        "drive must be a qsim_b200 Signal (Cos/Sin/Const/...): opaque callables cannot cross the C ABI; "
        "wrap an arbitrary closure with tabulated(fn, t_start, t_end)")


def dipole_drive(q, drive, rotation_rate: Scalar = 0.0) -> ParametricHamiltonian:
    """t -> drive(t) * X(q, rotation_rate*t) (operators.jl:220-226)."""
    sig = _as_signal(drive)

    def ham(t, params=None):
        pulse = sig(t, params)
        return pulse * X(q, resolve(rotation_rate, params) * t)

    if not isinstance(rotation_rate, Param) and rotation_rate == 0:
        terms = [LoweredTerm(TERM_SCALAR, sig, atom_a=raising(q) + lowering(q))]
    else:
        terms = [LoweredTerm(TERM_ROTATING, sig, atom_a=raising(q), atom_b=lowering(q), rate=rotation_rate)]
    return ParametricHamiltonian(ham, terms)


def rwa_dipole(q, drive) -> ParametricHamiltonian:
    """t -> real(drive(t))*X(q) + imag(drive(t))*Y(q) (operators.jl:242-250)."""
    x_ham, y_ham = X(q), Y(q)
    if isinstance(drive, ComplexSignal):
        re, im = drive.re, drive.im
        terms = [LoweredTerm(TERM_SCALAR, re, atom_a=x_ham.astype(complex)),
                 LoweredTerm(TERM_SCALAR, im, atom_a=y_ham)]
    else:
        re, im = _as_signal(drive), None
        terms = [LoweredTerm(TERM_SCALAR, re, atom_a=x_ham.astype(complex))]

    def ham(t, params=None):
        h = re(t, params) * x_ham
        return h + (im(t, params) * y_ham if im is not None else 0.0 * y_ham)

    return ParametricHamiltonian(ham, terms)


def scaled_operator(op: np.ndarray, drive) -> ParametricHamiltonian:
    """t -> drive(t) * op for a fixed operator `op` on the subsystems the term is added to: the generic raw
    closure `add_hamiltonian!(cqs, t -> f(t) * M, qs)` (composite_systems.jl:40-50).  With drive = Param(i)
    (or Const(Param(i))) this is a *static* term whose strength is swept per trajectory, e.g. a coupling
    g * X([q0, q1]) with g a sweep column."""
    sig = _as_signal(drive)
    mat = np.asarray(op, dtype=complex)

    def ham(t, params=None):
        return sig(t, params) * mat

    return ParametricHamiltonian(ham, [LoweredTerm(TERM_SCALAR, sig, atom_a=mat)])


def parametric_drive(q, drive) -> ParametricHamiltonian:
    """t -> hamiltonian(q, drive(t)) (operators.jl:265-268): a frequency drive on a
    DuffingTransmon (systems.jl:205-206) or a flux drive on a PerturbativeTransmon
    (systems.jl:213-216), optionally wrapped in a RotatingFrameSystem (:248-252)."""
    sig = _as_signal(drive)
    base, rot = q, 0.0
    if isinstance(q, RotatingFrameSystem):
        base, rot = q.system, q.rotation_rate
    if isinstance(base, DuffingTransmon):
        term = LoweredTerm(TERM_DUFFING, sig, anharm=base.spec.anharmonicity, rot=rot, dim=base.dimension)
    elif isinstance(base, PerturbativeTransmon):
        s = base.spec
        term = LoweredTerm(TERM_TRANSMON, sig, rot=rot, EC=s.EC, EJ1=s.EJ1, EJ2=s.EJ2,
                           num_terms=base.num_terms, dim=base.dimension)
    elif isinstance(base, DiagonalChargeBasisTransmon):
        # systems.jl:231-237: the reference diagonalises the num_terms x num_terms charge-basis Hamiltonian on every
        # call; the level energies depend on the flux only through cos(pi phi) and are tabulated once here
        s = base.spec
        if any(isinstance(x, Param) for x in (s.EC, s.EJ1, s.EJ2)):
            raise TypeError("parametric_drive: the spec of a DiagonalChargeBasisTransmon must be literal "
                            "(its spectrum is tabulated on the host)")
        term = LoweredTerm(TERM_SPECTRUM, sig, rot=rot, EC=s.EC, EJ1=s.EJ1, EJ2=s.EJ2, num_terms=base.num_terms,
                           dim=base.dimension,
                           level_tables=charge_spectrum_tables(s.EC, s.EJ1, s.EJ2, base.num_terms, base.dimension))
    else:
        raise TypeError(f"parametric_drive: no hamiltonian(q, x) kernel form for {type(base).__name__}")

    def ham(t, params=None):
        return hamiltonian(q, sig(t, params))

    return ParametricHamiltonian(ham, [term])


def xy_exchange_drive(qs, coupling: Scalar, rates: Sequence[Scalar]) -> ParametricHamiltonian:
    """t -> coupling * X_Y(qs, rates .* t): the raw closures of benchmarks.jl:184,193 and
    integration_tests.jl:97.  For two systems X_Y([a,b],[ta,tb]) =
    2(e^{2 pi i (ta-tb)} a'b + h.c.), a single rotating term at rate ra - rb."""
    if len(qs) != 2:
        raise ValueError("xy_exchange_drive needs exactly two systems")
    qa, qb = qs
    ra, rb = rates
    if isinstance(ra, Param) or isinstance(rb, Param):
        raise TypeError("xy_exchange_drive: frame rates must be literals")

    def ham(t, params=None):
        return resolve(coupling, params) * X_Y([qa, qb], [ra * t, rb * t])

    a_up = np.kron(raising(qa), lowering(qb))   # a'b
    a_dn = np.kron(lowering(qa), raising(qb))   # ab'
    terms = [LoweredTerm(TERM_ROTATING, Const(coupling), atom_a=2.0 * a_up, atom_b=2.0 * a_dn, rate=ra - rb)]
    return ParametricHamiltonian(ham, terms)


class ParametricLindblad:
    """L(t) = scale(t) * op: time-dependent collapse operator (composite_systems.jl:59-62)."""

    def __init__(self, op: np.ndarray, scale):
        self.op = np.asarray(op, dtype=complex)
        self.scale = _as_signal(scale)

    def __call__(self, t, params=None) -> np.ndarray:
        return self.scale(t, params) * self.op

    def max_param_index(self) -> int:
        return self.scale.max_param_index()
