"""DiagonalChargeBasisTransmon (/root/reference/src/systems.jl:131-152, 220-237).

The reference diagonalises the 101 x 101 charge-basis Hamiltonian inside every right-hand-side call of a flux-driven
DiagonalChargeBasisTransmon.  The library tabulates the level energies once per problem as functions of
cos(pi phi) (QSIM_TERM_SPECTRUM); the C oracle keeps the reference's per-call eigenvalue computation (bisection
on the Sturm sequence), and the closure-driven numpy oracle calls LAPACK through the host mirror -- three
independent evaluations of the same Hamiltonian."""
import numpy as np
import pytest

from oracle import c_oracle as co
from oracle import np_oracle as npo
from qsim_b200 import (ChargeBasisTransmon, CompositeQSystem, Cos, DiagonalChargeBasisTransmon, DuffingSpec,
                       DuffingTransmon, Param, PerturbativeTransmon, RotatingFrameSystem, TransmonSpec, X,
                       add_hamiltonian, charge_basis_levels, charge_spectrum_tables, duffing_hamiltonian,
                       fit_transmon, hamiltonian, parametric_drive)
from qsim_b200._ffi import Problem
from qsim_b200.composite import add_parametric_hamiltonians, composite_hamiltonian

SPEC = TransmonSpec(0.172, 12.71, 3.69)


def _isapprox(a, b, rtol):
    """Julia's isapprox on matrices: norm(a - b) <= rtol * max(norm(a), norm(b))."""
    return np.linalg.norm(a - b) <= rtol * max(np.linalg.norm(a), np.linalg.norm(b))


def test_perturbative_vs_charge_basis():
    """test_systems.jl:38-47."""
    pt = PerturbativeTransmon("q", 3, SPEC)
    ct = DiagonalChargeBasisTransmon("q", 3, SPEC)
    for phi in np.linspace(0.0, 1.0, 10):
        ham_pt, ham_ct = hamiltonian(pt, phi), hamiltonian(ct, phi)
        assert _isapprox(ham_pt + ham_ct[0, 0] * np.eye(3), ham_ct, 1e-5)


def test_charge_basis_hamiltonian_and_constructors():
    """systems.jl:220-229: 4 EC n^2 on the diagonal, -EJ(phi)/2 next to it; :135, :148-149 assertions."""
    h = hamiltonian(ChargeBasisTransmon("q", 5, SPEC), 0.25)
    EJ = np.sqrt(SPEC.EJ1 ** 2 + SPEC.EJ2 ** 2)        # cos(pi/2) = 0
    assert np.allclose(np.diag(h), 4 * SPEC.EC * np.array([4.0, 1.0, 0.0, 1.0, 4.0]))
    assert np.allclose(np.diag(h, 1), -0.5 * EJ) and np.allclose(np.diag(h, -1), -0.5 * EJ)
    full = np.sort(np.linalg.eigvalsh(hamiltonian(ChargeBasisTransmon("q", 101, SPEC), 0.1)))[:4]
    assert np.allclose(np.diag(hamiltonian(DiagonalChargeBasisTransmon("q", 4, SPEC), 0.1)), full, rtol=0, atol=1e-11)
    with pytest.raises(AssertionError):
        ChargeBasisTransmon("q", 4, SPEC)
    with pytest.raises(AssertionError):
        DiagonalChargeBasisTransmon("q", 3, SPEC, num_terms=100)
    with pytest.raises(AssertionError):
        DiagonalChargeBasisTransmon("q", 7, SPEC, num_terms=5)


def test_fit_transmon_charge_basis_model():
    """test_systems.jl:49-92 with model = DiagonalChargeBasisTransmon."""
    ham = duffing_hamiltonian(5.0, -0.2, 3)
    s = fit_transmon(5.0, 5.0, -0.2, DiagonalChargeBasisTransmon)
    assert s.EJ2 == 0.0
    h1 = hamiltonian(DiagonalChargeBasisTransmon("Test", 3, s))
    assert _isapprox(ham + h1[0, 0] * np.eye(3), h1, 1e-7)
    s = fit_transmon(5.0, 4.0, -0.2, DiagonalChargeBasisTransmon, 101)
    assert s.EJ2 != 0.0
    lev0 = np.diag(hamiltonian(DiagonalChargeBasisTransmon("Test", 3, s), 0.0))
    lev5 = np.diag(hamiltonian(DiagonalChargeBasisTransmon("Test", 3, s), 0.5))
    assert lev0[1] - lev0[0] == pytest.approx(5.0, rel=1e-8)
    assert (lev0[2] - lev0[1]) - (lev0[1] - lev0[0]) == pytest.approx(-0.2, rel=1e-7)
    assert lev5[1] - lev5[0] == pytest.approx(4.0, rel=1e-8)


def test_spectrum_tables_match_the_eigenvalues():
    for sp in (SPEC, TransmonSpec(0.2, 15.0, 15.0), TransmonSpec(0.18, 20.0, 0.0)):   # asymmetric, symmetric, fixed
        tabs = charge_spectrum_tables(sp.EC, sp.EJ1, sp.EJ2, 101, 4)
        phis = np.linspace(0.0, 1.0, 257)
        lev = charge_basis_levels(sp.EC, sp.EJ1, sp.EJ2, phis, 101, 4)
        for k, tb in enumerate(tabs):
            assert tb.n_seg <= 256
            assert np.max(np.abs(tb.eval(np.cos(np.pi * phis)) - lev[:, k])) <= 2e-11


def _flux_driven_system(rotating=False, sweep=True):
    """Fixed-frequency Duffing transmon coupled to a flux-modulated DiagonalChargeBasisTransmon (the parametric
    gate of benchmark/benchmarks.jl:120-155 with the tunable transmon in the charge-basis model)."""
    q0 = DuffingTransmon("q0", 3, DuffingSpec(3.94, -0.18))
    q1 = DiagonalChargeBasisTransmon("q1", 3, SPEC)
    if rotating:
        q1 = RotatingFrameSystem(q1, 4.1)
    cqs = CompositeQSystem([q0, q1])
    add_hamiltonian(cqs, q0)
    add_hamiltonian(cqs, 0.006 * X([q0, q1]), [q0, q1])
    amp = Param(0) if sweep else 0.3
    add_hamiltonian(cqs, parametric_drive(q1, Cos(amp=amp, freq=0.12, phase=0.3, offset=0.05)), q1)
    return cqs


@pytest.mark.parametrize("rotating", [False, True])
def test_lowered_spectrum_term_evaluates_like_the_closure(rotating):
    cqs = _flux_driven_system(rotating)
    row = np.array([0.27])
    P, O = Problem(cqs), co.OracleProblem(cqs)
    assert len(P.spec.tables) == 3            # one table per level
    for t in (0.0, 1.3, 4.71, 10.0):
        want = composite_hamiltonian(cqs)
        add_parametric_hamiltonians(want, cqs, t, row)          # LAPACK through the host mirror
        assert np.max(np.abs(O.eval(0, t, row) - want)) <= 1e-11     # bisection on the Sturm sequence
        assert np.max(np.abs(P.eval(0, t, row) - want)) <= 1e-11     # tables


def test_closure_oracle_pins_the_charge_term_oracle():
    cqs = _flux_driven_system(sweep=False)
    ts = np.linspace(0.0, 3.0, 4)
    want, _ = npo.unitary_propagator(cqs, ts)
    got, st = co.OracleProblem(cqs).solve(co.UNITARY_PROPAGATOR, ts, mode=0)
    assert st["retcode"][0] == 0
    assert np.max(np.abs(got[0] - np.asarray(want))) <= 1e-9


def test_spectrum_term_argument_checks():
    from qsim_b200 import _abi
    from qsim_b200._ffi import QsimError
    q1 = DiagonalChargeBasisTransmon("q1", 3, TransmonSpec(Param(0), 12.0, 3.0))
    with pytest.raises(TypeError):
        parametric_drive(q1, Cos(amp=0.1, freq=0.1))
    # a level without a table is refused by the library
    cqs = _flux_driven_system()
    spec = Problem(cqs).spec
    lt = spec.terms[0][0]
    lt.level_tables = lt.level_tables[:2]
    with pytest.raises(QsimError) as e:
        Problem(spec)
    assert e.value.code == _abi.ERR_ARG


@pytest.mark.gpu
@pytest.mark.parametrize("rotating", [False, True])
def test_charge_basis_flux_drive_gpu_parity(rotating):
    cqs = _flux_driven_system(rotating)
    ts = np.linspace(0.0, 12.0, 13)
    params = np.linspace(0.05, 0.35, 6)[:, None]
    P = Problem(cqs)
    got, st = P.solve("unitary_propagator", ts, params=params)
    want, ost = co.OracleProblem(cqs).solve(co.UNITARY_PROPAGATOR, ts, params, mode=0, n_threads=6)
    assert np.all(st["retcode"] == 0)
    assert np.array_equal(st["n_accept"], ost["n_accept"]) and np.array_equal(st["n_reject"], ost["n_reject"])
    assert np.max(np.abs(got - want) / np.maximum(np.abs(want), 1e-3)) <= 1e-8
    from qsim_b200 import _abi
    direct, std = P.solve("unitary_propagator", ts, params=params, opts=_abi.default_opts(flags=_abi.FLAG_DIRECT_DRIVES))
    assert np.array_equal(st["n_accept"], std["n_accept"])
    assert np.max(np.abs(got - direct)) <= 1e-10
