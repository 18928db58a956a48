"""Pin the CPU oracles before they may judge kernels ("oracle acceptance").

The reference ships no golden vectors and cannot run here (no Julia), so the pins are
 (1) its own hot-path known-answer tests at their tolerances
     (/root/reference/test/test_time_evolution.jl, test/integration_tests.jl),
 (2) the Runge-Kutta order conditions on the restated Tsit5 constants,
 (3) agreement with tight-tolerance solves / expm within the solver's expected global error,
 (4) bitwise independence of the final state from the save grid (test_time_evolution.jl:75-78),
 (5) agreement between the closure-driven numpy restatement and the descriptor-driven C one.
"""
import math

import numpy as np
import pytest
from scipy.linalg import block_diag, expm
from scipy.special import jv

from oracle import c_oracle as co
from oracle import np_oracle as npo
from qsim_b200 import (CompositeQSystem, Cos, DuffingSpec, DuffingTransmon, HermitianSpec, LiteralHermitian, Param,
                       Sin, X_Y, _abi, add_hamiltonian, add_lindblad, decay, dephasing, dipole_drive,
                       parametric_drive, rwa_dipole)
from qsim_b200 import workloads as W

UP, US, MP, MS = co.UNITARY_PROPAGATOR, co.UNITARY_STATE, co.ME_PROPAGATOR, co.ME_STATE


# ---- (2) order conditions -------------------------------------------------------
def test_tsit5_order_conditions():
    c, a, bt, r = co.tsit5_constants()
    assert np.array_equal(c, np.array(npo.C)) and np.array_equal(bt, np.array(npo.BTILDE))
    for i in range(7):
        assert np.array_equal(a[i, :i], np.array(npo.A[i]))
    A = np.zeros((7, 7))
    A[:, :6] = a
    b = A[6]                                   # FSAL: the 7th row is b
    assert np.allclose(A.sum(axis=1), c, atol=1e-15)
    one = np.ones(7)
    assert b @ one == pytest.approx(1, abs=1e-14)
    assert b @ c == pytest.approx(1 / 2, abs=1e-14)
    assert b @ c ** 2 == pytest.approx(1 / 3, abs=1e-14)
    assert b @ c ** 3 == pytest.approx(1 / 4, abs=1e-14)
    assert b @ c ** 4 == pytest.approx(1 / 5, abs=1e-14)
    assert b @ (A @ c) == pytest.approx(1 / 6, abs=1e-14)
    assert b @ (A @ c ** 2) == pytest.approx(1 / 12, abs=1e-14)
    assert b @ (c * (A @ c)) == pytest.approx(1 / 8, abs=1e-14)
    assert b @ (A @ (A @ c)) == pytest.approx(1 / 24, abs=1e-14)
    assert b @ (A @ c ** 3) == pytest.approx(1 / 20, abs=1e-14)
    assert b @ (c ** 2 * (A @ c)) == pytest.approx(1 / 10, abs=1e-14)
    assert b @ (A @ (A @ (A @ c))) == pytest.approx(1 / 120, abs=1e-14)
    # embedded estimator: btilde = b - bhat sums to zero and kills orders <= 4 quadrature
    assert bt.sum() == pytest.approx(0, abs=1e-15)
    for p in (1, 2, 3):
        assert bt @ c ** p == pytest.approx(0, abs=1e-14)
    # dense output: b(1) = b, and quadrature conditions at interior thetas
    assert np.allclose(npo.interp_weights(1.0), b, atol=1e-14)
    for th in (0.3, 0.7):
        w = np.array(npo.interp_weights(th))
        assert w.sum() == pytest.approx(th, abs=1e-14)
        assert w @ c == pytest.approx(th ** 2 / 2, abs=1e-14)
        assert w @ c ** 2 == pytest.approx(th ** 3 / 3, abs=1e-14)
        assert w @ c ** 3 == pytest.approx(th ** 4 / 4, abs=1e-14)


# ---- (1) the reference's known-answer tests --------------------------------------
def _literal_system():
    m1 = np.array([[0, 1], [1, 0]], dtype=complex)
    m2 = np.array([[0, -1j, 0], [1j, .1, 2 - 1j], [0, 2 + 1j, 0]])
    mat = block_diag(m1, m2, np.array([[3.0]]))
    lh = LiteralHermitian("label", HermitianSpec(mat))
    cqs = CompositeQSystem([lh])
    add_hamiltonian(cqs, lh)
    return cqs, mat


def test_unitary_propagator_time_independent():
    """test_time_evolution.jl:5-23, atol=1e-4 per matrix."""
    cqs, mat = _literal_system()
    ts = np.linspace(0, 10, 100)
    us, st = co.OracleProblem(cqs).solve(UP, ts)
    assert st["retcode"][0] == 0 and st["n_reject"][0] == 0
    for u, t in zip(us[0], ts):
        assert np.allclose(u, expm(-2j * np.pi * mat * t), atol=1e-4, rtol=0)
    us_np, st_np = npo.unitary_propagator(cqs, ts)
    assert st_np.n_accept == st["n_accept"][0]
    assert max(np.abs(a - b).max() for a, b in zip(us_np, us[0])) < 1e-12


def _iswap_system():
    """test_time_evolution.jl:25-50."""
    freq_fixed, anharm, drive_freq = 4.0, -0.2, 0.3
    q0 = DuffingTransmon("q0", 3, DuffingSpec(0.0, anharm))
    q1 = DuffingTransmon("q1", 3, DuffingSpec(freq_fixed, anharm))
    cqs = CompositeQSystem([q0, q1])
    add_hamiltonian(cqs, q1)
    a, b = freq_fixed + drive_freq, drive_freq
    add_hamiltonian(cqs, parametric_drive(q0, Cos(amp=b, freq=drive_freq, offset=a)), q0)
    g = 0.01
    add_hamiltonian(cqs, .5 * g * X_Y([q0, q1]), [q0, q1])
    g_eff = g * jv(1, b / drive_freq)
    return cqs, g_eff, 1 / (4 * g_eff)


def test_unitary_state_time_dependent():
    """test_time_evolution.jl:52-62: populations vs RWA, rtol=.03 on the vector norm."""
    cqs, g_eff, t_final = _iswap_system()
    ts = np.linspace(0, t_final, 25)
    psi10 = np.zeros(9, dtype=complex)
    psi10[3] = 1  # |1,0>: index 1*3+0
    psis, st = co.OracleProblem(cqs).solve(US, ts, init=psi10)
    pop10 = np.abs(psis[0][:, 3]) ** 2
    pop01 = np.abs(psis[0][:, 1]) ** 2
    rwa = np.cos(2 * np.pi * g_eff * ts) ** 2
    assert np.linalg.norm(pop10 - rwa) <= .03 * max(np.linalg.norm(pop10), np.linalg.norm(rwa))
    assert np.linalg.norm(pop01 - (1 - rwa)) <= .03 * max(np.linalg.norm(pop01), np.linalg.norm(1 - rwa))


def test_propagator_state_consistency_and_saveat_independence():
    """test_time_evolution.jl:65-79: populations agree to 1e-4; the scalar-t helper is BITWISE equal
    to the last entry of the full solve, i.e. the step sequence does not depend on saveat."""
    cqs, g_eff, t_final = _iswap_system()
    P = co.OracleProblem(cqs)
    ts = np.linspace(0, t_final, 3)
    us, _ = P.solve(UP, ts)
    psi0 = np.zeros(9, dtype=complex)
    psi0[3] = 1
    psis, _ = P.solve(US, ts, init=psi0)
    pop_state = np.abs(psis[0][:, 3]) ** 2
    pop_prop = np.array([abs(np.vdot(psi0, u @ psi0)) ** 2 for u in us[0]])
    assert np.allclose(pop_state, pop_prop, rtol=1e-4)
    u_end, _ = P.solve(UP, np.array([0.0, ts[-1]]))
    assert np.array_equal(u_end[0, -1], us[0, -1])
    psi_end, _ = P.solve(US, np.array([0.0, ts[-1]]), init=psi0)
    assert np.array_equal(psi_end[0, -1], psis[0, -1])
    us_dense, _ = P.solve(UP, np.linspace(0, t_final, 57))
    assert np.array_equal(us_dense[0, -1], us[0, -1])


def _driven_decay_system():
    """test_time_evolution.jl:121-135."""
    q0 = DuffingTransmon("q0", 3, DuffingSpec(5.0, -0.2))
    cqs = CompositeQSystem([q0])
    add_hamiltonian(cqs, q0)
    add_hamiltonian(cqs, dipole_drive(q0, Cos(amp=0.02, freq=5.0)), q0)
    add_lindblad(cqs, decay(q0, 1 / (2 * np.pi * 50.)), [q0])
    return cqs


def test_me_propagator_vs_me_state():
    """test_time_evolution.jl:137-152, rtol=1e-4; scalar-t helpers bitwise equal; faithful (dense
    Kronecker) and structured RHS agree."""
    cqs = _driven_decay_system()
    P = co.OracleProblem(cqs)
    ts = np.linspace(0, 100, 101)
    rho0 = np.zeros((3, 3), dtype=complex)
    rho0[0, 0] = 1
    us, st = P.solve(MP, ts, mode=co.FAITHFUL)
    rhos, _ = P.solve(MS, ts, init=rho0)
    for u, rho in zip(us[0], rhos[0]):
        rp = (u @ rho0.ravel(order="F")).reshape(3, 3, order="F")
        assert np.linalg.norm(rp - rho) <= 1e-4 * max(np.linalg.norm(rp), np.linalg.norm(rho))
    u_end, _ = P.solve(MP, np.array([0.0, 100.0]), mode=co.FAITHFUL)
    assert np.array_equal(u_end[0, -1], us[0, -1])
    rho_end, _ = P.solve(MS, np.array([0.0, 100.0]), init=rho0)
    assert np.array_equal(rho_end[0, -1], rhos[0, -1])
    us_s, st_s = P.solve(MP, ts, mode=co.STRUCTURED)
    assert st_s["n_accept"][0] == st["n_accept"][0]
    assert np.abs(us_s - us).max() < 1e-10   # roundoff over ~2.3e4 steps
    # trace preservation: vec(I)^T S = vec(I)^T
    vi = np.eye(3).ravel(order="F")
    assert np.abs(vi @ us[0, -1] - vi).max() < 1e-5


def test_me_numpy_closure_restatement_agrees():
    cqs = _driven_decay_system()
    P = co.OracleProblem(cqs)
    ts = np.linspace(0, 5, 6)
    us, st = P.solve(MP, ts, mode=co.FAITHFUL)
    us_np, st_np = npo.me_propagator(cqs, ts)
    assert st_np.n_accept == st["n_accept"][0] and st_np.n_reject == st["n_reject"][0]
    assert max(np.abs(a - b).max() for a, b in zip(us_np, us[0])) < 1e-12
    rho0 = np.diag([0.5, 0.5, 0]).astype(complex)
    rhos, _ = P.solve(MS, ts, init=rho0)
    rhos_np, _ = npo.me_state(cqs, ts, rho0)
    assert max(np.abs(a - b).max() for a, b in zip(rhos_np, rhos[0])) < 1e-12


def test_decay_and_dephasing_known_answer():
    """test_time_evolution.jl:155-182: rho22 vs exp(-t/T1)/2 at rtol 1e-10, rho12 vs exp(-t/T2)/2 at 1e-2."""
    q0 = DuffingTransmon("q0", 3, DuffingSpec(5.0, -0.2))
    cqs = CompositeQSystem([q0])
    add_hamiltonian(cqs, q0)
    T1, Tphi = 50.0, 200.0
    T2 = 1 / (.5 / T1 + 1 / Tphi)
    add_lindblad(cqs, decay(q0, 1 / (2 * np.pi * T1)), [q0])
    add_lindblad(cqs, dephasing(q0, 1 / (2 * np.pi * Tphi)), [q0])
    ts = np.linspace(0, 100, 101)
    psi0 = np.array([1, 1, 0], dtype=complex) / math.sqrt(2)
    rhos, st = co.OracleProblem(cqs).solve(MS, ts, init=np.outer(psi0, psi0.conj()))
    assert st["retcode"][0] == 0
    assert np.allclose(rhos[0][:, 1, 1], np.exp(-ts / T1) / 2, rtol=1e-10, atol=0)
    assert np.allclose(np.abs(rhos[0][:, 0, 1]), np.exp(-ts / T2) / 2, rtol=1e-2, atol=0)


def test_ramsey_and_rabi():
    """integration_tests.jl:8-27 (Ramsey, rtol 1e-3/1e-4) and :32-63 (Rabi lab + RWA, rtol 1e-2)."""
    q0 = DuffingTransmon("q0", 3, DuffingSpec(5.0, -0.2))
    cqs = CompositeQSystem([q0])
    add_hamiltonian(cqs, q0)
    ts = np.linspace(0, 1, 201)
    psi = np.array([1, 1, 0], dtype=complex) / math.sqrt(2)
    P = co.OracleProblem(cqs)
    psis, _ = P.solve(US, ts, init=psi)
    expected = 0.5 + 0.5 * np.cos(2 * np.pi * 5.0 * ts)
    sig = np.real(psis[0] @ psi.conj())
    assert np.linalg.norm(sig - expected) <= 1e-3 * np.linalg.norm(expected)
    us, _ = P.solve(UP, ts)
    sig = np.array([np.real(np.vdot(psi, u @ psi)) for u in us[0]])
    assert np.linalg.norm(sig - expected) <= 1e-4 * np.linalg.norm(expected)

    nut = 0.02
    ts = np.linspace(0, 100, 101)
    g_expected = 0.5 + 0.5 * np.cos(2 * np.pi * nut * ts)
    for factory in (dipole_drive, rwa_dipole):
        c2 = CompositeQSystem([q0])
        add_hamiltonian(c2, q0)
        add_hamiltonian(c2, factory(q0, Cos(amp=nut, freq=5.0)), q0)
        psis, st = co.OracleProblem(c2).solve(US, ts, init=np.array([1, 0, 0], dtype=complex))
        g = np.abs(psis[0][:, 0]) ** 2
        e = np.abs(psis[0][:, 1]) ** 2
        assert np.linalg.norm(g - g_expected) <= 1e-2 * np.linalg.norm(g_expected)
        assert np.linalg.norm(e - (1 - g_expected)) <= 1e-2 * np.linalg.norm(g_expected)


# ---- (3) tight-tolerance truth and (5) numpy vs C on the benchmark system --------------
def test_cfg1_short_against_tight_solve_and_numpy():
    cqs, params, _, _ = W.config("cfg1")
    ts = np.arange(0.0, 4.0)
    P = co.OracleProblem(cqs)
    us, st = P.solve(UP, ts, params, mode=co.FAITHFUL)
    us_np, st_np = npo.unitary_propagator(cqs, ts, params=params[0])
    assert (st_np.n_accept, st_np.n_reject, st_np.n_rhs) == (st["n_accept"][0], st["n_reject"][0], st["n_rhs"][0])
    assert st["n_rhs"][0] == 3 + 6 * (st["n_accept"][0] + st["n_reject"][0])
    assert st_np.dt_init == st["dt_init"][0]
    assert max(np.abs(a - b).max() for a, b in zip(us_np, us[0])) < 1e-11
    tight, _ = P.solve(UP, ts, params, opts=_abi.default_opts(reltol=1e-13, abstol=1e-15))
    err = np.abs(us - tight).max()
    assert 1e-7 < err < 5e-4          # the reference tolerances' own global error scale
    u = tight[0, -1]
    assert np.abs(u.conj().T @ u - np.eye(9)).max() < 1e-10


def test_sweep_threads_and_per_trajectory_ts():
    cqs, _, _, _ = W.config("cfg2")
    params = W.amp_freq_grid(3, 2)
    P = co.OracleProblem(cqs)
    ts = np.array([0.0, 0.5, 1.5])
    a, sa = P.solve(UP, ts, params, n_threads=1)
    b, sb = P.solve(UP, ts, params, n_threads=4)
    assert np.array_equal(a, b) and np.array_equal(sa, sb)
    ts2 = np.stack([ts * (1 + 0.1 * i) for i in range(len(params))])
    c, _ = P.solve(UP, ts2, params)
    assert np.array_equal(c[0], a[0]) and not np.array_equal(c[1], a[1])
    with pytest.raises(ValueError):
        P.solve(UP, np.array([0.0, 2.0, 1.0]), params)
