"""Floquet wrappers: host helpers against the reference's own tests
(/root/reference/test/test_time_evolution.jl:81-95,185-196, the jldoctest at src/time_evolution.jl:270-273),
and the wrappers around a propagator function (oracle on CPU, CUDA library under -m gpu)."""
import math

import numpy as np
import pytest
from scipy.special import erf, jv

from oracle import c_oracle as co
from qsim_b200 import (CompositeQSystem, Cos, DuffingSpec, DuffingTransmon, Signal, X_Y, add_hamiltonian,
                       choose_times_floquet, decompose_times_floquet, dipole_drive, floquet_propagator,
                       floquet_rise_fall_propagator, parametric_drive, unique_tol)


def test_unique_tol():
    l0 = np.array([1.0, 2.0, 3.0])
    l2 = np.concatenate([l0, l0 + 1e-5])
    a, inds = unique_tol(l2, 1e-6)
    assert np.array_equal(a, np.round(l2 / 1e-6) * 1e-6) and list(inds) == [0, 1, 2, 3, 4, 5]
    a, inds = unique_tol(l2, 1e-4)
    assert np.allclose(a, l0, atol=1e-12) and list(inds) == [0, 1, 2, 0, 1, 2]


def test_decompose_doctest_and_choose_times():
    # Julia's range 0.25:0.1:1 has exactly-rounded elements; with binary floats use a tolerance
    q, rem, inds = decompose_times_floquet(0.25 + 0.1 * np.arange(8), 0.25, rtol=1e-9)
    assert list(q) == [0, 0, 0, 1, 1, 2, 2, 2]
    assert np.allclose(rem, [0.25, 0.3, 0.35, 0.4, 0.45])
    assert list(inds + 1) == [1, 3, 5, 2, 4, 1, 3, 5]
    # test_time_evolution.jl:81-95
    g_eff = 0.01 * jv(1, 1.0)
    t_final, t_period = 1 / (4 * g_eff), 1 / 0.3
    dt = t_final / 10
    times = choose_times_floquet(t_final / 2, t_final, t_period, dt)
    assert times[0] > 0 and times[-1] <= t_final and t_final / 2 in times
    times = np.concatenate([[0.0], times])
    q, rem, inds = decompose_times_floquet(times, t_period)
    assert len(q) == len(times) and len(rem) <= len(times) and len(inds) == len(times)
    assert np.allclose(t_period * q + rem[inds], times)


def _oracle_propagator(cqs, ts):
    out, st = co.OracleProblem(cqs).solve(co.UNITARY_PROPAGATOR, np.asarray(ts, dtype=float))
    assert st["retcode"][0] == 0
    return list(out[0])


def _rabi():
    q0 = DuffingTransmon("q0", 3, DuffingSpec(5.0, -0.2))
    cqs = CompositeQSystem([q0])
    add_hamiltonian(cqs, q0)
    add_hamiltonian(cqs, dipole_drive(q0, Cos(amp=0.02, freq=5.0)), q0)
    return cqs


def test_floquet_propagator_rabi_cpu():
    """integration_tests.jl:66-75: Floquet with period 1/qubit_freq reproduces the Rabi flops."""
    cqs = _rabi()
    ts = np.linspace(0, 40, 41)
    us = floquet_propagator(_oracle_propagator, 1 / 5.0)(cqs, ts)
    psi = np.array([1, 0, 0], dtype=complex)
    g = np.array([abs((u @ psi)[0]) ** 2 for u in us])
    g_expected = 0.5 + 0.5 * np.cos(2 * np.pi * 0.02 * ts)
    assert np.linalg.norm(g - g_expected) <= 1e-2 * np.linalg.norm(g_expected)
    direct = _oracle_propagator(cqs, ts)
    assert max(np.abs(a - b).max() for a, b in zip(us, direct)) < 2e-4     # both carry the solver's own error


def _pulseshape_system():
    """test_time_evolution.jl:97-113."""
    g = 0.01
    a, b, drive_freq = 4.3, 0.3, 0.3
    t_final = 1 / (4 * g * jv(1, b / drive_freq))
    q0 = DuffingTransmon("q0", 3, DuffingSpec(0.0, -0.2))
    q1 = DuffingTransmon("q1", 3, DuffingSpec(4.0, -0.2))
    cqs = CompositeQSystem([q0, q1])
    add_hamiltonian(cqs, q1)
    add_hamiltonian(cqs, .5 * g * X_Y([q0, q1]), [q0, q1])
    rise_time = .1 * t_final
    fwhm = 0.25 * rise_time
    sigma = 0.5 * fwhm / math.sqrt(2 * math.log(2))
    pulse = Cos(amp=b, freq=drive_freq, offset=a).erf_squared(fwhm, t_final - fwhm, sigma, rest=a + b)
    # the descriptor is the closure of test_time_evolution.jl:107-111
    for t in (0.0, 3.3, 17.0, t_final):
        env = 0.5 * (erf((t - fwhm) / sigma) - erf((t - (t_final - fwhm)) / sigma))
        assert pulse(t) == pytest.approx(env * (a + b * math.cos(2 * math.pi * drive_freq * t)) + (1 - env) * (a + b),
                                         rel=1e-14)
    add_hamiltonian(cqs, parametric_drive(q0, pulse), q0)
    return cqs, t_final, 1 / drive_freq, rise_time


@pytest.mark.gpu
def test_floquet_with_pulseshape_gpu():
    """test_time_evolution.jl:97-119: floquet_rise_fall_propagator(unitary_propagator, ...) vs the direct
    solve, rtol 1e-5 per matrix, 500 times -- both through the CUDA library."""
    from qsim_b200 import unitary_propagator
    cqs, t_final, t_period, rise_time = _pulseshape_system()
    times = np.linspace(0.0, t_final, 500)
    prop = floquet_rise_fall_propagator(unitary_propagator, t_period, rise_time, rise_time)
    us_floquet = prop(cqs, times)
    us_direct = unitary_propagator(cqs, times)
    assert len(us_floquet) == 500
    for uf, ud in zip(us_floquet, us_direct):
        assert np.linalg.norm(uf - ud) <= 1e-5 * max(np.linalg.norm(uf), np.linalg.norm(ud))
    want, _ = co.OracleProblem(cqs).solve(co.UNITARY_PROPAGATOR, times)
    assert np.max(np.abs(us_direct - want[0]) / np.maximum(np.abs(want[0]), 1e-3)) <= 1e-8


@pytest.mark.gpu
def test_floquet_device_path_equals_host_composition():
    """The one-call device path (segments in one batched launch, powers and products on the device) against the
    generic composition of the same library solves on the host: same segments, so agreement to rounding."""
    from qsim_b200 import unitary_propagator
    cqs, t_final, t_period, rise_time = _pulseshape_system()
    times = np.linspace(0.0, t_final, 123)
    dev = floquet_rise_fall_propagator(unitary_propagator, t_period, rise_time, rise_time)(cqs, times)
    host = floquet_rise_fall_propagator(lambda c, t: unitary_propagator(c, t), t_period, rise_time, rise_time)(cqs, times)
    assert dev.shape == host.shape == (123, 9, 9)
    assert np.max(np.abs(dev - host)) < 1e-11
    # plain floquet_propagator on a strictly periodic system (integration_tests.jl:66-75)
    cq = _rabi()
    ts = np.linspace(0, 40, 41)
    dev = floquet_propagator(unitary_propagator, 1 / 5.0)(cq, ts)
    host = floquet_propagator(lambda c, t: unitary_propagator(c, t), 1 / 5.0)(cq, ts)
    assert np.max(np.abs(dev - host)) < 1e-11
    direct = unitary_propagator(cq, ts)
    assert np.max(np.abs(dev - direct)) < 2e-4


@pytest.mark.gpu
def test_floquet_sweep_cfg2_per_trajectory_periods():
    """cfg2 through one period + powers: the drive frequency is a sweep axis, so every point has its own period
    1/f.  200 ns of lab-frame evolution become one 7.5-9.1 ns period per point; agreement with the direct solve at
    the reference's own Floquet tolerance (test_time_evolution.jl:118: 1e-5 per matrix), speed-up printed."""
    import time
    from qsim_b200 import Context, floquet_sweep, unitary_propagator
    from qsim_b200 import workloads as W
    cqs, params, _, _ = W.config("cfg2")
    pr = np.ascontiguousarray(params[np.linspace(0, len(params) - 1, 296).astype(int)])
    ts = np.arange(0.0, 201.0, 4.0)
    ctx = Context(0)
    floquet_sweep(cqs, ts[:3], 1.0 / pr[:4, 1], params=pr[:4], ctx=ctx)      # builds the device tables
    t0 = time.perf_counter()
    fl, st = floquet_sweep(cqs, ts, 1.0 / pr[:, 1], params=pr, ctx=ctx, return_stats=True)
    t_fl = time.perf_counter() - t0
    t0 = time.perf_counter()
    direct, sd = unitary_propagator(cqs, ts, params=pr, ctx=ctx, return_stats=True)
    t_direct = time.perf_counter() - t0
    assert fl.shape == direct.shape == (296, len(ts), 9, 9)
    assert np.all(st["retcode"] == 0)
    for a, b in zip(fl.reshape(-1, 9, 9), direct.reshape(-1, 9, 9)):
        assert np.linalg.norm(a - b) <= 1e-5 * max(np.linalg.norm(a), np.linalg.norm(b))
    assert st["n_accept"].sum() < 0.1 * sd["n_accept"].sum()
    print(f"cfg2 Floquet sweep: {t_direct / t_fl:.1f}x faster than the direct solve "
          f"({int(sd['n_accept'].sum())} -> {int(st['n_accept'].sum())} steps), max |diff| {np.abs(fl - direct).max():.2e}")


@pytest.mark.gpu
def test_floquet_me_propagator():
    """The wrappers compose whatever the propagator function returns: superoperators of a driven, decaying transmon."""
    from qsim_b200 import add_lindblad, decay, floquet_sweep, me_propagator
    q0 = DuffingTransmon("q0", 3, DuffingSpec(5.0, -0.2))
    cqs = CompositeQSystem([q0])
    add_hamiltonian(cqs, q0)
    add_hamiltonian(cqs, dipole_drive(q0, Cos(amp=0.02, freq=5.0)), q0)
    add_lindblad(cqs, decay(q0, 1e-3), [q0])
    ts = np.linspace(0.0, 12.0, 25)
    # 60 periods: the per-period solver error adds up coherently in U(tau)^n, so compare at tight solver tolerances
    # (the composition itself is exact) and, at the reference's tolerances, at the level those imply
    from qsim_b200 import _abi
    tight = _abi.default_opts()
    tight.reltol, tight.abstol = 1e-11, 1e-13
    fl = floquet_sweep(cqs, ts, 1 / 5.0, which="me_propagator", opts=tight)
    direct = me_propagator(cqs, ts, opts=tight)
    assert fl.shape == (25, 9, 9)
    for a, b in zip(fl, direct):
        assert np.linalg.norm(a - b) <= 1e-8 * max(np.linalg.norm(a), np.linalg.norm(b))
    fl = floquet_sweep(cqs, ts, 1 / 5.0, which="me_propagator")
    direct = me_propagator(cqs, ts)
    for a, b in zip(fl, direct):
        assert np.linalg.norm(a - b) <= 5e-4 * max(np.linalg.norm(a), np.linalg.norm(b))
