"""Tabulated envelopes (QSIM_ENV_TABLE / qsim_problem_add_table): the path for
drives that are opaque closures in the reference (add_hamiltonian!(cqs, t -> ..., q),
/root/reference/src/composite_systems.jl:40-50).  The closure is sampled on the host into a piecewise Chebyshev
series; the closure-driven numpy oracle (which calls the closure itself) pins the accuracy of the table, the
descriptor-driven C oracle and the CUDA library evaluate the same table."""
import math
from dataclasses import replace

import numpy as np
import pytest

from oracle import c_oracle as co
from oracle import np_oracle as npo
from qsim_b200 import (CompositeQSystem, Cos, DuffingSpec, DuffingTransmon, Param, PerturbativeTransmon, Table,
                       TransmonSpec, X, add_hamiltonian, add_lindblad, decay, dipole_drive, parametric_drive,
                       rwa_dipole, tabulated)
from qsim_b200._ffi import Problem
from qsim_b200.composite import add_parametric_hamiltonians, composite_hamiltonian
from qsim_b200.operators import ParametricLindblad
from qsim_b200.signals import CARRIER_COS


def _drag(t):
    """A DRAG-like quadrature: derivative of a gaussian, an envelope outside the closed descriptor family."""
    x = (np.asarray(t) - 12.0) / 4.0
    return -x * np.exp(-0.5 * x * x)


def _ramp(t):
    return 0.2 * np.tanh((np.asarray(t) - 5.0) / 1.5) + 0.05 * np.sin(0.9 * np.asarray(t)) ** 2


def test_table_matches_closure():
    for fn, t0, t1 in ((_drag, 0.0, 24.0), (_ramp, 0.0, 30.0), (lambda t: math.exp(-t / 7.0), 0.0, 20.0)):
        tab = Table.from_function(fn, t0, t1)
        ts = np.linspace(t0, t1, 4001)
        want = np.array([float(fn(float(t))) for t in ts])
        assert np.max(np.abs(tab.eval(ts) - want)) <= 5e-14 * np.max(np.abs(want))
        # constant continuation outside the tabulated range
        assert tab.eval(t0 - 3.0) == pytest.approx(want[0], rel=1e-13)
        assert tab.eval(t1 + 3.0) == pytest.approx(want[-1], rel=1e-13)
    with pytest.raises(ValueError):
        Table.from_function(lambda t: np.sign(np.asarray(t) - 0.7), 0.0, 2.0, max_seg=64)     # discontinuous


def _shaped_pulse_system():
    """Qubit driven by a shaped microwave pulse I(t) cos(wt) + Q(t) sin(wt) with tabulated I/Q envelopes, coupled
    to a flux-tunable transmon whose flux follows a tabulated ramp; amplitude is a sweep parameter."""
    q0 = DuffingTransmon("q0", 3, DuffingSpec(5.0, -0.2))
    q1 = PerturbativeTransmon("q1", 3, TransmonSpec(0.172, 12.71, 3.69))
    cqs = CompositeQSystem([q0, q1])
    add_hamiltonian(cqs, q0)
    add_hamiltonian(cqs, 0.006 * X([q0, q1]), [q0, q1])
    gauss = tabulated(lambda t: np.exp(-0.5 * ((np.asarray(t) - 12.0) / 4.0) ** 2), 0.0, 24.0, amp=Param(0))
    i_quad = replace(gauss, carrier=CARRIER_COS, freq=5.0)
    q_quad = replace(tabulated(_drag, 0.0, 24.0, amp=0.01), carrier=CARRIER_COS, freq=5.0, phase=-math.pi / 2)
    add_hamiltonian(cqs, dipole_drive(q0, i_quad), q0)
    add_hamiltonian(cqs, dipole_drive(q0, q_quad), q0)
    add_hamiltonian(cqs, parametric_drive(q1, tabulated(_ramp, 0.0, 30.0)), q1)
    return cqs


def test_lowered_tables_evaluate_like_the_closures():
    cqs = _shaped_pulse_system()
    row = np.array([0.03])
    P, O = Problem(cqs), co.OracleProblem(cqs)
    for t in (0.0, 3.7, 12.0, 23.99, 26.5):
        want = composite_hamiltonian(cqs)
        add_parametric_hamiltonians(want, cqs, t, row)          # calls the closures
        for got in (P.eval(0, t, row), O.eval(0, t, row)):
            assert np.max(np.abs(got - want)) <= 1e-12
    # the same table shared by two signals is registered once
    assert len(P.spec.tables) == 3


def test_closure_oracle_pins_the_table_oracle():
    cqs = _shaped_pulse_system()
    ts = np.linspace(0.0, 6.0, 7)
    row = np.array([0.03])
    want, wst = npo.unitary_propagator(cqs, ts, params=row)
    got, st = co.OracleProblem(cqs).solve(co.UNITARY_PROPAGATOR, ts, row[None, :], mode=0)
    assert st["retcode"][0] == 0
    assert np.max(np.abs(got[0] - np.asarray(want))) <= 1e-9


def test_tabulated_gaussian_equals_gaussian_descriptor():
    """test_time_evolution.jl:45 style gaussian pulse, once through ENV_GAUSSIAN and once as a tabulated closure."""
    def system(sig):
        q0 = DuffingTransmon("q0", 3, DuffingSpec(5.0, -0.2))
        cqs = CompositeQSystem([q0])
        add_hamiltonian(cqs, q0)
        add_hamiltonian(cqs, rwa_dipole(q0, sig), q0)
        return cqs
    a = system(Cos(amp=0.05, freq=0.0).gaussian(10.0, 3.0))
    b = system(tabulated(lambda t: np.exp(-0.5 * ((np.asarray(t) - 10.0) / 3.0) ** 2), 0.0, 20.0, amp=0.05))
    ts = np.linspace(0.0, 20.0, 11)
    ua, _ = co.OracleProblem(a).solve(co.UNITARY_PROPAGATOR, ts)
    ub, _ = co.OracleProblem(b).solve(co.UNITARY_PROPAGATOR, ts)
    assert np.max(np.abs(ua - ub)) <= 1e-10


@pytest.mark.gpu
def test_tabulated_drives_gpu_parity():
    cqs = _shaped_pulse_system()
    ts = np.linspace(0.0, 26.0, 27)
    params = np.linspace(0.01, 0.05, 9)[:, None]
    P = Problem(cqs)
    got, st = P.solve("unitary_propagator", ts, params=params)
    want, ost = co.OracleProblem(cqs).solve(co.UNITARY_PROPAGATOR, ts, params, mode=0, n_threads=4)
    assert np.all(st["retcode"] == 0)
    assert np.array_equal(st["n_accept"], ost["n_accept"]) and np.array_equal(st["n_reject"], ost["n_reject"])
    assert np.max(np.abs(got - want) / np.maximum(np.abs(want), 1e-3)) <= 1e-8
    # drive windows on (default) and off agree
    from qsim_b200 import _abi
    direct, std = P.solve("unitary_propagator", ts, params=params, opts=_abi.default_opts(flags=_abi.FLAG_DIRECT_DRIVES))
    assert np.array_equal(st["n_accept"], std["n_accept"])
    assert np.max(np.abs(got - direct)) <= 1e-10


@pytest.mark.gpu
def test_tabulated_lindblad_rate_gpu_parity():
    """Time-dependent collapse operator L(t) = s(t) a with a tabulated s(t) (composite_systems.jl:59-62)."""
    q0 = DuffingTransmon("q0", 3, DuffingSpec(5.0, -0.2))
    cqs = CompositeQSystem([q0])
    add_hamiltonian(cqs, q0)
    add_hamiltonian(cqs, dipole_drive(q0, Cos(amp=0.02, freq=5.0)), q0)
    rate = tabulated(lambda t: 0.1 * (1.0 + 0.5 * np.tanh((np.asarray(t) - 2.0) / 0.7)), 0.0, 5.0)
    add_lindblad(cqs, ParametricLindblad(decay(q0, 1.0), rate), [q0])
    ts = np.linspace(0.0, 5.0, 6)
    from qsim_b200 import me_propagator
    got = me_propagator(cqs, ts)
    want, _ = co.OracleProblem(cqs).solve(co.ME_PROPAGATOR, ts, mode=0)
    assert np.max(np.abs(got - want[0]) / np.maximum(np.abs(want[0]), 1e-3)) <= 1e-8
