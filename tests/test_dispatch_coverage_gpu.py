"""Every kernel family the dispatcher can choose, on systems of many sizes: random dense-ish Hermitian systems
(LiteralHermitian, like /root/reference/test/test_time_evolution.jl:5-23) with a driven transmon-like subsystem,
all four entry points, checked against the oracle (1e-8, identical step counts).  The sizes walk through
lane-groups (n <= 16), multi-warp resident teams (17..32), streamed teams (33..96) and row-split teams (> 96)."""
import numpy as np
import pytest

from oracle import c_oracle as co
from qsim_b200 import (CompositeQSystem, Cos, HermitianSpec, LiteralHermitian, Param, add_hamiltonian, add_lindblad,
                       scaled_operator)
from qsim_b200._ffi import Problem

WHICH = {"unitary_propagator": 0, "unitary_state": 1, "me_propagator": 2, "me_state": 3}


def _random_system(d, rng, band=2, lindblad=False):
    """A d-level system: banded Hermitian H0 (frequencies ~0.3 GHz apart so a few ns resolve many steps), a
    driven coupling operator with a swept amplitude, optionally two collapse operators."""
    h = np.diag(min(0.3, 6.0 / d) * np.arange(d) + 0.05 * rng.normal(size=d)).astype(complex)
    for k in range(1, band + 1):
        off = 0.02 * (rng.normal(size=d - k) + 1j * rng.normal(size=d - k))
        h += np.diag(off, k) + np.diag(off.conj(), -k)
    q = LiteralHermitian("q", HermitianSpec(h))
    cqs = CompositeQSystem([q])
    add_hamiltonian(cqs, q)
    x = np.diag(np.sqrt(np.arange(1, d)), 1)
    add_hamiltonian(cqs, scaled_operator(x + x.T, Cos(amp=Param(0), freq=0.31, phase=0.4)), q)
    if lindblad:
        add_lindblad(cqs, 0.05 * x, [q])
        add_lindblad(cqs, 0.03 * np.diag(np.arange(d)).astype(complex), [q])
    return cqs


def _check(cqs, fn, ts, params, init=None):
    got, st = Problem(cqs).solve(fn, ts, params=params, init=init)
    want, ost = co.OracleProblem(cqs).solve(WHICH[fn], ts, params, init=init, mode=co.STRUCTURED, n_threads=8)
    assert np.all(st["retcode"] == 0)
    assert np.array_equal(st["n_accept"], ost["n_accept"]) and np.array_equal(st["n_reject"], ost["n_reject"]), fn
    err = float(np.max(np.abs(got - want) / np.maximum(np.abs(want), 1e-3)))
    assert err <= 1e-8, (fn, err)
    return Problem(cqs).kernel_name(fn)


@pytest.mark.gpu
@pytest.mark.parametrize("d", [2, 3, 4, 5, 7, 8, 10, 11, 13, 16, 17, 20, 24, 25, 29, 32, 33, 40, 64, 65, 80, 96])
def test_unitary_sizes(d):
    rng = np.random.default_rng(20261019 + d)
    cqs = _random_system(d, rng)
    params = np.array([[0.01], [0.03], [0.02]])
    ts = np.linspace(0.0, 2.0, 4)
    name = _check(cqs, "unitary_propagator", ts, params)
    psi0 = rng.normal(size=d) + 1j * rng.normal(size=d)
    psi0 /= np.linalg.norm(psi0)
    name_s = _check(cqs, "unitary_state", ts, params, init=psi0)
    assert name.startswith("tsit5_") and name_s.startswith("tsit5_")


@pytest.mark.gpu
@pytest.mark.parametrize("d", [97, 130, 200])
def test_unitary_row_split_sizes(d):
    """Propagators and states longer than 96 entries: rows spread over 2-3 warps, columns streamed."""
    rng = np.random.default_rng(4242 + d)
    cqs = _random_system(d, rng)
    params = np.array([[0.02], [0.01]])
    ts = np.array([0.0, 0.4, 0.8])
    assert "rowsplit" in _check(cqs, "unitary_propagator", ts, params)
    psi0 = rng.normal(size=d) + 1j * rng.normal(size=d)
    psi0 /= np.linalg.norm(psi0)
    assert "rowsplit" in _check(cqs, "unitary_state", ts, params, init=psi0)


@pytest.mark.gpu
@pytest.mark.parametrize("d", [2, 3, 4, 5, 6, 8, 9, 10, 12, 16])
def test_lindblad_sizes(d):
    rng = np.random.default_rng(777 + d)
    cqs = _random_system(d, rng, lindblad=True)
    params = np.array([[0.01], [0.03]])
    ts = np.linspace(0.0, 0.3 if d > 12 else (0.6 if d > 8 else 1.5), 3)
    _check(cqs, "me_propagator", ts, params)
    rho0 = rng.normal(size=(d, d)) + 1j * rng.normal(size=(d, d))
    rho0 = rho0 @ rho0.conj().T
    rho0 /= np.trace(rho0)
    _check(cqs, "me_state", ts, params, init=rho0)


@pytest.mark.gpu
@pytest.mark.parametrize("d", [12, 24, 27, 40, 48])
def test_dense_hamiltonians(d):
    """Fully dense Hermitian H (d entries per row): up to 16 per row run in the register-resident kernels, longer
    rows through the row-split kernels whatever the dimension."""
    rng = np.random.default_rng(99 + d)
    cqs = _random_system(d, rng, band=d - 1)
    params = np.array([[0.02], [0.03]])
    ts = np.array([0.0, 0.5, 1.0])
    name = _check(cqs, "unitary_propagator", ts, params)
    assert ("rowsplit" in name) == (d > 16)
    psi0 = np.zeros(d, dtype=complex)
    psi0[1] = 1
    _check(cqs, "unitary_state", ts, params, init=psi0)


def test_too_many_distinct_entries_is_reported():
    """A dense 70-level Hamiltonian has 4900 distinct entries: more value slots than the kernels' 16-bit packed
    offsets address.  The library says so (QSIM_ERR_UNSUPPORTED) instead of computing something else."""
    from qsim_b200 import _abi
    from qsim_b200._ffi import QsimError
    cqs = _random_system(70, np.random.default_rng(1), band=69)
    with pytest.raises(QsimError) as e:
        Problem(cqs).kernel_name("unitary_propagator")
    assert e.value.code == _abi.ERR_UNSUPPORTED


@pytest.mark.gpu
@pytest.mark.parametrize("d", [4, 6, 9])
def test_lindblad_dense_hamiltonians(d):
    """Dense H under the Lindblad generator: 2d - 1 commutator entries per row plus the dissipator -- long rows at
    d^2 = 36 and 81 go through the row-split kernels with a one-warp team."""
    rng = np.random.default_rng(5 + d)
    cqs = _random_system(d, rng, band=d - 1, lindblad=True)
    params = np.array([[0.02]])
    ts = np.array([0.0, 0.4, 0.8])
    _check(cqs, "me_propagator", ts, params)
    rho0 = np.zeros((d, d), dtype=complex)
    rho0[1, 1] = 1
    _check(cqs, "me_state", ts, params, init=rho0)
