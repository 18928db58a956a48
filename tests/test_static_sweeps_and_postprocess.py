"""Swept *static* parameters (couplings and qubit frequencies as sweep columns, the
static Hamiltonian differing per trajectory) and the post-processing helpers around the selected output."""
import numpy as np
import pytest

from oracle import c_oracle as co
from oracle import np_oracle as npo
from qsim_b200 import (CompositeQSystem, Const, DuffingSpec, DuffingTransmon, Param, PerturbativeTransmon, Sin,
                       TransmonSpec, X, add_hamiltonian, average_gate_fidelity, parametric_drive, populations,
                       scaled_operator, subspace_block, subspace_indices)
from qsim_b200._ffi import Problem


def _system():
    """cfg2's gate with the coupling g = Param(0) and q0's frequency = Param(1) swept; drive fixed."""
    q0 = DuffingTransmon("q0", 3, DuffingSpec(3.94015, -0.1807))
    q1 = PerturbativeTransmon("q1", 3, TransmonSpec(0.172, 16.4, 0.55))
    cqs = CompositeQSystem([q0, q1])
    add_hamiltonian(cqs, parametric_drive(q0, Const(Param(1))), q0)              # hamiltonian(q0, f): systems.jl:205
    add_hamiltonian(cqs, parametric_drive(q1, Sin(amp=0.323, freq=0.1221)), q1)
    add_hamiltonian(cqs, scaled_operator(X([q0, q1]), Param(0)), [q0, q1])
    return cqs


def test_static_sweep_oracles_agree():
    cqs = _system()
    params = np.array([[0.004, 3.90], [0.008, 3.98]])
    ts = np.linspace(0.0, 3.0, 4)
    got, st = co.OracleProblem(cqs).solve(co.UNITARY_PROPAGATOR, ts, params, mode=0)
    for i in range(2):
        want, _ = npo.unitary_propagator(cqs, ts, params=params[i])
        assert np.max(np.abs(got[i] - np.asarray(want))) <= 1e-10
    # the static slots are per-trajectory constants, not time-dependent ones: the kernel keeps value mode 2
    assert "_vm2_" in Problem(cqs).kernel_name("unitary_propagator")


def test_subspace_helpers():
    idx = subspace_indices([3, 3])
    assert list(idx) == [0, 1, 3, 4]
    sel, shape = subspace_block([3, 3])
    assert shape == (4, 4) and sel[1] == (0, 1) and sel[4] == (1, 0) and len(sel) == 16
    u = np.eye(4, dtype=complex)
    assert average_gate_fidelity(u, np.eye(4)) == pytest.approx(1.0)
    assert average_gate_fidelity(0.5 * u, np.eye(4)) == pytest.approx((1.0 + 4.0) / 20.0)
    assert np.allclose(populations(np.array([3 + 4j])), [25.0])


def test_superoperator_and_kraus_mix():
    from qsim_b200 import kraus_mix, superoperator
    rng = np.random.default_rng(20261019)
    us = np.linalg.qr(rng.normal(size=(5, 3, 3)) + 1j * rng.normal(size=(5, 3, 3)))[0]
    rho = rng.normal(size=(3, 3)) + 1j * rng.normal(size=(3, 3))
    s = superoperator(us[0])
    assert np.allclose((s @ rho.reshape(-1, order="F")).reshape(3, 3, order="F"), us[0] @ rho @ us[0].conj().T)
    w = np.array([0.1, 0.2, 0.3, 0.25, 0.15])
    mixed = (kraus_mix(us, w) @ rho.reshape(-1, order="F")).reshape(3, 3, order="F")
    assert np.allclose(mixed, sum(wi * u @ rho @ u.conj().T for wi, u in zip(w, us)))
    # trace preservation of the mixture: vec(I)' S = vec(I)'
    eye = np.eye(3).reshape(-1, order="F")
    assert np.allclose(eye @ kraus_mix(us, w), eye)


@pytest.mark.gpu
def test_static_sweep_gpu_parity_and_fidelity():
    cqs = _system()
    g, f = np.meshgrid(np.linspace(0.003, 0.009, 5), np.linspace(3.90, 3.98, 7), indexing="ij")
    params = np.ascontiguousarray(np.stack([g.ravel(), f.ravel()], axis=1))
    ts = np.linspace(0.0, 10.0, 6)
    P = Problem(cqs)
    got, st = P.solve("unitary_propagator", ts, params=params)
    want, ost = co.OracleProblem(cqs).solve(co.UNITARY_PROPAGATOR, ts, params, mode=0, n_threads=4)
    assert np.array_equal(st["n_accept"], ost["n_accept"])
    assert np.max(np.abs(got - want) / np.maximum(np.abs(want), 1e-3)) <= 1e-8
    # qubit-subspace block selected on the device, fidelity against the identity computed on 16 numbers
    sel, shape = subspace_block(cqs.dims)
    sub, _ = P.solve("unitary_propagator", ts, params=params, select=sel)
    sub = sub.reshape(sub.shape[:-1] + shape)
    idx = subspace_indices(cqs.dims)
    assert np.array_equal(sub, got[:, :, idx][:, :, :, idx])
    fid = average_gate_fidelity(sub, np.eye(4))
    assert fid.shape == (35, 6) and np.allclose(fid[:, 0], 1.0) and np.all(fid <= 1.0 + 1e-12)
