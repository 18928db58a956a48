"""Reduced output (qsim_solve_selected): selected entries / populations of every
saved result must equal the same entries of the full result bit for bit -- the solve is the same kernel, only
the store differs -- for the resident, streamed and row-split kernels and all four entry points."""
import numpy as np
import pytest

from qsim_b200 import _abi, workloads as W
from qsim_b200._ffi import Problem, QsimError


def _cfg(name, n):
    cqs, params, ts, fn = W.config(name)
    idx = np.linspace(0, len(params) - 1, n).astype(int)
    return cqs, np.ascontiguousarray(params[idx]), fn


def test_selection_argument_errors():
    cqs, params, fn = _cfg("cfg2", 2)
    P = Problem(cqs)
    with pytest.raises(QsimError) as e:                     # no ctx on a CPU box / or bad selection on a GPU box
        P.solve(fn, np.array([0.0, 1.0]), params=params, select=[(0, 0), (99, 0)], abs2=True)
    assert e.value.code in (_abi.ERR_ARG, _abi.ERR_NO_DEVICE)


@pytest.mark.gpu
def test_selected_entries_match_full_result_unitary():
    cqs, params, fn = _cfg("cfg2", 37)
    ts = np.linspace(0.0, 6.0, 13)
    P = Problem(cqs)
    full, st = P.solve(fn, ts, params=params)
    sel = [(3, 3), (1, 3), (3, 1), (0, 0), (8, 8), (4, 2)]
    got, st2 = P.solve(fn, ts, params=params, select=sel)
    pops, st3 = P.solve(fn, ts, params=params, select=sel, abs2=True)
    want = np.stack([full[:, :, r, c] for r, c in sel], axis=-1)
    assert got.shape == (37, 13, 6) and pops.dtype == np.float64
    assert np.array_equal(got, want)
    assert np.allclose(pops, want.real ** 2 + want.imag ** 2, rtol=4e-16, atol=0)    # one FMA of difference
    assert np.array_equal(st["n_accept"], st2["n_accept"]) and np.array_equal(st["n_accept"], st3["n_accept"])
    with pytest.raises(QsimError):
        P.solve(fn, ts, params=params, select=[(0, 0), (0, 0)])          # duplicates
    with pytest.raises(QsimError):
        P.solve(fn, ts, params=params, select=[(9, 0)])                  # out of range


@pytest.mark.gpu
def test_selected_entries_states_and_lindblad():
    from qsim_b200 import me_propagator, me_state, unitary_state
    # unitary_state: components of psi
    cqs, params, _ = _cfg("cfg2", 5)
    ts = np.linspace(0.0, 3.0, 4)
    psi0 = np.zeros(9, dtype=complex)
    psi0[3] = 1
    full = unitary_state(cqs, ts, psi0, params=params)
    got = unitary_state(cqs, ts, psi0, params=params, select=[3, 1], abs2=True)
    assert np.allclose(got, np.abs(full[:, :, [3, 1]]) ** 2, rtol=4e-16, atol=0)
    # streamed kernel (d^2 = 81 superoperator) and me_state by density-matrix indices
    cqs, params, _ = _cfg("cfg4", 3)
    ts = np.array([0.0, 0.5, 1.0])
    full = me_propagator(cqs, ts, params=params)
    sel = [(0, 0), (40, 40), (10, 30), (80, 0)]
    got = me_propagator(cqs, ts, params=params, select=sel)
    assert np.array_equal(got, np.stack([full[:, :, r, c] for r, c in sel], axis=-1))
    rho0 = np.zeros((9, 9), dtype=complex)
    rho0[3, 3] = 1
    full = me_state(cqs, ts, rho0, params=params)
    got = me_state(cqs, ts, rho0, params=params, select=[(3, 3), (1, 1), (3, 1)])
    assert np.array_equal(got, np.stack([full[:, :, 3, 3], full[:, :, 1, 1], full[:, :, 3, 1]], axis=-1))


@pytest.mark.gpu
def test_selected_entries_rowsplit():
    from qsim_b200 import me_propagator
    cqs, params, _ = _cfg("cfg5", 2)
    ts = np.array([0.0, 0.05])
    full = me_propagator(cqs, ts, params=params)
    sel = [(0, 0), (364, 364), (728, 1), (100, 600)]
    got = me_propagator(cqs, ts, params=params, select=sel)
    assert np.array_equal(got, np.stack([full[:, :, r, c] for r, c in sel], axis=-1))


@pytest.mark.gpu
def test_multi_context_sweep_matches_single_context():
    """qsim_solve_multi: the sweep split over several contexts (distinct devices when the box has them, two
    contexts on device 0 otherwise) gives bitwise the results of one context, for ragged splits too."""
    import torch
    from qsim_b200._ffi import Context
    ndev = torch.cuda.device_count()
    ctxs = [Context(i % ndev) for i in range(max(2, min(ndev, 4)))]
    cqs, params, fn = _cfg("cfg2", 301)
    ts = np.linspace(0.0, 4.0, 5)
    P = Problem(cqs)
    one, st1 = P.solve(fn, ts, params=params, ctx=ctxs[0])
    many, stn = P.solve(fn, ts, params=params, ctx=ctxs)
    assert np.array_equal(one, many) and np.array_equal(st1, stn)
    # per-trajectory save grids and initial states travel with their block
    from qsim_b200 import unitary_state
    psi0 = np.zeros((301, 9), dtype=complex)
    psi0[np.arange(301), np.arange(301) % 9] = 1
    tsn = ts[None, :] + 0.01 * np.arange(301)[:, None]
    a, _ = P.solve("unitary_state", tsn, params=params, init=psi0, ctx=ctxs[0])
    b, _ = P.solve("unitary_state", tsn, params=params, init=psi0, ctx=ctxs)
    assert np.array_equal(a, b)
    with pytest.raises(QsimError):
        P.solve(fn, ts, params=params, ctx=[ctxs[0], ctxs[0]])
    for c in ctxs:
        c.close()
