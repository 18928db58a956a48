"""Component pruning of propagators (csrc/qsim_components.cpp): when the generator's graph splits into connected
components the library integrates only the component blocks; the entries that couple different components are
exactly zero in the reference as well (/root/reference/src/time_evolution.jl:31-37, 98-111: dense products that
only ever add 0 * x there) and are written as zeros.  Checked here: the plan, parity with the oracle (which
integrates the full dense state) at 1e-8 with identical step counts, exact zeros, agreement with the un-pruned
kernels of the same library, selected output, failed trajectories."""
import os

import numpy as np
import pytest
from scipy.sparse import csr_matrix
from scipy.sparse.csgraph import connected_components

from oracle import c_oracle as co
from qsim_b200 import (CompositeQSystem, Cos, HermitianSpec, LiteralHermitian, Param, add_hamiltonian, add_lindblad,
                       scaled_operator)
from qsim_b200 import workloads as W
from qsim_b200._ffi import Problem

WHICH = {"unitary_propagator": 0, "unitary_state": 1, "me_propagator": 2, "me_state": 3}


def _pattern_components(P, kind, row):
    pat = None
    for t in (0.13, 0.77, 3.3):
        A = P.eval(kind, t, row) != 0
        pat = A if pat is None else (pat | A)
    nc, lab = connected_components(csr_matrix(pat | pat.T), directed=False)
    return nc, lab


@pytest.mark.parametrize("name,kernel,ncomp,ratio", [("cfg2", "r1c2", 2, 0.5), ("cfg3", "r1c3", 2, 0.5),
                                                      ("cfg4", "r3c1", 2, 0.5), ("cfg5", "rowsplit", 13, 0.1452)])
def test_plan_of_benchmark_configs(name, kernel, ncomp, ratio):
    """The lab-frame transmon gates of benchmarks.jl conserve the parity of the total excitation number: two
    components, half of the propagator's entries, half of the generator multiply-adds; the rotating-frame exchange form
    with decay and dephasing (cfg5) conserves the excitation-number difference of |i><j|: 13 components, 14.5 %."""
    cqs, params, _, which = W.config(name)
    P = Problem(cqs)
    info = P.plan_info(which)
    assert info["components"] == ncomp
    nc, _ = _pattern_components(P, 0 if which.startswith("unitary") else 1, params[len(params) // 3])
    assert nc == ncomp
    assert kernel in P.kernel_name(which)
    vm = info["value_mode"]
    full = (4.0 if vm >= 1 else 8.0) * info["nnz"] * info["n"]
    assert abs(P.flops_per_rhs(which) / full - ratio) < 0.02
    # state evolution takes an arbitrary initial state: never pruned
    assert P.plan_info("unitary_state")["components"] == 0
    os.environ["QSIM_NO_PRUNE"] = "1"
    try:
        assert Problem(cqs).plan_info(which)["components"] == 0
    finally:
        del os.environ["QSIM_NO_PRUNE"]


def _block_system(sizes, rng, lindblad=False, permute=True, band=2):
    """Levels in blocks that nothing couples (after a random permutation of the levels): banded Hermitian blocks, a
    driven operator that is block-diagonal too, optionally collapse operators inside the blocks."""
    d = int(sum(sizes))
    h = np.diag(min(0.3, 6.0 / d) * np.arange(d) + 0.05 * rng.normal(size=d)).astype(complex)
    x = np.zeros((d, d))
    lo = 0
    for s in sizes:
        for k in range(1, band + 1):
            if s > k:
                off = 0.02 * (rng.normal(size=s - k) + 1j * rng.normal(size=s - k))
                idx = np.arange(lo, lo + s - k)
                h[idx, idx + k] += off
                h[idx + k, idx] += off.conj()
        if s > 1:
            idx = np.arange(lo, lo + s - 1)
            x[idx, idx + 1] = np.sqrt(np.arange(1, s))
        lo += s
    perm = rng.permutation(d) if permute else np.arange(d)
    pm = np.eye(d)[perm]
    h = pm @ h @ pm.T
    x = pm @ x @ pm.T
    q = LiteralHermitian("q", HermitianSpec(h))
    cqs = CompositeQSystem([q])
    add_hamiltonian(cqs, q)
    add_hamiltonian(cqs, scaled_operator(x + x.T, Cos(amp=Param(0), freq=0.31, phase=0.4)), q)
    if lindblad:
        add_lindblad(cqs, 0.05 * x, [q])
        add_lindblad(cqs, 0.03 * (pm @ np.diag(np.arange(d)).astype(complex) @ pm.T), [q])
    return cqs


def _check(cqs, fn, ts, params, n_comp_min=2):
    P = Problem(cqs)
    info = P.plan_info(fn)
    assert info["components"] >= n_comp_min, info
    shape = P.state_shape(WHICH[fn])
    buf = np.full((len(params), len(ts), int(np.prod(shape))), np.nan + 1j * np.nan)   # every entry must be written
    got, st = P.solve(fn, ts, params=params, out=buf)
    assert not np.any(np.isnan(got.real))
    want, ost = co.OracleProblem(cqs).solve(WHICH[fn], ts, params, mode=co.STRUCTURED, n_threads=8)
    assert np.all(st["retcode"] == 0)
    assert np.array_equal(st["n_accept"], ost["n_accept"]) and np.array_equal(st["n_reject"], ost["n_reject"])
    err = float(np.max(np.abs(got - want) / np.maximum(np.abs(want), 1e-3)))
    assert err <= 1e-8, err
    # entries between components: exact zeros in the oracle and here
    assert np.array_equal(got == 0, want == 0) or np.all(got[want == 0] == 0)
    assert np.count_nonzero(want == 0) > 0
    # the un-pruned kernels of the same library: same steps, same numbers up to the rounding of the error norms
    os.environ["QSIM_NO_PRUNE"] = "1"
    try:
        full, fst = Problem(cqs).solve(fn, ts, params=params)
    finally:
        del os.environ["QSIM_NO_PRUNE"]
    assert np.array_equal(st["n_accept"], fst["n_accept"]) and np.array_equal(st["n_reject"], fst["n_reject"])
    assert float(np.max(np.abs(got - full))) < 1e-10
    return P.kernel_name(fn)


@pytest.mark.gpu
@pytest.mark.parametrize("sizes", [[5, 4], [1, 1], [2, 1, 3], [7, 7, 2], [14, 13], [16, 16], [20, 9, 3], [12, 11, 10, 9],
                                   [30, 20, 10, 4], [32, 32, 32], [41, 40], [50, 30, 10, 5, 1], [64, 32], [90, 4, 2]])
def test_unitary_block_systems(sizes):
    rng = np.random.default_rng(5150 + 7 * len(sizes) + sum(sizes))
    cqs = _block_system(sizes, rng)
    params = np.array([[0.01], [0.03], [0.02]])
    d = sum(sizes)
    ts = np.linspace(0.0, 2.0 if d <= 40 else 1.0, 4)
    name = _check(cqs, "unitary_propagator", ts, params, n_comp_min=len(sizes))
    print(sizes, name)


@pytest.mark.gpu
@pytest.mark.parametrize("sizes", [[2, 1], [2, 2], [3, 2, 1], [5, 4], [6, 3]])
def test_lindblad_block_systems(sizes):
    """Blocks of levels -> the superoperator couples rho_ij only within (block of i, block of j) classes."""
    rng = np.random.default_rng(808 + sum(sizes) + len(sizes))
    cqs = _block_system(sizes, rng, lindblad=True)
    params = np.array([[0.01], [0.03]])
    d = sum(sizes)
    ts = np.linspace(0.0, 0.6 if d > 6 else 1.5, 3)
    name = _check(cqs, "me_propagator", ts, params)
    print(sizes, name)


@pytest.mark.gpu
def test_benchmark_gates_pruned_vs_full():
    """cfg2 (two transmons, 20 ns with dense-output points inside steps) and cfg4 (its Lindblad form) against the
    oracle and the un-pruned kernels."""
    cqs, _, _, _ = W.config("cfg2")
    params = W.amp_freq_grid(64, 64)[[0, 777, 4095]]
    assert "r1c2" in _check(cqs, "unitary_propagator", np.linspace(0.0, 20.0, 47), params)
    cqs4, p4, _, _ = W.config("cfg4")
    _check(cqs4, "me_propagator", np.array([0.0, 0.7, 1.5]), p4[[0, 8191]])


@pytest.mark.gpu
def test_selected_output_and_failed_trajectories():
    cqs, _, _, _ = W.config("cfg2")
    params = W.amp_freq_grid(64, 64)[[5, 4000]]
    ts = np.arange(0.0, 4.0)
    P = Problem(cqs)
    full, _ = P.solve("unitary_propagator", ts, params=params)
    rows = np.array([0, 1, 3, 4, 4, 8, 2], dtype=np.int32)
    cols = np.array([0, 1, 1, 4, 5, 8, 7], dtype=np.int32)
    pairs = np.stack([rows, cols], axis=1)
    sel, _ = P.solve("unitary_propagator", ts, params=params, select=pairs,
                     out=np.full((2, len(ts), len(rows)), np.nan + 0j))
    assert np.array_equal(sel, full[:, :, rows, cols])
    assert np.any(sel == 0) and np.any(sel != 0)
    p2, _ = P.solve("unitary_propagator", ts, params=params, select=pairs, abs2=True, out=np.full((2, len(ts), len(rows)), np.nan))
    assert np.array_equal(p2, np.abs(full[:, :, rows, cols]) ** 2) or np.allclose(p2, np.abs(full[:, :, rows, cols]) ** 2, rtol=1e-15)
    # a trajectory that runs out of steps: every entry of the unreached save points is NaN, structural zeros included
    from qsim_b200 import _abi
    o = _abi.default_opts()
    o.maxiters = 50
    bad, st = P.solve("unitary_propagator", ts, params=params, opts=o)
    assert np.all(st["retcode"] != 0)
    assert np.all(np.isnan(bad[:, -1].real))
    assert np.array_equal(bad[:, 0], np.broadcast_to(np.eye(9), (2, 9, 9)))


@pytest.mark.gpu
@pytest.mark.parametrize("sizes,band", [([70, 60], 2), ([100, 50, 30, 20], 2), ([120, 40, 5, 3, 1, 1], 2), ([20, 20], 19),
                                        ([30, 25, 18], 24)])
def test_row_split_block_systems(sizes, band):
    """State vectors longer than 96 entries, or rows longer than 16 entries: the row-split kernels integrate several
    units (component, column) per pass."""
    rng = np.random.default_rng(31337 + sum(sizes) + band)
    cqs = _block_system(sizes, rng, band=band)
    params = np.array([[0.02], [0.01]])
    ts = np.array([0.0, 0.4, 0.8])
    name = _check(cqs, "unitary_propagator", ts, params, n_comp_min=len(sizes))
    assert "rowsplit" in name
    print(sizes, name)


def _check_state(cqs, fn, ts, params, init, expect_pruned, capfd):
    P = Problem(cqs)
    n = int(np.prod(P.state_shape(WHICH[fn])))
    buf = np.full((len(params), len(ts), n), np.nan + 1j * np.nan)
    os.environ["QSIM_DEBUG"] = "1"
    try:
        got, st = P.solve(fn, ts, params=params, init=init, out=buf)
    finally:
        del os.environ["QSIM_DEBUG"]
    err_text = capfd.readouterr().err
    assert ("state form:" in err_text) == expect_pruned, err_text[-400:]
    assert not np.any(np.isnan(got.real))
    want, ost = co.OracleProblem(cqs).solve(WHICH[fn], ts, params, init=init, mode=co.STRUCTURED, n_threads=8)
    assert np.all(st["retcode"] == 0)
    assert np.array_equal(st["n_accept"], ost["n_accept"]) and np.array_equal(st["n_reject"], ost["n_reject"])
    err = float(np.max(np.abs(got - want) / np.maximum(np.abs(want), 1e-3)))
    assert err <= 1e-8, err
    assert np.all(got[want == 0] == 0)
    os.environ["QSIM_NO_PRUNE"] = "1"
    try:
        full, fst = Problem(cqs).solve(fn, ts, params=params, init=init)
    finally:
        del os.environ["QSIM_NO_PRUNE"]
    assert np.array_equal(st["n_accept"], fst["n_accept"]) and np.array_equal(st["n_reject"], fst["n_reject"])
    assert float(np.max(np.abs(got - full))) < 1e-10
    return got


@pytest.mark.gpu
def test_state_evolution_on_touched_components(capfd):
    """unitary_state / me_state: only the components the initial state touches are integrated (state forms); an initial
    state that touches every component runs on the full state."""
    # cfg2 system: |01> lies in the odd-parity block (4 of 9 levels); a superposition of |00> and |01> touches both
    cqs, _, _, _ = W.config("cfg2")
    params = W.amp_freq_grid(64, 64)[[0, 2017, 4095]]
    ts = np.linspace(0.0, 6.0, 7)
    psi = np.zeros(9, dtype=complex)
    psi[1] = 1.0
    got = _check_state(cqs, "unitary_state", ts, params, psi, True, capfd)
    assert np.all(got[:, :, [0, 2, 4, 6, 8]] == 0)
    both = np.zeros(9, dtype=complex)
    both[0] = both[1] = np.sqrt(0.5)
    _check_state(cqs, "unitary_state", ts, params, both, False, capfd)
    # per-trajectory initial states in different blocks: the union is integrated (here: everything)
    per = np.zeros((3, 9), dtype=complex)
    per[0, 1] = per[1, 0] = per[2, 3] = 1.0
    _check_state(cqs, "unitary_state", ts, params, per, False, capfd)
    per[1, 0], per[1, 5] = 0.0, 1.0
    _check_state(cqs, "unitary_state", ts, params, per, True, capfd)
    # cfg4 system (Lindblad, d^2 = 81): rho0 = |01><01| lies in the (odd, odd) + (even, even) class: 41 of 81 entries
    cqs4, p4, _, _ = W.config("cfg4")
    rho = np.zeros((9, 9), dtype=complex)
    rho[1, 1] = 1.0
    _check_state(cqs4, "me_state", np.array([0.0, 0.8, 1.6]), p4[[0, 8191]], rho, True, capfd)
    # cfg5 system (d^2 = 729, 13 components): rho0 = |010><010| touches the 141-row component only (row-split, one pass)
    cqs5, p5, _, _ = W.config("cfg5")
    rho5 = np.zeros((27, 27), dtype=complex)
    rho5[3, 3] = 1.0
    _check_state(cqs5, "me_state", np.array([0.0, 0.1, 0.2]), p5[[0, 65535]], rho5, True, capfd)


@pytest.mark.gpu
@pytest.mark.parametrize("sizes", [[5, 4], [20, 9, 3], [50, 30, 10, 5, 1], [100, 50, 30, 20], [120, 110]])
def test_state_block_systems(sizes, capfd):
    rng = np.random.default_rng(99 + sum(sizes))
    cqs = _block_system(sizes, rng)
    d = sum(sizes)
    params = np.array([[0.02], [0.01]])
    ts = np.array([0.0, 0.5, 1.0])
    # an initial state inside the component of level 0 (whatever the permutation made of it)
    P = Problem(cqs)
    _, lab = _pattern_components(P, 0, params[0])
    psi = np.where(lab == lab[0], rng.normal(size=d) + 1j * rng.normal(size=d), 0.0)
    psi /= np.linalg.norm(psi)
    _check_state(cqs, "unitary_state", ts, params, psi, True, capfd)


@pytest.mark.gpu
def test_abstol_zero_integrates_the_full_state():
    """abstol = 0 makes the reference's error norm 0/0 on exactly-zero entries; the library then integrates the full
    state (no pruning), so it behaves exactly like its un-pruned kernels."""
    from qsim_b200 import _abi
    cqs, _, _, _ = W.config("cfg2")
    params = W.amp_freq_grid(64, 64)[[7, 3000]]
    ts = np.arange(0.0, 3.0)
    o = _abi.default_opts()
    o.abstol = 0.0
    a, sa = Problem(cqs).solve("unitary_propagator", ts, params=params, opts=o)
    os.environ["QSIM_NO_PRUNE"] = "1"
    try:
        b, sb = Problem(cqs).solve("unitary_propagator", ts, params=params, opts=o)
    finally:
        del os.environ["QSIM_NO_PRUNE"]
    assert np.array_equal(sa, sb)
    assert np.array_equal(a, b, equal_nan=True)


@pytest.mark.parametrize("sizes,lindblad,band", [([5, 4], False, 2), ([1, 1], False, 2), ([7, 7, 2], False, 2), ([14, 13], False, 2),
                                                 ([12, 11, 10, 9], False, 2), ([32, 32, 32], False, 2), ([41, 40], False, 2),
                                                 ([50, 30, 10, 5, 1], False, 2), ([90, 4, 2], False, 2), ([3, 2, 1], True, 2),
                                                 ([6, 3], True, 2), ([70, 60], False, 2), ([120, 40, 5, 3, 1, 1], False, 2),
                                                 ([20, 20], False, 19), ([30, 25, 18], False, 24)])
def test_component_form_tables_cover_the_blocks(sizes, lindblad, band, capfd):
    """No GPU needed: the packing of (component, column) units into warps / row-split passes is checked by the builder's
    own self-check (QSIM_DEBUG): every entry inside a component block is held by exactly one (batch, local row, column
    slot), nothing outside is, for every kernel shape the chooser tries."""
    rng = np.random.default_rng(424242 + sum(sizes) + band)
    cqs = _block_system(sizes, rng, lindblad=lindblad, band=band)
    fn = "me_propagator" if lindblad else "unitary_propagator"
    os.environ["QSIM_DEBUG"] = "1"
    try:
        P = Problem(cqs)
        info = P.plan_info(fn)
        name = P.kernel_name(fn)
    finally:
        del os.environ["QSIM_DEBUG"]
    err = capfd.readouterr().err
    checks = [ln for ln in err.splitlines() if "self-check" in ln]
    assert checks, err[-500:]
    assert all(ln.rstrip().endswith(" 0 bad entries") for ln in checks), checks
    assert info["components"] >= len(sizes)
    n_state = info["n"]
    kind = 1 if lindblad else 0
    nc, lab = _pattern_components(P, kind, np.array([0.02]))
    assert nc == info["components"]
    # executed multiply-adds: the blocks' share of the generator's
    sizes_c = np.bincount(lab)
    assert P.flops_per_rhs(fn) < (4.0 if info["value_mode"] >= 1 else 8.0) * info["nnz"] * n_state
    assert ("rowsplit" in name) == (n_state > 96 or info["max_row"] > 16)
    assert sizes_c.max() <= n_state
