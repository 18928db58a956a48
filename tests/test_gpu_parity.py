"""GPU parity: the CUDA library (through the C ABI) vs the CPU oracle on identical inputs.

Tolerance: elementwise |a-b| / max(|b|, 1e-3) <= 1e-8 (BASELINE.json north_star: "agree elementwise
within 1e-8 relative at matched solver tolerances"; denominator floor per SURVEY.md section 8d), with
identical accepted/rejected step counts -- i.e. the same step sequence as the oracle.
"""
import math

import numpy as np
import pytest

from oracle import c_oracle as co
from qsim_b200 import (CompositeQSystem, Cos, DuffingSpec, DuffingTransmon, HermitianSpec, LiteralHermitian, Param,
                       PerturbativeTransmon, Problem, RotatingFrameSystem, Sin, TransmonSpec, X_Y, _abi,
                       add_hamiltonian, add_lindblad, decay, dephasing, dipole_drive, me_propagator, me_state,
                       parametric_drive, rwa_dipole, unitary_propagator, unitary_state, xy_exchange_drive)
from qsim_b200 import workloads as W

pytestmark = pytest.mark.gpu

RTOL = 1e-8
FLOOR = 1e-3


def max_rel_err(a, b):
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), FLOOR)))


def check_parity(cqs, which, ts, params=None, init=None, opts=None, steps_exact=True):
    P, O = Problem(cqs), co.OracleProblem(cqs)
    got, gs = P.solve(which, ts, params=params, init=init, opts=opts)
    want, ws = O.solve(_ffi_which(which), ts, params=params, init=init, opts=opts, n_threads=8)
    assert got.shape == want.shape
    assert np.all(gs["retcode"] == 0) and np.all(ws["retcode"] == 0)
    err = max_rel_err(got, want)
    assert err <= RTOL, f"max rel err {err:.3e}"
    if steps_exact:
        assert np.array_equal(gs["n_accept"], ws["n_accept"])
        assert np.array_equal(gs["n_reject"], ws["n_reject"])
        assert np.array_equal(gs["n_rhs"], ws["n_rhs"])
    assert np.allclose(gs["dt_init"], ws["dt_init"], rtol=1e-12)
    assert np.array_equal(gs["t_final"], ws["t_final"])
    return got, gs, err


def _ffi_which(name):
    return {"unitary_propagator": 0, "unitary_state": 1, "me_propagator": 2, "me_state": 3}[name]


# ---------------------------------------------------------------- unitary
def test_cfg1_benchmark_system_short():
    """benchmarks.jl:125-154 (n=2), first 12 ns, 13 saves."""
    cqs, params, _, _ = W.config("cfg1")
    got, st, err = check_parity(cqs, "unitary_propagator", np.arange(0.0, 13.0), params)
    u = got[0, -1]
    assert np.abs(u.conj().T @ u - np.eye(9)).max() < 1e-3   # reference tolerances' own unitarity error


def test_cfg2_sweep_corners_and_interior():
    cqs, _, _, _ = W.config("cfg2")
    params = W.amp_freq_grid(64, 64)[[0, 63, 2017, 4032, 4095]]
    check_parity(cqs, "unitary_propagator", np.arange(0.0, 9.0), params)


def test_cfg3_three_transmons():
    """d = 27 (BASELINE config 3), 20 ns = ~8.3k steps per trajectory, 21 saves: corners and interior of the grid."""
    cqs, _, _, _ = W.config("cfg3")
    params = W.amp_freq_grid(128, 128)[[0, 5000, 8191, 16383]]
    got, st, err = check_parity(cqs, "unitary_propagator", np.arange(0.0, 21.0), params)
    assert st["n_accept"].min() > 8000
    print(f"cfg3 20 ns: max rel err {err:.2e}, steps {st['n_accept'].min()}..{st['n_accept'].max()}")


def test_literal_hermitian_dense_and_states():
    """test_time_evolution.jl:5-23: dense complex 6x6 H, 100 save points."""
    from scipy.linalg import block_diag, expm
    m1 = np.array([[0, 1], [1, 0]], dtype=complex)
    m2 = np.array([[0, -1j, 0], [1j, .1, 2 - 1j], [0, 2 + 1j, 0]])
    mat = block_diag(m1, m2, np.array([[3.0]]))
    lh = LiteralHermitian("label", HermitianSpec(mat))
    cqs = CompositeQSystem([lh])
    add_hamiltonian(cqs, lh)
    ts = np.linspace(0, 10, 100)
    got, _, _ = check_parity(cqs, "unitary_propagator", ts)
    for u, t in zip(got[0], ts):
        assert np.allclose(u, expm(-2j * np.pi * mat * t), atol=1e-4, rtol=0)
    us = unitary_propagator(cqs, ts)
    assert us.shape == (100, 6, 6) and np.array_equal(us, got[0])
    psi = np.arange(1, 7, dtype=complex) / np.linalg.norm(np.arange(1, 7))
    check_parity(cqs, "unitary_state", ts, init=psi)


def _iswap_system():
    q0 = DuffingTransmon("q0", 3, DuffingSpec(0.0, -0.2))
    q1 = DuffingTransmon("q1", 3, DuffingSpec(4.0, -0.2))
    cqs = CompositeQSystem([q0, q1])
    add_hamiltonian(cqs, q1)
    add_hamiltonian(cqs, parametric_drive(q0, Cos(amp=0.3, freq=0.3, offset=4.3)), q0)
    add_hamiltonian(cqs, .5 * 0.01 * X_Y([q0, q1]), [q0, q1])
    return cqs


def test_saveat_independence_bitwise():
    """test_time_evolution.jl:75-78: unitary_propagator(cqs, t) == us_full[end], bitwise."""
    cqs = _iswap_system()
    t_final = 20.0
    us_full = unitary_propagator(cqs, np.linspace(0, t_final, 3))
    us_dense = unitary_propagator(cqs, np.linspace(0, t_final, 57))
    u_end = unitary_propagator(cqs, t_final)
    assert np.array_equal(u_end, us_full[-1]) and np.array_equal(u_end, us_dense[-1])
    psi0 = np.zeros(9, dtype=complex)
    psi0[3] = 1
    psis = unitary_state(cqs, np.linspace(0, t_final, 3), psi0)
    assert np.array_equal(unitary_state(cqs, t_final, psi0), psis[-1])
    pop_state = np.abs(psis[:, 3]) ** 2
    pop_prop = np.array([abs(np.vdot(psi0, u @ psi0)) ** 2 for u in us_full])
    assert np.allclose(pop_state, pop_prop, rtol=1e-4)


def test_drive_family_kitchen_sink():
    """Every descriptor kind at once: flux drive in a rotating frame, erf-squared blended frequency drive,
    gaussian lab-frame dipole with frame rotation, RWA dipole, swept exchange coupling."""
    q0 = DuffingTransmon("q0", 3, DuffingSpec(3.94015, -0.1807))
    q1 = RotatingFrameSystem.wrap(PerturbativeTransmon("q1", 3, TransmonSpec(0.172, 12.71, 3.69)), 3.9)
    q2 = DuffingTransmon("q2", 2, DuffingSpec(4.2, -0.22))
    cqs = CompositeQSystem([q0, q1, q2])
    add_hamiltonian(cqs, q0)
    add_hamiltonian(cqs, parametric_drive(q1, Sin(amp=Param(0), freq=0.12, offset=0.05)), q1)
    add_hamiltonian(cqs, parametric_drive(q0, Cos(amp=0.3, freq=0.3, offset=4.3).erf_squared(2.0, 9.0, 0.7, rest=4.6)), q0)
    add_hamiltonian(cqs, dipole_drive(q2, Cos(amp=0.02, freq=4.2).gaussian(5.0, 2.0), 4.0), q2)
    add_hamiltonian(cqs, rwa_dipole(q2, Cos(amp=0.01, freq=0.01)), q2)
    add_hamiltonian(cqs, xy_exchange_drive([q2, q0], Param(1), [0.2, 0.0]), [q2, q0])
    params = np.array([[0.31, 0.004], [0.1, 0.002]])
    check_parity(cqs, "unitary_propagator", np.linspace(0, 6, 7), params)
    psi = np.zeros(18, dtype=complex)
    psi[7] = 1
    check_parity(cqs, "unitary_state", np.linspace(0, 6, 7), params, init=psi)


def test_per_trajectory_save_grids():
    cqs, _, _, _ = W.config("cfg2")
    params = W.amp_freq_grid(2, 2)
    ts = np.stack([np.linspace(0, 2.0 + i, 5) for i in range(4)])
    got, st, _ = check_parity(cqs, "unitary_propagator", ts, params)
    assert np.array_equal(st["t_final"], ts[:, -1])


def test_edge_cases():
    cqs, params, _, _ = W.config("cfg1")
    # single save time / zero-length span: identity (the reference returns u0)
    u = unitary_propagator(cqs, np.array([0.0]), params=params)
    assert u.shape == (1, 1, 9, 9) and np.array_equal(u[0, 0], np.eye(9))
    u = unitary_propagator(cqs, np.array([1.0, 1.0]), params=params)
    assert np.array_equal(u[0, 0], np.eye(9)) and np.array_equal(u[0, 1], np.eye(9))
    # start time is ts[0], not 0
    a = unitary_propagator(cqs, np.array([3.0, 4.0]), params=params)
    b = unitary_propagator(cqs, np.array([0.0, 1.0]), params=params)
    assert not np.allclose(a[0, -1], b[0, -1])
    with pytest.raises(AssertionError):
        unitary_propagator(cqs, np.array([0.0, 2.0, 1.0]), params=params)
    # maxiters surfaces as a retcode, unreached saves are NaN
    P = Problem(cqs)
    out, st = P.solve("unitary_propagator", np.array([0.0, 5.0]), params=params, opts=_abi.default_opts(maxiters=10))
    assert st["retcode"][0] == _abi.RET_MAXITERS and np.isnan(out[0, -1]).all() and st["n_accept"][0] <= 10
    # ragged sweep sizes around the CTA/team granularity
    for n in (1, 3, 5, 149):
        pr = np.tile(params, (n, 1))
        pr[:, 0] = np.linspace(0.2, 0.45, n)
        got, st = P.solve("unitary_propagator", np.array([0.0, 0.5]), params=pr)
        want, _ = co.OracleProblem(cqs).solve(0, np.array([0.0, 0.5]), params=pr, n_threads=8)
        assert max_rel_err(got, want) <= RTOL


# ---------------------------------------------------------------- Lindblad
def _driven_decay_system():
    q0 = DuffingTransmon("q0", 3, DuffingSpec(5.0, -0.2))
    cqs = CompositeQSystem([q0])
    add_hamiltonian(cqs, q0)
    add_hamiltonian(cqs, dipole_drive(q0, Cos(amp=0.02, freq=5.0)), q0)
    add_lindblad(cqs, decay(q0, 1 / (2 * np.pi * 50.)), [q0])
    return cqs


def test_me_propagator_and_state_single_transmon():
    """test_time_evolution.jl:121-153."""
    cqs = _driven_decay_system()
    ts = np.linspace(0, 20, 21)
    rho0 = np.zeros((3, 3), dtype=complex)
    rho0[0, 0] = 1
    us, _, _ = check_parity(cqs, "me_propagator", ts)
    rhos, _, _ = check_parity(cqs, "me_state", ts, init=rho0)
    for u, rho in zip(us[0], rhos[0]):
        rp = (u @ rho0.ravel(order="F")).reshape(3, 3, order="F")
        assert np.linalg.norm(rp - rho) <= 1e-4 * max(np.linalg.norm(rp), np.linalg.norm(rho))
    assert np.array_equal(me_propagator(cqs, 20.0), us[0, -1])
    assert np.array_equal(me_state(cqs, 20.0, rho0), rhos[0, -1])


def test_decay_and_dephasing_known_answer():
    """test_time_evolution.jl:155-182."""
    q0 = DuffingTransmon("q0", 3, DuffingSpec(5.0, -0.2))
    cqs = CompositeQSystem([q0])
    add_hamiltonian(cqs, q0)
    T1, Tphi = 50.0, 200.0
    T2 = 1 / (.5 / T1 + 1 / Tphi)
    add_lindblad(cqs, decay(q0, 1 / (2 * np.pi * T1)), [q0])
    add_lindblad(cqs, dephasing(q0, 1 / (2 * np.pi * Tphi)), [q0])
    ts = np.linspace(0, 100, 101)
    psi0 = np.array([1, 1, 0], dtype=complex) / math.sqrt(2)
    rhos = me_state(cqs, ts, np.outer(psi0, psi0.conj()))
    assert np.allclose(rhos[:, 1, 1], np.exp(-ts / T1) / 2, rtol=1e-10, atol=0)
    assert np.allclose(np.abs(rhos[:, 0, 1]), np.exp(-ts / T2) / 2, rtol=1e-2, atol=0)
    check_parity(cqs, "me_state", ts, init=np.outer(psi0, psi0.conj()))


def test_cfg4_lindblad_two_transmons():
    """d^2 = 81 superoperator propagator (BASELINE config 4), swept T1/Tphi, 6 ns = ~1.4k steps per trajectory."""
    cqs, params, _, _ = W.config("cfg4")
    pr = params[[0, 4097, 8191]]
    ts = np.array([0.0, 1.5, 6.0])
    got, st, err = check_parity(cqs, "me_propagator", ts, pr)
    assert st["n_accept"].min() > 1300
    print(f"cfg4 6 ns: max rel err {err:.2e}, steps {st['n_accept'].min()}..{st['n_accept'].max()}")
    vi = np.eye(9).ravel(order="F")
    for s in got[:, -1]:
        assert np.abs(vi @ s - vi).max() < 1e-5          # trace preservation
    rho0 = np.zeros((9, 9), dtype=complex)
    rho0[3, 3] = 1
    rhos, _, _ = check_parity(cqs, "me_state", ts, pr, init=rho0)
    for s, rho in zip(got[:, -1], rhos[:, -1]):
        rp = (s @ rho0.ravel(order="F")).reshape(9, 9, order="F")
        assert np.linalg.norm(rp - rho) <= 1e-4


# ---------------------------------------------------------------- row-split kernels (state longer than 96)
def _d12_lindblad_system():
    qa = DuffingTransmon("qa", 3, DuffingSpec(4.0, -0.2))
    qb = DuffingTransmon("qb", 4, DuffingSpec(4.3, -0.25))
    cqs = CompositeQSystem([qa, qb])
    add_hamiltonian(cqs, qa)
    add_hamiltonian(cqs, qb)
    from qsim_b200 import X
    add_hamiltonian(cqs, 0.01 * X([qa, qb]), [qa, qb])
    add_hamiltonian(cqs, dipole_drive(qb, Cos(amp=Param(0), freq=4.3)), qb)
    add_lindblad(cqs, decay(qa, 2e-3), [qa])
    add_lindblad(cqs, decay(qb, 1e-3), [qb])
    add_lindblad(cqs, dephasing(qb, 3e-3), [qb])
    return cqs


def test_rowsplit_d12_lindblad_propagator_and_state():
    """d = 12: the superoperator state has 144 rows -> row-split team of 2 warps, streamed columns."""
    cqs = _d12_lindblad_system()
    assert "rowsplit" in Problem(cqs).kernel_name("me_propagator")
    params = np.array([[0.02], [0.05]])
    ts = np.array([0.0, 0.4, 1.0])
    got, st, err = check_parity(cqs, "me_propagator", ts, params)
    vi = np.eye(12).ravel(order="F")
    for s in got[:, -1]:
        assert np.abs(vi @ s - vi).max() < 1e-5
    rho0 = np.zeros((12, 12), dtype=complex)
    rho0[5, 5] = 1
    check_parity(cqs, "me_state", ts, params, init=rho0)
    # unitary evolution of the same system stays on the register-resident kernels
    check_parity(cqs, "unitary_propagator", ts, params)


def test_cfg5_three_transmon_lindblad():
    """d^2 = 729 (BASELINE config 5).  me_propagator against the oracle's structured mode (the commutator /
    sandwich form of time_evolution.jl:98-111 applied column by column) elementwise to 1e-8 with identical step
    counts, on two sweep points over a span of seven accepted steps; me_state against the oracle; and the
    properties the domain offers (trace preservation, agreement with me_state on a basis state)."""
    cqs, params, _, _ = W.config("cfg5")
    pr = params[[0, 40000]]
    ts = np.array([0.0, 0.25, 0.5])
    rho0 = np.zeros((27, 27), dtype=complex)
    rho0[12, 12] = 1          # |1,1,0>
    rhos, _, _ = check_parity(cqs, "me_state", ts, pr, init=rho0)
    P = Problem(cqs)
    assert "rowsplit" in P.kernel_name("me_propagator")
    sup, st, err = check_parity(cqs, "me_propagator", ts, pr)
    assert st["n_accept"].min() >= 3
    print(f"cfg5 me_propagator vs structured oracle: max rel err {err:.2e}, steps {st['n_accept'].tolist()}")
    vi = np.eye(27).ravel(order="F")
    for s, rho in zip(sup[:, -1], rhos[:, -1]):
        assert np.abs(vi @ s - vi).max() < 1e-5
        rp = (s @ rho0.ravel(order="F")).reshape(27, 27, order="F")
        assert np.abs(rp - rho).max() < 1e-5      # different step sequences: solver tolerance, not 1e-8


# ---------------------------------------------------------------- BASELINE.json sizes
def test_cfg2_full_length_subsample():
    """cfg2 at full length (ts = 0:200, ~65.6k steps per trajectory) on 16 points spread over the 64x64 grid:
    elementwise 1e-8 against the oracle at every one of the 201 save points, identical step counts."""
    cqs, params, ts, _ = W.config("cfg2")
    pr = params[np.linspace(0, len(params) - 1, 16).astype(int)]
    got, st, err = check_parity(cqs, "unitary_propagator", ts, pr)
    assert st["n_accept"].min() > 60000
    print(f"cfg2 full length: max rel err {err:.2e}, steps {st['n_accept'].min()}..{st['n_accept'].max()}")


def test_cfg2_full_sweep_properties():
    """The whole 4096-point sweep at full length: size-independent properties (no oracle at this size):
    every trajectory succeeds, start save is the identity, the propagators are unitary to the reference
    tolerances' own error, and a sharded run (2 x 2048, block-cyclic) is bitwise identical."""
    from qsim_b200 import shard_params
    cqs, params, ts, _ = W.config("cfg2")
    ts2 = np.array([0.0, 100.0, 200.0])
    P = Problem(cqs)
    full, st = P.solve("unitary_propagator", ts2, params=params)
    assert np.all(st["retcode"] == 0) and np.all(st["t_final"] == 200.0)
    assert np.array_equal(full[:, 0], np.broadcast_to(np.eye(9), (len(params), 9, 9)))
    u = full[:, -1]
    unit = np.abs(np.einsum("tji,tjk->tik", u.conj(), u) - np.eye(9)).max()
    assert unit < 5e-3
    for rank in range(2):
        loc, idx = shard_params(params, 2, rank)
        part, _ = P.solve("unitary_propagator", ts2, params=loc)
        assert np.array_equal(part, full[idx])


def test_drive_windows_vs_direct_evaluation():
    """The windowed Chebyshev drive interpolants (default) and direct evaluation of the drive descriptors at
    every stage time (QSIM_FLAG_DIRECT_DRIVES) must both meet the parity bar, take the same steps and agree
    with each other far below it -- including a 5 GHz lab-frame dipole, where windows span under one period."""
    direct = _abi.default_opts(flags=_abi.FLAG_DIRECT_DRIVES)
    cqs, params, _, _ = W.config("cfg2")
    pr = params[[7, 2222, 4000]]
    ts = np.arange(0.0, 31.0)
    a, sa, _ = check_parity(cqs, "unitary_propagator", ts, pr)
    b, sb, _ = check_parity(cqs, "unitary_propagator", ts, pr, opts=direct)
    assert np.array_equal(sa["n_accept"], sb["n_accept"]) and np.array_equal(sa["n_reject"], sb["n_reject"])
    assert max_rel_err(a, b) < 1e-10
    q0 = DuffingTransmon("q0", 3, DuffingSpec(5.0, -0.2))
    rabi = CompositeQSystem([q0])
    add_hamiltonian(rabi, q0)
    add_hamiltonian(rabi, dipole_drive(q0, Cos(amp=0.02, freq=5.0).gaussian(10.0, 4.0), 4.9), q0)
    ts = np.linspace(0.0, 20.0, 21)
    a, _, _ = check_parity(rabi, "unitary_propagator", ts)
    b, _, _ = check_parity(rabi, "unitary_propagator", ts, opts=direct)
    assert max_rel_err(a, b) < 1e-10
    cqs4, p4, _, _ = W.config("cfg4")
    a, _, _ = check_parity(cqs4, "me_propagator", np.array([0.0, 2.0]), p4[[5, 8000]])
    b, _, _ = check_parity(cqs4, "me_propagator", np.array([0.0, 2.0]), p4[[5, 8000]], opts=direct)
    assert max_rel_err(a, b) < 1e-10


@pytest.mark.gpu
def test_repeated_runs_are_bitwise_identical():
    """Reductions are fixed-order and inter-warp exchanges are barrier-ordered, so every kernel mode must
    reproduce itself bit for bit (a data race would show up as run-to-run differences): resident one-warp,
    resident multi-warp team, streamed, row-split."""
    from qsim_b200 import workloads as W
    from qsim_b200._ffi import Problem
    for name, n, tf in (("cfg2", 300, 3.0), ("cfg3", 150, 2.0), ("cfg4", 150, 0.5), ("cfg5", 4, 0.05)):
        cqs, params, _, fn = W.config(name)
        idx = np.linspace(0, len(params) - 1, n).astype(int)
        pr = np.ascontiguousarray(params[idx])
        ts = np.linspace(0.0, tf, 4)
        P = Problem(cqs)
        first, st0 = P.solve(fn, ts, params=pr)
        assert np.all(st0["retcode"] == 0)
        for _ in range(3):
            again, st = P.solve(fn, ts, params=pr)
            assert np.array_equal(first, again), name
            assert np.array_equal(st0, st), name


@pytest.mark.gpu
@pytest.mark.parametrize("env,tag", [("QSIM_NO_LEAN", "tsit5_r2c1_nnz5_t128_vm2"), ("QSIM_NO_PRUNE", "tsit5_lean_r1c3_nnz5_t128_vm2"),
                                     ("QSIM_MAX_MINB", "tsit5_lean_r1c2_nnz5_t128b3_vm2"), ("QSIM_TIERED", "lean_tiered_r3c1"),
                                     ("QSIM_SPLIT", "lean_split_r1c3_nnz3"), ("QSIM_ADDSPLIT", "lean_addsplit_r1c3_nnz3")])
def test_cfg2_kernel_variants(monkeypatch, env, tag):
    """The default cfg2 kernel is the lean row-per-lane one on the component form (two blocks of 5 x 5 and 4 x 4 entries,
    two columns per lane, 128 registers, 16 warps per SM).  The 255-register kernels (QSIM_NO_LEAN=1: two rows of one column per lane
    on the component form), its 168-register
    form (QSIM_MAX_MINB=3), the lean kernel on the full 9 x 9 state (QSIM_NO_PRUNE=1) and the three opt-in mappings -- three rows of one column per lane sorted into tiers
    (QSIM_TIERED=1), split rows with helper lanes (QSIM_SPLIT=1), additively split rows (QSIM_ADDSPLIT=1: the long row's
    state carried as two parts that add up to the physical entry) -- take the same step sequence and agree with it to
    rounding and with the oracle to 1e-8, including dense-output points inside steps (parked stage vectors)."""
    from qsim_b200 import workloads as W
    from qsim_b200._ffi import Problem
    cqs, params, _, fn = W.config("cfg2")
    pr = np.ascontiguousarray(params[np.linspace(0, len(params) - 1, 40).astype(int)])
    ts = np.linspace(0.0, 5.0, 11)
    P = Problem(cqs)
    assert "tsit5_lean_r1c2_nnz5_t128b4_vm2" in P.kernel_name(fn)
    base, st0 = P.solve(fn, ts, params=pr)
    monkeypatch.setenv(env, "3" if env == "QSIM_MAX_MINB" else "1")
    P2 = Problem(cqs)
    assert tag in P2.kernel_name(fn)
    alt, st1 = P2.solve(fn, ts, params=pr)
    assert np.all(st1["retcode"] == 0)
    assert np.array_equal(st0["n_accept"], st1["n_accept"]) and np.array_equal(st0["n_reject"], st1["n_reject"])
    assert np.max(np.abs(alt - base)) <= 1e-9
    want, _ = co.OracleProblem(cqs).solve(co.UNITARY_PROPAGATOR, ts, pr, mode=0, n_threads=4)
    assert np.max(np.abs(alt - want) / np.maximum(np.abs(want), 1e-3)) <= 1e-8
    assert np.max(np.abs(base - want) / np.maximum(np.abs(want), 1e-3)) <= 1e-8


@pytest.mark.gpu
def test_blocked_variant_opt_in(monkeypatch):
    """QSIM_BLOCKED=1 selects the blocked kernels (3 consecutive rows per lane, dense 3 x 3 coupling blocks with
    values in registers) for the 2-transmon propagators: same step sequence, results within rounding of the
    default row-per-lane kernel and within 1e-8 of the oracle."""
    from qsim_b200 import workloads as W
    from qsim_b200._ffi import Problem
    cqs, params, _, fn = W.config("cfg2")
    pr = np.ascontiguousarray(params[np.linspace(0, len(params) - 1, 40).astype(int)])
    ts = np.linspace(0.0, 5.0, 6)
    P = Problem(cqs)
    assert "_vm2_" in P.kernel_name(fn)
    base, st0 = P.solve(fn, ts, params=pr)
    monkeypatch.setenv("QSIM_BLOCKED", "1")
    P2 = Problem(cqs)
    assert "r3c1_nnz3" in P2.kernel_name(fn) and "_vm3_" in P2.kernel_name(fn)
    blk, st1 = P2.solve(fn, ts, params=pr)
    assert np.array_equal(st0["n_accept"], st1["n_accept"]) and np.array_equal(st0["n_reject"], st1["n_reject"])
    assert np.max(np.abs(blk - base)) <= 1e-9
    want, _ = co.OracleProblem(cqs).solve(co.UNITARY_PROPAGATOR, ts, pr, mode=0, n_threads=4)
    assert np.max(np.abs(blk - want) / np.maximum(np.abs(want), 1e-3)) <= 1e-8


@pytest.mark.gpu
def test_large_sweep_work_queue():
    """50 000 short trajectories in one launch: every sweep point is fetched from the work queue exactly once
    (results depend only on the point's parameters), nothing is skipped, and a sample agrees with the oracle."""
    from qsim_b200 import workloads as W
    from qsim_b200._ffi import Problem
    cqs, params, _, fn = W.config("cfg2")
    n = 50_000
    pr = np.ascontiguousarray(params[np.arange(n) % len(params)])
    ts = np.array([0.0, 0.25])
    P = Problem(cqs)
    got, st = P.solve(fn, ts, params=pr)
    assert np.all(st["retcode"] == 0) and np.all(st["t_final"] == 0.25)
    # rows n and n + 4096 carry the same parameters: identical results, bit for bit
    assert np.array_equal(got[:4096], got[4096:8192]) and np.array_equal(got[:4096], got[40960:45056])
    idx = np.array([0, 1, 4095, 4096, 17_123, 49_999])
    want, ost = co.OracleProblem(cqs).solve(co.UNITARY_PROPAGATOR, ts, pr[idx], mode=0, n_threads=4)
    assert np.array_equal(st["n_accept"][idx], ost["n_accept"])
    assert np.max(np.abs(got[idx] - want) / np.maximum(np.abs(want), 1e-3)) <= 1e-8


@pytest.mark.gpu
def test_cfg3_cfg4_cfg5_sweep_properties():
    """Size-independent properties on sweeps too large for the oracle: unitarity and U(t0) = I for the 3-transmon
    propagators (cfg3, 1184 points), trace preservation vec(I)' S = vec(I)' and S(t0) = I for the Lindblad
    superoperators (cfg4: d^2 = 81, 592 points; cfg5: d^2 = 729, 8 points), every trajectory successful, and
    independence of how the sweep is partitioned into launches (bitwise)."""
    def subset(name, n):
        cqs, params, _, fn = W.config(name)
        return cqs, np.ascontiguousarray(params[np.linspace(0, len(params) - 1, n).astype(int)]), fn
    # cfg3
    cqs, pr, fn = subset("cfg3", 1184)
    P = Problem(cqs)
    ts = np.array([0.0, 5.0, 10.0])
    u, st = P.solve(fn, ts, params=pr)
    assert np.all(st["retcode"] == 0) and np.all(st["t_final"] == 10.0)
    assert np.array_equal(u[:, 0], np.broadcast_to(np.eye(27), (len(pr), 27, 27)))
    assert np.abs(np.einsum("tji,tjk->tik", u[:, -1].conj(), u[:, -1]) - np.eye(27)).max() < 2e-3
    part, _ = P.solve(fn, ts, params=pr[500:900])
    assert np.array_equal(part, u[500:900])
    # cfg4 / cfg5: trace preservation of the superoperator
    for name, n, tf in (("cfg4", 592, 1.0), ("cfg5", 8, 0.1)):
        cqs, pr, fn = subset(name, n)
        d = cqs.dimension
        P = Problem(cqs)
        ts = np.array([0.0, tf])
        s, st = P.solve(fn, ts, params=pr)
        assert np.all(st["retcode"] == 0)
        assert np.array_equal(s[:, 0], np.broadcast_to(np.eye(d * d), (len(pr), d * d, d * d)))
        vec_eye = np.eye(d).reshape(-1, order="F")
        assert np.abs(np.einsum("i,tij->tj", vec_eye, s[:, -1]) - vec_eye).max() < 1e-6
        part, _ = P.solve(fn, ts, params=pr[n // 4: n // 2])
        assert np.array_equal(part, s[n // 4: n // 2])


@pytest.mark.gpu
def test_failed_trajectory_is_isolated():
    """A sweep point whose parameters are not finite fails on its own (retcode, NaN results) and leaves every
    other trajectory of the launch untouched -- the same outcome the oracle reports."""
    cqs, params, _, fn = W.config("cfg2")
    pr = np.ascontiguousarray(params[np.linspace(0, len(params) - 1, 64).astype(int)])
    ts = np.array([0.0, 1.0, 2.0])
    P = Problem(cqs)
    good, st_good = P.solve(fn, ts, params=pr)
    bad = pr.copy()
    bad[5, 0] = np.nan
    bad[40, 1] = np.inf
    got, st = P.solve(fn, ts, params=bad)
    want, ost = co.OracleProblem(cqs).solve(co.UNITARY_PROPAGATOR, ts, bad, n_threads=8)
    ok = np.ones(64, dtype=bool)
    ok[[5, 40]] = False
    assert np.array_equal(got[ok], good[ok]) and np.array_equal(st[ok], st_good[ok])
    assert np.all(st["retcode"][~ok] != 0) and np.array_equal(st["retcode"], ost["retcode"])
    assert np.isnan(got[~ok][:, -1]).all()


@pytest.mark.gpu
@pytest.mark.parametrize("kw", [
    dict(reltol=1e-9, abstol=1e-11),
    dict(reltol=1e-4, abstol=1e-6),
    dict(beta1=0.2, beta2=0.05, gamma=0.8),
    dict(qmin=0.5, qmax=3.0),
    dict(dtmax=0.004),
    dict(qoldinit=1e-2),
])
def test_solver_options_follow_the_oracle(kw):
    """Non-default tolerances and controller constants (the per-call power tables of the kernel are rebuilt from
    beta1, beta2): same accepted / rejected steps as the oracle and 1e-8 agreement."""
    cqs, params, _, fn = W.config("cfg2")
    pr = np.ascontiguousarray(params[[3, 2000, 4090]])
    check_parity(cqs, fn, np.linspace(0.0, 3.0, 4), pr, opts=_abi.default_opts(**kw))
