"""numpy restatement of the reference's time-evolution path -- TEST INFRASTRUCTURE.

PARITY UNPINNED at the 1e-8 level: Julia is not installed here, the reference
cannot run, and its integrator (OrdinaryDiffEq, compat "5.29", DiffEqBase "6.20",
/root/reference/Project.toml:8-20; Manifest.toml git-ignored) is neither vendored
nor version-pinned.  This file restates
  * the four right-hand sides exactly as the reference writes them, dense:
      unitary_propagator  /root/reference/src/time_evolution.jl:31-37
      unitary_state       /root/reference/src/time_evolution.jl:65-71
      me_propagator       /root/reference/src/time_evolution.jl:98-111  (Kronecker superoperator)
      me_state            /root/reference/src/time_evolution.jl:142-154
    driven through closures `ham(t)` like the reference (composite_systems.jl:46-50), and
  * OrdinaryDiffEq's published Tsit5 algorithm as called at time_evolution.jl:43,75,119,160
    (`solve(prob, Tsit5(); saveat=ts, save_start=true, reltol=1e-6, abstol=1e-8)`):
    Tsitouras 5(4) tableau, RMS error norm on complex entries, PI step controller,
    Hairer-Wanner initial step, 4th-order free interpolant for saveat.
It is pinned by the reference's own known-answer tests (tests/test_oracle_*.py), by the
Runge-Kutta order conditions, and by tight-tolerance solves -- not by golden vectors, which
the reference does not ship.  Only tests/, __graft_entry__.smoke() and bench.py's CPU
baseline may import this module; the product never does.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Callable, List, Optional, Sequence

import numpy as np

# ---- Tsit5 constants (OrdinaryDiffEq Tsit5ConstantCache, Float64) -------------
C = (0.0, 0.161, 0.327, 0.9, 0.9800255409045097, 1.0, 1.0)
A = (
    (),
    (0.161,),
    (-0.008480655492356989, 0.335480655492357),
    (2.8971530571054935, -6.359448489975075, 4.3622954328695815),
    (5.325864828439257, -11.748883564062828, 7.4955393428898365, -0.09249506636175525),
    (5.86145544294642, -12.92096931784711, 8.159367898576159, -0.071584973281401, -0.028269050394068383),
    (0.09646076681806523, 0.01, 0.4798896504144996, 1.379008574103742, -3.290069515436081, 2.324710524099774),
)
BTILDE = (-0.00178001105222577714, -0.0008164344596567469, 0.007880878010261995, -0.1447110071732629,
          0.5823571654525552, -0.45808210592918697, 0.015151515151515152)
# dense output: b_1(T) = T*(r11 + T*(r12 + T*(r13 + T*r14))), b_i(T) = T^2*(ri2 + T*(ri3 + T*ri4))
R = (
    (1.0, -2.763706197274826, 2.9132554618219126, -1.0530884977290216),
    (0.13169999999999998, -0.2234, 0.1017),
    (3.9302962368947516, -5.941033872131505, 2.490627285651253),
    (-12.411077166933676, 30.33818863028232, -16.548102889244902),
    (37.50931341651104, -88.1789048947664, 47.37952196281928),
    (-27.896526289197286, 65.09189467479366, -34.87065786149661),
    (1.5, -4.0, 2.5),
)


def interp_weights(theta: float):
    b = [theta * (R[0][0] + theta * (R[0][1] + theta * (R[0][2] + theta * R[0][3])))]
    t2 = theta * theta
    for i in range(1, 7):
        b.append(t2 * (R[i][0] + theta * (R[i][1] + theta * R[i][2])))
    return b


@dataclass
class Options:
    reltol: float = 1e-6
    abstol: float = 1e-8
    gamma: float = 0.9
    qmin: float = 0.2
    qmax: float = 10.0
    beta1: float = 7 / 50
    beta2: float = 2 / 25
    qoldinit: float = 1e-4
    dtmax: float = 0.0
    maxiters: int = 1_000_000


@dataclass
class Stats:
    retcode: int = 0
    n_accept: int = 0
    n_reject: int = 0
    n_rhs: int = 0
    dt_init: float = 0.0
    t_final: float = 0.0


def _rms(x: np.ndarray) -> float:
    return math.sqrt(float(np.sum(np.abs(x) ** 2)) / x.size)


def tsit5_solve(f: Callable[[np.ndarray, float], np.ndarray], u0: np.ndarray, ts: Sequence[float],
                opts: Optional[Options] = None):
    """Integrate du/dt = f(u, t) from ts[0], saving at every ts[k].  Returns (list of u, Stats)."""
    o = opts or Options()
    ts = [float(t) for t in ts]
    assert all(ts[i] <= ts[i + 1] for i in range(len(ts) - 1)), "ts must be sorted"
    st = Stats()
    t0, tend = ts[0], ts[-1]
    u = np.array(u0, dtype=complex)
    out: List[Optional[np.ndarray]] = [None] * len(ts)
    nsave = 0
    while nsave < len(ts) and ts[nsave] <= t0:   # save_start=true
        out[nsave] = u.copy()
        nsave += 1
    st.t_final = t0
    if tend <= t0:
        return out, st
    dtmax = o.dtmax if o.dtmax > 0 else (tend - t0)

    # ---- initial step (DiffEqBase ode_determine_initdt) ----
    sk = o.abstol + np.abs(u) * o.reltol
    f0 = f(u, t0)
    st.n_rhs += 1
    d0 = _rms(u / sk)
    d1 = _rms(f0 / sk)
    if d0 < 1e-5 or d1 < 1e-5:
        dt0 = 1e-6
    else:
        dt0 = (d0 / d1) / 100
    dt0 = min(dt0, dtmax)
    u1 = u + dt0 * f0
    f1 = f(u1, t0 + dt0)
    st.n_rhs += 1
    d2 = _rms((f1 - f0) / sk) / dt0
    md = max(d1, d2)
    if md <= 1e-15:
        dt1 = max(1e-6, dt0 * 1e-3)
    else:
        dt1 = 10.0 ** (-(2 + math.log10(md)) / 5)
    dt = min(100 * dt0, dt1, dtmax)
    st.dt_init = dt

    k = [None] * 7
    k[0] = f(u, t0)   # the integrator's own first-stage evaluation (FSAL start)
    st.n_rhs += 1
    t = t0
    qold = o.qoldinit
    iters = 0
    while t < tend:
        iters += 1
        if iters > o.maxiters:
            st.retcode = 1
            break
        dt = min(dt, dtmax)
        dt = min(dt, tend - t)
        if not (dt > abs(t) * 2.220446049250313e-16):
            st.retcode = 2
            break
        for i in range(1, 6):
            acc = A[i][0] * k[0]
            for j in range(1, i):
                acc = acc + A[i][j] * k[j]
            k[i] = f(u + dt * acc, t + C[i] * dt)
        acc = A[6][0] * k[0]
        for j in range(1, 6):
            acc = acc + A[6][j] * k[j]
        unew = u + dt * acc
        k[6] = f(unew, t + dt)
        st.n_rhs += 6
        acc = BTILDE[0] * k[0]
        for j in range(1, 7):
            acc = acc + BTILDE[j] * k[j]
        utilde = dt * acc
        resid = utilde / (o.abstol + np.maximum(np.abs(u), np.abs(unew)) * o.reltol)
        eest = _rms(resid)
        if not math.isfinite(eest):
            st.retcode = 3
            break
        # PI controller (stepsize_controller!)
        if eest == 0.0:
            q = 1.0 / o.qmax
            q11 = 0.0
        else:
            q11 = eest ** o.beta1
            q = q11 / (qold ** o.beta2)
            q = max(1.0 / o.qmax, min(1.0 / o.qmin, q / o.gamma))
        if eest <= 1.0:
            st.n_accept += 1
            if q == 1.0:   # qsteady_min == qsteady_max == 1
                q = 1.0
            qold = max(eest, o.qoldinit)
            dtnew = dt / q
            tnew = t + dt
            if abs(tnew - tend) < 10 * np.spacing(max(t, tend)):
                tnew = tend
            while nsave < len(ts) and ts[nsave] <= tnew:
                if ts[nsave] == tnew:
                    out[nsave] = unew.copy()
                else:
                    b = interp_weights((ts[nsave] - t) / dt)
                    acc = k[0] * b[0]
                    for j in range(1, 7):
                        acc = acc + k[j] * b[j]
                    out[nsave] = u + dt * acc
                nsave += 1
            t = tnew
            u = unew
            k[0] = k[6]
            dt = dtnew
        else:
            st.n_reject += 1
            dt = dt / min(1.0 / o.qmin, q11 / o.gamma)
    st.t_final = t
    return out, st


# ---- the four right-hand sides, dense, as the reference writes them -----------
def _parametric_adder(cqs, params):
    from qsim_b200.composite import add_parametric_hamiltonians

    def add(ham, t):
        add_parametric_hamiltonians(ham, cqs, t, params)
    return add


def _embedded(cqs, op, idxs):
    from qsim_b200.composite import embed_add
    d = cqs.dimension
    m = np.zeros((d, d), dtype=complex, order="F")
    embed_add(m, op, idxs)
    return m


def _lindblad_ops(cqs, t, params):
    from qsim_b200.operators import ParametricLindblad
    ops = [_embedded(cqs, op, idxs) for op, idxs in cqs.fixed_Ls]
    for fn, idxs in cqs.parametric_Ls:
        m = fn(t, params) if isinstance(fn, ParametricLindblad) else fn(t)
        ops.append(_embedded(cqs, m, idxs))
    return ops


def unitary_propagator(cqs, ts, params=None, opts=None):
    from qsim_b200.composite import composite_hamiltonian
    h0 = composite_hamiltonian(cqs)
    add = _parametric_adder(cqs, params)

    def ode(u, t):
        ham = h0.copy(order="F")
        add(ham, t)
        ham *= -2j * math.pi
        return ham @ u

    return tsit5_solve(ode, np.eye(cqs.dimension, dtype=complex), ts, opts)


def unitary_state(cqs, ts, psi0, params=None, opts=None):
    from qsim_b200.composite import composite_hamiltonian
    h0 = composite_hamiltonian(cqs)
    add = _parametric_adder(cqs, params)

    def ode(psi, t):
        ham = h0.copy(order="F")
        add(ham, t)
        ham *= -2j * math.pi
        return ham @ psi

    return tsit5_solve(ode, np.asarray(psi0, dtype=complex), ts, opts)


def me_propagator(cqs, ts, params=None, opts=None):
    from qsim_b200.composite import composite_hamiltonian
    h0 = composite_hamiltonian(cqs)
    add = _parametric_adder(cqs, params)
    d = cqs.dimension
    eye = np.eye(d, dtype=complex)

    def ode(u, t):
        ham = h0.copy(order="F")
        add(ham, t)
        du = (-2j * math.pi * (np.kron(eye, ham) - np.kron(ham.T, eye))) @ u
        for lm in _lindblad_ops(cqs, t, params):
            ldl = lm.conj().T @ lm
            du = du + (2 * math.pi * (np.kron(lm.conj(), lm) - 0.5 * np.kron(eye, ldl)
                                      - 0.5 * np.kron(ldl.T, eye))) @ u
        return du

    return tsit5_solve(ode, np.eye(d * d, dtype=complex), ts, opts)


def me_state(cqs, ts, rho0, params=None, opts=None):
    from qsim_b200.composite import composite_hamiltonian
    h0 = composite_hamiltonian(cqs)
    add = _parametric_adder(cqs, params)

    def ode(rho, t):
        ham = h0.copy(order="F")
        add(ham, t)
        drho = -2j * math.pi * (ham @ rho - rho @ ham)
        for lm in _lindblad_ops(cqs, t, params):
            lh = lm.conj().T
            drho = drho + 2 * math.pi * (lm @ rho @ lh - 0.5 * (lh @ lm @ rho) - 0.5 * (rho @ lh @ lm))
        return drho

    return tsit5_solve(ode, np.asarray(rho0, dtype=complex), ts, opts)
