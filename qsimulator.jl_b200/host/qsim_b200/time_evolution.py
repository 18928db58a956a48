"""The four solver entry points, same names and argument meaning as the reference.

  unitary_propagator(cqs, ts)        /root/reference/src/time_evolution.jl:29-47
  unitary_state(cqs, ts, psi0)       /root/reference/src/time_evolution.jl:63-79
  me_propagator(cqs, ts)             /root/reference/src/time_evolution.jl:96-123
  me_state(cqs, ts, rho0)            /root/reference/src/time_evolution.jl:140-164

Like the reference: `ts` must be sorted (AssertionError otherwise), integration starts at ts[0]
from the identity / the given state, one result per ts[k]; a scalar `t` means ts = [0.0, t] and
returns the last result.  The reference's sweeps are user-level loops that rebuild the system per
point; here a sweep is ONE call: pass `params` (n_traj x n_params, columns referenced by Param(i) in
the drives) and get an extra leading trajectory axis.  Everything runs in the CUDA library through
the C ABI (include/qsim_b200.h); there is no CPU fallback.
"""
from __future__ import annotations

import numbers
from typing import Optional

import numpy as np

from . import _abi
from ._ffi import Context, Problem, default_context
from .composite import CompositeQSystem

def _problem_for(cqs: CompositeQSystem, n_params: Optional[int]) -> Problem:
    """The compiled qsim_problem is cached on the CompositeQSystem, keyed by the term-list lengths so
    that adding a term invalidates it (the lists are append-only, composite_systems.jl:29-62)."""
    sig = (len(cqs.fixed_Hs), len(cqs.parametric_Hs), len(cqs.fixed_Ls), len(cqs.parametric_Ls), n_params)
    cached = getattr(cqs, "_qsim_problem", None)
    if cached is not None and cached[0] == sig:
        return cached[1]
    prob = Problem(cqs, n_params)
    cqs._qsim_problem = (sig, prob)
    return prob


def _check_ts(ts):
    if isinstance(ts, numbers.Real):
        return np.array([0.0, float(ts)]), True
    ts = np.asarray(ts, dtype=np.float64)
    if ts.ndim not in (1, 2) or ts.shape[-1] < 1:
        raise ValueError("ts must be a non-empty vector (or one row per trajectory)")
    assert np.all(np.diff(ts, axis=-1) >= 0), "issorted(ts)"   # @assert issorted(ts)
    return ts, False


def _solve(name, cqs, ts, init, params, opts, ctx, return_stats, select=None, abs2=False):
    ts, scalar = _check_ts(ts)
    n_params = None if params is None else np.atleast_2d(params).shape[1]
    prob = _problem_for(cqs, n_params)
    out, stats = prob.solve(name, ts, params=params, init=init, opts=opts, ctx=ctx, select=select, abs2=abs2)
    if params is None:
        out = out[0]
        res = out[-1] if scalar else out
    else:
        res = out[:, -1] if scalar else out
    return (res, stats) if return_stats else res


def unitary_propagator(cqs: CompositeQSystem, ts, params=None, opts: Optional[_abi.SolverOpts] = None,
                       ctx: Optional[Context] = None, return_stats: bool = False, select=None, abs2: bool = False):
    """Solve dU/dt = -2 pi i H(t) U, U(ts[0]) = I.  Returns U at every ts[k]: array [n_ts, d, d]
    (or [n_traj, n_ts, d, d] with `params`).

    select=[(row, col), ...] keeps only those entries of every U(ts[k]) on the device side (result
    [..., n_ts, len(select)]); with abs2=True their squared moduli -- the transition populations
    abs2(u[i, j]) that the reference's sweeps reduce each propagator to (test_time_evolution.jl:72)."""
    return _solve("unitary_propagator", cqs, ts, None, params, opts, ctx, return_stats, select, abs2)


def unitary_state(cqs: CompositeQSystem, ts, psi0, params=None, opts=None, ctx=None, return_stats: bool = False,
                  select=None, abs2: bool = False):
    """Solve d psi/dt = -2 pi i H(t) psi from psi0.  Returns [n_ts, d]; select=[i, ...] keeps components."""
    return _solve("unitary_state", cqs, ts, np.asarray(psi0, dtype=np.complex128), params, opts, ctx, return_stats,
                  select, abs2)


def me_propagator(cqs: CompositeQSystem, ts, params=None, opts=None, ctx=None, return_stats: bool = False,
                  select=None, abs2: bool = False):
    """Solve the vectorised Lindblad master equation for the d^2 x d^2 propagator (column-major vec).
    Returns [n_ts, d^2, d^2]."""
    return _solve("me_propagator", cqs, ts, None, params, opts, ctx, return_stats, select, abs2)


def me_state(cqs: CompositeQSystem, ts, rho0, params=None, opts=None, ctx=None, return_stats: bool = False,
             select=None, abs2: bool = False):
    """Solve d rho/dt = -2 pi i [H, rho] + 2 pi sum (L rho L' - 1/2 {L'L, rho}) from rho0.  Returns [n_ts, d, d];
    select=[(row, col), ...] keeps those entries of rho."""
    return _solve("me_state", cqs, ts, np.asarray(rho0, dtype=np.complex128), params, opts, ctx, return_stats,
                  select, abs2)
