"""Floquet propagation: U(n*tau + dt) = U(dt) U(tau)^n for tau-periodic systems.

Same names and meaning as /root/reference/src/time_evolution.jl:166-360
  floquet_propagator(propagator_func, t_period[, rtol])                     :188-228
  floquet_rise_fall_propagator(propagator_func, t_period, rise, fall[, rtol]) :312-346
  choose_times_floquet / decompose_times_floquet / unique_tol               :238-252, 275-300, 355-360
but organised for sweeps.  The reference wraps a propagator function and, per system, runs up to three solves
and a running matrix product on the host.  Here

  floquet_sweep(cqs, ts, t_period, params=..., rise_time=..., fall_time=...)

does a whole parameter sweep in ONE library call (qsim_floquet_propagator, csrc/qsim_floquet.cu): every segment of
every sweep point is a trajectory of a single batched solver launch with per-trajectory save grids -- each point
may have its own period, e.g. when the drive frequency is a sweep axis -- and the powers and products are formed
on the device; only the requested propagators come back.  The reference-style wrappers returned by
`floquet_propagator(unitary_propagator, tau)` route to that call when they wrap this package's
`unitary_propagator` / `me_propagator`, and otherwise (any callable (cqs, ts) -> propagators, like the reference
allows) combine the callable's results with batched numpy products.
"""
from __future__ import annotations

import math
from typing import Callable, Optional, Sequence, Tuple

import numpy as np


# ---- time bookkeeping (host, O(n_ts)) ------------------------------------------------------------------------
def unique_tol(ts: Sequence[float], dt: float) -> Tuple[np.ndarray, np.ndarray]:
    """Values snapped to the grid of spacing `dt`, de-duplicated in first-seen order, and for every input the index
    of its representative: a[inds] approximates ts to within dt/2 (time_evolution.jl:355-360)."""
    snapped = np.round(np.asarray(ts, dtype=float) / dt) * dt
    sorted_vals, first_seen, inverse = np.unique(snapped, return_index=True, return_inverse=True)
    order = np.argsort(first_seen, kind="stable")            # np.unique sorts by value: restore first-seen order
    rank = np.empty_like(order)
    rank[order] = np.arange(len(order))
    return sorted_vals[order], rank[inverse]


def _eps(x: float) -> float:
    return float(np.spacing(abs(x)))


def decompose_times_floquet(ts: Sequence[float], t_period: float, rtol: Optional[float] = None):
    """ts = quotients * t_period + unique_remainders[unique_inds] with as few remainders as possible, sorted, and
    unique_remainders[0] == ts[0] (time_evolution.jl:275-300; rtol defaults to eps(t_period))."""
    rtol = _eps(t_period) if rtol is None else rtol
    ts = np.asarray(ts, dtype=float)
    shifted = ts - ts[0]
    remainders = np.mod(shifted, t_period)
    quotients = np.round((shifted - remainders) / t_period)   # fld, consistent with mod by construction
    vals, inds = unique_tol(remainders + ts[0], rtol * t_period)
    order = np.argsort(vals, kind="stable")
    rank = np.empty_like(order)
    rank[order] = np.arange(len(order))
    return quotients, vals[order], rank[inds]


def choose_times_floquet(center: float, width: float, t_period: float, dt: float) -> np.ndarray:
    """An odd number of equally spaced times around `center` (which is one of them) whose spacing divides the
    period, so that few distinct remainders occur (time_evolution.jl:238-252)."""
    assert width >= 0 and dt >= 0
    if width == 0 or dt == 0:
        return np.array([center], dtype=float)
    step = t_period / math.ceil(t_period / dt)
    count = int(math.floor(width / step))
    count += (count + 1) % 2                                  # odd, for symmetry about the centre
    half = step * count / 2
    times = np.linspace(center - half, center + half, count)
    times[(count - 1) // 2] = center
    return times


# ---- the sweep-wide call -------------------------------------------------------------------------------------
def floquet_sweep(cqs, ts, t_period, params=None, which: str = "unitary_propagator", rise_time: float = 0.0,
                  fall_time: float = 0.0, rtol: Optional[float] = None, opts=None, ctx=None, return_stats: bool = False):
    """Propagators at every ts[k] for systems that are periodic with `t_period` (a scalar, or one period per sweep
    point) between ts[0] + rise_time and ts[-1] - fall_time.  Returns [n_ts, n, n], or [n_traj, n_ts, n, n] with
    `params`, n = d (unitary_propagator) or d^2 (me_propagator)."""
    from .time_evolution import _problem_for
    ts = np.asarray(ts, dtype=np.float64)
    assert np.all(np.diff(ts, axis=-1) >= 0), "issorted(ts)"
    n_params = None if params is None else np.atleast_2d(params).shape[1]
    prob = _problem_for(cqs, n_params)
    out, stats = prob.floquet(which, ts, t_period, params=params, rise_time=rise_time, fall_time=fall_time,
                              rtol=0.0 if rtol is None else rtol, opts=opts, ctx=ctx)
    res = out[0] if params is None else out
    return (res, stats) if return_stats else res


# ---- reference-style wrappers --------------------------------------------------------------------------------
def _library_entry(propagator_func) -> Optional[str]:
    from . import time_evolution as te
    if propagator_func is te.unitary_propagator:
        return "unitary_propagator"
    if propagator_func is te.me_propagator:
        return "me_propagator"
    return None


def _compose_periodic(propagator_func: Callable, cqs, ts: np.ndarray, t_period: float, rtol: float) -> np.ndarray:
    """Generic path for an arbitrary propagator function: one call for the unique remainders plus one period, then
    all powers and products at once (stacked matrices)."""
    quotients, remainders, inds = decompose_times_floquet(ts, t_period, rtol)
    stack = np.asarray(propagator_func(cqs, np.concatenate([remainders, [t_period + ts[0]]])))
    u_period = stack[-1]
    levels = np.unique(quotients).astype(int)
    powers = np.stack([np.linalg.matrix_power(u_period, int(v)) for v in levels])
    return stack[inds] @ powers[np.searchsorted(levels, quotients.astype(int))]


def floquet_propagator(propagator_func: Callable, t_period: float, rtol: Optional[float] = None) -> Callable:
    """A propagator function for systems periodic with `t_period` (time_evolution.jl:188-228)."""
    entry = _library_entry(propagator_func)
    tol = _eps(t_period) if rtol is None else rtol

    def p(cqs, ts, **kw):
        ts = np.asarray(ts, dtype=float)
        assert np.all(np.diff(ts) >= 0), "issorted(ts)"
        if entry is not None:
            return floquet_sweep(cqs, ts, t_period, which=entry, rtol=tol, **kw)
        return _compose_periodic(propagator_func, cqs, ts, t_period, tol)

    return p


def floquet_rise_fall_propagator(propagator_func: Callable, t_period: float, rise_time: float, fall_time: float,
                                 rtol: Optional[float] = None) -> Callable:
    """Floquet evolution between a non-periodic rise at the start of `ts` and a fall at its end
    (time_evolution.jl:312-346)."""
    assert t_period >= 0 and rise_time >= 0 and fall_time >= 0
    entry = _library_entry(propagator_func)
    tol = _eps(t_period) if rtol is None else rtol

    def p(cqs, ts, **kw):
        ts = np.asarray(ts, dtype=float)
        t0, t1 = ts[0] + rise_time, ts[-1] - fall_time
        assert t0 <= t1
        if entry is not None:
            return floquet_sweep(cqs, ts, t_period, which=entry, rise_time=rise_time, fall_time=fall_time, rtol=tol, **kw)
        before, after = ts < t0, ts > t1
        rise = np.asarray(propagator_func(cqs, np.concatenate([ts[before], [t0]])))
        middle = _compose_periodic(propagator_func, cqs, np.concatenate([[t0], ts[~before & ~after], [t1]]), t_period, tol)
        fall = np.asarray(propagator_func(cqs, np.concatenate([[t1], ts[after]])))
        u_t0 = rise[-1]
        u_t1 = middle[-1] @ u_t0
        return np.concatenate([rise[:-1], middle[1:-1] @ u_t0, fall[1:] @ u_t1])

    return p
