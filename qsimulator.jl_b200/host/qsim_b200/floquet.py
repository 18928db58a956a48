"""Floquet wrappers: the main in-package callers of the hot path.

Host-side mirror of /root/reference/src/time_evolution.jl:166-360:
  floquet_propagator            :188-220   U(n*tau + dt) = U(dt) U(tau)^n for a tau-periodic system
  choose_times_floquet          :238-252
  decompose_times_floquet       :275-293
  floquet_rise_fall_propagator  :312-338   non-periodic rise / fall segments around a periodic middle
  unique_tol                    :355-360
A `propagator_func` is any callable (cqs, ts) -> sequence of propagators with the identity at ts[0]
(`unitary_propagator`, `me_propagator`); the wrappers only reorganise which times it is asked for and
combine the results with matrix powers and products, exactly like the reference.
"""
from __future__ import annotations

import math
from typing import Callable, List, Sequence, Tuple

import numpy as np


def unique_tol(ts: Sequence[float], dt: float) -> Tuple[np.ndarray, np.ndarray]:
    """Values rounded to multiples of `dt`, de-duplicated in first-seen order, and for every input the
    index of its representative (time_evolution.jl:355-360)."""
    vals = np.round(np.asarray(ts, dtype=float) / dt) * dt
    unique_vals: List[float] = []
    inds = np.empty(len(vals), dtype=int)
    for i, v in enumerate(vals):
        for j, uv in enumerate(unique_vals):
            if uv == v:
                inds[i] = j
                break
        else:
            unique_vals.append(v)
            inds[i] = len(unique_vals) - 1
    return np.array(unique_vals), inds


def decompose_times_floquet(ts: Sequence[float], t_period: float, rtol: float = None):
    """ts = quotients * t_period + unique_remainders[unique_inds], unique_remainders[0] == ts[0]
    (time_evolution.jl:275-293)."""
    if rtol is None:
        rtol = float(np.spacing(t_period))      # eps(t_period)
    ts = np.asarray(ts, dtype=float)
    quotients = np.floor_divide(ts - ts[0], t_period)
    remainders = np.mod(ts - ts[0], t_period) + ts[0]
    unique_remainders, unique_inds = unique_tol(remainders, rtol * t_period)
    sort_inds = np.argsort(unique_remainders, kind="stable")
    unique_remainders = unique_remainders[sort_inds]
    unique_inds = np.argsort(sort_inds, kind="stable")[unique_inds]
    return quotients, unique_remainders, unique_inds


def choose_times_floquet(center: float, width: float, t_period: float, dt: float) -> np.ndarray:
    """Times that maximise remainder overlap for Floquet integration; `center` is one of them
    (time_evolution.jl:238-252)."""
    assert width >= 0 and dt >= 0
    if width == 0 or dt == 0:
        return np.array([center], dtype=float)
    dt = t_period / math.ceil(t_period / dt)
    num_times = int(math.floor(width / dt))
    num_times += 1 if num_times % 2 == 0 else 0
    width = dt * num_times
    times = np.linspace(center - width / 2, center + width / 2, num_times)
    times[int(math.ceil(num_times / 2)) - 1] = center
    return times


def floquet_propagator(propagator_func: Callable, t_period: float, rtol: float = None) -> Callable:
    """A propagator function for systems periodic with `t_period` (time_evolution.jl:188-220)."""
    if rtol is None:
        rtol = float(np.spacing(t_period))

    def p(cqs, ts):
        ts = np.asarray(ts, dtype=float)
        assert np.all(np.diff(ts) >= 0), "issorted(ts)"
        quotients, unique_remainders, unique_inds = decompose_times_floquet(ts, t_period, rtol)
        us_remainders = propagator_func(cqs, np.concatenate([unique_remainders, [t_period + ts[0]]]))
        u_period = np.asarray(us_remainders[-1])
        us = []
        quotient = quotients[0]
        u_floquet = np.linalg.matrix_power(u_period, int(quotient))
        for i in range(len(ts)):
            if quotients[i] > quotient:
                u_floquet = u_floquet @ np.linalg.matrix_power(u_period, int(quotients[i] - quotient))
                quotient = quotients[i]
            us.append(np.asarray(us_remainders[unique_inds[i]]) @ u_floquet)
        return us

    return p


def floquet_rise_fall_propagator(propagator_func: Callable, t_period: float, rise_time: float, fall_time: float,
                                 rtol: float = None) -> Callable:
    """Floquet evolution with non-periodic rise and fall segments (time_evolution.jl:312-338)."""
    assert t_period >= 0 and rise_time >= 0 and fall_time >= 0
    floquet_prop = floquet_propagator(propagator_func, t_period, rtol)

    def p(cqs, ts):
        ts = np.asarray(ts, dtype=float)
        t0, t1 = ts[0] + rise_time, ts[-1] - fall_time
        assert t0 <= t1
        us_risetime = propagator_func(cqs, np.concatenate([ts[ts < t0], [t0]]))
        u0 = np.asarray(us_risetime[-1])
        us = [np.asarray(u) for u in us_risetime[:-1]]
        us_floquet = floquet_prop(cqs, np.concatenate([[t0], ts[(t0 <= ts) & (ts <= t1)], [t1]]))
        u1 = us_floquet[-1] @ u0
        us.extend(u @ u0 for u in us_floquet[1:-1])
        us_falltime = propagator_func(cqs, np.concatenate([[t1], ts[t1 < ts]]))
        us.extend(np.asarray(u) @ u1 for u in us_falltime[1:])
        return us

    return p
