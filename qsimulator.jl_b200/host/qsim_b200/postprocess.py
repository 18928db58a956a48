"""What the reference's sweeps keep of a propagator (/root/reference/doc/Examples.ipynb
cells 14-15, 31, 34-38, test/test_time_evolution.jl:56-57,72): transition populations abs2(u[i, j]), the block on
the computational (qubit) subspace, and the average gate fidelity of that block against a target.

The selection itself runs on the device (`select=` / `abs2=` of the solver entry points, qsim_solve_selected):
these helpers build the index lists and do the few flops that remain on the selected entries."""
from __future__ import annotations

from typing import Sequence

import numpy as np


def subspace_indices(dims: Sequence[int], levels: int = 2) -> np.ndarray:
    """Composite basis indices (last subsystem fastest, composite_systems.jl:86-128) of the states whose every
    subsystem is below `levels`: the 2^n computational states of n transmons."""
    idx = [0]
    for d in dims:
        idx = [i * d + k for i in idx for k in range(min(levels, d))]
    return np.array(idx, dtype=np.int32)


def subspace_block(dims: Sequence[int], levels: int = 2):
    """(select, shape): the `select=` list for the block of a propagator on the computational subspace, and the
    shape to give the selected entries: `u_sub = out.reshape(out.shape[:-1] + shape)` is indexed [row, col]."""
    idx = subspace_indices(dims, levels)
    sel = [(int(r), int(c)) for r in idx for c in idx]
    return sel, (len(idx), len(idx))


def populations(u: np.ndarray) -> np.ndarray:
    """abs2 of propagator entries or state amplitudes (what `abs2=True` returns directly)."""
    return u.real ** 2 + u.imag ** 2


def average_gate_fidelity(u_sub: np.ndarray, target: np.ndarray) -> np.ndarray:
    """Average gate fidelity of the (generally non-unitary, leakage) subspace block against a unitary target:
    F = (Tr(M M') + |Tr M|^2) / (n (n + 1)), M = target' u_sub.  Broadcasts over leading axes."""
    n = target.shape[0]
    m = np.conj(target.T) @ u_sub
    tr = np.trace(m, axis1=-2, axis2=-1)
    return (np.real(np.sum(m * np.conj(m), axis=(-2, -1))) + np.abs(tr) ** 2) / (n * (n + 1))


def superoperator(u: np.ndarray) -> np.ndarray:
    """conj(U) (x) U: the column-major vec superoperator of rho -> U rho U' (the convention of me_propagator,
    time_evolution.jl:85-87).  Broadcasts over leading axes."""
    uc = np.conj(u)
    return np.einsum("...ac,...bd->...abcd", uc, u).reshape(u.shape[:-2] + (u.shape[-2] ** 2, u.shape[-1] ** 2))


def kraus_mix(us: np.ndarray, weights=None) -> np.ndarray:
    """Average channel of a sweep over a noise parameter: sum_i w_i conj(U_i) (x) U_i over the first axis of
    `us` (doc/Examples.ipynb cells 34-38: quasi-static flux noise as a mixture of unitaries)."""
    us = np.asarray(us)
    w = np.full(us.shape[0], 1.0 / us.shape[0]) if weights is None else np.asarray(weights, dtype=float)
    return np.tensordot(w, superoperator(us), axes=(0, 0))
