"""CompositeQSystem: term lists and Hamiltonian assembly.

Host-side mirror of /root/reference/src/composite_systems.jl (SURVEY.md
section 8 rows a6-a9): the struct :7-16, add_hamiltonian! :29-43, add_lindblad!
:53-62, add_parametric_hamiltonians! :46-50, embed_add! :65-72, embed :86-128,
embed_indices :136-148, hamiltonian(cqs) :151-157.

Kronecker convention: the LAST subsystem is the fastest index (:95-110).
Where the reference finds the embedding positions by embedding a label matrix
and scanning it, this computes them directly from the mixed-radix digits; the
result is the same index lists (0-based here, column-major linear indices).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Callable, List, Optional, Sequence, Tuple, Union

import numpy as np

from .operators import (LoweredTerm, ParametricHamiltonian, ParametricLindblad, TERM_DUFFING,
                        TERM_ROTATING, TERM_SCALAR, TERM_TRANSMON)
from .signals import Signal
from .systems import QSystem, dimension, hamiltonian as _sys_hamiltonian, label


@dataclass
class CompositeQSystem:
    subsystems: List[QSystem]
    fixed_Hs: List[Tuple[np.ndarray, list]] = field(default_factory=list)
    parametric_Hs: List[Tuple[Callable, list]] = field(default_factory=list)
    fixed_Ls: List[Tuple[np.ndarray, list]] = field(default_factory=list)
    parametric_Ls: List[Tuple[Callable, list]] = field(default_factory=list)
    # acting_on (subsystem positions) per term, kept for the C-ABI lowering
    _parametric_H_acting: List[List[int]] = field(default_factory=list)
    _fixed_L_acting: List[List[int]] = field(default_factory=list)
    _parametric_L_acting: List[List[int]] = field(default_factory=list)

    def __post_init__(self):
        self.subsystems = list(self.subsystems)

    @property
    def dimension(self) -> int:
        return int(np.prod([dimension(q) for q in self.subsystems]))

    @property
    def dims(self) -> List[int]:
        return [dimension(q) for q in self.subsystems]


def find_indices(cqs: CompositeQSystem, s) -> List[int]:
    """Positions of subsystem(s) `s` (object, label, or list of either) (:22-26)."""
    if isinstance(s, (list, tuple)):
        out: List[int] = []
        for x in s:
            out.extend(find_indices(cqs, x))
        return out
    if isinstance(s, str):
        return [i for i, sub in enumerate(cqs.subsystems) if label(sub) == s]
    return [i for i, sub in enumerate(cqs.subsystems) if sub == s]


def _digit_strides(dims: Sequence[int]) -> List[int]:
    strides = [1] * len(dims)
    for i in range(len(dims) - 2, -1, -1):
        strides[i] = strides[i + 1] * dims[i + 1]
    return strides


def embed_indices(acting_on: Sequence[int], dims: Sequence[int]) -> List[np.ndarray]:
    """For every column-major linear index `ct` of the small operator on `acting_on`
    (in that order), the column-major linear indices of the big matrix where it lands."""
    dims = list(dims)
    acting_on = list(acting_on)
    if len(set(acting_on)) != len(acting_on):
        raise ValueError("acting_on has repeated subsystems")
    D = int(np.prod(dims))
    strides = _digit_strides(dims)
    sub_dims = [dims[a] for a in acting_on]
    da = int(np.prod(sub_dims))
    sub_strides = _digit_strides(sub_dims)
    rest = [i for i in range(len(dims)) if i not in acting_on]
    # offsets of the identity part (all combinations of the untouched digits)
    rest_offsets = np.zeros(1, dtype=np.int64)
    for r in rest:
        rest_offsets = (rest_offsets[:, None] + (np.arange(dims[r]) * strides[r])[None, :]).ravel()
    rest_offsets.sort()
    # big-space offset of each small basis index
    small_to_big = np.zeros(da, dtype=np.int64)
    for a in range(da):
        off = 0
        for j, s in enumerate(acting_on):
            digit = (a // sub_strides[j]) % sub_dims[j]
            off += digit * strides[s]
        small_to_big[a] = off
    out = []
    for ct in range(da * da):
        r, c = ct % da, ct // da
        rows = small_to_big[r] + rest_offsets
        cols = small_to_big[c] + rest_offsets
        out.append(np.sort(rows + cols * D))
    return out


def embed(m: np.ndarray, acting_on: Sequence[int], dims: Sequence[int]) -> np.ndarray:
    """Embed operator `m` on subsystem positions `acting_on` into the full space (:86-128)."""
    m = np.asarray(m)
    dims = list(dims)
    D = int(np.prod(dims))
    if m.shape[0] != int(np.prod([dims[a] for a in acting_on])):
        raise AssertionError("Oops! Dimensions of operator do not match dims argument.")
    M = np.zeros(D * D, dtype=m.dtype)
    flat = m.ravel(order="F")
    for ct, idxs in enumerate(embed_indices(acting_on, dims)):
        M[idxs] = flat[ct]
    return M.reshape((D, D), order="F")


def embed_add(op: np.ndarray, added_op: np.ndarray, expand_idxs) -> None:
    """op[idx] += added_op[ct] for idx in expand_idxs[ct] (:65-72); op is modified in place."""
    flat = op.reshape(-1, order="F") if op.flags.f_contiguous else None
    if flat is None:
        raise ValueError("embed_add needs a column-major (Fortran-ordered) target")
    small = np.asarray(added_op).ravel(order="F")
    for ct, idxs in enumerate(expand_idxs):
        flat[idxs] += small[ct]


def _acting(cqs, acting_on) -> List[int]:
    idx = find_indices(cqs, acting_on)
    if not idx:
        raise ValueError("acting_on does not name a subsystem of this CompositeQSystem")
    return idx


def add_hamiltonian(cqs: CompositeQSystem, ham, acting_on=None, phi: Optional[float] = None) -> None:
    """add_hamiltonian!(cqs, matrix, acting_on) / (cqs, q) / (cqs, q, phi) / (cqs, function, acting_on)."""
    if isinstance(ham, QSystem):
        q = ham
        if phi is None and isinstance(acting_on, (int, float)):
            phi = acting_on  # add_hamiltonian!(cqs, q, phi) (:36)
        mat = _sys_hamiltonian(q) if phi is None else _sys_hamiltonian(q, phi)
        idxs = embed_indices(_acting(cqs, q), cqs.dims)
        cqs.fixed_Hs.append((np.asarray(mat), idxs))
        return
    act = _acting(cqs, acting_on)
    idxs = embed_indices(act, cqs.dims)
    if callable(ham):
        cqs.parametric_Hs.append((ham, idxs))
        cqs._parametric_H_acting.append(act)
    else:
        cqs.fixed_Hs.append((np.asarray(ham), idxs))


def add_lindblad(cqs: CompositeQSystem, lind_op, acting_on) -> None:
    act = _acting(cqs, acting_on)
    idxs = embed_indices(act, cqs.dims)
    if callable(lind_op):
        cqs.parametric_Ls.append((lind_op, idxs))
        cqs._parametric_L_acting.append(act)
    else:
        cqs.fixed_Ls.append((np.asarray(lind_op), idxs))
        cqs._fixed_L_acting.append(act)


def composite_hamiltonian(cqs: CompositeQSystem) -> np.ndarray:
    """Static part of the Hamiltonian (:151-157)."""
    d = cqs.dimension
    ham = np.zeros((d, d), dtype=complex, order="F")
    for new_ham, idxs in cqs.fixed_Hs:
        embed_add(ham, new_ham, idxs)
    return ham


def add_parametric_hamiltonians(ham: np.ndarray, cqs: CompositeQSystem, t: float, params=None) -> None:
    """(:46-50); `params` is the sweep row for descriptor-carrying terms."""
    for ham_func, idxs in cqs.parametric_Hs:
        mat = ham_func(t, params) if isinstance(ham_func, ParametricHamiltonian) else ham_func(t)
        embed_add(ham, mat, idxs)


# ---- lowering to the C ABI ---------------------------------------------------
@dataclass
class ProblemSpec:
    """Everything qsim_problem_* needs: the final term lists, embedded."""
    d: int
    n_params: int
    h0: np.ndarray
    terms: List[Tuple[LoweredTerm, Optional[np.ndarray], Optional[np.ndarray], Optional[np.ndarray]]]
    lindblads: List[Tuple[np.ndarray, Optional[Signal]]]
    tables: List = None     # distinct signals.Table objects referenced by the terms (qsim_problem_add_table order)


def _levels(cqs: CompositeQSystem, pos: int) -> np.ndarray:
    dims = cqs.dims
    stride = _digit_strides(dims)[pos]
    return ((np.arange(cqs.dimension) // stride) % dims[pos]).astype(np.int32)


def lower(cqs: CompositeQSystem, n_params: Optional[int] = None) -> ProblemSpec:
    d = cqs.dimension
    dims = cqs.dims
    terms = []
    max_p = -1
    for (fn, _), act in zip(cqs.parametric_Hs, cqs._parametric_H_acting):
        if not isinstance(fn, ParametricHamiltonian):
            raise TypeError(
                "time-dependent term is an opaque callable; the CUDA path needs descriptor drives "
                "(dipole_drive / rwa_dipole / parametric_drive / xy_exchange_drive with Signal arguments)")
        max_p = max(max_p, fn.max_param_index())
        for lt in fn.lowered():
            if lt.kind in (TERM_SCALAR, TERM_ROTATING):
                a = np.asfortranarray(embed(np.asarray(lt.atom_a, dtype=complex), act, dims))
                b = None if lt.atom_b is None else np.asfortranarray(
                    embed(np.asarray(lt.atom_b, dtype=complex), act, dims))
                terms.append((lt, a, b, None))
            else:
                if len(act) != 1:
                    raise ValueError("parametric_drive acts on exactly one subsystem")
                terms.append((lt, None, None, _levels(cqs, act[0])))
    lind = []
    for (op, _), act in zip(cqs.fixed_Ls, cqs._fixed_L_acting):
        lind.append((np.asfortranarray(embed(np.asarray(op, dtype=complex), act, dims)), None))
    for (fn, _), act in zip(cqs.parametric_Ls, cqs._parametric_L_acting):
        if not isinstance(fn, ParametricLindblad):
            raise TypeError("time-dependent Lindblad operator is an opaque callable; use ParametricLindblad")
        max_p = max(max_p, fn.max_param_index())
        lind.append((np.asfortranarray(embed(fn.op, act, dims)), fn.scale))
    if n_params is None:
        n_params = max_p + 1
    if n_params <= max_p:
        raise ValueError(f"n_params={n_params} but a term uses Param({max_p})")
    tables = []
    for sig in [t[0].signal for t in terms] + [sc for _, sc in lind if sc is not None]:
        if sig.table is not None and all(sig.table is not tb for tb in tables):
            tables.append(sig.table)
    for lt in [t[0] for t in terms]:
        for tb in lt.level_tables or []:     # spectra of DiagonalChargeBasisTransmon drives
            if all(tb is not x for x in tables):
                tables.append(tb)
    return ProblemSpec(d=d, n_params=n_params, h0=composite_hamiltonian(cqs), terms=terms, lindblads=lind,
                       tables=tables)
