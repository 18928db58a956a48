"""ctypes wrapper around oracle/libqsim_oracle.so -- TEST INFRASTRUCTURE (see qsim_oracle.c).

Same inputs as the product's C ABI (a lowered CompositeQSystem, a parameter table, a save grid)
so that tests can run both on identical data.  Never imported by the product.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Optional

import numpy as np

from qsim_b200 import _abi
from qsim_b200.composite import CompositeQSystem, ProblemSpec, lower

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libqsim_oracle.so")

UNITARY_PROPAGATOR, UNITARY_STATE, ME_PROPAGATOR, ME_STATE = 0, 1, 2, 3
FAITHFUL, STRUCTURED = 0, 1


def build(force: bool = False) -> str:
    src = os.path.join(HERE, "qsim_oracle.c")
    hdr = os.path.join(HERE, "..", "include", "qsim_b200.h")
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < max(
            os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["make", "-C", HERE, "-B", "libqsim_oracle.so"], stdout=subprocess.DEVNULL)
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()     # no-op when libqsim_oracle.so is newer than its sources
        L = C.CDLL(LIB_PATH)
        dp = C.POINTER(C.c_double)
        L.qso_problem_create.restype = C.c_void_p
        L.qso_problem_create.argtypes = [C.c_int, C.c_int]
        L.qso_problem_destroy.argtypes = [C.c_void_p]
        L.qso_problem_set_h0.argtypes = [C.c_void_p, dp]
        L.qso_problem_add_term.argtypes = [C.c_void_p, C.POINTER(_abi.TermStruct), dp, dp, C.POINTER(C.c_int32)]
        L.qso_problem_add_lindblad.argtypes = [C.c_void_p, dp, C.POINTER(_abi.SignalStruct)]
        L.qso_problem_add_charge_term.argtypes = [C.c_void_p, C.POINTER(_abi.TermStruct), C.POINTER(C.c_int32), C.c_int]
        L.qso_problem_add_table.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_int, C.c_int, dp]
        L.qso_problem_add_table.restype = C.c_int
        L.qso_problem_eval.argtypes = [C.c_void_p, C.c_int, C.c_double, dp, dp]
        L.qso_tsit5_constants.argtypes = [dp, dp, dp, dp]
        L.qso_solve.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int64, dp, dp, C.c_int, C.c_int64, dp, C.c_int64,
                                C.POINTER(_abi.SolverOpts), dp, C.c_void_p, C.c_int]
        _lib = L
    return _lib


class OracleProblem:
    def __init__(self, spec_or_cqs, n_params: Optional[int] = None):
        spec = spec_or_cqs if isinstance(spec_or_cqs, ProblemSpec) else lower(spec_or_cqs, n_params)
        self.spec = spec
        self.d = spec.d
        self.n_params = spec.n_params
        L = lib()
        self._h = L.qso_problem_create(spec.d, spec.n_params)
        p, keep = _abi.cplx_ptr(spec.h0)
        L.qso_problem_set_h0(self._h, p)
        table_ids = {}
        for tb in spec.tables or []:
            table_ids[id(tb)] = L.qso_problem_add_table(self._h, tb.t_start, tb.t_end, tb.n_seg, tb.n_coef,
                                                        tb.coef.ctypes.data_as(C.POINTER(C.c_double)))
        for lt, a, b, levels in spec.terms:
            ts = _abi.make_term(lt, table_ids)
            pa, ka = _abi.cplx_ptr(a)
            pb, kb = _abi.cplx_ptr(b)
            pl = None
            if levels is not None:
                lv = np.ascontiguousarray(levels, dtype=np.int32)
                pl = lv.ctypes.data_as(C.POINTER(C.c_int32))
            if lt.kind == _abi.TERM_SPECTRUM:
                # the oracle diagonalises the charge-basis Hamiltonian on every RHS call, like the reference
                L.qso_problem_add_charge_term(self._h, C.byref(ts), pl, int(lt.dim))
                continue
            L.qso_problem_add_term(self._h, C.byref(ts), pa, pb, pl)
        for op, scale in spec.lindblads:
            po, ko = _abi.cplx_ptr(op)
            if scale is None:
                L.qso_problem_add_lindblad(self._h, po, None)
            else:
                ss = _abi.make_signal(scale, table_ids)
                L.qso_problem_add_lindblad(self._h, po, C.byref(ss))

    def __del__(self):
        try:
            if self._h:
                lib().qso_problem_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def eval(self, kind: int, t: float, param_row=None) -> np.ndarray:
        n = self.d if kind == 0 else self.d * self.d
        out = np.zeros((n, n), dtype=np.complex128, order="F")
        pr, kr = _abi.dbl_ptr(param_row)
        lib().qso_problem_eval(self._h, kind, float(t), pr, out.ctypes.data_as(C.POINTER(C.c_double)))
        return out

    def solve(self, which: int, ts, params=None, init=None, opts=None, mode: int = STRUCTURED, n_threads: int = 1):
        """Returns (out, stats): out[n_traj, n_ts, ...] complex (inner matrices column-major -> numpy
        arrays indexed [row, col]), stats structured array."""
        d = self.d
        ts = np.ascontiguousarray(ts, dtype=np.float64)
        if params is None:
            params = np.zeros((1, max(self.n_params, 1)))
        params = np.ascontiguousarray(np.atleast_2d(params), dtype=np.float64)
        n_traj = params.shape[0]
        if self.n_params and params.shape[1] != self.n_params:
            raise ValueError("params has the wrong number of columns")
        if ts.ndim == 2:
            if ts.shape[0] != n_traj:
                raise ValueError("per-trajectory ts needs one row per trajectory")
            n_ts, ts_stride = ts.shape[1], ts.shape[1]
        else:
            n_ts, ts_stride = ts.shape[0], 0
        shape = {0: (d, d), 1: (d,), 2: (d * d, d * d), 3: (d, d)}[which]
        n = int(np.prod(shape))
        pin, kin, init_stride = None, None, 0
        if which in (1, 3):
            init = np.asarray(init, dtype=np.complex128)
            if init.ndim == len(shape) + 1:
                kin = np.ascontiguousarray(np.stack([np.asfortranarray(x).ravel(order="F") for x in init]))
                init_stride = n
            else:
                kin = np.ascontiguousarray(init.ravel(order="F"))
            pin = kin.ctypes.data_as(C.POINTER(C.c_double))
        out = np.empty((n_traj, n_ts, n), dtype=np.complex128)
        stats = np.zeros(n_traj, dtype=_abi.STATS_DTYPE)
        o = opts if opts is not None else _abi.default_opts()
        rc = lib().qso_solve(self._h, which, mode, n_traj, params.ctypes.data_as(C.POINTER(C.c_double)),
                             ts.ctypes.data_as(C.POINTER(C.c_double)), n_ts, ts_stride, pin, init_stride,
                             C.byref(o), out.ctypes.data_as(C.POINTER(C.c_double)), stats.ctypes.data_as(C.c_void_p),
                             int(n_threads))
        if rc != 0:
            raise ValueError("oracle rejected the arguments (ts not sorted?)")
        if len(shape) == 2:
            out = out.reshape(n_traj, n_ts, shape[1], shape[0]).transpose(0, 1, 3, 2)
        return out, stats


def tsit5_constants():
    c = np.zeros(7); a = np.zeros((7, 6)); bt = np.zeros(7); r = np.zeros((7, 4))
    dp = C.POINTER(C.c_double)
    lib().qso_tsit5_constants(c.ctypes.data_as(dp), a.ctypes.data_as(dp), bt.ctypes.data_as(dp), r.ctypes.data_as(dp))
    return c, a, bt, r
