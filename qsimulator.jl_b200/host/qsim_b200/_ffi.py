"""ctypes binding of libqsim_b200.so (include/qsim_b200.h).

The library is the product: if it is missing, or no CUDA device is usable, every solver call
raises -- there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import numpy as np

from . import _abi
from .composite import CompositeQSystem, ProblemSpec, lower

PKG_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
LIB_PATH = os.environ.get("QSIM_B200_LIB", os.path.join(PKG_ROOT, "lib", "libqsim_b200.so"))

EXPORTS = [
    "qsim_version", "qsim_last_error", "qsim_default_opts", "qsim_create", "qsim_destroy", "qsim_device_info",
    "qsim_synchronize", "qsim_alloc_pinned", "qsim_free_pinned", "qsim_problem_create", "qsim_problem_destroy",
    "qsim_problem_set_h0", "qsim_problem_add_term", "qsim_problem_add_spectrum_term", "qsim_problem_add_lindblad", "qsim_problem_add_table",
    "qsim_problem_finalize",
    "qsim_problem_eval", "qsim_unitary_propagator", "qsim_unitary_state", "qsim_me_propagator", "qsim_me_state",
    "qsim_solve_selected", "qsim_solve_multi", "qsim_problem_flops_per_rhs", "qsim_problem_kernel_name", "qsim_launch_count",
    "qsim_peak_fp64", "qsim_last_kernel_ms", "qsim_problem_plan_info", "qsim_problem_dims", "qsim_floquet_propagator",
]


class QsimError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"qsim_b200 error {code}: {msg}")
        self.code = code


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise QsimError(_abi.ERR_NO_DEVICE,
                        f"CUDA library not built: {LIB_PATH} (run __graft_entry__.build()); there is no CPU fallback")
    L = C.CDLL(LIB_PATH)
    dp = C.POINTER(C.c_double)
    vp = C.c_void_p
    L.qsim_version.restype = C.c_int
    L.qsim_last_error.restype = C.c_char_p
    L.qsim_default_opts.argtypes = [C.POINTER(_abi.SolverOpts)]
    L.qsim_default_opts.restype = None
    L.qsim_create.argtypes = [C.c_int, C.POINTER(vp)]
    L.qsim_destroy.argtypes = [vp]
    L.qsim_device_info.argtypes = [vp, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int), dp,
                                   C.POINTER(C.c_int64)]
    L.qsim_synchronize.argtypes = [vp]
    L.qsim_alloc_pinned.argtypes = [C.c_size_t, C.POINTER(vp)]
    L.qsim_free_pinned.argtypes = [vp]
    L.qsim_problem_create.argtypes = [C.c_int, C.c_int, C.POINTER(vp)]
    L.qsim_problem_destroy.argtypes = [vp]
    L.qsim_problem_set_h0.argtypes = [vp, dp]
    L.qsim_problem_add_term.argtypes = [vp, C.POINTER(_abi.TermStruct), dp, dp, C.POINTER(C.c_int32)]
    L.qsim_problem_add_lindblad.argtypes = [vp, dp, C.POINTER(_abi.SignalStruct)]
    L.qsim_problem_add_spectrum_term.argtypes = [vp, C.POINTER(_abi.SignalStruct), C.POINTER(_abi.Slot),
                                                 C.POINTER(C.c_int32), C.c_int, C.POINTER(C.c_int32)]
    L.qsim_problem_add_table.argtypes = [vp, C.c_double, C.c_double, C.c_int, C.c_int, dp, C.POINTER(C.c_int32)]
    L.qsim_problem_finalize.argtypes = [vp]
    L.qsim_problem_eval.argtypes = [vp, C.c_int, C.c_double, dp, dp]
    sig_prop = [vp, vp, C.c_int64, vp, vp, C.c_int, C.c_int64, C.POINTER(_abi.SolverOpts), vp, vp]
    sig_state = [vp, vp, C.c_int64, vp, vp, C.c_int, C.c_int64, vp, C.c_int64, C.POINTER(_abi.SolverOpts), vp, vp]
    L.qsim_unitary_propagator.argtypes = sig_prop
    L.qsim_me_propagator.argtypes = sig_prop
    L.qsim_unitary_state.argtypes = sig_state
    L.qsim_me_state.argtypes = sig_state
    L.qsim_solve_selected.argtypes = [vp, vp, C.c_int, C.c_int64, vp, vp, C.c_int, C.c_int64, vp, C.c_int64,
                                      C.POINTER(_abi.SolverOpts), C.c_int, vp, vp, C.c_int, vp, vp]
    L.qsim_solve_multi.argtypes = [C.POINTER(vp), C.c_int, vp, C.c_int, C.c_int64, vp, vp, C.c_int, C.c_int64, vp,
                                   C.c_int64, C.POINTER(_abi.SolverOpts), vp, vp]
    L.qsim_problem_flops_per_rhs.argtypes = [vp, C.c_int, dp]
    L.qsim_problem_kernel_name.argtypes = [vp, C.c_int, C.c_char_p, C.c_int]
    L.qsim_launch_count.argtypes = [vp]
    L.qsim_launch_count.restype = C.c_int64
    L.qsim_peak_fp64.argtypes = [vp, C.c_double, dp]
    L.qsim_last_kernel_ms.argtypes = [vp, dp]
    L.qsim_problem_dims.argtypes = [vp, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.qsim_floquet_propagator.argtypes = [vp, vp, C.c_int, C.c_int64, vp, vp, C.c_int, C.c_int64, vp, C.c_int64, C.c_double,
                                          C.c_double, C.c_double, C.POINTER(_abi.SolverOpts), vp, vp]
    L.qsim_problem_plan_info.argtypes = [vp, C.c_int, C.POINTER(C.c_int64)]
    _lib = L
    return L


def check(rc: int):
    if rc != 0:
        raise QsimError(rc, lib().qsim_last_error().decode())


class Context:
    """One CUDA device + stream (qsim_create)."""

    def __init__(self, device: int = 0):
        self._h = C.c_void_p()
        check(lib().qsim_create(int(device), C.byref(self._h)))
        self.device = device

    def close(self):
        if getattr(self, "_h", None):
            lib().qsim_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def info(self) -> dict:
        sm, ma, mi, mem = C.c_int(), C.c_int(), C.c_int(), C.c_int64()
        clk = C.c_double()
        check(lib().qsim_device_info(self._h, C.byref(sm), C.byref(ma), C.byref(mi), C.byref(clk), C.byref(mem)))
        return {"sm_count": sm.value, "cc": (ma.value, mi.value), "sm_clock_mhz": clk.value, "mem_bytes": mem.value}

    def synchronize(self):
        check(lib().qsim_synchronize(self._h))

    def launch_count(self) -> int:
        return int(lib().qsim_launch_count(self._h))

    def last_kernel_ms(self) -> float:
        """Device time of the kernels of the most recent solver call on this context (CUDA events)."""
        v = C.c_double()
        check(lib().qsim_last_kernel_ms(self._h, C.byref(v)))
        return v.value

    def peak_fp64(self, ms_target: float = 200.0) -> float:
        v = C.c_double()
        check(lib().qsim_peak_fp64(self._h, float(ms_target), C.byref(v)))
        return v.value


def device_available() -> bool:
    """False only when qsim_create reports QSIM_ERR_NO_DEVICE (no usable CUDA device); any other failure raises."""
    try:
        Context(0).close()
        return True
    except QsimError as e:
        if e.code == _abi.ERR_NO_DEVICE and os.path.exists(LIB_PATH):
            return False
        raise


_default_ctx: Optional[Context] = None


def default_context() -> Context:
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = Context(int(os.environ.get("LOCAL_RANK", "0")) if os.environ.get("QSIM_USE_LOCAL_RANK") else 0)
    return _default_ctx


class PinnedArray:
    """numpy view over page-locked host memory (qsim_alloc_pinned)."""

    def __init__(self, shape, dtype):
        self.shape = tuple(int(x) for x in shape)
        self.dtype = np.dtype(dtype)
        nbytes = max(int(np.prod(self.shape)) * self.dtype.itemsize, 1)
        self._p = C.c_void_p()
        check(lib().qsim_alloc_pinned(nbytes, C.byref(self._p)))
        buf = (C.c_char * nbytes).from_address(self._p.value)
        self.array = np.frombuffer(buf, dtype=self.dtype, count=int(np.prod(self.shape))).reshape(self.shape)

    def free(self):
        if getattr(self, "_p", None) and self._p.value:
            self.array = None
            lib().qsim_free_pinned(self._p)
            self._p = C.c_void_p()

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


dp_t = C.POINTER(C.c_double)
WHICH = {"unitary_propagator": 0, "unitary_state": 1, "me_propagator": 2, "me_state": 3}


class Problem:
    """qsim_problem built from a CompositeQSystem's final term lists (host-only until solved)."""

    def __init__(self, spec_or_cqs, n_params: Optional[int] = None):
        spec = spec_or_cqs if isinstance(spec_or_cqs, ProblemSpec) else lower(spec_or_cqs, n_params)
        self.spec = spec
        self.d = spec.d
        self.n_params = spec.n_params
        L = lib()
        self._h = C.c_void_p()
        check(L.qsim_problem_create(spec.d, spec.n_params, C.byref(self._h)))
        p, keep = _abi.cplx_ptr(spec.h0)
        check(L.qsim_problem_set_h0(self._h, p))
        table_ids = {}
        for tb in spec.tables or []:
            tid = C.c_int32()
            check(L.qsim_problem_add_table(self._h, tb.t_start, tb.t_end, tb.n_seg, tb.n_coef,
                                           tb.coef.ctypes.data_as(dp_t), C.byref(tid)))
            table_ids[id(tb)] = tid.value
        for lt, a, b, levels in spec.terms:
            ts = _abi.make_term(lt, table_ids)
            pa, ka = _abi.cplx_ptr(a)
            pb, kb = _abi.cplx_ptr(b)
            pl, lv = None, None
            if levels is not None:
                lv = np.ascontiguousarray(levels, dtype=np.int32)
                pl = lv.ctypes.data_as(C.POINTER(C.c_int32))
            if lt.kind == _abi.TERM_SPECTRUM:
                ids = np.array([table_ids[id(tb)] for tb in lt.level_tables], dtype=np.int32)
                check(L.qsim_problem_add_spectrum_term(self._h, C.byref(ts.signal), C.byref(ts.rot), pl, len(ids),
                                                       ids.ctypes.data_as(C.POINTER(C.c_int32))))
                continue
            check(L.qsim_problem_add_term(self._h, C.byref(ts), pa, pb, pl))
        for op, scale in spec.lindblads:
            po, ko = _abi.cplx_ptr(op)
            if scale is None:
                check(L.qsim_problem_add_lindblad(self._h, po, None))
            else:
                ss = _abi.make_signal(scale, table_ids)
                check(L.qsim_problem_add_lindblad(self._h, po, C.byref(ss)))
        check(L.qsim_problem_finalize(self._h))

    def close(self):
        if getattr(self, "_h", None):
            lib().qsim_problem_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def eval(self, kind: int, t: float, param_row=None) -> np.ndarray:
        n = self.d if kind == 0 else self.d * self.d
        out = np.zeros((n, n), dtype=np.complex128, order="F")
        pr, kr = _abi.dbl_ptr(param_row)
        check(lib().qsim_problem_eval(self._h, kind, float(t), pr, out.ctypes.data_as(C.POINTER(C.c_double))))
        return out

    def flops_per_rhs(self, which) -> float:
        v = C.c_double()
        check(lib().qsim_problem_flops_per_rhs(self._h, WHICH.get(which, which), C.byref(v)))
        return v.value

    PLAN_KEYS = ("n", "ncols", "nnz", "max_row", "n_slots", "n_dyn_slots", "n_chan", "team_warps", "teams_per_cta",
                 "block", "smem_bytes", "streamed", "scratch_bytes_per_team", "value_mode", "rowsplit", "components")

    def plan_info(self, which) -> dict:
        """Generator and launch-plan facts of one entry point (qsim_problem_plan_info)."""
        v = (C.c_int64 * 16)()
        check(lib().qsim_problem_plan_info(self._h, WHICH.get(which, which), v))
        return {k: int(x) for k, x in zip(self.PLAN_KEYS, v) if not k.startswith("_")}

    def kernel_name(self, which) -> str:
        buf = C.create_string_buffer(128)
        check(lib().qsim_problem_kernel_name(self._h, WHICH.get(which, which), buf, 128))
        return buf.value.decode()

    def state_shape(self, which: int):
        d = self.d
        return {0: (d, d), 1: (d,), 2: (d * d, d * d), 3: (d, d)}[which]

    def floquet(self, which, ts, t_period, params=None, rise_time: float = 0.0, fall_time: float = 0.0, rtol: float = 0.0,
                opts=None, ctx: Optional["Context"] = None, out=None):
        """Floquet propagation of the whole sweep in one call (qsim_floquet_propagator): the system of trajectory i
        is periodic with t_period[i] (a scalar: one period for all) between ts[0] + rise_time and ts[-1] - fall_time.
        Returns (out[n_traj, n_ts, n, n] indexed [row, col], stats)."""
        which = WHICH.get(which, which)
        if which not in (0, 2):
            raise ValueError("Floquet propagation applies to unitary_propagator and me_propagator")
        ctx = ctx or default_context()
        ts = np.ascontiguousarray(ts, dtype=np.float64)
        if params is None:
            params = np.zeros((1, max(self.n_params, 1)))
        params = np.ascontiguousarray(np.atleast_2d(params), dtype=np.float64)
        n_traj = params.shape[0]
        if self.n_params and params.shape[1] != self.n_params:
            raise ValueError(f"params needs {self.n_params} columns")
        if ts.ndim == 2:
            if ts.shape[0] != n_traj:
                raise ValueError("per-trajectory ts needs one row per trajectory")
            n_ts, ts_stride = ts.shape[1], ts.shape[1]
        else:
            n_ts, ts_stride = ts.shape[0], 0
        per = np.ascontiguousarray(np.atleast_1d(t_period), dtype=np.float64)
        if per.shape not in ((1,), (n_traj,)):
            raise ValueError("t_period must be a scalar or one period per trajectory")
        shape = self.state_shape(which)
        n = int(np.prod(shape))
        if out is None:
            out = np.empty((n_traj, n_ts, n), dtype=np.complex128)
        elif out.shape != (n_traj, n_ts, n) or out.dtype != np.complex128 or not out.flags.c_contiguous:
            raise ValueError("out buffer has the wrong shape / dtype / layout")
        stats = np.zeros(n_traj, dtype=_abi.STATS_DTYPE)
        o = opts if opts is not None else _abi.default_opts()
        check(lib().qsim_floquet_propagator(ctx._h, self._h, which, n_traj, params.ctypes.data, ts.ctypes.data, n_ts, ts_stride,
                                            per.ctypes.data, 0 if per.shape[0] == 1 and n_traj != 1 else (1 if per.shape[0] > 1 else 0),
                                            float(rise_time), float(fall_time), float(rtol), C.byref(o), out.ctypes.data,
                                            stats.ctypes.data))
        return out.reshape(n_traj, n_ts, shape[1], shape[0]).transpose(0, 1, 3, 2), stats

    # -- raw call on caller-provided buffers (host numpy arrays, or device pointers as ints) --
    def solve_raw(self, ctx: Context, which: int, n_traj: int, params_ptr, ts_ptr, n_ts: int, ts_stride: int,
                  init_ptr, init_stride: int, out_ptr, stats_ptr, opts: Optional[_abi.SolverOpts] = None):
        o = opts if opts is not None else _abi.default_opts()
        L = lib()
        if which == 0:
            rc = L.qsim_unitary_propagator(ctx._h, self._h, n_traj, params_ptr, ts_ptr, n_ts, ts_stride, C.byref(o),
                                           out_ptr, stats_ptr)
        elif which == 2:
            rc = L.qsim_me_propagator(ctx._h, self._h, n_traj, params_ptr, ts_ptr, n_ts, ts_stride, C.byref(o),
                                      out_ptr, stats_ptr)
        elif which == 1:
            rc = L.qsim_unitary_state(ctx._h, self._h, n_traj, params_ptr, ts_ptr, n_ts, ts_stride, init_ptr,
                                      init_stride, C.byref(o), out_ptr, stats_ptr)
        else:
            rc = L.qsim_me_state(ctx._h, self._h, n_traj, params_ptr, ts_ptr, n_ts, ts_stride, init_ptr, init_stride,
                                 C.byref(o), out_ptr, stats_ptr)
        check(rc)

    def solve(self, which, ts, params=None, init=None, opts=None, ctx: Optional[Context] = None, out=None,
              select=None, abs2: bool = False):
        """Batched solve through the C ABI with host buffers.  Returns (out, stats) with
        out[n_traj, n_ts, ...] complex, inner matrices indexed [row, col].

        select = [(row, col), ...] (or [row, ...] for unitary_state): only those entries of every saved result
        are produced (qsim_solve_selected); out is then [n_traj, n_ts, len(select)], complex or, with abs2,
        real squared moduli -- the populations the reference's sweeps keep of a propagator.

        ctx may be a list of Contexts (one per device): the sweep is then split into contiguous blocks, one per
        device, inside the library (qsim_solve_multi)."""
        which = WHICH.get(which, which)
        multi = list(ctx) if isinstance(ctx, (list, tuple)) else None
        if multi is not None and select is not None:
            raise ValueError("select= is not available together with a list of contexts")
        ctx = None if multi is not None else (ctx or default_context())
        ts = np.ascontiguousarray(ts, dtype=np.float64)
        if params is None:
            params = np.zeros((1, max(self.n_params, 1)))
        params = np.ascontiguousarray(np.atleast_2d(params), dtype=np.float64)
        n_traj = params.shape[0]
        if self.n_params and params.shape[1] != self.n_params:
            raise ValueError(f"params needs {self.n_params} columns")
        if ts.ndim == 2:
            if ts.shape[0] != n_traj:
                raise ValueError("per-trajectory ts needs one row per trajectory")
            n_ts, ts_stride = ts.shape[1], ts.shape[1]
        else:
            n_ts, ts_stride = ts.shape[0], 0
        shape = self.state_shape(which)
        n = int(np.prod(shape))
        init_ptr, keep_init, init_stride = None, None, 0
        if which in (1, 3):
            if init is None:
                raise ValueError("initial state required")
            init = np.asarray(init, dtype=np.complex128)
            if init.ndim == len(shape) + 1:
                keep_init = np.ascontiguousarray(np.stack([np.asfortranarray(x).ravel(order="F") for x in init]))
                init_stride = n
            else:
                if init.shape != shape:
                    raise ValueError(f"initial state must have shape {shape}")
                keep_init = np.ascontiguousarray(init.ravel(order="F"))
            init_ptr = keep_init.ctypes.data
        stats = np.zeros(n_traj, dtype=_abi.STATS_DTYPE)
        if select is not None:
            sel = np.asarray(select, dtype=np.int32)
            if sel.ndim == 1:
                rows, cols = np.ascontiguousarray(sel), None
            else:
                rows, cols = np.ascontiguousarray(sel[:, 0]), np.ascontiguousarray(sel[:, 1])
            want = (n_traj, n_ts, len(rows))
            dtype = np.float64 if abs2 else np.complex128
            if out is None:
                out = np.empty(want, dtype=dtype)
            elif out.shape != want or out.dtype != dtype or not out.flags.c_contiguous:
                raise ValueError("out buffer has the wrong shape / dtype / layout")
            o = opts if opts is not None else _abi.default_opts()
            check(lib().qsim_solve_selected(ctx._h, self._h, which, n_traj, params.ctypes.data, ts.ctypes.data, n_ts,
                                            ts_stride, init_ptr, init_stride, C.byref(o), len(rows), rows.ctypes.data,
                                            cols.ctypes.data if cols is not None else None, 1 if abs2 else 0,
                                            out.ctypes.data, stats.ctypes.data))
            return out, stats
        if out is None:
            out = np.empty((n_traj, n_ts, n), dtype=np.complex128)
        elif out.shape != (n_traj, n_ts, n) or out.dtype != np.complex128 or not out.flags.c_contiguous:
            raise ValueError("out buffer has the wrong shape / dtype / layout")
        if multi is not None:
            handles = (C.c_void_p * len(multi))(*[c._h.value for c in multi])
            o = opts if opts is not None else _abi.default_opts()
            check(lib().qsim_solve_multi(handles, len(multi), self._h, which, n_traj, params.ctypes.data, ts.ctypes.data,
                                         n_ts, ts_stride, init_ptr, init_stride, C.byref(o), out.ctypes.data,
                                         stats.ctypes.data))
        else:
            self.solve_raw(ctx, which, n_traj, params.ctypes.data, ts.ctypes.data, n_ts, ts_stride, init_ptr,
                           init_stride, out.ctypes.data, stats.ctypes.data, opts)
        res = out
        if len(shape) == 2:
            res = out.reshape(n_traj, n_ts, shape[1], shape[0]).transpose(0, 1, 3, 2)
        return res, stats
