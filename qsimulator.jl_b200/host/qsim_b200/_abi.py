"""ctypes mirror of include/qsim_b200.h (struct layouts and enum values)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from .operators import LoweredTerm
from .signals import Param, Signal

ABI_VERSION = 1

OK, ERR_ARG, ERR_NO_DEVICE, ERR_CUDA, ERR_UNSUPPORTED, ERR_STATE, ERR_ALLOC = 0, -1, -2, -3, -4, -5, -6
RET_SUCCESS, RET_MAXITERS, RET_DT_UNDERFLOW, RET_NONFINITE = 0, 1, 2, 3
FLAG_DEVICE_PTRS = 1
FLAG_DIRECT_DRIVES = 2
TERM_SPECTRUM = 4


class Slot(C.Structure):
    _fields_ = [("param", C.c_int32), ("_pad", C.c_int32), ("value", C.c_double)]


class SignalStruct(C.Structure):
    _fields_ = [("carrier", C.c_int32), ("envelope", C.c_int32), ("blend", C.c_int32), ("table", C.c_int32),
                ("offset", Slot), ("amp", Slot), ("freq", Slot), ("phase", Slot),
                ("t1", Slot), ("t2", Slot), ("sigma", Slot), ("rest", Slot)]


class TermStruct(C.Structure):
    _fields_ = [("kind", C.c_int32), ("num_terms", C.c_int32), ("signal", SignalStruct),
                ("rate", Slot), ("anharm", Slot), ("rot", Slot), ("EC", Slot), ("EJ1", Slot), ("EJ2", Slot)]


class SolverOpts(C.Structure):
    _fields_ = [("reltol", C.c_double), ("abstol", C.c_double), ("gamma", C.c_double), ("qmin", C.c_double),
                ("qmax", C.c_double), ("beta1", C.c_double), ("beta2", C.c_double), ("qoldinit", C.c_double),
                ("dtmax", C.c_double), ("maxiters", C.c_int64), ("flags", C.c_int32), ("_pad", C.c_int32),
                ("stream", C.c_void_p)]


class TrajStats(C.Structure):
    _fields_ = [("retcode", C.c_int32), ("n_accept", C.c_int32), ("n_reject", C.c_int32), ("n_rhs", C.c_int32),
                ("dt_init", C.c_double), ("t_final", C.c_double)]


STATS_DTYPE = np.dtype([("retcode", "<i4"), ("n_accept", "<i4"), ("n_reject", "<i4"), ("n_rhs", "<i4"),
                        ("dt_init", "<f8"), ("t_final", "<f8")])
assert STATS_DTYPE.itemsize == C.sizeof(TrajStats) == 32


def default_opts(**overrides) -> SolverOpts:
    """The literals the reference hard-codes (time_evolution.jl:43) plus OrdinaryDiffEq's Tsit5 defaults."""
    o = SolverOpts(reltol=1e-6, abstol=1e-8, gamma=0.9, qmin=0.2, qmax=10.0, beta1=7 / 50, beta2=2 / 25,
                   qoldinit=1e-4, dtmax=0.0, maxiters=1_000_000, flags=0, stream=None)
    for k, v in overrides.items():
        if not hasattr(o, k):
            raise TypeError(f"unknown solver option {k!r}")
        setattr(o, k, v)
    return o


def make_slot(x) -> Slot:
    if isinstance(x, Param):
        return Slot(param=int(x.index), value=0.0)
    return Slot(param=-1, value=float(x))


def make_signal(s: Signal, table_ids=None) -> SignalStruct:
    """table_ids: {id(Table) -> 1-based id in the problem being built} for tabulated envelopes."""
    table = 0
    if s.table is not None:
        if table_ids is None or id(s.table) not in table_ids:
            raise ValueError("tabulated envelope was not registered with the problem")
        table = table_ids[id(s.table)]
    return SignalStruct(carrier=s.carrier, envelope=s.envelope, blend=s.blend, table=table,
                        offset=make_slot(s.offset), amp=make_slot(s.amp), freq=make_slot(s.freq),
                        phase=make_slot(s.phase), t1=make_slot(s.t1), t2=make_slot(s.t2),
                        sigma=make_slot(s.sigma), rest=make_slot(s.rest))


def make_term(t: LoweredTerm, table_ids=None) -> TermStruct:
    return TermStruct(kind=t.kind, num_terms=int(t.num_terms), signal=make_signal(t.signal, table_ids),
                      rate=make_slot(t.rate), anharm=make_slot(t.anharm), rot=make_slot(t.rot),
                      EC=make_slot(t.EC), EJ1=make_slot(t.EJ1), EJ2=make_slot(t.EJ2))


def cplx_ptr(a):
    """Pointer to a column-major complex128 array as double*; returns (pointer, keepalive)."""
    if a is None:
        return None, None
    arr = np.asfortranarray(np.asarray(a, dtype=np.complex128))
    return arr.ctypes.data_as(C.POINTER(C.c_double)), arr


def dbl_ptr(a):
    if a is None:
        return None, None
    arr = np.ascontiguousarray(np.asarray(a, dtype=np.float64))
    return arr.ctypes.data_as(C.POINTER(C.c_double)), arr
