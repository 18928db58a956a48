"""Drive descriptors: the closed family of scalar time dependences.

The reference passes opaque Julia closures `t -> Real` into `dipole_drive`,
`rwa_dipole` and `parametric_drive` (/root/reference/src/operators.jl:220-268).
Closures cannot cross the C ABI, so drives are `Signal` objects here: callable
like the closure (`s(t)`), and lowerable to the `qsim_signal` struct of
include/qsim_b200.h.  Any field may be a `Param(i)`: column i of the
per-trajectory parameter row of a sweep.

  carrier = 1 | cos((2*pi*freq)*t + phase) | sin((2*pi*freq)*t + phase)
  env     = 1 | exp(-0.5*((t-t1)/sigma)**2) | 0.5*(erf((t-t1)/sigma) - erf((t-t2)/sigma))
              | a tabulated closure (piecewise Chebyshev series, `tabulated(fn, t0, t1)`)
  blend NONE: s = offset + amp*env*carrier
  blend REST: s = env*(offset + amp*carrier) + (1-env)*rest

Forms covered: benchmarks.jl:111,138,185;
test_time_evolution.jl:45,107-113,130; integration_tests.jl:40,102,146; doc notebooks.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field, replace
from typing import Callable, Optional, Sequence, Union

import numpy as np

CARRIER_NONE, CARRIER_COS, CARRIER_SIN = 0, 1, 2
ENV_NONE, ENV_GAUSSIAN, ENV_ERF_SQUARED, ENV_TABLE = 0, 1, 2, 3
BLEND_NONE, BLEND_REST = 0, 1


@dataclass(frozen=True)
class Param:
    """Column `index` of the sweep's per-trajectory parameter row."""
    index: int

    def resolve(self, params):
        if params is None:
            raise ValueError(f"Param({self.index}) needs a parameter row")
        return float(params[self.index])


Scalar = Union[float, int, Param]


def resolve(x: Scalar, params=None) -> float:
    return x.resolve(params) if isinstance(x, Param) else float(x)


def max_param_index(*xs) -> int:
    m = -1
    for x in xs:
        if isinstance(x, Param):
            m = max(m, x.index)
        elif isinstance(x, Signal):
            m = max(m, x.max_param_index())
    return m


class Table:
    """A real function of time sampled into a piecewise Chebyshev series (qsim_problem_add_table): n_seg equal
    segments over [t_start, t_end], n_coef coefficients each, end values continued outside.  This is how an
    opaque closure `t -> Real` of the reference (composite_systems.jl:40-50) reaches the device.  `fn` is kept
    so that the host mirror still calls the closure itself, like the reference does."""

    def __init__(self, t_start: float, t_end: float, coef: np.ndarray, fn: Optional[Callable] = None):
        self.t_start, self.t_end = float(t_start), float(t_end)
        self.coef = np.ascontiguousarray(coef, dtype=np.float64)
        if self.coef.ndim != 2 or not (2 <= self.coef.shape[1] <= 16) or not (self.t_end > self.t_start):
            raise ValueError("Table needs coef[n_seg, 2..16] and t_start < t_end")
        self.fn = fn

    @property
    def n_seg(self) -> int:
        return self.coef.shape[0]

    @property
    def n_coef(self) -> int:
        return self.coef.shape[1]

    def eval(self, t):
        """The piecewise polynomial itself (same lookup and Clenshaw recurrence as the device and the oracle)."""
        t = np.asarray(t, dtype=np.float64)
        u = (t - self.t_start) * (self.n_seg / (self.t_end - self.t_start))
        k = np.clip(np.floor(u).astype(np.int64), 0, self.n_seg - 1)
        x = np.clip(2.0 * (u - k) - 1.0, -1.0, 1.0)
        c = self.coef[k]                       # [..., n_coef]
        b1 = np.zeros_like(x)
        b2 = np.zeros_like(x)
        for j in range(self.n_coef - 1, 0, -1):
            b1, b2 = 2.0 * x * b1 + (c[..., j] - b2), b1
        return x * b1 + (c[..., 0] - b2)

    @staticmethod
    def from_function(fn: Callable, t_start: float, t_end: float, n_coef: int = 12, tol: float = 1e-14,
                      max_seg: int = 1 << 16, atol: float = 0.0) -> "Table":
        """Sample `fn` on [t_start, t_end]: the segment count doubles until the series matches `fn` at 2*n_coef
        off-node points per segment to max(tol * max|fn|, atol)."""
        def f(ts):
            try:
                v = np.asarray(fn(ts), dtype=np.float64)
                if v.shape == ts.shape:
                    return v
            except Exception:
                pass
            return np.array([float(fn(float(x))) for x in ts.ravel()]).reshape(ts.shape)

        N = n_coef
        j = np.arange(N)
        nodes = np.cos(np.pi * (j + 0.5) / N)                              # x_j
        dct = (2.0 / N) * np.cos(np.pi * np.outer(np.arange(N), j + 0.5) / N)
        dct[0] *= 0.5
        chk = np.cos(np.pi * (np.arange(2 * N) + 0.25) / (2 * N))
        n_seg = 1
        while True:
            h = (t_end - t_start) / n_seg
            left = t_start + h * np.arange(n_seg)
            coef = f(left[:, None] + 0.5 * h * (1.0 + nodes[None, :])) @ dct.T
            tab = Table(t_start, t_end, coef, fn)
            tc = left[:, None] + 0.5 * h * (1.0 + chk[None, :])
            want = f(tc)
            err = np.max(np.abs(tab.eval(tc) - want))
            scale = max(np.max(np.abs(want)), 1e-300)
            if err <= max(tol * scale, atol) or n_seg >= max_seg:
                if err > max(1e-9 * scale, atol):
                    raise ValueError(f"could not tabulate the function to 1e-9 with {n_seg} segments (error {err / scale:.1e})")
                return tab
            n_seg *= 2


@dataclass(frozen=True)
class Signal:
    carrier: int = CARRIER_NONE
    envelope: int = ENV_NONE
    blend: int = BLEND_NONE
    offset: Scalar = 0.0
    amp: Scalar = 1.0
    freq: Scalar = 0.0
    phase: Scalar = 0.0
    t1: Scalar = 0.0
    t2: Scalar = 0.0
    sigma: Scalar = 1.0
    rest: Scalar = 0.0
    table: Optional[Table] = None      # ENV_TABLE

    FIELDS = ("offset", "amp", "freq", "phase", "t1", "t2", "sigma", "rest")

    def max_param_index(self) -> int:
        return max_param_index(*(getattr(self, f) for f in self.FIELDS))

    def __call__(self, t, params=None) -> float:
        r = lambda x: resolve(x, params)
        if self.carrier == CARRIER_NONE:
            car = 1.0
        else:
            arg = (2 * math.pi * r(self.freq)) * t + r(self.phase)
            car = math.cos(arg) if self.carrier == CARRIER_COS else math.sin(arg)
        if self.envelope == ENV_NONE:
            env = 1.0
        elif self.envelope == ENV_GAUSSIAN:
            x = (t - r(self.t1)) / r(self.sigma)
            env = math.exp(-0.5 * (x * x))
        elif self.envelope == ENV_TABLE:
            # the closure itself when there is one (what the reference would call), else the table
            tb = self.table
            env = float(tb.fn(min(max(t, tb.t_start), tb.t_end))) if tb.fn is not None else float(tb.eval(t))
        else:
            s = r(self.sigma)
            env = 0.5 * (math.erf((t - r(self.t1)) / s) - math.erf((t - r(self.t2)) / s))
        if self.blend == BLEND_NONE:
            return r(self.offset) + r(self.amp) * env * car
        return env * (r(self.offset) + r(self.amp) * car) + (1 - env) * r(self.rest)

    # envelope combinators -------------------------------------------------
    def gaussian(self, center: Scalar, sigma: Scalar) -> "Signal":
        """s -> offset + amp*exp(-0.5((t-center)/sigma)^2)*carrier."""
        return replace(self, envelope=ENV_GAUSSIAN, t1=center, sigma=sigma)

    def erf_squared(self, t1: Scalar, t2: Scalar, sigma: Scalar, rest: Optional[Scalar] = None) -> "Signal":
        """Flat-top pulse 0.5*(erf((t-t1)/sigma) - erf((t-t2)/sigma)); with `rest`
        the blended form of test_time_evolution.jl:107-111."""
        if rest is None:
            return replace(self, envelope=ENV_ERF_SQUARED, t1=t1, t2=t2, sigma=sigma)
        return replace(self, envelope=ENV_ERF_SQUARED, blend=BLEND_REST, t1=t1, t2=t2, sigma=sigma, rest=rest)


def tabulated(fn: Callable, t_start: float, t_end: float, amp: Scalar = 1.0, offset: Scalar = 0.0,
              n_coef: int = 12, tol: float = 1e-14) -> Signal:
    """t -> offset + amp*fn(t) for an arbitrary real closure `fn` on [t_start, t_end] (constant continuation
    outside): the generic path for the reference's opaque drive closures.  Combine with a carrier through
    dataclasses.replace(sig, carrier=..., freq=...) for shaped microwave pulses."""
    return Signal(envelope=ENV_TABLE, amp=amp, offset=offset,
                  table=Table.from_function(fn, t_start, t_end, n_coef=n_coef, tol=tol))


def Const(value: Scalar) -> Signal:
    return Signal(carrier=CARRIER_NONE, offset=0.0, amp=value)


def Cos(amp: Scalar = 1.0, freq: Scalar = 0.0, phase: Scalar = 0.0, offset: Scalar = 0.0) -> Signal:
    """t -> offset + amp*cos(2*pi*freq*t + phase)."""
    return Signal(carrier=CARRIER_COS, amp=amp, freq=freq, phase=phase, offset=offset)


def Sin(amp: Scalar = 1.0, freq: Scalar = 0.0, phase: Scalar = 0.0, offset: Scalar = 0.0) -> Signal:
    """t -> offset + amp*sin(2*pi*freq*t + phase)."""
    return Signal(carrier=CARRIER_SIN, amp=amp, freq=freq, phase=phase, offset=offset)


@dataclass(frozen=True)
class ComplexSignal:
    """A complex drive re(t) + i*im(t) for rwa_dipole (operators.jl:242-250)."""
    re: Signal
    im: Signal = field(default_factory=lambda: Const(0.0))

    def __call__(self, t, params=None) -> complex:
        return complex(self.re(t, params), self.im(t, params))

    def max_param_index(self) -> int:
        return max(self.re.max_param_index(), self.im.max_param_index())
