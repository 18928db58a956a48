"""Drive descriptors: the closed family of scalar time dependences.

The reference passes opaque Julia closures `t -> Real` into `dipole_drive`,
`rwa_dipole` and `parametric_drive` (/root/reference/src/operators.jl:220-268).
Closures cannot cross the C ABI, so drives are `Signal` objects here: callable
like the closure (`s(t)`), and lowerable to the `qsim_signal` struct of
include/qsim_b200.h.  Any field may be a `Param(i)`: column i of the
per-trajectory parameter row of a sweep.

  carrier = 1 | cos((2*pi*freq)*t + phase) | sin((2*pi*freq)*t + phase)
  env     = 1 | exp(-0.5*((t-t1)/sigma)**2) | 0.5*(erf((t-t1)/sigma) - erf((t-t2)/sigma))
  blend NONE: s = offset + amp*env*carrier
  blend REST: s = env*(offset + amp*carrier) + (1-env)*rest

Forms covered: benchmarks.jl:111,138,185;
test_time_evolution.jl:45,107-113,130; integration_tests.jl:40,102,146; doc notebooks.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field, replace
from typing import Optional, Sequence, Union

CARRIER_NONE, CARRIER_COS, CARRIER_SIN = 0, 1, 2
ENV_NONE, ENV_GAUSSIAN, ENV_ERF_SQUARED = 0, 1, 2
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
