"""CPU-side checks of the product library: it loads, exports every symbol include/qsim_b200.h
declares, compiles problems on the host, and refuses to solve without a CUDA device (no fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

from oracle import c_oracle as co
from qsim_b200 import (CompositeQSystem, Cos, DuffingSpec, DuffingTransmon, Param, PerturbativeTransmon, Problem,
                       QsimError, RotatingFrameSystem, Sin, TransmonSpec, _abi, _ffi, add_hamiltonian, add_lindblad,
                       decay, dephasing, dipole_drive, parametric_drive, rwa_dipole, xy_exchange_drive)
from qsim_b200 import workloads as W

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "qsim_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(qsim_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    L = ctypes.CDLL(_ffi.LIB_PATH)
    declared = _declared_symbols()
    assert len(declared) >= 24
    for name in declared:
        assert hasattr(L, name), f"{name} declared in include/qsim_b200.h but not exported"
    assert set(_ffi.EXPORTS) == set(declared)
    assert _ffi.lib().qsim_version() == _abi.ABI_VERSION


def test_default_opts_match_reference_literals():
    o = _abi.SolverOpts()
    _ffi.lib().qsim_default_opts(ctypes.byref(o))
    ref = _abi.default_opts()
    for f, _ in _abi.SolverOpts._fields_:
        if f != "stream":
            assert getattr(o, f) == getattr(ref, f), f
    assert (o.reltol, o.abstol) == (1e-6, 1e-8)        # time_evolution.jl:43


def _kitchen_sink():
    q0 = DuffingTransmon("q0", 3, DuffingSpec(3.94015, -0.1807))
    q1 = RotatingFrameSystem.wrap(PerturbativeTransmon("q1", 3, TransmonSpec(0.172, 12.71, 3.69)), 3.9)
    q2 = DuffingTransmon("q2", 2, DuffingSpec(4.2, -0.22))
    cqs = CompositeQSystem([q0, q1, q2])
    add_hamiltonian(cqs, q0)
    add_hamiltonian(cqs, parametric_drive(q1, Sin(amp=Param(0), freq=0.12, offset=0.05)), q1)
    add_hamiltonian(cqs, parametric_drive(q0, Cos(amp=0.3, freq=0.3, offset=4.3).erf_squared(2.0, 9.0, 0.7, rest=4.6)), q0)
    add_hamiltonian(cqs, dipole_drive(q2, Cos(amp=0.02, freq=4.2).gaussian(5.0, 2.0), 4.0), q2)
    add_hamiltonian(cqs, rwa_dipole(q2, Cos(amp=0.01, freq=0.01)), q2)
    add_hamiltonian(cqs, xy_exchange_drive([q2, q0], Param(1), [0.2, 0.0]), [q2, q0])
    add_lindblad(cqs, decay(q0, 1e-3), [q0])
    add_lindblad(cqs, dephasing(q2, 2e-3), [q2])
    return cqs


@pytest.mark.parametrize("builder", ["kitchen", "cfg1", "cfg3", "cfg4"])
def test_compiled_generator_matches_oracle_assembly(builder):
    """qsim_problem_eval (sparse slots/channels, host) vs the oracle's dense per-RHS assembly
    (composite_systems.jl:46-50, time_evolution.jl:104-109)."""
    if builder == "kitchen":
        cqs, row = _kitchen_sink(), np.array([0.31, 0.004])
    else:
        cqs, params, _, _ = W.config(builder)
        row = params[len(params) // 3]
    P, O = Problem(cqs), co.OracleProblem(cqs)
    for t in (0.0, 0.37, 5.55, 171.3):
        h, ho = P.eval(0, t, row), O.eval(0, t, row)
        assert np.abs(h - ho).max() < 1e-13 * max(1.0, np.abs(ho).max())
        if cqs.dimension <= 18:
            g, go = P.eval(1, t, row), O.eval(1, t, row)
            assert np.abs(g - go).max() < 1e-12 * max(1.0, np.abs(go).max())


def test_kernel_dispatch_names_and_flops():
    cqs, _, _, _ = W.config("cfg2")
    P = Problem(cqs)
    # one-warp resident teams run the lean kernels; the propagator is integrated on its component form (two
    # columns per lane, 128 registers, 16 warps per SM)
    assert P.kernel_name("unitary_propagator").startswith("tsit5_lean_r1c2")
    assert "resident" in P.kernel_name("unitary_propagator")
    assert "streamed" in P.kernel_name("me_propagator")
    # executed flop never exceed the dense count 8 d^3
    assert 0 < P.flops_per_rhs("unitary_propagator") <= 8 * 9 ** 3
    assert 0 < P.flops_per_rhs("unitary_state") <= 8 * 9 ** 2
    # value modes: real Hamiltonian with static couplings -> 2 (imaginary generator, static off-diagonals in
    # registers); Lindblad generator of real H and real collapse operators -> 4 (off-diagonal entries purely real
    # or purely imaginary, complex diagonal); rotating-frame exchange terms are complex -> 0
    assert "_vm2_" in P.kernel_name("unitary_propagator")
    c4 = Problem(W.config("cfg4")[0])
    assert "_vm4_" in c4.kernel_name("me_propagator") and "_vm4_" in c4.kernel_name("me_state")
    assert "_vm0_" in Problem(W.config("cfg5")[0]).kernel_name("me_propagator")


def test_argument_errors():
    cqs, _, _, _ = W.config("cfg2")
    spec = Problem(cqs).spec
    with pytest.raises(ValueError):
        from qsim_b200.composite import lower
        lower(cqs, n_params=1)          # drive uses Param(1)
    L = _ffi.lib()
    h = ctypes.c_void_p()
    assert L.qsim_problem_create(0, 0, ctypes.byref(h)) == _abi.ERR_ARG
    assert b"dimension" in L.qsim_last_error()


def test_no_cpu_fallback():
    """Without a CUDA device the product refuses to run; it never routes through the oracle."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA device present")
    h = ctypes.c_void_p()
    rc = _ffi.lib().qsim_create(0, ctypes.byref(h))
    assert rc == _abi.ERR_NO_DEVICE
    assert b"no CPU fallback" in _ffi.lib().qsim_last_error()
    from qsim_b200 import unitary_propagator
    cqs, params, ts, _ = W.config("cfg1")
    with pytest.raises(QsimError):
        unitary_propagator(cqs, ts, params=params)
    src = "".join(open(os.path.join(ROOT, "qsimulator.jl_b200", "host", "qsim_b200", f)).read()
                  for f in os.listdir(os.path.join(ROOT, "qsimulator.jl_b200", "host", "qsim_b200")) if f.endswith(".py"))
    assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M)
    csrc = os.path.join(ROOT, "qsimulator.jl_b200", "csrc")
    for f in os.listdir(csrc):
        if f.endswith((".cu", ".cuh", ".cpp", ".h")):
            assert "oracle/" not in open(os.path.join(csrc, f)).read()
