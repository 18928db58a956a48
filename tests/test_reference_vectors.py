"""Golden vectors from the REAL reference (QSimulator.jl + OrdinaryDiffEq through baseline/run_reference.jl).

The build image has no Julia, so the directory tests/golden/reference_vectors/ is absent and these tests skip; the
first person with a Julia install runs

    julia --project=/path/to/QSimulator.jl baseline/run_reference.jl tests/golden/reference_vectors

and this file then pins the oracle (CPU) and the CUDA library (-m gpu) to the reference's own output at the
1e-8 level of BASELINE.json's north star -- closing the "parity unpinned" gap of DESIGN.md section 5."""
import importlib.util
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
VEC = os.path.join(ROOT, "tests", "golden", "reference_vectors")
spec = importlib.util.spec_from_file_location("compare_with_reference", os.path.join(ROOT, "baseline", "compare_with_reference.py"))
cmp_mod = importlib.util.module_from_spec(spec)
spec.loader.exec_module(cmp_mod)

present = [n for n in cmp_mod.CASES if os.path.exists(os.path.join(VEC, n + ".meta"))]
needs_vectors = pytest.mark.skipif(not present, reason="no reference vectors (needs a Julia install: baseline/run_reference.jl)")


@needs_vectors
@pytest.mark.parametrize("name", present or ["none"])
def test_oracle_matches_reference(name):
    from oracle import c_oracle as co
    from qsim_b200 import workloads as W
    cfg, which = cmp_mod.CASES[name]
    params, ts, ref = cmp_mod.load(VEC, name)
    got, st = co.OracleProblem(W.config(cfg)[0]).solve(which, ts, params, mode=1 if cfg == "cfg5" else 0, n_threads=8)
    assert np.all(st["retcode"] == 0)
    assert cmp_mod.rel(got, ref) <= 1e-8


@needs_vectors
@pytest.mark.gpu
@pytest.mark.parametrize("name", present or ["none"])
def test_library_matches_reference(name):
    from oracle import c_oracle as co
    from qsim_b200 import workloads as W
    from qsim_b200._ffi import Problem
    cfg, which = cmp_mod.CASES[name]
    params, ts, ref = cmp_mod.load(VEC, name)
    fn = "unitary_propagator" if which == co.UNITARY_PROPAGATOR else "me_propagator"
    got, st = Problem(W.config(cfg)[0]).solve(fn, ts, params=params)
    assert np.all(st["retcode"] == 0)
    assert cmp_mod.rel(got, ref) <= 1e-8
