"""Committed regression fixtures (tests/golden/oracle_regression.npz, written by tests/golden/make_golden.py):
oracle outputs frozen at round 1.  CPU: the oracle still reproduces them (step counts exactly, values to 1e-12);
-m gpu: the CUDA library matches the frozen values to 1e-8 with the same step counts."""
import importlib.util
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
spec = importlib.util.spec_from_file_location("make_golden", os.path.join(HERE, "golden", "make_golden.py"))
mg = importlib.util.module_from_spec(spec)
spec.loader.exec_module(mg)
GOLD = np.load(os.path.join(HERE, "golden", "oracle_regression.npz"))


@pytest.mark.parametrize("name", list(mg.CASES))
def test_oracle_reproduces_golden(name):
    pr, ts, out, st = mg.compute(name)
    assert np.array_equal(pr, GOLD[name + "_params"]) and np.array_equal(ts, GOLD[name + "_ts"])
    assert np.array_equal(st["n_accept"], GOLD[name + "_n_accept"])
    assert np.array_equal(st["n_reject"], GOLD[name + "_n_reject"])
    assert np.max(np.abs(out - GOLD[name + "_out"])) <= 1e-12


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(mg.CASES))
def test_gpu_matches_golden(name):
    from qsim_b200 import workloads as W
    from qsim_b200._ffi import Problem
    cfg = mg.CASES[name][0]
    cqs, _, _, fn = W.config(cfg)
    got, st = Problem(cqs).solve(fn, GOLD[name + "_ts"], params=GOLD[name + "_params"])
    want = GOLD[name + "_out"]
    assert np.array_equal(st["n_accept"], GOLD[name + "_n_accept"])
    assert np.array_equal(st["n_reject"], GOLD[name + "_n_reject"])
    assert np.max(np.abs(got - want) / np.maximum(np.abs(want), 1e-3)) <= 1e-8
