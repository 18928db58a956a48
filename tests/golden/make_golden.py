"""Regression fixtures: outputs of the CPU oracle (faithful mode) on small versions of the BASELINE.json configs.

These are NOT reference outputs -- the reference is a Julia package and cannot run in this image (DESIGN.md
section 5; baseline/run_reference.jl produces real ones for anyone with Julia).  They freeze what the oracle
computed when the round-1 parity numbers were taken, so that a later change to the oracle, the host lowering or
the libm underneath shows up as a diff, and they give the GPU tests a check that does not depend on the oracle
library being rebuilt identically.

    python tests/golden/make_golden.py        # rewrites tests/golden/oracle_regression.npz
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, "qsimulator.jl_b200", "host")]
import numpy as np

from oracle import c_oracle as co
from qsim_b200 import workloads as W

CASES = {
    # name: (config, sweep rows, save grid)
    "cfg1": ("cfg1", [0], np.array([0.0, 0.5, 1.0, 5.0, 10.0])),
    "cfg2": ("cfg2", [0, 63, 2080, 4032, 4095], np.array([0.0, 1.0, 2.0, 5.0])),
    "cfg3": ("cfg3", [0, 16383], np.array([0.0, 0.5, 2.0])),
    "cfg4": ("cfg4", [0, 8191], np.array([0.0, 0.25, 1.0])),
}
WHICH = {"unitary_propagator": co.UNITARY_PROPAGATOR, "me_propagator": co.ME_PROPAGATOR}


def compute(name):
    cfg, rows, ts = CASES[name]
    cqs, params, _, fn = W.config(cfg)
    pr = np.ascontiguousarray(params[rows])
    out, st = co.OracleProblem(cqs).solve(WHICH[fn], ts, pr, mode=co.FAITHFUL, n_threads=os.cpu_count() or 1)
    return pr, ts, out, st


def main():
    data = {}
    for name in CASES:
        pr, ts, out, st = compute(name)
        data[name + "_params"] = pr
        data[name + "_ts"] = ts
        data[name + "_out"] = out
        data[name + "_n_accept"] = st["n_accept"]
        data[name + "_n_reject"] = st["n_reject"]
    path = os.path.join(HERE, "oracle_regression.npz")
    np.savez_compressed(path, **data)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
