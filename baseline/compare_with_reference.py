"""Compare the oracle (and, with --gpu on a CUDA box, the library) with golden vectors produced by the real
reference through baseline/run_reference.jl.  Prints max |a-b| / max(|b|, 1e-3) per case.

    python baseline/compare_with_reference.py reference_vectors [--gpu]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "qsimulator.jl_b200", "host")]
import numpy as np

from oracle import c_oracle as co
from qsim_b200 import workloads as W

CASES = {"cfg1": ("cfg1", co.UNITARY_PROPAGATOR), "cfg2_short": ("cfg2", co.UNITARY_PROPAGATOR),
         "cfg3_short": ("cfg3", co.UNITARY_PROPAGATOR), "cfg4_short": ("cfg4", co.ME_PROPAGATOR),
         "cfg5_short": ("cfg5", co.ME_PROPAGATOR)}


def load(d, name):
    n_traj, n_params, n_ts, n_rows, n_cols = (int(x) for x in open(os.path.join(d, name + ".meta")).read().split())
    params = np.fromfile(os.path.join(d, name + ".params"), dtype="<f8").reshape(n_traj, n_params)
    ts = np.fromfile(os.path.join(d, name + ".ts"), dtype="<f8")
    out = np.fromfile(os.path.join(d, name + ".out"), dtype="<c16").reshape(n_traj, n_ts, n_cols, n_rows)
    return params, ts, out.transpose(0, 1, 3, 2)      # -> [traj, k, row, col]


def rel(a, b):
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-3)))


def main():
    d = sys.argv[1]
    gpu = "--gpu" in sys.argv[2:]
    for name, (cfg, which) in CASES.items():
        if not os.path.exists(os.path.join(d, name + ".meta")):
            continue
        params, ts, ref = load(d, name)
        cqs = W.config(cfg)[0]
        # (cfg5: the structured mode; the faithful dense Kronecker form takes minutes per step)
        got, st = co.OracleProblem(cqs).solve(which, ts, params, mode=1 if cfg == "cfg5" else 0,
                                              n_threads=os.cpu_count() or 1)
        line = f"{name}: oracle vs reference {rel(got, ref):.3e} (steps {st['n_accept'].tolist()})"
        if gpu:
            from qsim_b200._ffi import Problem
            fn = "unitary_propagator" if which == co.UNITARY_PROPAGATOR else "me_propagator"
            g, _ = Problem(cqs).solve(fn, ts, params=params)
            line += f"; CUDA library vs reference {rel(g, ref):.3e}"
        print(line)


if __name__ == "__main__":
    main()
