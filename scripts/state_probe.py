"""Development probe (GPU box): unitary_state sweeps (one column per trajectory) through the host-buffer C ABI.
usage: python scripts/state_probe.py [cfg2|cfg3] [points] [t_final]"""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "qsimulator.jl_b200", "host")]
import numpy as np
from qsim_b200 import Context, Problem
from qsim_b200 import workloads as W

name = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
tf = float(sys.argv[3]) if len(sys.argv) > 3 else 20.0
ctx = Context(0)
cqs, params, _, _ = W.config(name)
pr = np.ascontiguousarray(params[np.linspace(0, len(params) - 1, n).astype(int)])
ts = np.arange(0.0, tf + 1.0)
psi0 = np.zeros(cqs.dimension, dtype=complex)
psi0[1] = 1.0
P = Problem(cqs)
P.solve("unitary_state", ts[:2], params=pr[:8], init=psi0, ctx=ctx)
out, st = P.solve("unitary_state", ts, params=pr, init=psi0, ctx=ctx)
print(json.dumps({"cfg": name, "kernel": P.kernel_name("unitary_state"), "n_traj": n, "kernel_ms": ctx.last_kernel_ms(),
                  "steps": int(st["n_accept"].sum() + st["n_reject"].sum()), "norm_err": float(np.abs(np.linalg.norm(out[:, -1], axis=-1) - 1).max())}))
