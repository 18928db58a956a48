"""Development probe (GPU box): unitary_state sweeps (one column per trajectory) through the host-buffer C ABI.
usage: python scripts/state_probe.py [cfg2|cfg3|cfg4|cfg5] [points] [t_final]   (cfg4 / cfg5: me_state from |0..1><0..1|)"""
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
d = cqs.dimension
fn = "me_state" if name in ("cfg4", "cfg5") else "unitary_state"
if fn == "me_state":
    psi0 = np.zeros((d, d), dtype=complex)
    psi0[1, 1] = 1.0
    ts = np.array([0.0, tf])
else:
    psi0 = np.zeros(d, dtype=complex)
    psi0[1] = 1.0
P = Problem(cqs)
P.solve(fn, ts[:2], params=pr[:8], init=psi0, ctx=ctx)
out, st = P.solve(fn, ts, params=pr, init=psi0, ctx=ctx)
last = out[:, -1]
inv = (np.abs(np.trace(last.reshape(-1, d, d), axis1=1, axis2=2) - 1).max() if fn == "me_state"
       else np.abs(np.linalg.norm(last, axis=-1) - 1).max())
print(json.dumps({"cfg": name, "entry": fn, "kernel": P.kernel_name(fn), "n_traj": n, "kernel_ms": ctx.last_kernel_ms(),
                  "steps": int(st["n_accept"].sum() + st["n_reject"].sum()), "norm_or_trace_err": float(inv),
                  "pruned": not bool(os.environ.get("QSIM_NO_PRUNE"))}))
