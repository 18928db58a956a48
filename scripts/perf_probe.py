"""Development probe (GPU box): quick timings of the configs through the host-buffer C ABI.
usage: perf_probe.py cfg:points:t_final[:sel] ...   (":sel": deliver only the qubit-subspace block, as bench.py does for cfg3 / cfg5)"""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "qsimulator.jl_b200", "host")]
import numpy as np
from qsim_b200 import Context, Problem, _abi
from qsim_b200 import workloads as W

ctx = Context(0)
print("device", ctx.info())
print("peak fp64 TFLOP/s", ctx.peak_fp64(200.0))
which = sys.argv[1:] or ["cfg2:256:20", "cfg2:4096:200", "cfg3:256:10", "cfg4:256:2"]
for item in which:
    parts = item.split(":")
    name, n, tf = parts[0], int(parts[1]), float(parts[2])
    use_sel = len(parts) > 3 and parts[3] == "sel"
    cqs, params, ts, fn = W.config(name)
    idx = np.linspace(0, len(params) - 1, n).astype(int)
    pr = np.ascontiguousarray(params[idx])
    ts = np.arange(0.0, tf + 1.0) if fn == "unitary_propagator" else np.array([0.0, tf])
    P = Problem(cqs)
    select = None
    if use_sel:
        sys.path.insert(0, ROOT)
        from bench import qubit_subspace_selection
        select = qubit_subspace_selection([q.dimension for q in cqs.subsystems], fn == "me_propagator")
    P.solve(fn, ts[:2], params=pr[:8], ctx=ctx, select=select)  # warm-up / upload
    t0 = time.perf_counter()
    out, st = P.solve(fn, ts, params=pr, ctx=ctx, select=select)
    dt = time.perf_counter() - t0
    kms = ctx.last_kernel_ms()
    steps = int(st["n_accept"].sum() + st["n_reject"].sum())
    flops = P.flops_per_rhs(fn) * 6.0 * steps
    print(json.dumps({"cfg": name, "n_traj": n, "t_final": tf, "kernel": P.kernel_name(fn), "seconds": dt, "kernel_ms": kms,
                      "traj_per_s": n / dt, "steps": steps, "steps_per_s": steps / dt, "rejects": int(st["n_reject"].sum()),
                      "spmv_tflops": flops / dt / 1e12, "retcodes": np.unique(st["retcode"]).tolist()}))
