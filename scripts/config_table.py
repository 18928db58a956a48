"""Per-config kernel timings (GPU box): every BASELINE.json config through the device-pointer C ABI with CUDA
events, next to the CPU oracle (faithful mode, all host threads) on a small sample of the same points.
Not the benchmark (bench.py measures cfg2 at full size); this is the table in profiles/README.md.

usage: python scripts/config_table.py [cfgN:points:t_final ...] [--cpu-points K | --no-cpu] > gpurun_out/config_table.jsonl
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "qsimulator.jl_b200", "host")]
import numpy as np
import torch

from oracle import c_oracle as co
from qsim_b200 import Context, Problem, _abi
from qsim_b200 import workloads as W

DEFAULT = ["cfg1:1:200", "cfg2:4096:200", "cfg3:1184:50", "cfg4:1184:5", "cfg5:148:1"]
ORACLE_FN = {"unitary_propagator": co.UNITARY_PROPAGATOR, "me_propagator": co.ME_PROPAGATOR}


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    cpu_points = 0
    no_cpu = "--no-cpu" in sys.argv[1:]
    for a in sys.argv[1:]:
        if a.startswith("--cpu-points="):
            cpu_points = int(a.split("=")[1])
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    ctx = Context(0)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    peak = ctx.peak_fp64(200.0)
    cores = os.cpu_count() or 1
    for item in args or DEFAULT:
        name, n, tf = item.split(":")
        n, tf = int(n), float(tf)
        cqs, params, _, fn = W.config(name)
        idx = np.linspace(0, len(params) - 1, n).astype(int)
        pr = np.ascontiguousarray(params[idx])
        ts = np.arange(0.0, tf + 1.0) if fn == "unitary_propagator" else np.array([0.0, tf])
        P = Problem(cqs)
        which = {"unitary_propagator": 0, "me_propagator": 2}[fn]
        nst = int(np.prod(P.state_shape(which)))
        opts = _abi.default_opts(flags=_abi.FLAG_DEVICE_PTRS, stream=stream.cuda_stream)
        d_par = torch.from_numpy(pr).to(dev)
        d_ts = torch.from_numpy(ts).to(dev)
        d_out = torch.empty(n * len(ts) * nst, dtype=torch.complex128, device=dev)
        d_st = torch.zeros(n * 32, dtype=torch.uint8, device=dev)

        def launch():
            P.solve_raw(ctx, which, n, d_par.data_ptr(), d_ts.data_ptr(), len(ts), 0, None, 0, d_out.data_ptr(),
                        d_st.data_ptr(), opts)

        launch()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        launch()
        e1.record(stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        st = np.frombuffer(d_st.cpu().numpy().tobytes(), dtype=_abi.STATS_DTYPE)
        steps = int(st["n_accept"].sum() + st["n_reject"].sum())
        flops = P.flops_per_rhs(which) * (6.0 * steps + 2.0 * n)
        row = {"cfg": name, "entry": fn, "n_traj": n, "t_final_ns": tf, "n_saves": len(ts), "kernel": P.kernel_name(which),
               "kernel_ms": ms, "traj_per_s": n / (ms * 1e-3), "attempted_steps": steps,
               "steps_per_traj": steps / n, "executed_tflops": flops / (ms * 1e-3) / 1e12,
               "frac_of_fp64_peak": flops / (ms * 1e-3) / 1e12 / peak, "retcodes": np.unique(st["retcode"]).tolist()}
        k = 0 if no_cpu else cpu_points or {"cfg1": 1, "cfg2": 2 * cores, "cfg3": cores, "cfg4": cores, "cfg5": 0}[name]
        if k > 0:
            kidx = np.linspace(0, n - 1, min(k, n)).astype(int)
            O = co.OracleProblem(cqs)
            t0 = time.perf_counter()
            want, ost = O.solve(ORACLE_FN[fn], ts, pr[kidx], mode=0, n_threads=min(cores, len(kidx)))
            secs = time.perf_counter() - t0
            got = d_out.view(n, len(ts), nst)[torch.from_numpy(kidx).to(dev)].cpu().numpy()
            got = got.reshape(want.shape[0], want.shape[1], want.shape[3], want.shape[2]).transpose(0, 1, 3, 2)
            err = float(np.max(np.abs(got - want) / np.maximum(np.abs(want), 1e-3)))
            row.update({"cpu_traj_per_s": len(kidx) / secs, "cpu_threads": min(cores, len(kidx)), "cpu_sample": len(kidx),
                        "cpu_seconds": secs, "max_rel_err_vs_oracle": err,
                        "same_step_counts": bool(np.array_equal(ost["n_accept"], st["n_accept"][kidx]))})
        print(json.dumps(row), flush=True)


if __name__ == "__main__":
    main()
