#!/usr/bin/env python
"""Benchmark of the time-evolution hot path: BASELINE.json's metric, headline on its configs[1], the other
configs at their stated sizes in a `configs` block of the same JSON line.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--configs cfg3,cfg4,cfg5|none]

Headline workload ("step"): one pass of batched `unitary_propagator` over the 4096-point amplitude x frequency
sweep of the 2-transmon lab-frame parametric gate (benchmark/benchmarks.jl:120-155 system, ts = 0:200), i.e.
4096 propagator trajectories, each returning 201 9x9 matrices.  N > 1 (torchrun, one rank per GPU): weak scaling,
the sweep grows to N x 4096 points, sharded block-cyclically by parameter index with no data-path collective.

value  = propagators/s with parameters resident in HBM and results left in HBM (CUDA events on the launch stream)
e2e    = propagators/s through the public host-buffer API: H2D of the parameter table and save grid from pinned
         memory, kernel, results and per-trajectory stats delivered into the caller's page-locked host arrays (the
         kernel writes them in place over PCIe while it integrates: qsim_api.cu, "zero-copy")
roofline: FP64 pipe.  achieved = executed algorithmic flop of the sparse generator application (4 or 8 flop x
         stored entries x columns per RHS, 6 RHS per attempted step + 2, from the solver's own step counts) /
         kernel time; peak = DFMA peak measured in the same run (MEASURED_PEAKS.json has no FP64 entry).
         traffic: DRAM bytes of one launch, measured in this run by an ncu child process (measure_traffic); the
         algorithmic bytes are stated beside it and profiles/traffic.json keeps the committed capture as a cross-check.
cpu_baseline / --impl reference: the C oracle in reference-faithful mode on the host cores (threads over the
         sweep, the analogue of Threads.@threads) -- a C restatement of the reference algorithm, not Julia.

configs block (cfg3, cfg4, cfg5 of BASELINE.json): each config's whole sweep, strong-sharded over the N ranks (the
sweep points dealt round-robin after a fixed pseudo-random permutation, so every shard samples the whole grid), ONE
call through the host-buffer API per rank with page-locked buffers.  e2e = points / wall time of that call (max over ranks), value = points / device time of its kernels
(qsim_last_kernel_ms, max over ranks); roofline as above; max_rel_err = elementwise error against the oracle on two
points over a short span (identical step counts required); cpu = oracle on a stated, time-truncated sample.
cfg3 and cfg5 deliver, of each saved result, the block of the computational (qubit) subspace (qsim_solve_selected: 8 x 8
of the 27 x 27 propagators, 64 x 64 of the 729 x 729 superoperators): the full results of those sweeps are 38 GB and
557 GB.  cfg4 delivers every 81 x 81 superoperator whole.
"""
import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [ROOT, os.path.join(ROOT, "qsimulator.jl_b200", "host")]

import numpy as np  # noqa: E402

METRIC = "propagators/sec (FP64 sweep)"
WORKLOAD = ("cfg2: unitary_propagator, 2 transmons d=9 lab-frame parametric 2Q gate, 64x64 amplitude x frequency "
            "sweep (amp 0.20-0.45 Phi0, f 110-134 MHz), ts=0:200 (201 saves)")
N_AMP, N_FREQ = 64, 64


SMALL = False  # --small: a reduced sweep for profiler captures only (never a bench value)
POINTS = 0     # --points N: development override of the sweep size (never a bench value)


def build_workload(world: int, rank: int):
    from qsim_b200 import workloads as W
    cqs = W.lab_frame_parametric_gate(2)
    if SMALL:
        grid = W.amp_freq_grid(37, 48 * world)   # 1776 points: 12 one-warp teams per SM
        ts = np.arange(0.0, 21.0)
    else:
        grid = W.amp_freq_grid(N_AMP, N_FREQ * world)
        ts = np.arange(0.0, 201.0)
    from qsim_b200 import shard_params
    params, _ = shard_params(grid, world, rank)   # block-cyclic by parameter index
    if POINTS:                                    # development only: occupancy experiments
        params = np.ascontiguousarray(np.resize(params, (POINTS, params.shape[1])))
    return cqs, params, ts


class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        busy = [s for s in sm if s > 0]
        return {"sm_mhz": float(np.median(busy)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


def cpu_reference_run(n_traj: int, threads: int, mode: int = 0):
    """Oracle (C restatement of the reference algorithm) on `n_traj` points of the same sweep; returns
    (seconds, stats)."""
    from oracle import c_oracle as co
    cqs, params, ts = build_workload(1, 0)
    idx = np.linspace(0, len(params) - 1, n_traj).astype(int)
    pr = np.ascontiguousarray(params[idx])
    O = co.OracleProblem(cqs)
    O.solve(co.UNITARY_PROPAGATOR, np.array([0.0, 0.1]), pr[:threads], mode=mode, n_threads=threads)  # warm
    t0 = time.perf_counter()
    _, st = O.solve(co.UNITARY_PROPAGATOR, ts, pr, mode=mode, n_threads=threads)
    return time.perf_counter() - t0, st


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    n_sample = max(cores, min(4096, 12 * cores))   # ~8 s of CPU work per step (0.7 s per trajectory per core)
    for _ in range(args.warmup):
        cpu_reference_run(min(n_sample, cores), cores)
    times = []
    for _ in range(args.steps):
        dt, _ = cpu_reference_run(n_sample, cores)
        times.append(dt)
    per_step = float(np.mean(times))
    value = n_sample / per_step
    sample = f"{n_sample} of the 4096 sweep points (evenly spaced), full ts=0:200, per step"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "propagators/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": per_step * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "note": "C restatement of the reference algorithm (oracle, faithful mode: dense "
                   "per-RHS assembly + dense zgemm, same Tsit5 controller) on host threads -- not Julia (not installed)"},
        "cpu_baseline": {"value": value, "unit": "propagators/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "propagators/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def pruning_info(prob, which, info):
    """Component pruning (csrc/qsim_components.cpp): connected components of the generator graph, the fraction of the
    generator multiply-adds (= of the propagator's entries, to within the blocks' density) that is integrated, and
    the flop one RHS call of the un-pruned sparse form would execute."""
    per = 4.0 if info["value_mode"] >= 1 else 8.0
    full = per * info["nnz"] * (info["n"] if which in (0, 2, "unitary_propagator", "me_propagator") else 1)
    f = prob.flops_per_rhs(which)
    return {"components": info["components"], "integrated_fraction": f / full if full > 0 else 1.0,
            "flops_per_rhs_full_state": full}


def measure_traffic(args):
    """roofline.traffic, measured in this run: DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) of ONE launch of
    the benchmark kernel on the benchmark workload, from a child process of this script under ncu (two metrics, one
    replay pass; the child repeats the device-resident launch of the timed region and ncu profiles its second launch).
    Nothing timed runs under the profiler.  Returns (bytes or None, note)."""
    import shutil
    ncu = shutil.which("ncu") or ("/usr/local/cuda/bin/ncu" if os.path.exists("/usr/local/cuda/bin/ncu") else None)
    if ncu is None:
        return None, "ncu not found on this box"
    cmd = [ncu, "--metrics", "dram__bytes_read.sum,dram__bytes_write.sum", "--clock-control", "none", "--print-units", "base",
           "-k", "regex:solve_kernel", "-s", "1", "-c", "1", "--csv", sys.executable, os.path.abspath(__file__), "--traffic-child"]
    if args.small:
        cmd.append("--small")
    if args.points:
        cmd += ["--points", str(args.points)]
    try:
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    except (OSError, subprocess.TimeoutExpired) as e:
        return None, f"ncu child failed: {type(e).__name__}"
    unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    total, seen = 0.0, 0
    import csv as _csv
    for row in _csv.reader(r.stdout.splitlines()):
        if len(row) >= 3 and row[-3] in ("dram__bytes_read.sum", "dram__bytes_write.sum") and row[-2] in unit:
            try:
                total += float(row[-1].replace(",", "")) * unit[row[-2]]
                seen += 1
            except ValueError:
                pass
    if seen != 2:
        return None, f"ncu child gave no DRAM counters (exit {r.returncode}): {(r.stderr or r.stdout)[-160:]!r}"
    return total, ("dram__bytes_read.sum + dram__bytes_write.sum of one launch of this kernel on this workload, measured by an "
                   "ncu child process of this run (second of two device-resident launches, L2 flushed before it)")


def recorded_traffic(kernel, n_traj, n_ts):
    """Cross-check only: DRAM bytes per launch from the committed `ncu --set full` capture of this kernel on this
    workload (profiles/traffic.json), or None when no capture matches."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            for rec in json.load(f):
                if rec["kernel"] == kernel and rec["points"] == n_traj and rec["n_ts"] == n_ts:
                    return rec["dram_bytes_read"] + rec["dram_bytes_write"], rec["source"]
    except (OSError, ValueError, KeyError):
        pass
    return None, None


# ------------------------------------------------------------------------------------------------------------
# The other BASELINE.json configs, each at its stated size, strong-sharded over the ranks.
# `t_cpu`: span of the time-truncated CPU sample; `t_par`: span of the parity check against the oracle.
EXTRA = {
    "cfg3": dict(label="cfg3: unitary_propagator, 3 transmons d=27 lab-frame parametric gate, 128x128 amplitude x "
                       "frequency sweep, ts=0:200 (201 saves); of each saved propagator the 8x8 block of the qubit "
                       "subspace is delivered (qsim_solve_selected)", t_cpu=2.0, t_par=3.0, cpu_mode=0),
    "cfg4": dict(label="cfg4: me_propagator, 2 transmons d^2=81 with decay and dephasing on both (n_L=4), 128x64 "
                       "T1 x Tphi sweep, ts=[0,20]", t_cpu=0.25, t_par=1.0, cpu_mode=0),
    "cfg5": dict(label="cfg5: me_propagator, 3 transmons d^2=729 rotating frame with decay and dephasing on all three "
                       "(n_L=6), 16^4 amplitude x frequency x T1 x Tphi sweep, ts=[0,10]; of each superoperator the "
                       "64x64 block of the qubit subspace is kept (qsim_solve_selected)", t_cpu=0.05, t_par=0.15,
                 cpu_mode=1),
}


def qubit_subspace_selection(dims, superoperator):
    """(row, col) pairs of the block acting inside the computational subspace (levels 0 and 1 of every subsystem):
    of a propagator U the 2^k x 2^k block, of a superoperator the 4^k x 4^k block over vec indices i + d j of
    |i><j| (column-major vec)."""
    d = int(np.prod(dims))
    comp = []
    for idx in range(d):
        digits, r = [], idx
        for dim in reversed(dims):
            digits.append(r % dim)
            r //= dim
        if all(x < 2 for x in digits):
            comp.append(idx)
    if not superoperator:
        return np.array([(r, c) for c in comp for r in comp], dtype=np.int32)
    vec = [i + d * j for j in comp for i in comp]
    return np.array([(r, c) for c in vec for r in vec], dtype=np.int32)


def run_extra_config(name, ctx, world, rank, dev, dist, peak_tflops, deadline, run_cpu):
    """One BASELINE config at its stated size; returns this config's dict for rank 0 (None elsewhere)."""
    import torch
    from oracle import c_oracle as co
    from qsim_b200 import PinnedArray, Problem, _abi
    from qsim_b200 import workloads as W
    spec = EXTRA[name]
    cqs, params_all, ts, fn = W.config(name)
    which = {"unitary_propagator": 0, "me_propagator": 2}[fn]
    n_total = len(params_all)
    prob = Problem(cqs)
    # cfg3 (201 saves of 27 x 27: 38 GB for the sweep) and cfg5 (729 x 729: 557 GB) deliver the qubit-subspace block
    # of every saved result, which is what a gate sweep keeps of a propagator; cfg4 delivers every superoperator whole
    select = (qubit_subspace_selection([q.dimension for q in cqs.subsystems], name == "cfg5")
              if name in ("cfg3", "cfg5") else None)
    per_rank = n_total // world
    info = prob.plan_info(which)
    # the ranks' shards: sweep points dealt round-robin after a fixed pseudo-random permutation of the sweep order, so
    # that every shard samples the whole grid (dealing the grid order itself gives rank r the same two values of the
    # fastest grid axis, and step counts depend on them: 5 % imbalance on cfg5 at 8 ranks)
    order = np.random.default_rng(20260101).permutation(n_total) if world > 1 else np.arange(n_total)
    shard = params_all[order[rank::world]]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce_max(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0])

    def reduce_min(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
        return float(t[0])

    def reduce_sum(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t[0])

    # warm-up + probe: two points per SM over the full span (builds the device tables, measures the rate)
    n_probe = min(per_rank, 2 * ctx.info()["sm_count"])
    probe = np.ascontiguousarray(shard[:n_probe])
    t0 = time.perf_counter()
    prob.solve(which, ts, params=probe, ctx=ctx, select=select)
    probe_s = reduce_max(time.perf_counter() - t0)
    # every rank runs its whole shard (N = 1: the whole config, cfg5's 65 536 points included); the budget guard halves
    # the shard only when the probe predicts that the call would not fit the time that is left
    n_run = per_rank
    full = True
    budget = max(20.0, deadline - time.perf_counter())
    while n_run > n_probe and probe_s * n_run / n_probe * 0.9 > budget:
        n_run //= 2
        full = False
    params = np.ascontiguousarray(shard[:n_run])
    n_state = len(select) if select is not None else int(np.prod(prob.state_shape(which)))
    pin_out = PinnedArray((n_run, len(ts), n_state), np.complex128)
    pin_par = PinnedArray(params.shape, np.float64)
    pin_par.array[...] = params
    barrier()
    t0 = time.perf_counter()
    _, st = prob.solve(which, ts, params=pin_par.array, ctx=ctx, out=pin_out.array, select=select)
    wall = time.perf_counter() - t0
    kernel_ms = ctx.last_kernel_ms()
    barrier()
    wall_max, kernel_ms_max, kernel_ms_min = reduce_max(wall), reduce_max(kernel_ms), reduce_min(kernel_ms)
    ok = bool(np.all(st["retcode"] == 0))
    attempts = reduce_sum(float(st["n_accept"].sum() + st["n_reject"].sum()))
    n_done = int(reduce_sum(float(n_run)))
    all_ok = reduce_sum(0.0 if ok else 1.0) == 0.0
    flops = prob.flops_per_rhs(which) * (6.0 * attempts + 2.0 * n_done)
    # invariant against truth on a few results: unitarity (cfg3) / trace preservation (cfg4) of the final propagator
    inv = None
    if name == "cfg4":
        d = cqs.dimension
        vi = np.eye(d).ravel(order="F")
        s = pin_out.array[:: max(1, n_run // 8), -1].reshape(-1, d * d, d * d)   # stored column-major: s[k] = S^T
        inv = {"trace_preservation_err": float(max(np.abs(x @ vi - vi).max() for x in s))}
    del pin_out
    if rank != 0:
        return None
    res = {
        "workload": spec["label"], "kernel": prob.kernel_name(which), "points": n_done, "points_in_config": n_total,
        "full_config": bool(full and n_done == n_total),
        "sharding": f"round-robin over {world} rank(s) after a fixed permutation of the sweep order, no collective",
        "value": n_done / (kernel_ms_max * 1e-3), "e2e": n_done / wall_max, "unit": "propagators/s",
        "kernel_s": kernel_ms_max * 1e-3, "kernel_s_fastest_rank": kernel_ms_min * 1e-3, "call_s": wall_max, "all_succeeded": all_ok,
        "attempted_steps": int(attempts), "steps_per_point": attempts / max(n_done, 1),
        "h2d_bytes": int(params.nbytes + ts.nbytes) * world, "d2h_bytes": int(n_done * len(ts) * n_state * 16 + n_done * 32),
        # (achieved: per GPU -- the ranks' flop summed, over the slowest rank's kernel time, divided by the ranks)
        "roofline": {"bound": "fp64", "achieved": flops / (kernel_ms_max * 1e-3) / 1e12 / world, "peak": peak_tflops,
                     "unit": "TFLOP/s", "frac": flops / (kernel_ms_max * 1e-3) / 1e12 / world / peak_tflops,
                     "flops_per_rhs": prob.flops_per_rhs(which), "generator_nnz": info["nnz"],
                     "pruning": pruning_info(prob, which, info),
                     "full_state_equivalent_tflops": pruning_info(prob, which, info)["flops_per_rhs_full_state"]
                                                     * (6.0 * attempts + 2.0 * n_done) / (kernel_ms_max * 1e-3) / 1e12 / world,
                     "scratch_stream_gbs": (info["scratch_bytes_per_team"] * attempts / max(kernel_ms_max * 1e-3, 1e-9) / 1e9
                                            if info["streamed"] else 0.0)},
    }
    if not res["full_config"]:
        res["note"] = "reduced by the time-budget guard: value/e2e are rates per point"
    if inv:
        res.update(inv)
    # parity against the oracle on two points, short span (elementwise, identical step counts)
    idx = [0, n_total - 1]
    tsp = np.array([0.0, spec["t_par"] / 2, spec["t_par"]])
    got, gs = prob.solve(which, tsp, params=params_all[idx], ctx=ctx)
    want, ws = co.OracleProblem(cqs).solve(which, tsp, params=params_all[idx], mode=spec["cpu_mode"] or 1, n_threads=2)
    res["max_rel_err"] = float(np.max(np.abs(got - want) / np.maximum(np.abs(want), 1e-3)))
    res["same_step_counts"] = bool(np.array_equal(gs["n_accept"], ws["n_accept"]) and
                                   np.array_equal(gs["n_reject"], ws["n_reject"]))
    res["parity_sample"] = f"points 0 and {n_total - 1}, span {spec['t_par']} ns, {int(gs['n_accept'].min())}+ steps"
    if run_cpu:
        cores = os.cpu_count() or 1
        cidx = np.linspace(0, n_total - 1, cores).astype(int)
        O = co.OracleProblem(cqs)
        t0 = time.perf_counter()
        O.solve(which, np.array([0.0, spec["t_cpu"]]), params=params_all[cidx], mode=spec["cpu_mode"], n_threads=cores)
        secs = time.perf_counter() - t0
        t_final = float(ts[-1])
        res["cpu"] = {"value": cores / (secs * t_final / spec["t_cpu"]), "unit": "propagators/s", "cores": cores,
                      "kind": "port", "sample": f"{cores} points over the first {spec['t_cpu']} ns ({secs:.1f} s), "
                      f"scaled by {t_final / spec['t_cpu']:.0f} to the {t_final:.0f} ns span (steps are uniform in time); "
                      + ("reference-faithful mode (dense assembly, dense GEMM / Kronecker superoperator)" if spec["cpu_mode"] == 0
                         else "structured mode (commutator / sandwich form per column): the faithful dense 729^3 "
                              "Kronecker GEMMs take minutes per step")}
    return res


def run_fanout(world, n_ts, d, cqs, ts):
    """Single-process fan-out (qsim_solve_multi), the path a Julia caller uses: rank 0 drives all N devices from one
    call with pageable host buffers (the other ranks idle at the barrier).  Returns propagators/s."""
    from qsim_b200 import Context, Problem
    from qsim_b200 import workloads as W
    ctxs = [Context(i) for i in range(world)]
    prob = Problem(cqs)
    params = W.amp_freq_grid(N_AMP, N_FREQ * world)
    out = np.empty((len(params), n_ts, d * d), dtype=np.complex128)
    out[...] = 0          # touch the pages: the measurement is the call, not the first-touch faults
    t0 = time.perf_counter()
    prob.solve(0, ts, params=params, ctx=ctxs, out=out)       # first call: device result buffers and staging are allocated
    first = time.perf_counter() - t0
    t0 = time.perf_counter()
    _, st = prob.solve(0, ts, params=params, ctx=ctxs, out=out)
    secs = time.perf_counter() - t0
    assert np.all(st["retcode"] == 0)
    for c in ctxs:
        c.close()
    return len(params) / secs, secs, first


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--configs", default="cfg3,cfg4,cfg5", help="other BASELINE configs to run (comma list, or none)")
    ap.add_argument("--budget-s", type=float, default=400.0, help="time budget for the --configs block")
    ap.add_argument("--small", action="store_true", help="reduced sweep (1776 points, 20 ns) for ncu captures")
    ap.add_argument("--points", type=int, default=0, help="development: override the number of sweep points")
    ap.add_argument("--no-traffic", action="store_true", help="skip the ncu child process that measures roofline.traffic")
    ap.add_argument("--traffic-child", action="store_true", help=argparse.SUPPRESS)
    args = ap.parse_args()
    global SMALL, POINTS
    SMALL = args.small
    POINTS = args.points
    if args.impl == "reference":
        run_reference(args)
        return

    import torch
    import torch.distributed as dist
    from qsim_b200 import Context, PinnedArray, Problem, _abi

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the hot path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    cqs, params, ts = build_workload(world, rank)
    n_traj, n_ts, d = params.shape[0], ts.shape[0], cqs.dimension
    ctx = Context(local_rank)
    prob = Problem(cqs)
    which = 0
    stream = torch.cuda.Stream(device=dev)   # a real (non-NULL) stream: the C ABI treats NULL as "the ctx stream"
    torch.cuda.set_stream(stream)
    opts = _abi.default_opts(flags=_abi.FLAG_DEVICE_PTRS, stream=stream.cuda_stream)

    d_params = torch.from_numpy(params).to(dev)
    d_ts = torch.from_numpy(ts).to(dev)
    d_out = torch.empty(n_traj * n_ts * d * d, dtype=torch.complex128, device=dev)
    d_stats = torch.zeros(n_traj * 32, dtype=torch.uint8, device=dev)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def launch():
        prob.solve_raw(ctx, which, n_traj, d_params.data_ptr(), d_ts.data_ptr(), n_ts, 0, None, 0,
                       d_out.data_ptr(), d_stats.data_ptr(), opts)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    if args.traffic_child:
        # child of measure_traffic(): the same device-resident launch twice (L2 flushed before each); ncu profiles the second
        for _ in range(2):
            flush.zero_()
            launch()
        torch.cuda.synchronize()
        return

    peak_tflops = ctx.peak_fp64(300.0)
    for _ in range(max(args.warmup, 3)):
        flush.zero_()
        launch()
    torch.cuda.synchronize()

    l0 = ctx.launch_count()
    sampler = ClockSampler(local_rank)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    for i in range(args.steps):
        flush.zero_()
        ev[i][0].record(stream)
        launch()
        ev[i][1].record(stream)
    barrier()
    clocks = sampler.stop()
    launches = ctx.launch_count() - l0
    kernel_ms = [a.elapsed_time(b) for a, b in ev]
    total_ms = float(sum(kernel_ms))

    stats = np.frombuffer(d_stats.cpu().numpy().tobytes(), dtype=_abi.STATS_DTYPE)
    assert np.all(stats["retcode"] == 0), "solver failures in the benchmark sweep"
    attempts = int(stats["n_accept"].sum() + stats["n_reject"].sum())
    flops_per_launch = prob.flops_per_rhs(which) * (6.0 * attempts + 2.0 * n_traj)
    dense_flops_per_launch = 8.0 * d ** 3 * (6.0 * attempts + 3.0 * n_traj)
    # invariant reported against truth: unitarity of the final propagators (a few trajectories)
    u_last = d_out.view(n_traj, n_ts, d, d)[:: max(1, n_traj // 16), -1].cpu().numpy().transpose(0, 2, 1)
    unit_err = float(max(np.abs(u.conj().T @ u - np.eye(d)).max() for u in u_last))
    del d_out

    # ---- end to end through the public host-buffer API ----
    e2e = None
    if not args.no_e2e:
        pin_out = PinnedArray((n_traj, n_ts, d * d), np.complex128)
        pin_par = PinnedArray(params.shape, np.float64)
        pin_par.array[...] = params
        h_opts = _abi.default_opts()
        prob.solve(which, ts, params=pin_par.array, opts=h_opts, ctx=ctx, out=pin_out.array)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            prob.solve(which, ts, params=pin_par.array, opts=h_opts, ctx=ctx, out=pin_out.array)
        barrier()
        e2e_s = time.perf_counter() - t0
        e2e = {"seconds": e2e_s, "h2d": int(params.nbytes + ts.nbytes), "d2h": int(pin_out.array.nbytes + n_traj * 32)}
        launches += args.steps
        # the results delivered into the caller's buffer are the device run's (same kernel, same inputs)
        e2e_ok = bool(np.isfinite(pin_out.array[-1, -1]).all() and abs(abs(pin_out.array[-1, 0, 0]) - 1.0) < 1e-12)
        del pin_out

    # max over ranks
    t_dev = torch.tensor([total_ms, (e2e or {}).get("seconds", 0.0)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_dev, op=dist.ReduceOp.MAX)
    total_ms_max, e2e_s_max = float(t_dev[0]), float(t_dev[1])

    cpu_base = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        n_sample = max(cores, min(4096, 24 * cores))   # ~15-20 s of CPU work
        secs, _ = cpu_reference_run(n_sample, cores)
        cpu_base = {"value": n_sample / secs, "unit": "propagators/s", "cores": cores, "kind": "port",
                    "sample": f"{n_sample} of the 4096 sweep points (evenly spaced), full ts=0:200, {secs:.1f} s; "
                              "C restatement of the reference algorithm (dense per-RHS assembly + zgemm), not Julia"}

    traffic, traffic_note = None, "not measured (multi-rank run or --no-traffic)"
    if rank == 0 and world == 1 and not args.no_traffic:
        traffic, traffic_note = measure_traffic(args)

    # ---- the other BASELINE configs (strong-sharded over the ranks) ----
    extra = {}
    names = [] if (args.configs in ("", "none") or SMALL or POINTS) else [c for c in args.configs.split(",") if c in EXTRA]
    deadline = time.perf_counter() + args.budget_s
    for i, nm in enumerate(names):
        # cfg5 is last and takes what is left of the budget; cfg3 / cfg4 are sized to their full sweeps
        try:
            r = run_extra_config(nm, ctx, world, rank, dev, dist, peak_tflops, deadline,
                                 run_cpu=(world == 1 and not args.no_cpu_baseline))
        except Exception as e:   # never lose the headline line to a secondary config
            r = {"error": f"{type(e).__name__}: {e}"} if rank == 0 else None
            if world > 1:
                raise
        if rank == 0:
            extra[nm] = r

    # ---- single-process fan-out over all N devices (qsim_solve_multi), rank 0 only ----
    fanout = None
    if world > 1 and not (SMALL or POINTS):
        # rank 0 drives all devices from one process; the other ranks must leave their GPUs and host cores alone
        # meanwhile: they wait in a gloo (socket) barrier, not in an NCCL one (a spinning kernel on every GPU) nor in
        # a busy-waiting synchronize
        barrier()
        idle = dist.new_group(backend="gloo")
        if rank == 0:
            try:
                v, secs, first = run_fanout(world, n_ts, d, cqs, ts)
                fanout = {"value": v, "unit": "propagators/s", "call_s": secs, "first_call_s": first, "contexts": world,
                          "what": "one qsim_solve_multi call from one process over all devices, block-cyclic shards, "
                                  "pageable caller buffers (pipelined device-to-host copies), N x 4096 cfg2 points"}
            except Exception as e:
                fanout = {"error": f"{type(e).__name__}: {e}"}
        dist.barrier(group=idle)
        barrier()

    if rank == 0:
        ms_per_step = total_ms_max / args.steps
        value = n_traj * world / (ms_per_step * 1e-3)
        achieved = flops_per_launch / (total_ms / args.steps * 1e-3) / 1e12
        xcheck, xsrc = recorded_traffic(prob.kernel_name(which), n_traj, len(ts))
        prune = pruning_info(prob, which, prob.plan_info(which))
        out_bytes = n_traj * n_ts * d * d * 16
        line = {
            "metric": METRIC, "value": value, "unit": "propagators/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD if not (SMALL or POINTS) else "DEVELOPMENT SUBSET (not a bench value) of " + WORKLOAD,
                       "points_per_gpu": n_traj, "sharding": "block-cyclic by parameter index, "
                       "no collective", "kernel": prob.kernel_name(which), "solver": "Tsit5 reltol=1e-6 abstol=1e-8 "
                       "(reference literals)", "l2": "256 MB flush write before every timed launch; each launch also "
                       "streams 1.07 GB of results", "attempted_steps_per_launch": attempts,
                       "unitarity_err_final": unit_err},
            "roofline": {"bound": "fp64", "achieved": achieved, "peak": peak_tflops, "unit": "TFLOP/s",
                         "frac": achieved / peak_tflops if peak_tflops > 0 else None,
                         "traffic": traffic, "traffic_unit": "bytes per launch", "traffic_note": traffic_note,
                         "algorithmic_bytes": int(out_bytes + params.nbytes + n_traj * 32),
                         "traffic_crosscheck": xcheck, "traffic_crosscheck_source": xsrc,
                         "peak_source": "DFMA micro-benchmark in this run (MEASURED_PEAKS.json has no FP64 figure; "
                                        "nominal 37.2)",
                         "flop_counted": "executed sparse generator multiply-adds only (%d flop each: the generator is "
                                         "%s), on the component blocks of the propagator only (the entries between "
                                         "components are exactly zero and are not integrated); RK stage combinations, "
                                         "error norm and drive evaluation not counted"
                                         % ((4, "purely imaginary, 2 DFMA per entry") if "_vm0" not in prob.kernel_name(which)
                                            else (8, "complex, 4 DFMA per entry")),
                         "pruning": prune,
                         # the same launch time against the sparse generator applied to the full state (what the kernels
                         # executed before component pruning) and against the reference's dense products
                         "full_state_equivalent_tflops": prune["flops_per_rhs_full_state"] * (6.0 * attempts + 2.0 * n_traj)
                                                         / (total_ms / args.steps * 1e-3) / 1e12,
                         "dense_equivalent_tflops": dense_flops_per_launch / (total_ms / args.steps * 1e-3) / 1e12,
                         # informational: + the Runge-Kutta vector arithmetic any Tsit5 must do per attempted step and
                         # complex state entry (21 stage-combination and 7 error-estimate real-by-complex multiply-adds
                         # at 4 flop, 6 stage arguments u + dt*(...) at 4 flop, ~10 flop of scaled error norm = 146 flop)
                         "tsit5_total_tflops": (flops_per_launch + 146.0 * d * d * prune["integrated_fraction"] * attempts)
                                               / (total_ms / args.steps * 1e-3) / 1e12,
                         "hbm_out_gbs": out_bytes / (total_ms / args.steps * 1e-3) / 1e9},
            "gpu_launches": int(launches), "clocks": clocks,
        }
        if e2e is not None:
            line["e2e"] = {"value": n_traj * world * args.steps / e2e_s_max, "unit": "propagators/s",
                           "h2d_bytes_per_step": e2e["h2d"], "d2h_bytes_per_step": e2e["d2h"],
                           "delivery": "results written into the caller's page-locked host buffer by the kernel (no "
                                       "device result buffer, no copy after the kernel)", "results_checked": e2e_ok}
        if cpu_base is not None:
            line["cpu_baseline"] = cpu_base
        if extra:
            line["configs"] = extra
        if fanout is not None:
            line["single_process_fanout"] = fanout
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
