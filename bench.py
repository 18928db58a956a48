#!/usr/bin/env python
"""Benchmark of the time-evolution hot path: BASELINE.json's metric on its configs[1].

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload ("step"): one pass of batched `unitary_propagator` over the 4096-point amplitude x frequency
sweep of the 2-transmon lab-frame parametric gate (benchmark/benchmarks.jl:120-155 system, ts = 0:200),
i.e. 4096 propagator trajectories, each returning 201 9x9 matrices.  N > 1 (torchrun, one rank per GPU):
weak scaling, the sweep grows to N x 4096 points, sharded block-cyclically by parameter index with no
data-path collective.

value  = propagators/s with parameters resident in HBM and results left in HBM (CUDA events on the launch stream)
e2e    = propagators/s through the public host-buffer API: H2D of the parameter table and save grid from
         pinned memory, kernel, D2H of every propagator and the per-trajectory stats
roofline: FP64 pipe.  achieved = executed algorithmic flop of the sparse generator application
         (8 flop x stored entries x columns per RHS, 6 RHS per attempted step + 2, from the solver's own step
         counts) / kernel time; peak = DFMA peak measured in the same run (MEASURED_PEAKS.json has no FP64 entry).
cpu_baseline / --impl reference: the C oracle in reference-faithful mode on the host cores (threads over the
         sweep, the analogue of Threads.@threads) -- a C restatement of the reference algorithm, not Julia.
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
        grid = W.amp_freq_grid(37, 32 * world)
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


def recorded_traffic(kernel, n_traj, n_ts):
    """DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) from the committed `ncu --set full`
    capture of this kernel on this workload (profiles/traffic.json), or None when no capture matches."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            for rec in json.load(f):
                if rec["kernel"] == kernel and rec["points"] == n_traj and rec["n_ts"] == n_ts:
                    return rec["dram_bytes_read"] + rec["dram_bytes_write"], rec["source"]
    except (OSError, ValueError, KeyError):
        pass
    return None, None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--small", action="store_true", help="reduced sweep (1184 points, 20 ns) for ncu captures")
    ap.add_argument("--points", type=int, default=0, help="development: override the number of sweep points")
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

    peak_tflops = ctx.peak_fp64(300.0)
    for _ in range(max(args.warmup, 3)):
        flush.zero_()
        launch()
    torch.cuda.synchronize()

    l0 = ctx.launch_count()
    sampler = ClockSampler(local_rank)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    t_wall0 = time.perf_counter()
    for i in range(args.steps):
        flush.zero_()
        ev[i][0].record(stream)
        launch()
        ev[i][1].record(stream)
    barrier()
    t_wall = time.perf_counter() - t_wall0
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

    if rank == 0:
        ms_per_step = total_ms_max / args.steps
        value = n_traj * world / (ms_per_step * 1e-3)
        achieved = flops_per_launch / (total_ms / args.steps * 1e-3) / 1e12
        traffic, traffic_src = recorded_traffic(prob.kernel_name(which), n_traj, len(ts))
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
                         "traffic": traffic, "traffic_unit": "DRAM bytes per launch", "traffic_source": traffic_src,
                         "peak_source": "DFMA micro-benchmark in this run (MEASURED_PEAKS.json has no FP64 figure; "
                                        "nominal 37.2)",
                         "flop_counted": "executed sparse generator multiply-adds only (%d flop each: the generator is "
                                         "%s); RK stage combinations, error norm and drive evaluation not counted"
                                         % ((4, "purely imaginary, 2 DFMA per entry") if "_vm0" not in prob.kernel_name(which)
                                            else (8, "complex, 4 DFMA per entry")),
                         "dense_equivalent_tflops": dense_flops_per_launch / (total_ms / args.steps * 1e-3) / 1e12,
                         # informational: + the Runge-Kutta vector arithmetic any Tsit5 must do per attempted step and
                         # complex state entry (21 stage-combination and 7 error-estimate real-by-complex multiply-adds
                         # at 4 flop, 6 stage arguments u + dt*(...) at 4 flop, ~10 flop of scaled error norm = 146 flop)
                         "tsit5_total_tflops": (flops_per_launch + 146.0 * d * d * attempts)
                                               / (total_ms / args.steps * 1e-3) / 1e12,
                         "hbm_out_gbs": d_out.numel() * 16 / (total_ms / args.steps * 1e-3) / 1e9},
            "gpu_launches": int(launches), "clocks": clocks,
        }
        if e2e is not None:
            line["e2e"] = {"value": n_traj * world * args.steps / e2e_s_max, "unit": "propagators/s",
                           "h2d_bytes_per_step": e2e["h2d"], "d2h_bytes_per_step": e2e["d2h"]}
        if cpu_base is not None:
            line["cpu_baseline"] = cpu_base
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
