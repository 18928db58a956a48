"""Time one host-buffer call with PAGEABLE caller buffers (pipelined delivery) next to the page-locked (zero-copy) call:
cfg2, 4096 points per context.  usage: python scripts/pageable_probe.py [n_ctx]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "qsimulator.jl_b200", "host")]
import numpy as np
from qsim_b200 import Context, PinnedArray, Problem
from qsim_b200 import workloads as W

n_ctx = int(sys.argv[1]) if len(sys.argv) > 1 else 1
cqs = W.lab_frame_parametric_gate(2)
ts = np.arange(0.0, 201.0)
params = W.amp_freq_grid(64, 64 * n_ctx)
prob = Problem(cqs)
ctxs = [Context(i) for i in range(n_ctx)]
arg = ctxs if n_ctx > 1 else ctxs[0]
out = np.empty((len(params), len(ts), 81), dtype=np.complex128)
out[...] = 0
prob.solve(0, ts, params=params[:64 * n_ctx], ctx=arg)
for rep in range(3):
    t0 = time.perf_counter()
    _, st = prob.solve(0, ts, params=params, ctx=arg, out=out)
    dt = time.perf_counter() - t0
    print(f"pageable, {n_ctx} ctx: {len(params)} points in {dt:.3f} s = {len(params) / dt:.0f} propagators/s, "
          f"kernel window {ctxs[0].last_kernel_ms():.0f} ms", flush=True)
ref = out[::97].copy()
if n_ctx == 1:
    pin = PinnedArray(out.shape, np.complex128)
    for rep in range(2):
        t0 = time.perf_counter()
        prob.solve(0, ts, params=params, ctx=arg, out=pin.array)
        dt = time.perf_counter() - t0
        print(f"page-locked: {len(params) / dt:.0f} propagators/s ({dt:.3f} s)")
    assert np.array_equal(pin.array[::97], ref)
