"""Development probe (GPU box): repeat the two-contexts-on-one-device fan-out into a page-locked buffer and compare
with a single-context run (tests/test_delivery_gpu.py::test_block_cyclic_fanout_two_contexts_one_device)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "qsimulator.jl_b200", "host")]
import numpy as np
from qsim_b200 import Context, PinnedArray, Problem
from qsim_b200 import workloads as W

cqs, params, _, _ = W.config("cfg2")
params = np.ascontiguousarray(params[np.linspace(0, len(params) - 1, 301).astype(int)])
ts = np.linspace(0.0, 3.0, 4)
prob = Problem(cqs)
ref, st_ref = prob.solve("unitary_propagator", ts, params=params, ctx=Context(0))
ctxs = [Context(0), Context(0)]
pin = PinnedArray((len(params), len(ts), 81), np.complex128)
for mode in ("pinned", "pageable"):
    for it in range(int(sys.argv[1]) if len(sys.argv) > 1 else 10):
        if mode == "pinned":
            pin.array[...] = np.nan
            got, st = prob.solve("unitary_propagator", ts, params=params, ctx=ctxs, out=pin.array)
        else:
            got, st = prob.solve("unitary_propagator", ts, params=params, ctx=ctxs)
        bad = np.argwhere(~(got == ref))
        if len(bad):
            trajs = np.unique(bad[:, 0])
            steps_differ = np.flatnonzero(st["n_accept"] != st_ref["n_accept"])
            print(mode, it, "MISMATCH", len(bad), "entries,", len(trajs), "trajectories, parity of trajectory index:",
                  np.unique(trajs % 2).tolist(), "max abs diff", float(np.nanmax(np.abs(got - ref))), "nan:", int(np.isnan(got.real).sum()),
                  "step counts differ on", len(steps_differ), "dt_init differ", int((st["dt_init"] != st_ref["dt_init"]).sum()),
                  "first bad", bad[0].tolist())
        else:
            print(mode, it, "ok")
