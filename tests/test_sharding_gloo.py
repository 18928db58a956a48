"""The N > 1 path on CPU: two gloo ranks shard a sweep block-cyclically, integrate their shards (with the
CPU oracle standing in for the GPU kernels, which need a device) and gather; the result must equal the
single-process sweep bit for bit, with every point owned by exactly one rank."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    for p in (ROOT, os.path.join(ROOT, "qsimulator.jl_b200", "host")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch.distributed as dist
    from oracle import c_oracle as co
    from qsim_b200 import gather_results, shard_params
    from qsim_b200 import workloads as W
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    cqs = W.lab_frame_parametric_gate(2)
    params = W.amp_freq_grid(3, 3)[:7]          # ragged: 7 points over 2 ranks
    ts = np.array([0.0, 0.5, 1.0])
    local, idx = shard_params(params, world, rank)
    out, _ = co.OracleProblem(cqs).solve(co.UNITARY_PROPAGATOR, ts, local)
    full = gather_results(out, len(params))
    dist.barrier()
    dist.destroy_process_group()
    q.put((rank, idx.tolist(), full))


def test_two_rank_sharded_sweep_matches_single_process():
    from oracle import c_oracle as co
    from qsim_b200 import shard_indices
    from qsim_b200 import workloads as W
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    owned = sorted(i for _, idx, _ in res for i in idx)
    assert owned == list(range(7))                      # a partition: every point exactly once
    cqs = W.lab_frame_parametric_gate(2)
    params = W.amp_freq_grid(3, 3)[:7]
    want, _ = co.OracleProblem(cqs).solve(co.UNITARY_PROPAGATOR, np.array([0.0, 0.5, 1.0]), params)
    for _, _, full in res:
        assert np.array_equal(full, want)
    assert np.array_equal(np.sort(np.concatenate([shard_indices(4096, 8, r) for r in range(8)])), np.arange(4096))
    with pytest.raises(ValueError):
        shard_indices(10, 2, 2)
