#!/bin/bash
# Round 2, component pruning: follow-up checks (one B200).
mkdir -p gpurun_out
{
timeout 600 python -m pytest tests/test_component_pruning.py tests/test_gpu_parity.py -q -m gpu 2>&1 | tail -8
python scripts/perf_probe.py cfg5:148:1 cfg5:296:1 cfg3:1184:50
} > gpurun_out/prune_check5.log 2>&1
tail -40 gpurun_out/prune_check5.log | cut -c1-600
