#!/bin/bash
# Round 2, component pruning: follow-up checks (one B200).
mkdir -p gpurun_out
{
timeout 600 python -m pytest tests/test_delivery_gpu.py tests/test_component_pruning.py -q -m gpu 2>&1 | tail -15
python scripts/perf_probe.py cfg3:1184:50 cfg2:4096:200
echo "--- r1c4"; QSIM_CF_SHAPE=r1c4 python scripts/perf_probe.py cfg3:1184:50
echo "--- r1c3"; QSIM_CF_SHAPE=r1c3 python scripts/perf_probe.py cfg3:1184:50
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --configs none
} > gpurun_out/prune_check4.log 2>&1
tail -40 gpurun_out/prune_check4.log | cut -c1-600
