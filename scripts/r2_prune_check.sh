#!/bin/bash
# Round 2, component pruning: parity + first timings (one B200).
mkdir -p gpurun_out
{
timeout 900 python -m pytest tests/test_component_pruning.py -x -q -m gpu 2>&1 | tail -15
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_selected_output.py -x -q -m gpu 2>&1 | tail -8
python scripts/perf_probe.py cfg2:4096:200 cfg3:1184:50 cfg4:1184:5
echo "--- MINB=3"; QSIM_MAX_MINB=3 python scripts/perf_probe.py cfg2:4096:200
echo "--- NO_PRUNE"; QSIM_NO_PRUNE=1 python scripts/perf_probe.py cfg2:4096:200 cfg3:1184:50 cfg4:1184:5
} > gpurun_out/prune_check.log 2>&1
tail -60 gpurun_out/prune_check.log
