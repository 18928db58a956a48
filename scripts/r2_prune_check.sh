#!/bin/bash
# Round 2, component pruning: parity + timings (one B200).
mkdir -p gpurun_out
{
timeout 900 python -m pytest tests/test_component_pruning.py -q -m gpu 2>&1 | tail -15
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_selected_output.py tests/test_dispatch_coverage_gpu.py tests/test_delivery_gpu.py -q -m gpu 2>&1 | tail -12
python scripts/perf_probe.py cfg2:4096:200 cfg3:1184:50 cfg4:1184:5 cfg5:148:1
echo "--- NO_PRUNE cfg5"; QSIM_NO_PRUNE=1 python scripts/perf_probe.py cfg5:148:1
} > gpurun_out/prune_check3.log 2>&1
tail -40 gpurun_out/prune_check3.log | cut -c1-400
