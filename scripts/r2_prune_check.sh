#!/bin/bash
# Round 2: kernel shapes of the cfg4 component form.
mkdir -p gpurun_out
{
python scripts/perf_probe.py cfg4:1184:5 2>&1 | tail -1
echo "--- r2c2"; QSIM_CF_SHAPE=r2c2 python scripts/perf_probe.py cfg4:1184:5 2>&1 | tail -1
echo "--- r2c1"; QSIM_CF_SHAPE=r2c1 python scripts/perf_probe.py cfg4:1184:5 2>&1 | tail -1
} > gpurun_out/prune_check12.log 2>&1
cat gpurun_out/prune_check12.log | cut -c1-330
