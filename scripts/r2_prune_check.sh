#!/bin/bash
# Round 2: row-split component passes, sixteen- against eight-warp teams, selected output as in bench.py.
mkdir -p gpurun_out
{
python scripts/perf_probe.py cfg5:148:1:sel cfg5:592:2:sel 2>&1 | tail -2
echo "--- QSIM_RS_R3"; QSIM_RS_R3=1 python scripts/perf_probe.py cfg5:148:1:sel cfg5:592:2:sel 2>&1 | tail -2
} > gpurun_out/prune_check11.log 2>&1
cat gpurun_out/prune_check11.log | cut -c1-330
