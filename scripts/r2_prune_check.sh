#!/bin/bash
# Round 2: operand chunk sizes of the cfg2 component-form kernel (alternate builds of qsim_inst_lean_c2.cu).
mkdir -p gpurun_out
{
for tag in b200 ch422 ch444 ch111; do
  echo "--- $tag"; QSIM_B200_LIB=$PWD/qsimulator.jl_b200/lib/libqsim_$tag.so python scripts/perf_probe.py cfg2:4096:200 2>&1 | tail -1
done
} > gpurun_out/prune_check8.log 2>&1
cat gpurun_out/prune_check8.log | cut -c1-300
