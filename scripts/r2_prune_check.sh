#!/bin/bash
# Round 2, component pruning: follow-up checks (one B200).
mkdir -p gpurun_out
{
timeout 900 python -m pytest tests/test_component_pruning.py -q -m gpu -k "state" 2>&1 | tail -30
python scripts/state_probe.py cfg2 4096 20 2>&1 | tail -2
QSIM_NO_PRUNE=1 python scripts/state_probe.py cfg2 4096 20 2>&1 | tail -2
python scripts/state_probe.py cfg3 2048 20 2>&1 | tail -2
QSIM_NO_PRUNE=1 python scripts/state_probe.py cfg3 2048 20 2>&1 | tail -2
} > gpurun_out/prune_check6.log 2>&1
tail -60 gpurun_out/prune_check6.log | cut -c1-400
