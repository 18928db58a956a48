# Round-2 (component pruning) evidence run, one B200: the whole GPU test suite, the default bench line, the launch list
# of the bench command, ncu --set full of the cfg2 kernel (reduced and full-size launch) and of the kernels behind
# cfg3 / cfg4 / cfg5.  Every ncu pass follows an un-profiled run of the same command that exited 0.
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r2g_pytest_gpu.log 2>&1; tail -3 gpurun_out/r2g_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2g_smoke.log 2>&1; tail -1 gpurun_out/r2g_smoke.log
timeout 900 python bench.py > gpurun_out/bench_r2g.json 2> gpurun_out/bench_r2g.err; tail -c 600 gpurun_out/bench_r2g.json
python bench.py --steps 2 --warmup 1 --no-cpu-baseline --configs none > gpurun_out/r2g_plain_bench.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2g_bench_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --configs none > gpurun_out/r2g_ncu_launches.log 2>&1
python bench.py --small --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --configs none > gpurun_out/r2g_plain_small.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:solve_kernel -s 1 -c 1 -f -o gpurun_out/prof_r2g_small python bench.py --small --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --configs none > gpurun_out/r2g_ncu_small.log 2>&1
python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --configs none > gpurun_out/r2g_plain_full1.log 2>&1 && \
timeout 600 ncu --set full --clock-control none -k regex:solve_kernel -s 3 -c 1 -f -o gpurun_out/prof_r2g_full python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --configs none > gpurun_out/r2g_ncu_full.log 2>&1
for spec in cfg3:296:10 cfg4:148:2 cfg5:148:0.3; do
  name=${spec%%:*}
  python scripts/config_table.py $spec --no-cpu > gpurun_out/r2g_${name}_plain.log 2>&1 && \
  timeout 500 ncu --set full --clock-control none -k regex:solve_kernel -s 1 -c 1 -f -o /tmp/prof_r2g_${name} python scripts/config_table.py $spec --no-cpu > gpurun_out/r2g_${name}_ncu.log 2>&1 && \
  ncu -i /tmp/prof_r2g_${name}.ncu-rep --page raw --csv > gpurun_out/prof_r2g_${name}_raw.csv 2>/dev/null
done
ls -la gpurun_out | tail -24
