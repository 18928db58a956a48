# Round-2 evidence run (GPU box): launch list of the bench command, ncu --set full of the cfg2 kernel on the reduced
# and the full-size launch, raw-page metrics of the kernels behind cfg3 / cfg4 / cfg5.  Every ncu pass follows an
# un-profiled run of the same command that exited 0.
set -x
python bench.py --steps 2 --warmup 1 --no-cpu-baseline --configs none > gpurun_out/r2f_plain_bench.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_bench_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --configs none > gpurun_out/r2f_ncu_launches.log 2>&1
python bench.py --small --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --configs none > gpurun_out/r2f_plain_small.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:solve_kernel -s 1 -c 1 -f -o gpurun_out/prof_r2_final_small python bench.py --small --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --configs none > gpurun_out/r2f_ncu_small.log 2>&1
python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --configs none > gpurun_out/r2f_plain_full1.log 2>&1 && \
timeout 600 ncu --set full --clock-control none -k regex:solve_kernel -s 3 -c 1 -f -o gpurun_out/prof_r2_final_full python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --configs none > gpurun_out/r2f_ncu_full.log 2>&1
for spec in cfg3:148:10 cfg4:148:2 cfg5:148:0.3; do
  name=${spec%%:*}
  python scripts/config_table.py $spec --no-cpu > gpurun_out/r2p_${name}_plain.log 2>&1 && \
  timeout 500 ncu --set full --clock-control none -k regex:solve_kernel -s 1 -c 1 -f -o /tmp/prof_r2_${name} python scripts/config_table.py $spec --no-cpu > gpurun_out/r2p_${name}_ncu.log 2>&1 && \
  ncu -i /tmp/prof_r2_${name}.ncu-rep --page raw --csv > gpurun_out/prof_r2_${name}_raw.csv 2>/dev/null
done
ls -la gpurun_out | tail -12
