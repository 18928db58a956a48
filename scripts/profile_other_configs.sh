# ncu --set full of the kernels behind cfg3 / cfg4 / cfg5 (small launches); only the raw-page CSV is kept
# (the .ncu-rep files together exceed what gpurun copies back)
set -x
for spec in cfg3:148:10 cfg4:148:2 cfg5:148:0.3; do
  name=${spec%%:*}
  python scripts/config_table.py $spec --no-cpu > gpurun_out/p_${name}_plain.log 2>&1 && \
  timeout 500 ncu --set full --clock-control none -k regex:solve_kernel -s 1 -c 1 -f -o /tmp/prof_r1_${name} python scripts/config_table.py $spec --no-cpu > gpurun_out/p_${name}_ncu.log 2>&1 && \
  ncu -i /tmp/prof_r1_${name}.ncu-rep --page raw --csv > gpurun_out/prof_r1_${name}_raw.csv 2>/dev/null
done
ls -la gpurun_out/prof_r1_cfg*_raw.csv
