set -x
python -m pytest tests -m gpu -x -q > gpurun_out/f_pytest.log 2>&1; tail -2 gpurun_out/f_pytest.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/f_smoke.log 2>&1; tail -1 gpurun_out/f_smoke.log
python bench.py --impl reference > gpurun_out/f_bench_ref.log 2>&1; tail -1 gpurun_out/f_bench_ref.log
python bench.py > gpurun_out/f_bench.log 2>&1; tail -1 gpurun_out/f_bench.log
python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/f_plain_bench.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1_final_bench_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/f_ncu_launches.log 2>&1
python bench.py --small --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/f_plain_small.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:solve_kernel -s 1 -c 1 -f -o gpurun_out/prof_r1_final_small python bench.py --small --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/f_ncu_small.log 2>&1
python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/f_plain_full1.log 2>&1 && \
timeout 600 ncu --set full --clock-control none -k regex:solve_kernel -s 3 -c 1 -f -o gpurun_out/prof_r1_final_full python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/f_ncu_full.log 2>&1
ls -la gpurun_out | tail -8
