# Round-2 final check (one B200): the whole GPU test suite, smoke(), the default bench line and the reference arm.
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r2h_pytest_gpu.log 2>&1; tail -3 gpurun_out/r2h_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2h_smoke.log 2>&1; tail -1 gpurun_out/r2h_smoke.log
timeout 900 python bench.py > gpurun_out/bench_r2h.json 2> gpurun_out/bench_r2h.err; tail -c 300 gpurun_out/bench_r2h.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r2h_reference.json 2>&1; tail -c 400 gpurun_out/bench_r2h_reference.json
