"""Write a text summary of an .ncu-rep (raw page key metrics + per-region instruction/sample breakdown).
usage: ncu_summary.py <rep> <inst object .o> <mangled kernel substring> <n_steps> <out.txt> [header...]"""
import csv, os, subprocess, sys, tempfile
rep, obj, kern, nsteps, out = sys.argv[1:6]
hdr_lines = sys.argv[6:]
tmp = tempfile.mkdtemp()
raw = os.path.join(tmp, "raw.csv"); src = os.path.join(tmp, "src.csv")
subprocess.run(f"ncu -i {rep} --page raw --csv > {raw} 2>/dev/null", shell=True, check=True)
subprocess.run(f"ncu -i {rep} --page source --csv > {src} 2>/dev/null", shell=True, check=True)
subprocess.run(f"cd {tmp} && cuobjdump -xelf all {os.path.abspath(obj)} > /dev/null && nvdisasm -g -c *.cubin > dis.txt 2>/dev/null", shell=True, check=True)
rows = list(csv.reader(open(raw))); h, u, v = rows[0], rows[1], rows[2]
want = ['gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'launch__waves_per_multiprocessor', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'sass__inst_executed_local_loads', 'sass__inst_executed_local_stores',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio']
with open(out, "w") as f:
    for l in hdr_lines: f.write("# " + l + "\n")
    for a, b, c in zip(h, u, v):
        if a in want: f.write(f"{a} [{b}] = {c}\n")
    f.write("\n# instructions per attempted step by source region (ncu source page x nvdisasm line info)\n")
    r = subprocess.run([sys.executable, os.path.join(os.path.dirname(__file__), "ncu_by_region.py"), src, os.path.join(tmp, "dis.txt"), kern, nsteps],
                       capture_output=True, text=True, cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    f.write(r.stdout)
print(open(out).read())
