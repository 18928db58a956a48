"""Summarise a full-size ncu capture into profiles/: key raw-page metrics, per-step instruction and wavefront counts,
and profiles/traffic.json (DRAM bytes per launch, bench.py's cross-check).
usage: full_capture_summary.py <rep> <plain bench log of the same command> <output name under profiles/>"""
import csv, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "prof_r1_final_full.ncu-rep")
plain_log = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "gpurun_out", "f_plain_full1.log")
raw = subprocess.run(f"ncu -i {rep} --page raw --csv 2>/dev/null", shell=True, capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h, u, v = rows[0], rows[1], rows[2]
d = {a: (b, c) for a, b, c in zip(h, u, v)}
want = ['gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'sass__inst_executed_local_loads', 'sass__inst_executed_local_stores',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio']
unit = {'Mbyte': 1e6, 'Gbyte': 1e9, 'Kbyte': 1e3, 'byte': 1}
rd = float(d['dram__bytes_read.sum'][1]) * unit[d['dram__bytes_read.sum'][0]]
wr = float(d['dram__bytes_write.sum'][1]) * unit[d['dram__bytes_write.sum'][0]]
plain = json.loads(open(plain_log).read().strip().splitlines()[-1])
steps = plain['config']['attempted_steps_per_launch']
kernel = plain['config']['kernel']
out_name = sys.argv[3] if len(sys.argv) > 3 else "r1_ncu_cfg2full_final.txt"
out = os.path.join(ROOT, "profiles", out_name)
with open(out, "w") as f:
    f.write(f"# ncu --set full --clock-control none, kernel {d['Kernel Name'][1].strip()} ({kernel})\n")
    f.write("# command: python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline  (-k regex:solve_kernel -s 3 -c 1: the timed full-size launch)\n")
    f.write(f"# 4096 trajectories x 200 ns, 201 saves each = {steps} attempted Tsit5 steps, 1 067 122 688 B of results\n")
    f.write(f"# same command without ncu: {plain['ms_per_step']:.1f} ms per launch\n")
    for a in want:
        if a in d:
            f.write(f"{a} [{d[a][0]}] = {d[a][1]}\n")
    f.write(f"\n# per attempted step: {float(d['smsp__inst_executed.sum'][1]) / steps:.0f} warp instructions, "
            f"{float(d['l1tex__data_pipe_lsu_wavefronts_mem_shared.sum'][1]) / steps:.0f} shared-memory wavefronts\n")
    f.write(f"# DRAM traffic per launch = {rd / 1e6:.1f} MB read + {wr / 1e6:.1f} MB written = {(rd + wr) / 1e9:.3f} GB vs 1.067 GB algorithmic\n"
            "# (the result stream; the few tens of MB on top are dirty lines of the 256 MB L2-flush buffer written before the\n"
            "# launch and evicted during it, and varies from capture to capture); no re-reads of the state.\n")
tj = os.path.join(ROOT, "profiles", "traffic.json")
recs = [r for r in (json.load(open(tj)) if os.path.exists(tj) else []) if r.get("kernel") != kernel]
recs.append({"kernel": kernel, "points": 4096, "n_ts": 201, "dram_bytes_read": rd, "dram_bytes_write": wr,
             "source": f"profiles/{out_name} (ncu --set full of the timed full-size launch)"})
json.dump(recs, open(tj, "w"), indent=1)
print(open(out).read())
