"""Hot spots of an ncu source-page CSV: per SASS instruction group (by CUDA source line via nvdisasm) the executed
instructions, stall samples and their dominant reasons, and shared-memory wavefronts (ideal / excessive).
usage: ncu_hot.py <src.csv from `ncu --page source --csv`> <object .o> <mangled kernel substring> [top N]"""
import collections, csv, os, re, subprocess, sys, tempfile
srccsv, obj, kern = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
tmp = tempfile.mkdtemp()
subprocess.run(f"cd {tmp} && cuobjdump -xelf all {os.path.abspath(obj)} > /dev/null && nvdisasm -g -c *.cubin > dis.txt 2>/dev/null",
               shell=True, check=True)
cur = None; infn = False; off2line = {}
for ln in open(os.path.join(tmp, "dis.txt")):
    if ln.startswith("//---") and ".text." in ln:
        infn = kern in ln
    if not infn: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m: cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*);", ln)
    if m: off2line[int(m.group(1), 16)] = cur
rows = list(csv.reader(open(srccsv)))
hdr = rows[1]; data = rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
base = int(data[0][ix["Address"]], 16)
agg = collections.defaultdict(lambda: collections.Counter())
tot = collections.Counter()
for r in data:
    off = int(r[ix["Address"]], 16) - base
    key = off2line.get(off, ("?", 0))
    a = agg[key]
    a["inst"] += int(r[ix["Instructions Executed"]]); a["samples"] += int(r[ix["# Samples"]])
    a["wf"] += int(r[ix["L1 Wavefronts Shared"]] or 0); a["wf_ideal"] += int(r[ix["L1 Wavefronts Shared Ideal"]] or 0)
    for s in stalls: a[s] += int(r[ix[s]] or 0)
for a in agg.values(): tot.update(a)
print(f"total inst {tot['inst']}  samples {tot['samples']}  smem wavefronts {tot['wf']} (ideal {tot['wf_ideal']})")
print("stall mix: " + ", ".join(f"{s[6:]} {tot[s]/max(tot['samples'],1)*100:.1f}%" for s in sorted(stalls, key=lambda s: -tot[s])[:8]))
for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["samples"])[:top]:
    top3 = ", ".join(f"{s[6:]} {a[s]/max(a['samples'],1)*100:.0f}%" for s in sorted(stalls, key=lambda s: -a[s])[:3])
    print(f"{k[0]}:{k[1]:<5d} inst {a['inst']/tot['inst']*100:5.1f}%  samples {a['samples']/tot['samples']*100:5.1f}%  "
          f"wf {a['wf']/max(tot['wf'],1)*100:5.1f}% (x{a['wf']/max(a['wf_ideal'],1):.2f})  {top3}")
