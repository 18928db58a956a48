"""Aggregate an ncu SASS-level source page by CUDA source line, using nvdisasm line info.
usage: ncu_by_line.py <src.csv from `ncu --page source --csv`> <nvdisasm -g -c output> <mangled kernel substring>"""
import csv, re, sys, collections
srccsv, dis, kern = sys.argv[1:4]
# offsets -> (file, line) from nvdisasm
cur = None; infn = False; off2line = {}
for ln in open(dis):
    if ln.startswith("//---") and ".text." in ln:
        infn = kern in ln
    if not infn: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m: cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*);", ln)
    if m: off2line[int(m.group(1), 16)] = (cur, m.group(2).strip())
rows = list(csv.reader(open(srccsv)))
hdr = rows[1]; ia = hdr.index("Address"); ii = hdr.index("Instructions Executed"); isamp = hdr.index("# Samples")
data = rows[2:]
base = int(data[0][ia], 16)
by = collections.defaultdict(lambda: [0, 0]); tot = 0; tots = 0
byop = collections.defaultdict(int)
for r in data:
    off = int(r[ia], 16) - base
    n = int(r[ii]); s = int(r[isamp])
    key, sass = off2line.get(off, (("?", 0), "?"))
    by[key][0] += n; by[key][1] += s; tot += n; tots += s
    byop[sass.split()[0] if not sass.startswith("@") else sass.split()[1]] += n
print("total inst", tot, "samples", tots)
for k, v in sorted(by.items(), key=lambda kv: -kv[1][0])[:45]:
    print(f"{k[0]}:{k[1]:<5d} inst {v[0]/tot*100:5.1f}%  samples {v[1]/max(tots,1)*100:5.1f}%")
print("--- by opcode")
for k, v in sorted(byop.items(), key=lambda kv: -kv[1])[:25]:
    print(f"{k:24s} {v/tot*100:5.1f}%")
