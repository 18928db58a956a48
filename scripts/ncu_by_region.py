"""Aggregate ncu SASS instruction counts into source regions (line ranges of qsim_solver.cuh / qsim_eval.h)."""
import csv, re, sys, collections
srccsv, dis, kern, nsteps = sys.argv[1], sys.argv[2], sys.argv[3], float(sys.argv[4])
src = open("qsimulator.jl_b200/csrc/qsim_solver.cuh").read().split("\n")
def find(pat, start=0):
    for i in range(start, len(src)):
        if pat in src[i]: return i + 1
    raise KeyError(pat)
marks = [("lds/sts helpers", find("ld.shared.v2.f64") - 3), ("team_sync/sum", find("void team_sync")), ("compute_tables", find("void compute_tables")),
         ("put_stage", find("void put_stage")), ("apply (SpMV)", find("void apply(")), ("write_out/init/save/ldst", find("void write_out")),
         ("run: setup+initdt", find("void run(")), ("run: step head", find("// ---- main loop ----")), ("run: stage combos", find("A21 * k1[e].x") - 2),
         ("run: error norm", find("BT1 * k1[e].x") - 2), ("run: streamed store", find("if (a.streamed) {", find("BT1 * k1[e].x"))),
         ("run: controller", find("n_attempt++;")), ("run: tail", find("st.n_rhs = 3 + 6")), ("kernel prologue", find("__global__ void"))]
def region(f, line):
    if f == "qsim_eval.h": return "evaluator (eval.h)"
    if f != "qsim_solver.cuh": return f"libdevice/{f}"
    name = "header"
    for nm, ln in marks:
        if line >= ln: name = nm
    return name
cur = None; infn = False; off2 = {}
for ln in open(dis):
    if ln.startswith("//---") and ".text." in ln: infn = kern in ln
    if not infn: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m: cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*);", ln)
    if m: off2[int(m.group(1), 16)] = (cur, m.group(2).strip())
rows = list(csv.reader(open(srccsv))); hdr = rows[1]
ia, ii, isamp = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples")
data = rows[2:]; base = int(data[0][ia], 16)
agg = collections.defaultdict(lambda: [0, 0, 0]); tot = 0; tots = 0
prev = "kernel prologue"
for r in data:
    key, sass = off2.get(int(r[ia], 16) - base, (None, "?"))
    reg = region(*key) if key else prev
    # libdevice code (no line info) inherits the last line marker, which nvdisasm repeats; keep as-is
    n, s = int(r[ii]), int(r[isamp])
    agg[reg][0] += n; agg[reg][1] += s; tot += n; tots += s
    op = sass.split()[1] if sass.startswith("@") else sass.split()[0]
    if op.startswith(("DFMA", "DMUL", "DADD")): agg[reg][2] += n
print(f"total {tot/nsteps:.0f} instr/step, {tots} samples")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f"{k:28s} {v[0]/nsteps:7.0f} instr/step ({v[0]/tot*100:4.1f}%)  fp64 {v[2]/nsteps:6.0f}  samples {v[1]/tots*100:4.1f}%")
