"""Aggregate ncu per-line stats into named kernel regions (line ranges read from section comments)."""
import collections, csv, glob, io, os, re, subprocess, sys, tempfile
rep, sub, per = sys.argv[1], sys.argv[2], float(sys.argv[3])
src = os.path.join(os.path.dirname(__file__), "..", "vsrevolution_b200", "csrc", "vsr_kernel.cuh")
marks = []
for i, l in enumerate(open(src), 1):
    m = re.search(r"// ---- ([^-]+?) -+\s*$", l) or re.search(r"^\s*// (welds, \w+ round|contacts of my corner masses|weld position solve|robot-wide convergence|narrowphase)", l) \
        or re.search(r"(auto detect = )", l) or re.search(r"^#define (SEL4)", l) or re.search(r"^(template <class R> struct M;)", l)
    if m: marks.append((i, m.group(1).strip()))
def region(f, ln):
    if "intrinsics" in f: return "shuffle intrinsics"
    if f != "vsr_kernel.cuh": return f
    name = "prologue"
    for i, n in marks:
        if ln >= i: name = f"{i}:{n}"
    return name
sys.path.insert(0, os.path.dirname(__file__))
from _sass import kernel_sass
txt = kernel_sass(sub)
inside = False; cur = None; line_of = {}; static = collections.Counter()
for line in txt:
    if line.startswith(".text."): inside = sub in line; continue
    if not inside: continue
    if line.startswith("\t.section"): inside = False; continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', line)
    if m: cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    m = re.match(r'\s+/\*([0-9a-f]{4,6})\*/', line)
    if m: line_of[int(m.group(1), 16)] = cur; static[region(*cur) if cur else "?"] += 1
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]; N = {n: i for i, n in enumerate(hdr)}
base = None; agg = collections.defaultdict(collections.Counter); tot = collections.Counter()
K = ("Instructions Executed", "Thread Instructions Executed", "# Samples", "stall_no_inst", "stall_wait", "stall_short_sb", "stall_long_sb", "stall_barrier", "stall_branch_resolving", "stall_not_selected", "stall_selected")
for r in rows[2:]:
    if len(r) < len(hdr) or not r[0].startswith("0x"): continue
    a = int(r[0], 16)
    if base is None: base = a
    ln = line_of.get(a - base)
    key = region(*ln) if ln else "?"
    for k in K:
        v = int(r[N[k]] or 0); agg[key][k] += v; tot[k] += v
S = tot["# Samples"]
print(f"total: {tot['Instructions Executed']/per:.0f} inst/warp-step, thread-eff {tot['Thread Instructions Executed']/tot['Instructions Executed']/32:.2f}, static {sum(static.values())} instrs")
print(f"{'region':58s} static inst/wstep thr-eff samples% | no_inst  wait  ssb   lsb   barr  branch notsel sel")
for key, c in sorted(agg.items(), key=lambda kv: -kv[1]["# Samples"]):
    ie = c["Instructions Executed"]
    if c["# Samples"] < 0.002 * S: continue
    print(f"{key[:58]:58s} {static[key]:5d} {ie/per:9.1f}  {c['Thread Instructions Executed']/max(ie,1)/32:5.2f}  {100*c['# Samples']/S:6.2f}  | "
          + " ".join(f"{100*c[k]/S:5.2f}" for k in K[3:]))
