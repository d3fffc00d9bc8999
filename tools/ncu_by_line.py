"""Join ncu's SASS-level source page (executed counts, stall samples) with nvdisasm line info.
usage: python tools/ncu_by_line.py <report.ncu-rep> <mangled-kernel-substring> <warps*steps normaliser> [bucket]"""
import collections, csv, glob, io, os, re, subprocess, sys, tempfile
rep, sub, per = sys.argv[1], sys.argv[2], float(sys.argv[3])
bucket = int(sys.argv[4]) if len(sys.argv) > 4 else 1
sys.path.insert(0, os.path.dirname(__file__))
from _sass import kernel_sass
txt = kernel_sass(sub)
inside = False; cur = None; line_of = {}
for line in txt:
    if line.startswith(".text."): inside = sub in line; continue
    if not inside: continue
    if line.startswith("\t.section"): inside = False; continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', line)
    if m: cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    m = re.match(r'\s+/\*([0-9a-f]{4,6})\*/', line)
    if m: line_of[int(m.group(1), 16)] = cur
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]; N = {n: i for i, n in enumerate(hdr)}
base = None; agg = collections.defaultdict(lambda: collections.Counter()); tot = collections.Counter()
for r in rows[2:]:
    if len(r) < len(hdr) or not r[0].startswith("0x"): continue
    a = int(r[0], 16)
    if base is None: base = a
    ln = line_of.get(a - base)
    key = (ln[0], ln[1] // bucket * bucket) if ln else ("?", 0)
    for k in ("Instructions Executed", "# Samples", "stall_no_inst", "stall_wait", "stall_short_sb", "stall_long_sb", "stall_barrier", "stall_branch_resolving", "Thread Instructions Executed"):
        v = int(r[N[k]] or 0); agg[key][k] += v; tot[k] += v
print(f"total: {tot['Instructions Executed']/per:.0f} inst/warp-step, thread-eff {tot['Thread Instructions Executed']/tot['Instructions Executed']/32:.2f}, samples {tot['# Samples']}")
S = tot["# Samples"]
print("file:line  inst/warp-step  thr-eff  samples%  [no_inst wait short_sb long_sb barrier branch]%")
for key, c in sorted(agg.items(), key=lambda kv: -kv[1]["# Samples"])[:int(os.environ.get("TOP", 45))]:
    ie = c["Instructions Executed"]
    print(f"{key[0]}:{key[1]:<5d} {ie/per:9.1f}  {c['Thread Instructions Executed']/max(ie,1)/32:5.2f}  {100*c['# Samples']/S:6.2f}   "
          + " ".join(f"{100*c[k]/S:5.2f}" for k in ("stall_no_inst", "stall_wait", "stall_short_sb", "stall_long_sb", "stall_barrier", "stall_branch_resolving")))
