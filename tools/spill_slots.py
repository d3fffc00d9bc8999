"""List local-memory spill slots of one kernel variant with the source lines of their stores/loads."""
import subprocess, re, glob, os, tempfile, collections, sys
sub = sys.argv[1]
sys.path.insert(0, os.path.dirname(__file__))
from _sass import kernel_sass
txt = kernel_sass(sub)
inside = False; cur = None
st = collections.defaultdict(list); ld = collections.defaultdict(list)
for line in txt:
    if line.startswith(".text."): inside = sub in line; continue
    if not inside: continue
    if line.startswith("\t.section"): inside = False; continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', line)
    if m: cur = int(m.group(2)) if "vsr_kernel" in m.group(1) else -1; continue
    m = re.search(r'(STL|LDL)[.\w]*\s+(.*);', line)
    if m:
        off = re.search(r'\[R1(\+0x[0-9a-f]+)?\]', m.group(2))
        o = int(off.group(1), 16) if off and off.group(1) else 0
        (st if m.group(1) == "STL" else ld)[o].append(cur)
for o in sorted(set(st) | set(ld)):
    print(f"slot {o:4d}: stores at lines {sorted(set(st[o]))}  loads at lines {sorted(set(ld[o]))}")
