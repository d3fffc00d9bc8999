"""Attribute SASS instructions of one kernel variant to source lines (needs -lineinfo).
usage: python tools/sass_lines.py <mangled-substring> [bucket]"""
import collections, re, subprocess, sys, tempfile, os, glob
sub = sys.argv[1]; bucket = int(sys.argv[2]) if len(sys.argv) > 2 else 10
sys.path.insert(0, os.path.dirname(__file__))
from _sass import kernel_sass
txt = kernel_sass(sub)
inside = False; cur = None; cnt = collections.Counter(); ops = collections.Counter()
for line in txt:
    if line.startswith(".text."):
        inside = sub in line; continue
    if not inside: continue
    if line.startswith("\t.section"): inside = False; continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', line)
    if m: cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    m = re.match(r'\s+/\*[0-9a-f]{4,6}\*/\s+(@!?U?P\d\s+)?([A-Z0-9_.]+)', line)
    if m and cur:
        cnt[cur] += 1; ops[m.group(2).split('.')[0]] += 1
tot = sum(cnt.values()); print("total instructions", tot)
b = collections.Counter()
for (f, l), c in cnt.items(): b[(f, l // bucket * bucket)] += c
for k, c in sorted(b.items(), key=lambda x: (x[0][0], x[0][1])): 
    if c >= 15: print(f"{k[0]}:{k[1]:4d}  {c}")
print(ops.most_common(25))
