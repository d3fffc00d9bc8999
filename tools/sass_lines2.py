"""SASS of one kernel from one object file, with source lines: python tools/sass_lines2.py <obj> <mangled-substring> [lo hi]
prints per-opcode counts, and the instructions attributed to source lines lo..hi."""
import collections, re, subprocess, sys, tempfile, glob, os
obj, sub = sys.argv[1], sys.argv[2]
lo, hi = (int(sys.argv[3]), int(sys.argv[4])) if len(sys.argv) > 4 else (0, -1)
d = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=d, stdout=subprocess.DEVNULL)
for cub in glob.glob(d + "/*.cubin"):
    t = subprocess.run(["nvdisasm", "--print-line-info", cub], capture_output=True, text=True).stdout.splitlines()
    inside = False; cur = None; out = []
    for line in t:
        if line.startswith(".text."): inside = sub in line; continue
        if not inside: continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', line)
        if m: cur = int(m.group(2)); continue
        m = re.match(r'\s+/\*([0-9a-f]{4,6})\*/\s+(.*?);', line)
        if m: out.append((cur, m.group(1), m.group(2).strip()))
    if not out: continue
    ops = collections.Counter(); per = collections.Counter()
    for l, a, ins in out:
        op = re.sub(r'^@!?U?P\d\s+', '', ins).split()[0].split('.')[0]
        ops[op] += 1; per[l] += 1
    print("total", len(out)); print(ops.most_common(24))
    print("lines with most instructions:", sorted(per.items(), key=lambda x: -x[1])[:12])
    for l, a, ins in out:
        if lo <= (l or 0) <= hi: print(l, a, ins)
