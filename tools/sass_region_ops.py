"""Static opcode histogram for a source-line range of one kernel variant. usage: sass_region_ops.py <substr> <lo> <hi>"""
import collections, glob, os, re, subprocess, sys, tempfile
sub, lo, hi = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
so = os.path.join(os.path.dirname(__file__), "..", "vsrevolution_b200", "libvsr_eval.so")
d = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=d, stdout=subprocess.DEVNULL)
cub = [f for f in glob.glob(d + "/*.cubin") if "names" not in f][0]
txt = subprocess.run(["nvdisasm", "--print-line-info", cub], capture_output=True, text=True).stdout.splitlines()
inside = False; cur = None; ops = collections.Counter(); n = 0
for line in txt:
    if line.startswith(".text."): inside = sub in line; continue
    if not inside: continue
    if line.startswith("\t.section"): inside = False; continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', line)
    if m: cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    m = re.match(r'\s+/\*[0-9a-f]{4,6}\*/\s+(@!?U?P\d\s+)?([A-Z0-9_.]+)', line)
    if m and cur and cur[0] == "vsr_kernel.cuh" and lo <= cur[1] <= hi:
        ops[m.group(2).split('.')[0]] += 1; n += 1
print(n, ops.most_common(30))
