"""nvdisasm text (with line info) of the cubin that holds a kernel whose mangled name contains `sub`.
One object file per launch shape lives in vsrevolution_b200/csrc/build (or $VSR_OBJ_DIR)."""
import glob, os, subprocess, tempfile


def kernel_sass(sub: str) -> list[str]:
    objdir = os.environ.get("VSR_OBJ_DIR") or os.path.join(os.path.dirname(__file__), "..", "vsrevolution_b200", "csrc", "build")
    for o in sorted(glob.glob(os.path.join(objdir, "inst*.o"))):
        d = tempfile.mkdtemp()
        subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(o)], cwd=d, stdout=subprocess.DEVNULL)
        for cub in glob.glob(d + "/*.cubin"):
            t = subprocess.run(["nvdisasm", "--print-line-info", cub], capture_output=True, text=True).stdout.splitlines()
            if any(l.startswith(".text.") and sub in l for l in t):
                return t
    raise SystemExit(f"no kernel matching {sub}")
