"""Regenerate tests/parity_bounds.json from gpurun_out/parity_measured.jsonl (written by the -m gpu tests on the B200 box):
  max       limit = 3 x measured (floor 1e-12)
  spearman  limit = 1 - 3 (1 - measured)
  min_int   limit = the measured value (deterministic step counts of seeded robots)
usage: python tests/tools/update_bounds.py"""
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
src = os.path.join(ROOT, "gpurun_out", "parity_measured.jsonl")
dst = os.path.join(ROOT, "tests", "parity_bounds.json")
table = json.load(open(dst)) if os.path.exists(dst) else {}
worst = {}
for line in open(src):
    r = json.loads(line)
    k = r["name"]
    if k not in worst:
        worst[k] = r
    elif r["kind"] == "max":
        worst[k] = r if r["value"] > worst[k]["value"] else worst[k]
    else:
        worst[k] = r if r["value"] < worst[k]["value"] else worst[k]
for k, r in worst.items():
    v = r["value"]
    if r["kind"] == "max":
        limit = max(3.0 * v, 1e-12)
    elif r["kind"] == "spearman":
        limit = 1.0 - 3.0 * (1.0 - v)
    else:
        limit = v
    table[k] = {"kind": r["kind"], "measured": v, "limit": float("%.4g" % limit) if r["kind"] != "min_int" else limit}
json.dump(table, open(dst, "w"), indent=1, sort_keys=True)
print(f"{len(worst)} statistics -> {dst}")
