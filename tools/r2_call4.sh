#!/bin/bash
# round 2, GPU call 4: per-lane rotation refresh (launch-shape independence restored), touching-mass register cache A/B, all tests,
# measured bounds, bench
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out
rm -f $O/parity_measured.jsonl
{
for r in 1 2; do for b in posv2b cmass; do VSR_LIB=build_ab/$b/libvsr_eval.so python tools/evolved_probe.py time "build=$b"; done; done
} > $O/c4_evolved.log 2>&1
cut -c1-200 $O/c4_evolved.log
python -m pytest tests -m gpu -q --tb=short > $O/c4_pytest.log 2>&1; echo "pytest rc $?" >> $O/c4_pytest.log
tail -12 $O/c4_pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 > $O/c4_bench.json 2> $O/c4_bench.err; echo "bench rc $?"; cut -c1-300 $O/c4_bench.json
tools/ab_cfg.sh 3 65536 base 2>&1 | tail -2
tools/ab_cfg.sh 4 16384 base 2>&1 | tail -2
