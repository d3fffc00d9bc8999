#!/bin/bash
# round 2, GPU call 7: same-box A/B of the kernel before / after the per-voxel sensor work, all tests, smoke, latency probe
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out
rm -f $O/parity_measured.jsonl
{
for r in 1 2; do
  VSR_LIB=build_ab/prevk/libvsr_eval.so python tools/evolved_probe.py time "build=prevk"
  python tools/evolved_probe.py time "build=intree"
done
} > $O/c7_evolved.log 2>&1
cut -c1-200 $O/c7_evolved.log
python -m pytest tests -m gpu -q --tb=short > $O/c7_pytest.log 2>&1; echo "pytest rc $?" >> $O/c7_pytest.log
tail -12 $O/c7_pytest.log | cut -c1-250
python __graft_entry__.py smoke 2>&1 | tail -2
tools/latency_probe.sh > $O/c7_latency.txt 2>&1; grep "X=1" $O/c7_latency.txt
