#!/bin/bash
# round 2, GPU call 2: new tests (vsr_eval_ex, locomoting parity, measured bounds), ncu source-level capture of the episode kernel on an
# EVOLVED population (position iterations dominate there), phase-barrier masks on evolved genomes
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out
rm -f $O/parity_measured.jsonl
python -m pytest tests -m gpu -q --tb=short > $O/c2_pytest.log 2>&1; echo "pytest rc $?" >> $O/c2_pytest.log
tail -25 $O/c2_pytest.log
python tools/evolved_probe.py time base > $O/c2_evolved.log 2>&1
{
for m in 115 99 627 611 123; do VSR_GENERIC_SCAFF=1 VSR_GENERIC_SENSORS=1 VSR_SYNC_MASK=$m python tools/evolved_probe.py time "generic mask=$m"; done
} >> $O/c2_evolved.log 2>&1
cat $O/c2_evolved.log | cut -c1-160
timeout 900 ncu --set full --clock-control none --import-source on -k regex:vsr_episode -s 6 -c 1 -f -o $O/r2_evolved python tools/evolved_probe.py time ncu > $O/c2_ncu.log 2>&1
ls -la $O/*.ncu-rep
timeout 300 python bench.py --steps 5 --warmup 3 > $O/c2_bench.json 2> $O/c2_bench.err; echo "bench rc $?"; cut -c1-600 $O/c2_bench.json
