#!/bin/bash
# round 2, GPU call 3: position-loop V2 (branch-free rotations / ropes, unrolled contact points) against V1 on random and evolved
# populations (bit-identity by fitness hash), ncu source-level capture on the evolved population, the fixed tests
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out
rm -f $O/parity_measured.jsonl
{
for r in 1 2; do for b in posv1 posv2; do VSR_LIB=build_ab/$b/libvsr_eval.so python tools/evolved_probe.py time "build=$b"; done; done
for m in 115 99 627; do VSR_GENERIC_SCAFF=1 VSR_GENERIC_SENSORS=1 VSR_SYNC_MASK=$m python tools/evolved_probe.py time "generic mask=$m"; done
VSR_LAUNCH_SHAPE=1 python tools/evolved_probe.py time "shape=1 (5x2)"
} > $O/c3_evolved.log 2>&1
cut -c1-220 $O/c3_evolved.log
python -m pytest tests -m gpu -q --tb=short > $O/c3_pytest.log 2>&1; echo "pytest rc $?" >> $O/c3_pytest.log
tail -15 $O/c3_pytest.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:vsr_episode -s 5 -c 1 -f -o $O/r2_evolved python tools/evolved_probe.py time ncu > $O/c3_ncu.log 2>&1
tail -3 $O/c3_ncu.log; ls -la $O/*.ncu-rep
