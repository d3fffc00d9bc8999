#!/bin/bash
# round 2, GPU call 1: green baseline after the ADVICE fixes, new bench blocks, locomoting-robot parity report, A/B of the FP32
# approximations and of the general narrowphase out of line, launch shapes / iteration pricing on random vs evolved populations
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $O/c1_smi.txt
python -m pytest tests -m gpu -x -q > $O/c1_pytest.log 2>&1; echo "pytest rc $?" >> $O/c1_pytest.log
python bench.py --steps 5 --warmup 3 > $O/c1_bench.json 2> $O/c1_bench.err; echo "bench rc $?" >> $O/c1_bench.err
python tools/evolved_probe.py make 20 > $O/c1_evolved.log 2>&1
{
for b in base noinline ieee libmsc norot libmtanh; do
  VSR_LIB=build_ab/$b/libvsr_eval.so python tools/evolved_probe.py time "build=$b"
done
for sh in 0 1 2 3; do VSR_LAUNCH_SHAPE=$sh python tools/evolved_probe.py time "shape=$sh"; done
for p in 0 1 2 3; do VSR_POS_ITERS=$p python tools/evolved_probe.py time "pos_iters=$p"; done
VSR_VEL_ITERS=0 VSR_POS_ITERS=0 python tools/evolved_probe.py time "vel=0,pos=0"
VSR_VEL_ITERS=5 python tools/evolved_probe.py time "vel=5"
} >> $O/c1_evolved.log 2>&1
for b in base noinline ieee libmsc norot libmtanh; do
  VSR_LIB=build_ab/$b/libvsr_eval.so python tools/approx_ab.py $b
done > $O/c1_approx_ab.jsonl 2> $O/c1_approx_ab.err
python tests/tools/parity_report.py > $O/c1_parity.log 2>&1; echo "parity rc $?" >> $O/c1_parity.log
tail -3 $O/c1_pytest.log; tail -2 $O/c1_bench.err; cat $O/c1_evolved.log | cut -c1-200; cut -c1-300 $O/c1_approx_ab.jsonl
