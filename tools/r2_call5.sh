#!/bin/bash
# round 2, GPU call 5: which FP32 approximation ends step-wise tracking of the oracle first, on the V2 kernel:
# r2a = shipped (tanhf + exact sin/cos refresh after the position iterations), r2b = fast tanh, r2c = no refresh, r2d = both off
# (= round-1 arithmetic), r2e = shipped + libm sincosf, r2f = shipped + correctly rounded rcp / rsqrt
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out
for b in r2a r2b r2c r2d r2e r2f; do
  VSR_LIB=build_ab/$b/libvsr_eval.so python tools/approx_ab.py $b
done > $O/c5_approx_ab.jsonl 2> $O/c5_approx_ab.err
cut -c1-400 $O/c5_approx_ab.jsonl
for b in r2a r2d; do VSR_LIB=build_ab/$b/libvsr_eval.so python tools/evolved_probe.py time "build=$b"; done > $O/c5_evolved.log 2>&1
cut -c1-200 $O/c5_evolved.log
