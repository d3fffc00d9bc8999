#!/bin/bash
# A/B of the packed-FP32 (FFMA2) velocity loop: head / f2 (packed spring pairs + weld rounds) / f2off (permuted staging only)
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out
{
for r in 1 2; do for b in ${VARIANTS:-head f2 f2off}; do VSR_LIB=build_ab/$b/libvsr_eval.so python tools/evolved_probe.py time "build=$b"; done; done
} > $O/c12_f2.log 2>&1
cut -c1-230 $O/c12_f2.log
