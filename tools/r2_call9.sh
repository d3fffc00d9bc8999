#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out
{
for r in 1 2; do for b in head novote poscache ropeskip mask83 mask51 mask243 mask627; do VSR_LIB=build_ab/$b/libvsr_eval.so python tools/evolved_probe.py time "build=$b"; done; done
} > $O/c9_evolved.log 2>&1
cut -c1-150 $O/c9_evolved.log
