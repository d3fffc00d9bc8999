#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out
{
for r in 1 2; do for b in ${VARIANTS:-head hmap}; do VSR_LIB=build_ab/$b/libvsr_eval.so python tools/cfg_probe.py "build=$b" ${WLS:-2,3,4}; done; done
} > $O/c16_cfg.log 2>&1
cut -c1-400 $O/c16_cfg.log
