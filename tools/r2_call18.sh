#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out
python -m pytest tests -m gpu -q --tb=short -k "launch_shape" > $O/c18_pytest.log 2>&1; tail -6 $O/c18_pytest.log | cut -c1-250
{
for r in 1 2; do
VSR_NO_WIDE_SHAPE=1 VSR_LIB=build_ab/w11/libvsr_eval.so python tools/cfg_probe.py "w10 n=16384" 3 16384
VSR_LIB=build_ab/w11/libvsr_eval.so python tools/cfg_probe.py "w11 n=16384" 3 16384
VSR_NO_WIDE_SHAPE=1 VSR_LIB=build_ab/w11/libvsr_eval.so python tools/cfg_probe.py "w10 n=65536" 3 65536
VSR_LIB=build_ab/w11/libvsr_eval.so python tools/cfg_probe.py "w11 n=65536" 3 65536
done
} > $O/c18_w11.log 2>&1
cut -c1-300 $O/c18_w11.log
