#!/bin/bash
# usage: tools/ab_build.sh <name> ["-DFLAG ..."]  -> build_ab/<name>/libvsr_eval.so (objects in build_ab/<name>/obj)
# run with VSR_LIB=build_ab/<name>/libvsr_eval.so (A/B inside one gpurun call; box-to-box noise is ~2%)
set -e
root=$(cd "$(dirname "$0")/.." && pwd)
mkdir -p "$root/build_ab/$1"
make -s -j8 -C "$root/vsrevolution_b200/csrc" OUT="$root/build_ab/$1/libvsr_eval.so" OBJ="$root/build_ab/$1/obj" EXTRA="$2" 2>&1 | grep -E "error|Li1ELi5ELi2ELi15E" -A3 | grep -E "error|spill|Used" || true
ls -la "$root/build_ab/$1/libvsr_eval.so"
