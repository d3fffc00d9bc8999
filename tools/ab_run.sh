#!/bin/bash
# usage (on the GPU box): tools/ab_run.sh name1 name2 ...  -> bench_short for every build_ab/<name>, two rounds interleaved
for round in 1 2; do
  for n in "$@"; do
    tools/bench_short.sh "VSR_LIB=build_ab/$n/libvsr_eval.so" | sed "s|VSR_LIB=build_ab/||; s|/libvsr_eval.so||"
  done
done
