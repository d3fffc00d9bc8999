#!/bin/bash
# round 2: evidence for profiles/ -- launch list of the default bench command, full ncu capture of the episode kernel on config 2
# (random genomes), the same on configs 3 / 4, latency probe
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out
python bench.py --steps 5 --warmup 3 > $O/f_bench.json 2> $O/f_bench.err; echo "bench rc $?"
python bench.py --impl reference --steps 2 --warmup 1 > $O/f_bench_ref.json 2> $O/f_bench_ref.err; echo "ref rc $?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-parity --no-other-configs > $O/f_ncu_list.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:vsr_episode -s 3 -c 1 -f -o $O/r2_prof_a python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-parity --no-other-configs --ga-generations 0 > $O/f_ncu_a.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:vsr_episode -s 3 -c 1 -f -o $O/r2_prof_cfg3 python bench.py --workload 3 --robots 16384 --steps 1 --warmup 3 --no-cpu-baseline --no-parity --no-other-configs --ga-generations 0 > $O/f_ncu_3.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:vsr_episode -s 3 -c 1 -f -o $O/r2_prof_cfg4 python bench.py --workload 4 --robots 16384 --steps 1 --warmup 3 --no-cpu-baseline --no-parity --no-other-configs --ga-generations 0 > $O/f_ncu_4.log 2>&1
ls -la $O/*.ncu-rep
