#!/bin/bash
# full validation of the current build: GPU test suite, smoke, both bench arms
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out
python -m pytest tests -m gpu -x -q > $O/c19_pytest.log 2>&1; echo "pytest rc $?"; tail -4 $O/c19_pytest.log | cut -c1-200
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/c19_smoke.log 2>&1; echo "smoke rc $?"; tail -2 $O/c19_smoke.log | cut -c1-300
python bench.py --impl reference --steps 2 --warmup 1 > $O/c19_bench_ref.json 2> $O/c19_bench_ref.err; echo "ref rc $?"
python bench.py > $O/c19_bench.json 2> $O/c19_bench.err; echo "bench rc $?"; cut -c1-600 $O/c19_bench.json
