#!/bin/bash
# regenerate tests/parity_bounds.json from scratch on the GPU box (loose defaults -> measure -> 3x measured), then re-run against it
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out
rm -f $O/parity_measured.jsonl tests/parity_bounds.json
python -m pytest tests -m gpu -q --tb=short > $O/bounds_pass1.log 2>&1; echo "pass1 rc $?"; tail -2 $O/bounds_pass1.log
python tests/tools/update_bounds.py
cp tests/parity_bounds.json $O/parity_bounds.json
cp $O/parity_measured.jsonl $O/parity_measured_pass1.jsonl
python -m pytest tests -m gpu -q --tb=short > $O/bounds_pass2.log 2>&1; echo "pass2 rc $?"; tail -4 $O/bounds_pass2.log
