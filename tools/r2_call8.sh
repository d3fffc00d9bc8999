#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out
rm -f $O/parity_measured.jsonl
python -m pytest tests -m gpu -q --tb=short > $O/c8_pytest.log 2>&1; echo "pytest rc $?" >> $O/c8_pytest.log
tail -14 $O/c8_pytest.log | cut -c1-250
python __graft_entry__.py smoke 2>&1 | tail -1
