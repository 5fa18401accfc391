#!/bin/bash
set -x
cd "$GRAFT_REPO_ROOT"
timeout 1500 python -m pytest tests -m gpu -q -s -k "full_size or attention or flat or tune or pipeline" > gpurun_out/r02o_pytest_sel.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02o_pytest_sel.log
grep -E "cosine deficit|passed|failed|rc=" gpurun_out/r02o_pytest_sel.log | tail -5
timeout 600 python bench.py --skip-secondary --skip-cpu-baseline --parity-queries 0 --query-tokens 4 --steps 20 > gpurun_out/r02o_n8shape.json 2> gpurun_out/r02o_n8shape.err; echo "rc=$?"
timeout 600 python bench.py --skip-secondary --skip-cpu-baseline --parity-queries 0 > gpurun_out/r02o_n1.json 2> gpurun_out/r02o_n1.err; echo "rc=$?"
