#!/bin/bash
# round-2 GPU session 28: the hypothesis edge-case test of the GPU index against the oracle
set -x
cd "$GRAFT_REPO_ROOT"
timeout 200 python -m pytest tests/test_ivf_gpu.py -m gpu -q -x -k "small_index_fuzz" > gpurun_out/r02ah_pytest_fuzz.log 2>&1; echo "pytest rc=$?"; tail -30 gpurun_out/r02ah_pytest_fuzz.log
