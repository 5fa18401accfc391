#!/bin/bash
set -x
cd "$GRAFT_REPO_ROOT"
timeout 300 python -m pytest tests/test_encoder_gpu.py -m gpu -q -x -k "tcgen05_gemm_matches_torch and (128-256-64 or 300-512-192)" > gpurun_out/r02t_pytest_quad_small.log 2>&1; echo "rc=$?" >> gpurun_out/r02t_pytest_quad_small.log
tail -4 gpurun_out/r02t_pytest_quad_small.log
timeout 600 python -m pytest tests/test_encoder_gpu.py -m gpu -q -x -k "tcgen05_gemm or fused_epilogues" > gpurun_out/r02t_pytest_quad.log 2>&1; echo "rc=$?" >> gpurun_out/r02t_pytest_quad.log
tail -4 gpurun_out/r02t_pytest_quad.log
timeout 300 python tools/gemm_bench.py 2048 > gpurun_out/r02t_gemm_2048.log 2>&1
timeout 300 python tools/gemm_bench.py 16384 > gpurun_out/r02t_gemm_16384.log 2>&1
