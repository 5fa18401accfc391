#!/bin/bash
# round-2 GPU session 22: ncu source-level capture of the O-proj (residual-add epilogue) GEMM
set -x
cd "$GRAFT_REPO_ROOT"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_tc_kernel -s 3 -c 1 -f -o gpurun_out/r02aa_oproj_m16384 python tools/gemm_one.py 2 16384 1536 1536 > gpurun_out/r02aa_ncu1.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_tc_kernel -s 3 -c 1 -f -o gpurun_out/r02aa_oproj_m2048 python tools/gemm_one.py 2 2048 1536 1536 > gpurun_out/r02aa_ncu2.log 2>&1
ls -la gpurun_out | tail -5
