#!/bin/bash
# round-2 GPU session 19: ncu of the grid-stride rmsnorm and the merge kernels
set -x
cd "$GRAFT_REPO_ROOT"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:rmsnorm_kernel -s 20 -c 2 -f -o gpurun_out/r02x_rms python bench.py --workload encode --seq-len 32 --encode-batch 512 --steps 1 --warmup 1 > gpurun_out/r02x_ncu1.log 2>&1
tail -3 gpurun_out/r02x_ncu1.log
ls -la gpurun_out | tail -4
