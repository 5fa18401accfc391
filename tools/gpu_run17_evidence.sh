#!/bin/bash
# round-2 GPU session 17: final-code evidence — ncu launch lists (default N=1 workload and the 8-GPU per-GPU shape),
# ncu --set full of the GEMMs, the two attention kernels and rmsnorm
set -x
cd "$GRAFT_REPO_ROOT"
NCUARGS="--skip-secondary --parity-queries 0 --skip-cpu-baseline --skip-e2e --no-kernel-events"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file gpurun_out/r02v_launches_raw.csv python bench.py --steps 2 --warmup 1 $NCUARGS > gpurun_out/r02v_ncu_launches.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file gpurun_out/r02v_launches_n8shape_raw.csv python bench.py --steps 2 --warmup 1 --query-tokens 4 $NCUARGS > gpurun_out/r02v_ncu_launches_n8shape.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_tc_kernel -s 500 -c 4 -f -o gpurun_out/r02v_gemm python bench.py --steps 1 --warmup 1 $NCUARGS > gpurun_out/r02v_ncu3.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_tc_kernel -s 500 -c 4 -f -o gpurun_out/r02v_gemm_n8shape python bench.py --steps 1 --warmup 1 --query-tokens 4 $NCUARGS > gpurun_out/r02v_ncu4.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"attention_tc_persistent|rmsnorm_kernel" -s 100 -c 3 -f -o gpurun_out/r02v_attn_rms python bench.py --steps 1 --warmup 1 $NCUARGS > gpurun_out/r02v_ncu5.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_tc_long -s 30 -c 1 -f -o gpurun_out/r02v_attn_long python bench.py --workload encode --seq-len 512 --steps 1 --warmup 1 > gpurun_out/r02v_ncu6.log 2>&1
ls -la gpurun_out/ | tail -14
