#!/bin/bash
# round-2 GPU session 6: full suite, default bench line, N=8-per-GPU-shape line, ncu launch list + full captures
set -x
cd "$GRAFT_REPO_ROOT"
python -m pytest tests -m gpu -q > gpurun_out/r02g_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02g_pytest.log
tail -4 gpurun_out/r02g_pytest.log
python bench.py > gpurun_out/r02g_bench_n1.json 2> gpurun_out/r02g_bench_n1.err; echo "bench rc=$?"; tail -3 gpurun_out/r02g_bench_n1.err
python bench.py --query-tokens 4 --steps 20 --skip-secondary --skip-cpu-baseline > gpurun_out/r02g_bench_n8shape.json 2> gpurun_out/r02g_bench_n8shape.err; echo "rc=$?"
python bench.py --corpus lattice --skip-secondary --skip-cpu-baseline > gpurun_out/r02g_bench_n1_lattice.json 2> gpurun_out/r02g_bench_n1_lattice.err; echo "rc=$?"
NCUARGS="--skip-secondary --parity-queries 0 --skip-cpu-baseline --skip-e2e --no-kernel-events"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file gpurun_out/r02g_launches_raw.csv python bench.py --steps 2 --warmup 1 $NCUARGS > gpurun_out/r02g_ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ivf_scan_ring_kernel -s 4 -c 1 -f -o gpurun_out/r02g_scan_fp16 python bench.py --steps 1 --warmup 1 $NCUARGS > gpurun_out/r02g_ncu1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ivf_scan_ring_kernel -s 2 -c 1 -f -o gpurun_out/r02g_scan_fp32 python bench.py --steps 1 --warmup 1 --two-stage 0 $NCUARGS > gpurun_out/r02g_ncu2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_tc_kernel -s 500 -c 4 -f -o gpurun_out/r02g_gemm python bench.py --steps 1 --warmup 1 $NCUARGS > gpurun_out/r02g_ncu3.log 2>&1
ls -la gpurun_out/ | tail -12
