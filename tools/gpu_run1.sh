#!/bin/bash
# round-2 GPU session 1: full GPU test suite, default bench line (N=1), ncu capture of the fp16 scan
set -x
cd "$GRAFT_REPO_ROOT"
python -m pytest tests -m gpu -x -q > gpurun_out/r02a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02a_pytest.log
tail -5 gpurun_out/r02a_pytest.log
python bench.py > gpurun_out/r02a_bench_n1.json 2> gpurun_out/r02a_bench_n1.err; echo "bench rc=$?"
tail -c 3000 gpurun_out/r02a_bench_n1.json
tail -5 gpurun_out/r02a_bench_n1.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ivf_scan16 -s 2 -c 1 -f -o gpurun_out/r02a_scan16 \
  python bench.py --steps 1 --warmup 1 --skip-secondary --parity-queries 0 --skip-cpu-baseline --skip-e2e --no-kernel-events > gpurun_out/r02a_ncu.log 2>&1
tail -3 gpurun_out/r02a_ncu.log
