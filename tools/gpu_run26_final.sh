#!/bin/bash
# last check of the round after the FFN-down tile heuristic change: GEMM / encoder tests, smoke(), default bench line
set -x
cd "$GRAFT_REPO_ROOT"
timeout 600 python -m pytest tests/test_encoder_gpu.py tests/test_pipeline_gpu.py -m gpu -q > gpurun_out/r02af_pytest_encoder.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02af_pytest_encoder.log
tail -3 gpurun_out/r02af_pytest_encoder.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02af_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r02af_smoke.log
tail -2 gpurun_out/r02af_smoke.log
timeout 600 python bench.py > gpurun_out/r02af_bench_n1.json 2> gpurun_out/r02af_bench_n1.err; echo "bench rc=$?"
