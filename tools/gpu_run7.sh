#!/bin/bash
set -x
cd "$GRAFT_REPO_ROOT"
python -m pytest tests/test_encoder_gpu.py tests/test_pipeline_gpu.py -m gpu -q -x > gpurun_out/r02i_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02i_pytest.log
tail -3 gpurun_out/r02i_pytest.log
COMMON="--skip-secondary --skip-cpu-baseline --parity-queries 0"
python bench.py $COMMON > gpurun_out/r02i_n1.json 2> gpurun_out/r02i_n1.err; echo "rc=$?"
python bench.py $COMMON --query-tokens 4 --steps 20 > gpurun_out/r02i_n8shape.json 2> gpurun_out/r02i_n8shape.err; echo "rc=$?"
python bench.py --workload encode --skip-cpu-baseline > gpurun_out/r02i_encode_s256.json 2> gpurun_out/r02i_encode.err; echo "rc=$?"
python tools/gemm_bench.py 16384 > gpurun_out/r02i_gemm_16384.log 2>&1
python tools/gemm_bench.py 2048 > gpurun_out/r02i_gemm_2048.log 2>&1
