#!/bin/bash
set -x
cd "$GRAFT_REPO_ROOT"
timeout 900 python -m pytest tests/test_encoder_gpu.py -m gpu -q -x -k "attention" > gpurun_out/r02l_pytest_attn.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02l_pytest_attn.log
tail -5 gpurun_out/r02l_pytest_attn.log
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r02l_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02l_pytest.log
tail -5 gpurun_out/r02l_pytest.log
timeout 600 python bench.py --workload encode --seq-len 512 --skip-cpu-baseline > gpurun_out/r02l_encode_s512.json 2> gpurun_out/r02l_encode_s512.err; echo "rc=$?"
timeout 600 python bench.py --workload encode --seq-len 384 --skip-cpu-baseline > gpurun_out/r02l_encode_s384.json 2> gpurun_out/r02l_encode_s384.err; echo "rc=$?"
timeout 600 python bench.py --skip-secondary --skip-cpu-baseline --parity-queries 0 --query-tokens 4 --steps 20 > gpurun_out/r02l_n8shape.json 2> gpurun_out/r02l_n8shape.err; echo "rc=$?"
