#!/bin/bash
set -x
cd "$GRAFT_REPO_ROOT"
timeout 900 python -m pytest tests/test_encoder_gpu.py -m gpu -q -x -k "attention" > gpurun_out/r02m_pytest_attn.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02m_pytest_attn.log
tail -5 gpurun_out/r02m_pytest_attn.log
timeout 600 python -m pytest tests/test_ivf_gpu.py -m gpu -q -x -k "tune or flat" > gpurun_out/r02m_pytest_flat.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02m_pytest_flat.log
tail -3 gpurun_out/r02m_pytest_flat.log
timeout 600 python bench.py --workload encode --seq-len 512 --skip-cpu-baseline > gpurun_out/r02m_encode_s512.json 2> gpurun_out/r02m_encode_s512.err; echo "rc=$?"
timeout 600 python bench.py --workload encode --seq-len 384 --skip-cpu-baseline > gpurun_out/r02m_encode_s384.json 2> gpurun_out/r02m_encode_s384.err; echo "rc=$?"
timeout 600 python bench.py --workload encode --seq-len 256 --skip-cpu-baseline > gpurun_out/r02m_encode_s256.json 2> gpurun_out/r02m_encode_s256.err; echo "rc=$?"
