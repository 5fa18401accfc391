#!/bin/bash
# round-2 GPU session 20: split-K of the residual-add GEMMs — parity tests, micro-benchmark, 8-GPU per-GPU shape A/B
set -x
cd "$GRAFT_REPO_ROOT"
timeout 600 python -m pytest tests/test_encoder_gpu.py -m gpu -q -x -k "gemm" > gpurun_out/r02y_pytest_gemm.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r02y_pytest_gemm.log
timeout 300 python tools/gemm_bench.py --split-k > gpurun_out/r02y_gemm_split_k.log 2>&1; echo "rc=$?"; cat gpurun_out/r02y_gemm_split_k.log
for ks in 1 0; do
timeout 600 python bench.py --query-tokens 4 --steps 20 --skip-secondary --skip-cpu-baseline --parity-queries 0 --gemm-ksplit $ks > gpurun_out/r02y_bench_n8shape_ks$ks.json 2> gpurun_out/r02y_bench_n8shape_ks$ks.err; echo "rc=$?"
done
python - <<'PY'
import json
for f in ["gpurun_out/r02y_bench_n8shape_ks1.json", "gpurun_out/r02y_bench_n8shape_ks0.json"]:
    try:
        j = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(j["value"]), j["ms_per_step"], j["clocks"]["sm_mhz"], json.dumps(j["phases_ms_per_step"]))
    except Exception as e:
        print(f, "ERR", e)
PY
