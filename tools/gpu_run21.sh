#!/bin/bash
# round-2 GPU session 21: residual-add epilogue as loads/adds/stores (ABSB_GEMM_RESIDUAL=ldst) against bulk tensor reductions
set -x
cd "$GRAFT_REPO_ROOT"
ABSB_GEMM_RESIDUAL=ldst timeout 600 python -m pytest tests/test_encoder_gpu.py -m gpu -q -x -k "gemm or tiny_encoder" > gpurun_out/r02z_pytest_gemm_ldst.log 2>&1; echo "pytest ldst rc=$?"; tail -3 gpurun_out/r02z_pytest_gemm_ldst.log
timeout 600 python -m pytest tests/test_encoder_gpu.py -m gpu -q -x -k "gemm" > gpurun_out/r02z_pytest_gemm.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02z_pytest_gemm.log
for T in 16384 2048; do
for mode in tma ldst tma ldst; do
echo "== T=$T residual=$mode"
ABSB_GEMM_RESIDUAL=$mode timeout 300 python tools/gemm_bench.py $T --epi-only 2>&1 | grep "f32+="
done
done
for mode in tma ldst; do
ABSB_GEMM_RESIDUAL=$mode timeout 600 python bench.py --query-tokens 4 --steps 20 --skip-secondary --skip-cpu-baseline --parity-queries 0 > gpurun_out/r02z_bench_n8shape_$mode.json 2> gpurun_out/r02z_bench_n8shape_$mode.err; echo "rc=$?"
ABSB_GEMM_RESIDUAL=$mode timeout 600 python bench.py --skip-secondary --skip-cpu-baseline --parity-queries 0 > gpurun_out/r02z_bench_n1_$mode.json 2> gpurun_out/r02z_bench_n1_$mode.err; echo "rc=$?"
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02z_bench_*.json")):
    try:
        j = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(j["value"]), j["ms_per_step"], j["clocks"]["sm_mhz"], json.dumps(j["phases_ms_per_step"]))
    except Exception as e:
        print(f, "ERR", e)
PY
