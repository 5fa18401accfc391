#!/bin/bash
# round-2 GPU session 18: grid-stride rmsnorm with next-row prefetch, merge_partials with four votes of loads in flight
set -x
cd "$GRAFT_REPO_ROOT"
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r02w_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02w_pytest.log
timeout 600 python bench.py --skip-secondary --skip-cpu-baseline > gpurun_out/r02w_bench_n1.json 2> gpurun_out/r02w_bench_n1.err; echo "bench rc=$?"
timeout 600 python bench.py --query-tokens 4 --steps 20 --skip-secondary --skip-cpu-baseline > gpurun_out/r02w_bench_n8shape.json 2> gpurun_out/r02w_bench_n8shape.err; echo "rc=$?"
python - <<'PY'
import json
for f in ["gpurun_out/r02w_bench_n1.json", "gpurun_out/r02w_bench_n8shape.json"]:
    try:
        j = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(j["value"]), j["ms_per_step"], j["clocks"], json.dumps(j["phases_ms_per_step"]), j["parity_sample"].get("ids_equal"))
    except Exception as e:
        print(f, "ERR", e)
PY
