#!/bin/bash
# round-2 GPU session 23 (2 GPUs): final code — the 2-GPU tests and the default bench line at N = 2
set -x
cd "$GRAFT_REPO_ROOT"
timeout 600 python -m pytest tests/test_ivf_gpu.py tests/test_pipeline_gpu.py -m gpu -q -k "distributed_build or peer or pipeline" > gpurun_out/r02ab_pytest_2gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02ab_pytest_2gpu.log
tail -3 gpurun_out/r02ab_pytest_2gpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 > gpurun_out/r02ab_bench_n2.json 2> gpurun_out/r02ab_bench_n2.err; echo "rc=$?"
tail -3 gpurun_out/r02ab_bench_n2.err
python - <<'PY'
import json
j = json.loads(open("gpurun_out/r02ab_bench_n2.json").read().strip().splitlines()[-1])
print(round(j["value"]), j["ms_per_step"], j["e2e"]["value"], j["clocks"], json.dumps(j["phases_ms_per_step"]), j["parity_sample"])
print(json.dumps(j["secondary"])[:1500])
PY
