#!/bin/bash
# round-2 GPU session 3: full GPU suite, pipelined bench at the N=1 shape and at the N=8 per-GPU shape (4-token queries)
set -x
cd "$GRAFT_REPO_ROOT"
python -m pytest tests -m gpu -q > gpurun_out/r02c_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02c_pytest.log
tail -8 gpurun_out/r02c_pytest.log
COMMON="--skip-secondary --skip-cpu-baseline --parity-queries 0"
# N=8 per-GPU load emulated on one GPU: 512 queries x 4 tokens = 2048 encoder tokens per step
for mode in "--pipeline 0" "--pipeline 1" "--pipeline 1 --no-coresident"; do
  tag=$(echo $mode | tr -d ' -')
  python bench.py $COMMON --query-tokens 4 --steps 20 $mode > gpurun_out/r02c_n8shape_$tag.json 2> gpurun_out/r02c_n8shape_$tag.err; echo "rc=$?"
  python -c "import json,sys; j=json.load(open('gpurun_out/r02c_n8shape_$tag.json')); print('$tag', j['value'], j['ms_per_step'], j['e2e'] and j['e2e']['value'], j['config'].get('ms_per_step_serial_same_run'), j['phases_ms_per_step'])"
  tail -2 gpurun_out/r02c_n8shape_$tag.err
done
for mode in "--pipeline 0" "--pipeline 1"; do
  tag=$(echo $mode | tr -d ' -')
  python bench.py $COMMON $mode > gpurun_out/r02c_n1_$tag.json 2> gpurun_out/r02c_n1_$tag.err; echo "rc=$?"
  python -c "import json,sys; j=json.load(open('gpurun_out/r02c_n1_$tag.json')); print('$tag', j['value'], j['ms_per_step'], j['e2e'] and j['e2e']['value'], j['config'].get('ms_per_step_serial_same_run'), j['phases_ms_per_step'])"
  tail -2 gpurun_out/r02c_n1_$tag.err
done
