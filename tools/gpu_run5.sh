#!/bin/bash
set -x
cd "$GRAFT_REPO_ROOT"
python -m pytest tests/test_encoder_gpu.py tests/test_pipeline_gpu.py -m gpu -q -x > gpurun_out/r02f_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02f_pytest.log
tail -3 gpurun_out/r02f_pytest.log
COMMON="--skip-secondary --skip-cpu-baseline --parity-queries 0"
for pdl in 1 0; do
 for mode in "--pipeline 0" "--pipeline 1"; do
  tag=$(echo $mode | tr -d ' -')_pdl$pdl
  ABSB_PDL=$pdl python bench.py $COMMON --query-tokens 4 --steps 20 $mode > gpurun_out/r02f_n8shape_$tag.json 2> gpurun_out/r02f_n8shape_$tag.err; echo "rc=$?"
  tail -2 gpurun_out/r02f_n8shape_$tag.err
 done
done
for mode in "--pipeline 0" "--pipeline 1"; do
  tag=$(echo $mode | tr -d ' -')
  python bench.py $COMMON $mode > gpurun_out/r02f_n1_$tag.json 2> gpurun_out/r02f_n1_$tag.err; echo "rc=$?"
  tail -2 gpurun_out/r02f_n1_$tag.err
done
ABSB_PDL=0 python bench.py $COMMON --pipeline 0 > gpurun_out/r02f_n1_pipeline0_pdl0.json 2> gpurun_out/r02f_n1_pipeline0_pdl0.err
