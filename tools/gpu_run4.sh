#!/bin/bash
set -x
cd "$GRAFT_REPO_ROOT"
python -m pytest tests/test_ivf_gpu.py tests/test_pipeline_gpu.py -m gpu -q -x > gpurun_out/r02e_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02e_pytest.log
tail -3 gpurun_out/r02e_pytest.log
python tools/overlap_timeline.py --budget-kb 162 > gpurun_out/r02e_timeline_evict.txt 2>&1
ABSB_SCAN_L2_DEFAULT=1 python tools/overlap_timeline.py --budget-kb 162 > gpurun_out/r02e_timeline_l2default.txt 2>&1
python tools/scan_sweep.py --random-queries --configs "0,0,0,0,512,0;1,4,3,1,512,0;1,4,3,1,512,1;1,8,3,1,512,1;1,8,2,1,512,1;2,0,0,0,512,1" > gpurun_out/r02e_sweep.jsonl 2> gpurun_out/r02e_sweep.err
