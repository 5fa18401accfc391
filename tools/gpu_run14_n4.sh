#!/bin/bash
set -x
cd "$GRAFT_REPO_ROOT"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 4 > gpurun_out/r02r_bench_n4.json 2> gpurun_out/r02r_bench_n4.err; echo "rc=$?"
tail -5 gpurun_out/r02r_bench_n4.err
