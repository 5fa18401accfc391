#!/bin/bash
set -x
cd "$GRAFT_REPO_ROOT"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 > gpurun_out/r02k_bench_n8.json 2> gpurun_out/r02k_bench_n8.err; echo "rc=$?"
tail -5 gpurun_out/r02k_bench_n8.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 --impl reference --steps 2 --warmup 1 > gpurun_out/r02k_ref_n8.json 2> gpurun_out/r02k_ref_n8.err; echo "rc=$?"
