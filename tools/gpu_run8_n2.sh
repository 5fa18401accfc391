#!/bin/bash
set -x
cd "$GRAFT_REPO_ROOT"
python -m pytest tests/test_ivf_gpu.py -m gpu -q -k "distributed_build or peer" > gpurun_out/r02j_pytest_2gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02j_pytest_2gpu.log
tail -3 gpurun_out/r02j_pytest_2gpu.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 > gpurun_out/r02j_bench_n2.json 2> gpurun_out/r02j_bench_n2.err; echo "rc=$?"
tail -5 gpurun_out/r02j_bench_n2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --exchange nccl --skip-secondary --parity-queries 8 > gpurun_out/r02j_bench_n2_nccl.json 2> gpurun_out/r02j_bench_n2_nccl.err; echo "rc=$?"
