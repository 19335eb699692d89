#!/bin/bash
set -u
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests -m gpu -q -k "group" ) > gpurun_out/pytest_2gpu_b.log 2>&1
tail -3 gpurun_out/pytest_2gpu_b.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --steps 5 --warmup 3 --cfg3-pairs 50000 --cfg4-pairs 500 --cfg5-reads 65536 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err
tail -c 300 gpurun_out/bench_2gpu.json; tail -2 gpurun_out/bench_2gpu.err
