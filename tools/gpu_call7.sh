#!/bin/bash
set -u
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu7.log 2>&1
tail -8 gpurun_out/pytest_gpu7.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench7.json 2> gpurun_out/bench7.err
tail -c 1500 gpurun_out/bench7.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches7.csv python bench.py --pairs 2048 --steps 1 --warmup 1 --no-cpu-baseline --em-pairs 2048 > gpurun_out/ncu7_run.log 2>&1
timeout 200 python tools/other_configs.py > gpurun_out/other_configs7.json 2> gpurun_out/other_configs7.err; cat gpurun_out/other_configs7.json
