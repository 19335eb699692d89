#!/bin/bash
set -u
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu6.log 2>&1
tail -8 gpurun_out/pytest_gpu6.log
timeout 600 python bench.py > gpurun_out/bench6.json 2> gpurun_out/bench6.err
tail -c 1500 gpurun_out/bench6.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches6.csv python bench.py --pairs 10000 --steps 1 --warmup 1 --no-cpu-baseline --em-pairs 256 > gpurun_out/ncu6_run.log 2>&1
timeout 200 python tools/e2e_breakdown.py 2>&1 | tail -3 | tee gpurun_out/e2e_breakdown6.log
