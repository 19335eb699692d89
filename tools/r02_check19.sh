#!/bin/bash
set -u
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q -x -k "column or synthetic or hmmer or lane" ) > gpurun_out/pytest_gpu19.log 2>&1
tail -6 gpurun_out/pytest_gpu19.log
( time timeout 1200 python bench.py --no-cpu-baseline ) > gpurun_out/bench19.json 2> gpurun_out/bench19.err
tail -c 300 gpurun_out/bench19.json; tail -3 gpurun_out/bench19.err
