#!/bin/bash
set -u
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
( time timeout 1800 python -m pytest tests -m gpu -q --durations=8 ) > gpurun_out/pytest_gpu12.log 2>&1
tail -14 gpurun_out/pytest_gpu12.log
