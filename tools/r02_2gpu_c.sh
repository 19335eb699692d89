#!/bin/bash
set -u
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests -m gpu -q -k "group" ) > gpurun_out/pytest_2gpu_c.log 2>&1
tail -4 gpurun_out/pytest_2gpu_c.log
