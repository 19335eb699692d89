#!/bin/bash
set -u
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
( time timeout 400 python -m pytest tests/test_host_cli.py tests/test_fit.py -m gpu -q ) > gpurun_out/pytest_gpu31.log 2>&1
tail -6 gpurun_out/pytest_gpu31.log | cut -c1-300
