#!/bin/bash
set -u
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
( time timeout 300 python -m pytest tests/test_host_cli.py -m gpu -q -k "downsample or post_trans" ) > gpurun_out/pytest_gpu30.log 2>&1
tail -25 gpurun_out/pytest_gpu30.log | cut -c1-300
