#!/bin/bash
set -u
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
timeout 600 python tools/debug_counts.py > gpurun_out/debug_counts.log 2>&1
tail -60 gpurun_out/debug_counts.log
( timeout 900 python -m pytest tests -m gpu -q -k "group or peaked or chunked or fit or cli" ) > gpurun_out/pytest_gpu2.log 2>&1
tail -30 gpurun_out/pytest_gpu2.log
