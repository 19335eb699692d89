#!/bin/bash
set -u
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
( time timeout 100 python -m pytest tests/test_host_cli.py tests/test_fit.py -m gpu -q -x -k "loglike_viterbi_counts or fit_bitnoise or align_matches" ) > gpurun_out/pytest_2gpu_d.log 2>&1
tail -4 gpurun_out/pytest_2gpu_d.log | cut -c1-200
