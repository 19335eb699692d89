#!/bin/bash
set -u
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu10.log 2>&1
tail -12 gpurun_out/pytest_gpu10.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
