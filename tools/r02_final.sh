#!/bin/bash
# the round's closing run on ONE GPU: the whole -m gpu suite, smoke(), then tools/profile_round.sh
set -u
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
( time timeout 1800 python -m pytest tests -m gpu -q --durations=6 ) > gpurun_out/pytest_gpu_final.log 2>&1
tail -12 gpurun_out/pytest_gpu_final.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
bash tools/profile_round.sh
