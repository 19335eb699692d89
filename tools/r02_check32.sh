#!/bin/bash
set -u
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
( time timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "kernel_cache" ) > gpurun_out/pytest_gpu32.log 2>&1
tail -12 gpurun_out/pytest_gpu32.log | cut -c1-300
mkdir -p gpurun_out/kc
for rep in 1 2 3; do ( time ./machineboss_b200/boss_b200 --preset dnapsw --input-chars ACGTACGTTG --output-chars ACGTTCGTG -L --kernel-cache gpurun_out/kc ) 2>&1 | grep -E "real|\[" | tr '\n' ' '; echo; done
rm -rf gpurun_out/kc
