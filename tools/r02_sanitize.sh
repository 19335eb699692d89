#!/bin/bash
set -u
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
which compute-sanitizer || ls /usr/local/cuda/bin | grep -i sanit
( time timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 86 --print-limit 20 python -m pytest tests -m gpu -q -x -k "synthetic_profile or strips_fitted or strip_widths or split_mode or impossible_reads or test_golden and dnapsw_small" ) > gpurun_out/sanitize.log 2>&1
echo "exit $?"
grep -c "Invalid\|out of bounds\|misaligned" gpurun_out/sanitize.log
grep -m 12 -A6 "Invalid\|misaligned" gpurun_out/sanitize.log | cut -c1-200
tail -8 gpurun_out/sanitize.log | cut -c1-200
P=2000 L=1000 VARIANTS='[{}]' timeout 300 python tools/jit_variants.py 2>/dev/null | cut -c1-300
