#!/bin/bash
set -u
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
./tools/pipe_peaks > gpurun_out/pipe_peaks.json 2> gpurun_out/pipe_peaks.err; tail -15 gpurun_out/pipe_peaks.json
( timeout 1200 python -m pytest tests -m gpu -q -x ) > gpurun_out/pytest_gpu4.log 2>&1
tail -8 gpurun_out/pytest_gpu4.log
VARIANTS='[{}, {"jit_unroll": 1}, {"jit_minblocks_linv": 3}, {"jit_minblocks_v": 4}]' timeout 600 python tools/jit_variants.py > gpurun_out/jit_variants2.jsonl 2> gpurun_out/jit_variants2.err
cat gpurun_out/jit_variants2.jsonl; tail -3 gpurun_out/jit_variants2.err
timeout 900 python bench.py > gpurun_out/bench4.json 2> gpurun_out/bench4.err
tail -c 6000 gpurun_out/bench4.json; tail -5 gpurun_out/bench4.err
