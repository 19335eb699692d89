#!/bin/bash
set -u
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -q -x -k "split or strip or ragged or set_b or long_pairs or golden and dnapsw" ) > gpurun_out/pytest_gpu6.log 2>&1
tail -8 gpurun_out/pytest_gpu6.log
# few pairs: split against no split
P=1250 VARIANTS='[{}, {"jit_split": 0}]' timeout 300 python tools/jit_variants.py > gpurun_out/jit_split_1250.jsonl 2> gpurun_out/jit_split.err
cat gpurun_out/jit_split_1250.jsonl
P=150 L=10000 VARIANTS='[{}, {"jit_split": 0}]' timeout 300 python tools/jit_variants.py > gpurun_out/jit_split_150x10k.jsonl 2>> gpurun_out/jit_split.err
cat gpurun_out/jit_split_150x10k.jsonl
P=10000 VARIANTS='[{}, {"jit_split": 1}]' timeout 300 python tools/jit_variants.py > gpurun_out/jit_split_10k.jsonl 2>> gpurun_out/jit_split.err
cat gpurun_out/jit_split_10k.jsonl; tail -3 gpurun_out/jit_split.err
# the lane sweep under ncu, source-level
READS=16384 LEN=120 VARIANTS='[{}]' timeout 600 ncu --set full --clock-control none --import-source on -k regex:lane2_kernel -c 1 -o gpurun_out/prof_lane2_forward \
   python tools/lane_variants.py > gpurun_out/ncu_lane2_run.log 2>&1
tail -3 gpurun_out/ncu_lane2_run.log
ls -la gpurun_out/*.ncu-rep
