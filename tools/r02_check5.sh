#!/bin/bash
set -u
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -q -x -k "lane or hmmer or unitindel or counter or group" ) > gpurun_out/pytest_gpu5.log 2>&1
tail -8 gpurun_out/pytest_gpu5.log
timeout 900 python tools/lane_variants.py > gpurun_out/lane_variants.jsonl 2> gpurun_out/lane_variants.err
cat gpurun_out/lane_variants.jsonl; grep "lane engine" gpurun_out/lane_variants.err | sort | uniq -c | head; tail -2 gpurun_out/lane_variants.err
MACHINE=PF00516_protpsw READS=32768 VARIANTS='[{}, {"lane_r": 2}, {"lane_r": 1}]' timeout 600 python tools/lane_variants.py > gpurun_out/lane_variants_c.jsonl 2>> gpurun_out/lane_variants.err
cat gpurun_out/lane_variants_c.jsonl
READS=262144 LEN=275 VARIANTS='[{}, {"lane_r": 2}]' timeout 600 python tools/lane_variants.py > gpurun_out/lane_variants_262k.jsonl 2>> gpurun_out/lane_variants.err
cat gpurun_out/lane_variants_262k.jsonl
