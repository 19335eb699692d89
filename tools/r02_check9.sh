#!/bin/bash
set -u
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
READS=8192 VARIANTS='[{"no_col": 1}, {}]' timeout 600 python tools/lane_variants.py > gpurun_out/col_variants_8k.jsonl 2> gpurun_out/col_variants.err
cat gpurun_out/col_variants_8k.jsonl; grep -i "column engine" gpurun_out/col_variants.err | sort | uniq -c | head
( time timeout 900 python -m pytest tests -m gpu -q -x --durations=5 -k "lane or hmmer or profile or cfg5 or config5" ) > gpurun_out/pytest_gpu9.log 2>&1
tail -12 gpurun_out/pytest_gpu9.log
READS=131072 VARIANTS='[{"no_col": 1}, {}, {"col_r": 2}, {"col_threads": 128}, {"col_minblocks": 3}, {"col_threads": 512, "col_minblocks": 1}]' timeout 900 python tools/lane_variants.py > gpurun_out/col_variants_131k.jsonl 2>> gpurun_out/col_variants.err
cat gpurun_out/col_variants_131k.jsonl
MACHINE=PF00516_protpsw READS=8192 VARIANTS='[{"no_col": 1}, {}, {"col_sil_regs": 100}]' timeout 900 python tools/lane_variants.py > gpurun_out/col_variants_comp.jsonl 2>> gpurun_out/col_variants.err
cat gpurun_out/col_variants_comp.jsonl
grep -i "column engine" gpurun_out/col_variants.err | sort | uniq -c | head -20
