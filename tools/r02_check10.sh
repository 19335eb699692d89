#!/bin/bash
set -u
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
READS=131072 VARIANTS='[{"col_c": 1}, {}, {"col_c": 3, "col_minblocks": 1}, {"col_c": 4}, {"col_c": 4, "col_sil_regs": 100, "col_minblocks": 1}, {"col_c": 2, "col_threads": 128, "col_minblocks": 4}, {"col_c": 2, "col_r": 2}]' timeout 900 python tools/lane_variants.py > gpurun_out/col_variants10_131k.jsonl 2> gpurun_out/col_variants10.err
cat gpurun_out/col_variants10_131k.jsonl
MACHINE=PF00516_protpsw READS=32768 VARIANTS='[{}, {"col_c": 2}, {"col_threads": 128}]' timeout 900 python tools/lane_variants.py > gpurun_out/col_variants10_comp.jsonl 2>> gpurun_out/col_variants10.err
cat gpurun_out/col_variants10_comp.jsonl
grep -i "column engine" gpurun_out/col_variants10.err | sort | uniq -c | head -30
( time timeout 600 python -m pytest tests -m gpu -q -x -k "lane or hmmer or profile or cfg5 or config5" ) > gpurun_out/pytest_gpu10.log 2>&1
tail -5 gpurun_out/pytest_gpu10.log
READS=32768 VARIANTS='[{}]' timeout 900 ncu --set full --import-source on --clock-control none -k regex:mb_k_col -c 2 -o gpurun_out/ncu_col -f python tools/lane_variants.py > gpurun_out/ncu_col_run.log 2>&1
tail -3 gpurun_out/ncu_col_run.log
READS=32768 VARIANTS='[{}]' timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_col.csv python tools/lane_variants.py > /dev/null 2>&1
grep -c . gpurun_out/launches_col.csv
