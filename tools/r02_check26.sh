#!/bin/bash
set -u
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mb_k_big_forward -c 1 -o gpurun_out/prof_mb_k_big_forward -f \
    python tools/bench_wide.py --machine prot2dna_dnapsw --pairs 1184 --li 300 --lo 1000 --engines 2 --no-trace --reps 1 > gpurun_out/ncu_big_run.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mb_k_big_viterbi$ -c 1 -o gpurun_out/prof_mb_k_big_viterbi -f \
    python tools/bench_wide.py --machine prot2dna_dnapsw --pairs 592 --li 300 --lo 1000 --engines 2 --reps 1 > gpurun_out/ncu_bigv_run.log 2>&1
ls -la gpurun_out/prof_mb_k_big*
