#!/bin/bash
# run under gpurun on ONE GPU: throughput of the other BASELINE configs' machines (parity is in tests/), one JSON line each
set -u
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
out=gpurun_out/other_configs_all.jsonl
: > $out
timeout 300 python tools/other_configs.py 2>/dev/null | head -1 >> $out
MB_WIDE_VERBOSE=1 timeout 300 python tools/bench_wide.py --machine prot2dna_dnapsw --pairs 592 --li 300 --lo 2000 --engines 2 --reps 2 2>/dev/null >> $out
timeout 300 python tools/bench_wide.py --machine prot2dna_dnapsw --pairs 1000 --li 300 --lo 10000 --engines 2 --no-trace --reps 2 2>/dev/null >> $out
timeout 300 python tools/bench_wide.py --machine hmmer_pf00516 --pairs 262144 --li 0 --lo 275 --engines 2 --no-trace --reps 2 2>/dev/null >> $out
timeout 300 python tools/bench_wide.py --machine hmmer_pf00516_protpsw --pairs 65536 --li 0 --lo 275 --engines 2 --no-trace --reps 2 2>/dev/null >> $out
cat $out | cut -c1-400
