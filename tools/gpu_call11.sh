#!/bin/bash
set -u
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu11.log 2>&1
tail -6 gpurun_out/pytest_gpu11.log
em() { timeout 300 python bench.py --no-cpu-baseline --pairs 2048 --steps 2 --warmup 3 --em-pairs 4096 2>>gpurun_out/err_sweep11.log | python -c "import json,sys; j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(j[\"em\"][\"pairs_per_s\"]), round(j[\"em\"][\"kernel_ms\"],2))"; }
( echo "em default"; em; echo "CNT=2"; MB_JIT_MINBLOCKS_CNT=2 em; echo "CNT=4"; MB_JIT_MINBLOCKS_CNT=4 em; echo "C=2 CNT=4"; MB_JIT_C=2 MB_JIT_CV=8 MB_JIT_MINBLOCKS_CNT=4 em; echo "C=2 CNT=5"; MB_JIT_C=2 MB_JIT_CV=8 MB_JIT_MINBLOCKS_CNT=5 em ) 2>&1 | tee gpurun_out/sweep11.log
MB_WIDE_VERBOSE=1 timeout 300 python tools/bench_wide.py --machine prot2dna_dnapsw --pairs 148 --li 300 --lo 10000 --engines 2 --no-trace > gpurun_out/wide11_cfg4.json 2> gpurun_out/wide11_cfg4.err
tail -c 600 gpurun_out/wide11_cfg4.json; tail -1 gpurun_out/wide11_cfg4.err
