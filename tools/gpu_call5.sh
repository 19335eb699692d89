#!/bin/bash
set -u
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu5.log 2>&1
tail -15 gpurun_out/pytest_gpu5.log
run() { timeout 300 python bench.py --no-cpu-baseline --steps 3 --warmup 3 --em-pairs 2048 2>>gpurun_out/err_sweep5.log | python -c "import json,sys; j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(j[\"forward_gcups\"]), round(j[\"viterbi_gcups\"]), round(j[\"value\"]), round(j[\"e2e\"][\"value\"]), round(j[\"roofline\"][\"forward\"][\"ms_per_launch\"],2), round(j[\"roofline\"][\"viterbi\"][\"ms_per_launch\"],2), round(j[\"em\"][\"pairs_per_s\"]), j[\"check\"])"; }
( echo "default"; run; echo "CV=4"; MB_JIT_CV=4 run; echo "V minblocks 2"; MB_JIT_MINBLOCKS_V=2 run; echo "V minblocks 4"; MB_JIT_MINBLOCKS_V=4 run ) 2>&1 | tee gpurun_out/sweep5.log
timeout 200 python tools/e2e_breakdown.py 2>&1 | tee gpurun_out/e2e_breakdown.log
timeout 200 python tools/other_configs.py > gpurun_out/other_configs5.json 2> gpurun_out/other_configs5.err; cat gpurun_out/other_configs5.json
