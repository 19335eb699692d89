#!/bin/bash
set -u
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu3.log 2>&1
tail -15 gpurun_out/pytest_gpu3.log
run() { timeout 300 python bench.py --no-cpu-baseline --steps 3 --warmup 3 --em-pairs 512 2>>gpurun_out/err_sweep3.log | python -c "import json,sys; j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(j[\"forward_gcups\"]), round(j[\"viterbi_gcups\"]), round(j[\"value\"]), round(j[\"e2e\"][\"value\"]), round(j[\"roofline\"][\"viterbi\"][\"ms_per_launch\"],2))"; }
( echo "default"; run; echo "TB_SIMPLE"; MB_JIT_TB_SIMPLE=1 run; echo "V minblocks 4"; MB_JIT_MINBLOCKS_V=4 run; echo "V minblocks 2"; MB_JIT_MINBLOCKS_V=2 run; echo "CV=4"; MB_JIT_CV=4 MB_JIT_MINBLOCKS_V=4 run ) 2>&1 | tee gpurun_out/sweep3.log
MB_WIDE_VERBOSE=1 timeout 300 python tools/bench_wide.py --machine hmmer_pf00516 --pairs 4096 --li 0 --lo 275 --engines 2 > gpurun_out/lane_cfg5a.json 2> gpurun_out/lane_cfg5a.err
tail -c 900 gpurun_out/lane_cfg5a.json; tail -2 gpurun_out/lane_cfg5a.err
MB_WIDE_VERBOSE=1 timeout 300 python tools/bench_wide.py --machine hmmer_pf00516 --pairs 65536 --li 0 --lo 275 --engines 2 --reps 1 > gpurun_out/lane_cfg5b.json 2> gpurun_out/lane_cfg5b.err
tail -c 900 gpurun_out/lane_cfg5b.json; tail -2 gpurun_out/lane_cfg5b.err
MB_WIDE_VERBOSE=1 timeout 400 python tools/bench_wide.py --machine hmmer_pf00516_protpsw --pairs 16384 --li 0 --lo 275 --engines 2 --reps 1 > gpurun_out/lane_cfg5c.json 2> gpurun_out/lane_cfg5c.err
tail -c 900 gpurun_out/lane_cfg5c.json; tail -2 gpurun_out/lane_cfg5c.err
