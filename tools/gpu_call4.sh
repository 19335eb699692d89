#!/bin/bash
set -u
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu4.log 2>&1
tail -15 gpurun_out/pytest_gpu4.log
run() { timeout 300 python bench.py --no-cpu-baseline --steps 3 --warmup 3 --em-pairs 2048 2>>gpurun_out/err_sweep4.log | python -c "import json,sys; j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(j[\"forward_gcups\"]), round(j[\"viterbi_gcups\"]), round(j[\"value\"]), round(j[\"e2e\"][\"value\"]), round(j[\"roofline\"][\"viterbi\"][\"ms_per_launch\"],2), round(j[\"em\"][\"pairs_per_s\"]))"; }
( echo "default"; run; echo "TB_TILED"; MB_JIT_TB_TILED=1 run; echo "V minblocks 2"; MB_JIT_MINBLOCKS_V=2 run ) 2>&1 | tee gpurun_out/sweep4.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches4.csv python bench.py --pairs 10000 --steps 1 --warmup 1 --no-cpu-baseline --em-pairs 256 > gpurun_out/ncu4_run.log 2>&1
MB_JIT_TB_TILED=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches4_tiled.csv python bench.py --pairs 10000 --steps 1 --warmup 1 --no-cpu-baseline --em-pairs 256 > gpurun_out/ncu4t_run.log 2>&1
for cfg in "4 32" "1 32" "2 32" "4 16" "4 48"; do set -- $cfg; echo "lane R=$1 warps=$2"; MB_LANE_R=$1 MB_LANE_WARPS=$2 timeout 300 python tools/bench_wide.py --machine hmmer_pf00516 --pairs 65536 --li 0 --lo 275 --engines 2 --reps 1 --no-trace 2>/dev/null | python -c "import json,sys; j=json.loads(sys.stdin.read())['engine2']; print(round(j['forward']['gcups'],1), round(j['viterbi_score']['gcups'],1))"; done 2>&1 | tee gpurun_out/lane4.log
MB_WIDE_VERBOSE=1 timeout 300 python tools/bench_wide.py --machine hmmer_pf00516 --pairs 65536 --li 0 --lo 275 --engines 2 --reps 1 > gpurun_out/lane4_cfg5b.json 2> gpurun_out/lane4_cfg5b.err
tail -c 700 gpurun_out/lane4_cfg5b.json
MB_WIDE_VERBOSE=1 timeout 400 python tools/bench_wide.py --machine hmmer_pf00516_protpsw --pairs 32768 --li 0 --lo 275 --engines 2 --reps 1 > gpurun_out/lane4_cfg5c.json 2> gpurun_out/lane4_cfg5c.err
tail -c 900 gpurun_out/lane4_cfg5c.json; tail -2 gpurun_out/lane4_cfg5c.err
