#!/bin/bash
set -u
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu2.log 2>&1
tail -3 gpurun_out/pytest_gpu2.log
for cfg in "4 5 4" "3 3 8" "3 4 8" "3 4 6" "4 4 4"; do set -- $cfg; echo "minblocks=$1 lin=$2 C=$3"; MB_JIT_MINBLOCKS=$1 MB_JIT_MINBLOCKS_LIN=$2 MB_JIT_C=$3 timeout 300 python bench.py --no-cpu-baseline --steps 3 --warmup 3 --em-pairs 512 2>>gpurun_out/err_sweep.log | python -c "import json,sys; j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(j[\"forward_gcups\"]), round(j[\"viterbi_gcups\"]), round(j[\"value\"]), round(j[\"e2e\"][\"value\"]), round(j[\"em\"][\"pairs_per_s\"]), j[\"em\"][\"kernel_ms\"])"; done 2>&1 | tee gpurun_out/sweep2.log
for k in mb_k_forward_lin mb_k_viterbi; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:^${k}\$ -c 1 -o gpurun_out/prof2_${k} \
      python bench.py --pairs 4736 --steps 1 --warmup 0 --no-cpu-baseline --em-pairs 256 > gpurun_out/ncu2_${k}_run.log 2>&1
done
ls -la gpurun_out/*.ncu-rep
