#!/bin/bash
# run under gpurun on ONE GPU: the whole -m gpu suite, the default bench, both arms, the wide engine's
# throughput on config-4 / config-5 style batches, the launch list and one full ncu capture of the wide sweep
set -u
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
tail -5 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -c 600 gpurun_out/bench.json
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
tail -c 300 gpurun_out/bench_reference.json
MB_WIDE_VERBOSE=1 timeout 300 python tools/bench_wide.py --machine prot2dna_dnapsw --pairs 148 --li 300 --lo 10000 --engines 2 > gpurun_out/wide_cfg4.json 2> gpurun_out/wide_cfg4.err
tail -c 900 gpurun_out/wide_cfg4.json; tail -2 gpurun_out/wide_cfg4.err
MB_WIDE_VERBOSE=1 timeout 300 python tools/bench_wide.py --machine hmmer_pf00516 --pairs 4096 --li 0 --lo 275 --engines 2 > gpurun_out/wide_cfg5.json 2> gpurun_out/wide_cfg5.err
tail -c 900 gpurun_out/wide_cfg5.json; tail -2 gpurun_out/wide_cfg5.err
timeout 300 python tools/other_configs.py > gpurun_out/other_configs.json 2> gpurun_out/other_configs.err
tail -c 900 gpurun_out/other_configs.json
timeout 300 ncu --set full --clock-control none --import-source on -k regex:wide_kernel -c 1 -o gpurun_out/prof_wide_forward \
    python tools/bench_wide.py --machine prot2dna_dnapsw --pairs 148 --li 300 --lo 1000 --engines 2 --no-trace --reps 1 > gpurun_out/ncu_wide_run.log 2>&1
ls -la gpurun_out/
