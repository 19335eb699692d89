#!/bin/bash
set -u
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q -x --durations=5 -k "split or post_trans or big or golden or chunked or set_b" ) > gpurun_out/pytest_gpu8.log 2>&1
tail -12 gpurun_out/pytest_gpu8.log
P=10000 L=1000 VARIANTS='[{}, {"jit_split": 0}]' timeout 300 python tools/jit_variants.py > gpurun_out/jit_variants8.jsonl 2> gpurun_out/jit_variants8.err
cat gpurun_out/jit_variants8.jsonl
P=150 L=10000 VARIANTS='[{}, {"jit_split": 0}]' timeout 300 python tools/jit_variants.py > gpurun_out/jit_split8_150x10k.jsonl 2>> gpurun_out/jit_variants8.err
cat gpurun_out/jit_split8_150x10k.jsonl
MACHINE=protpsw P=100000 L=300 VARIANTS='[{}, {"jit_narrow": 1}]' timeout 300 python tools/jit_variants.py > gpurun_out/jit_variants_prot8.jsonl 2>> gpurun_out/jit_variants8.err
cat gpurun_out/jit_variants_prot8.jsonl
( time timeout 900 python bench.py --cfg5-reads 4096 ) > gpurun_out/bench8.json 2> gpurun_out/bench8.err
tail -c 600 gpurun_out/bench8.json; tail -6 gpurun_out/bench8.err
