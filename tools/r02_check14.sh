#!/bin/bash
set -u
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
MACHINE=protpsw P=100000 L=300 VARIANTS='[{}, {"jit_cv": 10, "jit_narrow": 0}, {"jit_cv": 10, "jit_narrow": 0, "jit_minblocks_linv": 3}, {"jit_cv": 10, "jit_narrow": 0, "jit_minblocks_v": 2, "jit_minblocks_linv": 2}, {"jit_cv": 12, "jit_narrow": 0}]' timeout 600 python tools/jit_variants.py > gpurun_out/jit_variants_prot14.jsonl 2> gpurun_out/jit_variants14.err
cat gpurun_out/jit_variants_prot14.jsonl
P=10000 L=300 VARIANTS='[{}, {"jit_cv": 10, "jit_narrow": 0}]' timeout 600 python tools/jit_variants.py > gpurun_out/jit_variants_dna14.jsonl 2>> gpurun_out/jit_variants14.err
cat gpurun_out/jit_variants_dna14.jsonl
( time timeout 1200 python bench.py --no-cpu-baseline ) > gpurun_out/bench14.json 2> gpurun_out/bench14.err
tail -c 300 gpurun_out/bench14.json; tail -3 gpurun_out/bench14.err
