#!/bin/bash
set -u
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
MACHINE=protpsw P=100000 L=300 VARIANTS='[{}, {"jit_cv": 10, "jit_narrow": 0}, {"jit_cv": 10, "jit_narrow": 0, "jit_minblocks_linv": 3}, {"jit_cv": 10, "jit_narrow": 0, "jit_minblocks_v": 2, "jit_minblocks_linv": 2}, {"jit_cv": 12, "jit_narrow": 0, "jit_minblocks_linv": 3}]' timeout 600 python tools/jit_variants.py > gpurun_out/jit_variants_prot15.jsonl 2> gpurun_out/jit_variants15.err
cat gpurun_out/jit_variants_prot15.jsonl; tail -2 gpurun_out/jit_variants15.err
P=30000 L=300 VARIANTS='[{}, {"jit_cv": 10, "jit_narrow": 0}, {"jit_cv": 10, "jit_narrow": 0, "jit_minblocks_linv": 3}]' timeout 600 python tools/jit_variants.py > gpurun_out/jit_variants_dna15.jsonl 2>> gpurun_out/jit_variants15.err
cat gpurun_out/jit_variants_dna15.jsonl
READS=131072 VARIANTS='[{}, {"col_c": 1}]' timeout 900 python tools/lane_variants.py > gpurun_out/col_variants15.jsonl 2> gpurun_out/col_variants15.err
cat gpurun_out/col_variants15.jsonl
