#!/bin/bash
set -u
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -q -x -k "golden or strip or fitted or split or chunked or set_b or ragged or update_weights or group_of_one or cli or fit" ) > gpurun_out/pytest_gpu18.log 2>&1
tail -8 gpurun_out/pytest_gpu18.log
P=10000 L=1000 VARIANTS='[{}]' timeout 600 python tools/jit_variants.py > gpurun_out/jit_variants_dna18.jsonl 2> gpurun_out/jit_variants18.err
cat gpurun_out/jit_variants_dna18.jsonl
MACHINE=protpsw P=100000 L=300 VARIANTS='[{}]' timeout 600 python tools/jit_variants.py > gpurun_out/jit_variants_prot18.jsonl 2>> gpurun_out/jit_variants18.err
cat gpurun_out/jit_variants_prot18.jsonl
P=150 L=10000 VARIANTS='[{}]' timeout 300 python tools/jit_variants.py > gpurun_out/jit_split18.jsonl 2>> gpurun_out/jit_variants18.err
cat gpurun_out/jit_split18.jsonl
