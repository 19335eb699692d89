#!/bin/bash
set -u
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q -x -k "strip or fitted or golden or set_b or chunked or split" ) > gpurun_out/pytest_gpu16.log 2>&1
tail -12 gpurun_out/pytest_gpu16.log
MACHINE=protpsw P=100000 L=300 VARIANTS='[{}, {"jit_fit_c": 0}]' timeout 600 python tools/jit_variants.py > gpurun_out/jit_variants_prot16.jsonl 2> gpurun_out/jit_variants16.err
cat gpurun_out/jit_variants_prot16.jsonl; tail -2 gpurun_out/jit_variants16.err
P=10000 L=1000 VARIANTS='[{}]' timeout 600 python tools/jit_variants.py > gpurun_out/jit_variants_dna16.jsonl 2>> gpurun_out/jit_variants16.err
cat gpurun_out/jit_variants_dna16.jsonl
