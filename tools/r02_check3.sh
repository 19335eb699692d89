#!/bin/bash
set -u
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -q -x -k "golden and (dnapsw or protpsw or bit or unit or stutter or counter or translate) or ragged or strip or set_b or chunked or update_weights or long_pairs or dangerous" ) > gpurun_out/pytest_gpu3.log 2>&1
tail -15 gpurun_out/pytest_gpu3.log
timeout 600 python tools/jit_variants.py > gpurun_out/jit_variants.jsonl 2> gpurun_out/jit_variants.err
cat gpurun_out/jit_variants.jsonl; tail -3 gpurun_out/jit_variants.err
MACHINE=protpsw P=20000 L=300 VARIANTS='[{}, {"jit_narrow": 0}, {"jit_narrow": 1}, {"jit_narrow": 0, "jit_vit_intcmp": 0}]' timeout 300 python tools/jit_variants.py > gpurun_out/jit_variants_prot.jsonl 2>> gpurun_out/jit_variants.err
cat gpurun_out/jit_variants_prot.jsonl
