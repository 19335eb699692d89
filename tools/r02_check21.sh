#!/bin/bash
set -u
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q -x -k "golden or strip or fitted or chunked" ) > gpurun_out/pytest_gpu21.log 2>&1
tail -5 gpurun_out/pytest_gpu21.log
P=10000 L=1000 VARIANTS='[{}]' timeout 600 python tools/jit_variants.py 2>/dev/null | cut -c1-330
P=150 L=10000 VARIANTS='[{}]' timeout 300 python tools/jit_variants.py 2>/dev/null | cut -c1-330
