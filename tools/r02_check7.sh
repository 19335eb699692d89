#!/bin/bash
set -u
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q --durations=8 ) > gpurun_out/pytest_gpu7.log 2>&1
tail -22 gpurun_out/pytest_gpu7.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
( time timeout 1200 python bench.py ) > gpurun_out/bench7.json 2> gpurun_out/bench7.err
tail -c 1500 gpurun_out/bench7.json; tail -6 gpurun_out/bench7.err
MACHINE=protpsw P=100000 L=300 VARIANTS='[{}, {"jit_narrow": 0}, {"jit_narrow": 1}]' timeout 300 python tools/jit_variants.py > gpurun_out/jit_variants_prot2.jsonl 2> gpurun_out/jit_variants_prot2.err
cat gpurun_out/jit_variants_prot2.jsonl
READS=262144 LEN=275 VARIANTS='[{}, {"lane_warps_per_cta": 4}]' timeout 600 python tools/lane_variants.py > gpurun_out/lane_variants_262k.jsonl 2> gpurun_out/lane_variants2.err
cat gpurun_out/lane_variants_262k.jsonl; grep windowed gpurun_out/lane_variants2.err | sort | uniq -c
