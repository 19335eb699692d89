#!/bin/bash
set -u
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
P=10000 L=1000 VARIANTS='[{}, {"jit_tb_prefetch": 1}]' timeout 600 python tools/jit_variants.py 2>/dev/null | cut -c1-330
P=1250 L=1000 VARIANTS='[{}, {"jit_tb_prefetch": 1}]' timeout 600 python tools/jit_variants.py 2>/dev/null | cut -c1-330
python -c "
import bench, json; print(json.dumps(bench.run_ingest()))"
