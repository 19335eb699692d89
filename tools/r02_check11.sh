#!/bin/bash
set -u
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
READS=131072 VARIANTS='[{}, {"col_threads": 128, "col_minblocks": 5}, {"col_threads": 128, "col_minblocks": 6}, {"col_threads": 64, "col_minblocks": 8}, {"col_sil_regs": 0}, {"col_sil_regs": 0, "col_threads": 128, "col_minblocks": 6}]' timeout 900 python tools/lane_variants.py > gpurun_out/col_variants11_131k.jsonl 2> gpurun_out/col_variants11.err
cat gpurun_out/col_variants11_131k.jsonl
grep -i "column engine:.*CTA" gpurun_out/col_variants11.err | sort | uniq -c | head -30
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu11.log 2>&1
tail -6 gpurun_out/pytest_gpu11.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
bash tools/profile_round.sh
