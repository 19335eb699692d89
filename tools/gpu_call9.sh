#!/bin/bash
set -u
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu9.log 2>&1
tail -5 gpurun_out/pytest_gpu9.log
timeout 200 python tools/e2e_breakdown.py 2>&1 | tee gpurun_out/e2e_breakdown9.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench9.json 2> gpurun_out/bench9.err
python -c "
import json; d=json.loads(open('gpurun_out/bench9.json').read().strip().splitlines()[-1]); print(d['value'], d['e2e'], d['em']['pairs_per_s'])"
timeout 200 python tools/other_configs.py > gpurun_out/other_configs9.json 2> gpurun_out/other_configs9.err; cat gpurun_out/other_configs9.json
MB_JIT_NARROW=0 timeout 200 python tools/other_configs.py 2>/dev/null | head -1
