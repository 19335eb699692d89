#!/bin/bash
# HEAD of the round on ONE GPU: the whole -m gpu suite, smoke(), the default bench and its reference arm
set -u
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/pytest_gpu_final3.log 2>&1
tail -5 gpurun_out/pytest_gpu_final3.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -c 300 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
tail -c 300 gpurun_out/bench_reference.json
