#!/bin/bash
# run under gpurun on ONE GPU: the whole -m gpu suite, smoke(), the default bench of both arms, and the
# throughput of the other configs' machines (protpsw, the composed prot2dna => dnapsw, the PF00516 profiles)
set -u
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
tail -5 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -c 600 gpurun_out/bench.json
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
tail -c 300 gpurun_out/bench_reference.json
timeout 300 python tools/other_configs.py > gpurun_out/other_configs.json 2> gpurun_out/other_configs.err
cat gpurun_out/other_configs.json
MB_WIDE_VERBOSE=1 timeout 300 python tools/bench_wide.py --machine prot2dna_dnapsw --pairs 592 --li 300 --lo 2000 --engines 2 > gpurun_out/wide_cfg4.json 2> gpurun_out/wide_cfg4.err
tail -c 900 gpurun_out/wide_cfg4.json; tail -2 gpurun_out/wide_cfg4.err
MB_WIDE_VERBOSE=1 timeout 300 python tools/bench_wide.py --machine hmmer_pf00516 --pairs 262144 --li 0 --lo 275 --engines 2 --no-trace > gpurun_out/lane_cfg5.json 2> gpurun_out/lane_cfg5.err
tail -c 600 gpurun_out/lane_cfg5.json; tail -2 gpurun_out/lane_cfg5.err
MB_WIDE_VERBOSE=1 timeout 300 python tools/bench_wide.py --machine hmmer_pf00516_protpsw --pairs 65536 --li 0 --lo 275 --engines 2 --no-trace > gpurun_out/lane_cfg5b.json 2> gpurun_out/lane_cfg5b.err
tail -c 600 gpurun_out/lane_cfg5b.json; tail -2 gpurun_out/lane_cfg5b.err
