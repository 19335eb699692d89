#!/bin/bash
set -u
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q -x -k "big or prot2dna or composite" ) > gpurun_out/pytest_gpu23.log 2>&1
tail -4 gpurun_out/pytest_gpu23.log
timeout 900 python tools/bench_wide.py --machine prot2dna_dnapsw --pairs 592 --li 300 --lo 2000 --engines 2 --reps 2 > gpurun_out/big23.json 2> gpurun_out/big23.err
cat gpurun_out/big23.json; tail -2 gpurun_out/big23.err
