#!/bin/bash
# tools/profile_round.sh -- run under gpurun on ONE GPU: the bench, the ncu launch list of the same
# command, and one `ncu --set full` capture per hot kernel.  Outputs go to gpurun_out/ (scratch);
# tools/ncu_summary.py turns the reports into the summaries committed under profiles/.
set -u
mkdir -p gpurun_out
export MB_JIT_DUMP=gpurun_out/mb_jit_kernels.cu
python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -c 400 gpurun_out/bench.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches.csv \
    python bench.py --pairs 4736 --steps 2 --warmup 1 --no-cpu-baseline --em-pairs 2368 > gpurun_out/ncu_launches_run.log 2>&1
for k in mb_k_forward_lin mb_k_viterbi mb_k_fstore_lin mb_k_bcounts_lin; do
  ncu --set full --clock-control none --import-source on -k regex:^${k}\$ -c 1 -o gpurun_out/prof_${k} \
      python bench.py --pairs 4736 --steps 1 --warmup 0 --no-cpu-baseline --em-pairs 2368 > gpurun_out/ncu_${k}_run.log 2>&1
done
ls -la gpurun_out/*.ncu-rep
