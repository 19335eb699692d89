#!/bin/bash
# tools/profile_round.sh -- run under gpurun on ONE GPU: the bench, the ncu launch list of the same
# command, and one `ncu --set full` capture per hot kernel.  Outputs go to gpurun_out/ (scratch);
# tools/ncu_summary.py turns the reports into the summaries committed under profiles/.
set -u
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
rm -f gpurun_out/prof_*.ncu-rep
./tools/pipe_peaks > gpurun_out/pipe_peaks.json 2> gpurun_out/pipe_peaks.err
export MB_JIT_DUMP=gpurun_out/mb_jit_kernels.cu
python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -c 400 gpurun_out/bench.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
unset MB_JIT_DUMP
# the launch list of the default command's headline leg (10 000 pairs per step; the E-step leg on 2048)
ncu --metrics gpu__time_duration.sum --clock-control none -c 160 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-configs --em-pairs 2048 > gpurun_out/ncu_launches_run.log 2>&1
# full captures at the bench's own size, so that dram bytes per launch are the bench's
for k in mb_k_forward_lin mb_k_viterbi mb_k_fstore_lin mb_k_bcounts_lin; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:^${k}\$ -c 1 -o gpurun_out/prof_${k} \
      python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-configs --em-pairs 2048 > gpurun_out/ncu_${k}_run.log 2>&1
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:jit_traceback_kernel -c 1 -o gpurun_out/prof_jit_traceback_kernel \
    python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-configs --em-pairs 256 > gpurun_out/ncu_tb_run.log 2>&1
# the engines of the other configs: the generated thread-per-cell sweep (config 4 machine) ...
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mb_k_big_forward -c 1 -o gpurun_out/prof_mb_k_big_forward \
    python tools/bench_wide.py --machine prot2dna_dnapsw --pairs 1184 --li 300 --lo 1000 --engines 2 --no-trace --reps 1 > gpurun_out/ncu_big_run.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:mb_k_big_viterbi$' -c 1 -o gpurun_out/prof_mb_k_big_viterbi \
    python tools/bench_wide.py --machine prot2dna_dnapsw --pairs 592 --li 300 --lo 1000 --engines 2 --reps 1 > gpurun_out/ncu_bigv_run.log 2>&1
# ... and the profile sweeps (config 5 machine): the column engine's strip kernel, and the lane engine's windowed sweep it replaced
READS=32768 VARIANTS='[{}]' timeout 600 ncu --set full --clock-control none --import-source on -k regex:mb_k_col_sum -c 1 -o gpurun_out/prof_mb_k_col_sum \
    python tools/lane_variants.py > gpurun_out/ncu_col_run.log 2>&1
READS=32768 VARIANTS='[{}]' timeout 600 ncu --set full --clock-control none --import-source on -k regex:mb_k_col_max -c 1 -o gpurun_out/prof_mb_k_col_max \
    python tools/lane_variants.py > gpurun_out/ncu_colmax_run.log 2>&1
READS=32768 VARIANTS='[{"no_col": 1}]' timeout 600 ncu --set full --clock-control none --import-source on -k regex:lane2_kernel -c 1 -o gpurun_out/prof_lane2_kernel_forward \
    python tools/lane_variants.py > gpurun_out/ncu_lane_run.log 2>&1
ls -la gpurun_out/*.ncu-rep
