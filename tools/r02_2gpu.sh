#!/bin/bash
set -u
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
nvidia-smi -L
( time timeout 900 python -m pytest tests -m gpu -q -k "group or shard" ) > gpurun_out/pytest_2gpu.log 2>&1
tail -6 gpurun_out/pytest_2gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 --cfg3-pairs 20000 --cfg4-pairs 296 --cfg5-reads 32768 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err
tail -c 600 gpurun_out/bench_2gpu.json; tail -3 gpurun_out/bench_2gpu.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/bench_ref_2gpu.json 2>> gpurun_out/bench_2gpu.err
tail -c 300 gpurun_out/bench_ref_2gpu.json
# the one-process path: the CLI over both devices (E-step of a list through mb_group_counts with the NCCL all-reduce)
python - <<'PY' 2>&1 | tail -8
import sys, time, numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import bench
from machineboss_b200 import capi
mj = bench.dnapsw_machine()
x, xo, y, yo = bench.synth_batch(bench.SEED, 0, 8192, 1000, 1000, 4)
g = capi.Group()
print("group devices:", g.n_devices, "nccl:", g.uses_nccl)
gm = capi.GroupMachine(g, mj["n_states"], mj["n_in"], mj["n_out"], mj["src"], mj["dst"], mj["tin"], mj["tout"], mj["lw"])
gb = capi.GroupBatch(g, x=x, x_off=xo, y=y, y_off=yo)
for name, fn in (("forward", lambda: capi.group_forward(gm, gb)), ("viterbi", lambda: capi.group_viterbi(gm, gb)), ("counts", lambda: capi.group_counts(gm, gb))):
    fn(); t0 = time.time(); r = fn(); dt = time.time() - t0
    print("%s over %d devices: %.1f ms wall for 8192 pairs, kernel ms per device %s" % (name, g.n_devices, dt * 1e3, gb.last_kernel_ms()))
m = capi.Machine(mj["n_states"], mj["n_in"], mj["n_out"], mj["src"], mj["dst"], mj["tin"], mj["tout"], mj["lw"])
b = capi.Batch(x=x, x_off=xo, y=y, y_off=yo)
c1, ll1 = capi.counts(m, b)
c2, ll2 = capi.group_counts(gm, gb)
print("counts equal to one device: max rel diff %.2e, ll diff %.2e" % (float(np.max(np.abs(c1 - c2) / np.maximum(np.abs(c1), 1e-300))), float(np.max(np.abs(np.asarray(ll1) - np.asarray(ll2))))))
PY
