"""Where an end-to-end step (host buffers in, host buffers out) spends its time: per-phase host clock
around the C-ABI calls of bench.py's e2e leg."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from machineboss_b200 import capi

P = int(os.environ.get("P", "10000")); L = 1000
mj = bench.dnapsw_machine()
mach = capi.Machine(mj["n_states"], mj["n_in"], mj["n_out"], mj["src"], mj["dst"], mj["tin"], mj["tout"], mj["lw"])
x, x_off, y, y_off = bench.synth_batch(bench.SEED, 0, P, L, L, 4)
px = torch.from_numpy(x).pin_memory(); py = torch.from_numpy(y).pin_memory()
h_ll = torch.empty(P, dtype=torch.float64).pin_memory().numpy()
h_sc = torch.empty(P, dtype=torch.float64).pin_memory().numpy()
h_len = torch.empty(P, dtype=torch.int64).pin_memory().numpy()
h_off = torch.empty(P + 1, dtype=torch.int64).pin_memory().numpy()
h_paths = torch.empty(P * (2 * L + 2) * 2, dtype=torch.int32).pin_memory().numpy()
lib = capi.lib()
for it in range(6):
    torch.cuda.synchronize()
    t = [time.perf_counter()]
    b = capi.Batch(x=px.numpy(), x_off=x_off, y=py.numpy(), y_off=y_off); t.append(time.perf_counter())
    capi.forward_into(mach, b, h_ll); t.append(time.perf_counter())
    capi._check(lib.mb_viterbi(mach.h, b.h, capi._ptr(h_sc), capi._ptr(h_len))); t.append(time.perf_counter())
    kms, _ = b.last_kernel_ms()
    h_off[0] = 0; np.cumsum(h_len[:P], out=h_off[1:P + 1]); t.append(time.perf_counter())
    capi._check(lib.mb_viterbi_paths(b.h, capi._ptr(h_paths), capi._ptr(h_off))); t.append(time.perf_counter())
    b.close(); t.append(time.perf_counter())
    d = [round(1e3 * (t[i + 1] - t[i]), 2) for i in range(len(t) - 1)]
    print("iter %d: batch %.2f forward %.2f viterbi %.2f (kernels %.2f) cumsum %.2f paths_d2h %.2f (%.0f MB) close %.2f  total %.2f ms"
          % (it, d[0], d[1], d[2], kms, d[3], d[4], h_off[P] * 4 / 1e6, d[5], 1e3 * (t[-1] - t[0])))
