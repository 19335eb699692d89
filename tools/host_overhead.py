"""Wall time vs in-library kernel time of the two calls of a bench step (where does host time go?)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from machineboss_b200 import capi
mj = bench.dnapsw_machine()
mach = capi.Machine(mj["n_states"], mj["n_in"], mj["n_out"], mj["src"], mj["dst"], mj["tin"], mj["tout"], mj["lw"])
P = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
x, xo, y, yo = bench.synth_batch(12345, 0, P, 1000, 1000, 4)
b = capi.Batch(x=x, x_off=xo, y=y, y_off=yo)
for it in range(6):
    t0 = time.perf_counter(); ll = capi.forward(mach, b); t1 = time.perf_counter(); kf, _ = b.last_kernel_ms()
    sc, pl = capi.viterbi_lengths(mach, b); t2 = time.perf_counter(); kv, _ = b.last_kernel_ms()
    print("iter %d forward wall %.2f ms kernel %.2f | viterbi wall %.2f ms kernel %.2f" % (it, 1e3 * (t1 - t0), kf, 1e3 * (t2 - t1), kv))
