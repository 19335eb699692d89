#!/usr/bin/env python3
"""bench.py -- batched Forward + Viterbi throughput of preset dnapsw on synthetic 1 kb DNA pairs.

One "step" = one pass of the hot path over one batch: Forward log-likelihood (boss -L) and Viterbi
score + traceback (boss -V/-A) for every pair of the batch.  Metric (BASELINE.json): DP cell-state
updates per second, in GCUPS, summed over both sweeps and over all ranks; pairs/s beside it.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--pairs P] [--len L] [--impl reference]

N > 1: launched by torchrun, one rank per GPU; every rank processes its own P pairs (weak scaling,
no data-path collective; the E-step's count all-reduce is timed separately as "em").  Times are
taken per step between device synchronisations, L2 is flushed between steps, the maximum over
ranks is used, and SM clocks / throttle reasons are sampled with nvidia-smi during the timed region.

--impl reference times the reference's own CPU implementation (oracle/_ref/refdrv, the unmodified
reference sources; falls back to the C restatement if that binary is absent) on the host cores.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

METRIC = "forward_viterbi_gcups"
UNIT = "GCUPS"
SEED = 12345


# ---------------------------------------------------------------------------------------------
# workload: preset dnapsw, `boss -U` default parameters (src/constraints.cpp:65-75)
# ---------------------------------------------------------------------------------------------
def dnapsw_machine():
    """The flat evaluated dnapsw machine from the committed fixture (generated from the reference)."""
    with open(os.path.join(REPO, "tests", "golden", "dnapsw_synth64.json")) as f:
        j = json.load(f)["machine"]
    t = j["trans"]

    def num(v):
        return float("-inf") if v == "-Infinity" else float(v)
    return dict(n_states=j["nStates"], n_in=len(j["inAlphabet"]), n_out=len(j["outAlphabet"]),
                src=np.array([r[0] for r in t], np.int32), dst=np.array([r[1] for r in t], np.int32),
                tin=np.array([r[2] for r in t], np.int32), tout=np.array([r[3] for r in t], np.int32),
                lw=np.array([num(r[4]) for r in t], np.float64))


def _splitmix64(x):
    with np.errstate(over="ignore"):
        x = x + np.uint64(0x9E3779B97F4A7C15)
        x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return x ^ (x >> np.uint64(31))


def synth_batch(seed: int, first_pair: int, n_pairs: int, li: int, lo: int, n_sym: int):
    """iid uniform tokens; same stream as oracle/synth.h (pair k, stream which, position p)."""
    out = []
    for which, length in ((0, li), (1, lo)):
        with np.errstate(over="ignore"):
            k = (np.arange(first_pair, first_pair + n_pairs, dtype=np.uint64) * np.uint64(2) + np.uint64(which))
            base = np.uint64(seed) * np.uint64(0x9E3779B97F4A7C15) + k * np.uint64(0xD1B54A32D192ED03)
            p = base[:, None] + np.arange(length, dtype=np.uint64)[None, :]
        out.append((1 + (_splitmix64(p) % np.uint64(n_sym))).astype(np.uint8).reshape(-1))
    x_off = np.arange(n_pairs + 1, dtype=np.int64) * li
    y_off = np.arange(n_pairs + 1, dtype=np.int64) * lo
    return out[0], x_off, out[1], y_off


# ---------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, device: int):
        self.device, self.proc, self.lines = device, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for line in self.lines:
            f = [c.strip() for c in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------
# CPU reference arm
# ---------------------------------------------------------------------------------------------
def cpu_reference(li: int, lo: int, n_pairs: int, threads: int, seed: int = SEED):
    """Forward (rolling, as boss -L) + Viterbi with traceback on `n_pairs` synthetic pairs.

    Returns (gcups, pairs_per_s, kind, seconds)."""
    refdrv = os.path.join(REPO, "oracle", "_ref", "refdrv")
    cells = float(li + 1) * float(lo + 1) * 8 * n_pairs
    if os.path.exists(refdrv):
        r = subprocess.run([refdrv, "--machine", "preset:dnapsw", "--synth", "%d,%d,%d,%d" % (n_pairs, li, lo, seed),
                            "--do", "rolling,viterbi,path", "--threads", str(threads), "--quiet-results"],
                           check=True, capture_output=True, text=True)
        secs = float(json.loads(r.stdout)["seconds"])
        return 2 * cells / secs / 1e9, n_pairs / secs, "reference", secs
    # the C restatement (single-threaded per call; run `threads` processes' worth sequentially is
    # pointless, so time one thread and say so)
    sys.path.insert(0, os.path.join(REPO, "tests"))
    from helpers import FlatMachine, Oracle, load_golden, synth_tokens
    orc = Oracle(FlatMachine.from_json(load_golden("dnapsw_synth64")["machine"]))
    t0 = time.perf_counter()
    for k in range(n_pairs):
        x, y = synth_tokens(seed, k, 0, li, 4), synth_tokens(seed, k, 1, lo, 4)
        orc.forward(x, y)
        orc.viterbi(x, y)
    secs = time.perf_counter() - t0
    return 2 * cells / secs / 1e9, n_pairs / secs, "port", secs


def run_reference(args, rank: int, world: int):
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    n = max(threads, 8)
    for _ in range(args.warmup if args.warmup < 1 else 1):     # one bounded warm-up pass is enough on a CPU
        cpu_reference(args.len, args.len, min(n, threads), threads)
    vals, secs = [], []
    kind = "reference"
    for _ in range(args.steps):
        g, pps, kind, s = cpu_reference(args.len, args.len, n, threads)
        vals.append(g); secs.append(s)
    value = statistics.mean(vals)
    sample = "%d synthetic %dx%d dnapsw pairs per step (Forward rolling + Viterbi with traceback), %d threads" % (n, args.len, args.len, threads)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * statistics.mean(secs), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, n),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "pairs_per_s": n / statistics.mean(secs)}))


def workload_config(args, pairs_per_rank):
    return {"workload": "preset dnapsw (S=8, 34 transitions, boss -U default parameters), %d synthetic iid-uniform pairs of "
                        "%d x %d nt per GPU: Forward log-likelihood + Viterbi score and traceback for every pair"
                        % (pairs_per_rank, args.len, args.len),
            "pairs_per_gpu": pairs_per_rank, "len_in": args.len, "len_out": args.len, "seed": SEED,
            "l2": "256 MiB buffer written between timed steps (L2 flush)", "parallelism": "pairs sharded, %d rank(s)" % args.gpus}


# ---------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--pairs", type=int, default=int(os.environ.get("MB_BENCH_PAIRS", "10000")), help="pairs per GPU")
    ap.add_argument("--len", type=int, default=1000)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--engine", type=int, default=-1)
    ap.add_argument("--em-pairs", type=int, default=4096, help="pairs per GPU in the E-step (counts) leg")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from machineboss_b200 import capi

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; there is no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    capi.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    mj = dnapsw_machine()
    capi.set_engine(args.engine)
    mach = capi.Machine(mj["n_states"], mj["n_in"], mj["n_out"], mj["src"], mj["dst"], mj["tin"], mj["tout"], mj["lw"])
    capi.set_engine(-1)
    S = mj["n_states"]
    P = args.pairs
    x, x_off, y, y_off = synth_batch(SEED, rank * P, P, args.len, args.len, 4)
    # pinned host staging for the end-to-end leg
    px = torch.from_numpy(x).pin_memory()
    py = torch.from_numpy(y).pin_memory()
    cells = float(args.len + 1) * float(args.len + 1) * S * P      # cell-states per sweep per rank
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    batch = capi.Batch(x=x, x_off=x_off, y=y, y_off=y_off)
    kernel_ms = {"forward": [], "viterbi": []}
    launches = 0

    def step_resident():
        nonlocal launches
        ll = capi.forward(mach, batch)
        ms, n = batch.last_kernel_ms(); kernel_ms["forward"].append(ms); launches += n
        sc, plen = capi.viterbi_lengths(mach, batch)
        ms, n = batch.last_kernel_ms(); kernel_ms["viterbi"].append(ms); launches += n
        return ll, sc, plen

    # end-to-end leg: host buffers in, host buffers out, through the C ABI.  Inputs are copied from
    # pinned host memory, results land in pinned host memory the caller owns (allocated once, as a
    # service would): per-pair log-likelihoods, Viterbi scores, path lengths and the packed paths
    # (global transition ids, one byte each for a machine this small).
    path_cap = P * (2 * args.len + 2) * 2
    h_ll = torch.empty(P, dtype=torch.float64).pin_memory().numpy()
    h_sc = torch.empty(P, dtype=torch.float64).pin_memory().numpy()
    h_len = torch.empty(P, dtype=torch.int64).pin_memory().numpy()
    h_off = torch.empty(P + 1, dtype=torch.int64).pin_memory().numpy()
    # transition ids as bytes when the machine has at most 256 transitions (dnapsw: 34), mb_viterbi_paths_narrow
    id_dtype = torch.uint8 if len(mj["lw"]) <= 256 else torch.int32
    h_paths = torch.empty(path_cap, dtype=id_dtype).pin_memory().numpy()

    def step_e2e():
        b = capi.Batch(x=px.numpy(), x_off=x_off, y=py.numpy(), y_off=y_off)      # H2D from pinned memory
        capi.forward_into(mach, b, h_ll)
        total = capi.viterbi_into(mach, b, h_sc, h_len, h_off, h_paths)           # D2H of scores, lengths and packed paths
        b.close()
        return h_ll, h_sc, (h_paths[:total], h_off)

    for _ in range(args.warmup):
        step_resident()
    kernel_ms = {"forward": [], "viterbi": []}
    launches = 0

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    times = []
    for _ in range(args.steps):
        flush.fill_(1)
        barrier()
        t0 = time.perf_counter()
        ll, sc, plen = step_resident()
        torch.cuda.synchronize()
        times.append(time.perf_counter() - t0)
    total = sum(times)
    e2e_times = []
    d2h = 0
    for n in range(args.warmup + args.steps):      # W untimed steps first: a fresh batch's scratch comes from the allocator until the pool is warm
        flush.fill_(1)
        barrier()
        t0 = time.perf_counter()
        ll2, sc2, paths = step_e2e()
        torch.cuda.synchronize()
        if n >= args.warmup:
            e2e_times.append(time.perf_counter() - t0)
        d2h = ll2.nbytes + sc2.nbytes + paths[0].nbytes + paths[1].nbytes
    e2e_total = sum(e2e_times)
    clocks = sampler.stop() if rank == 0 else None
    assert np.array_equal(ll, ll2) and np.array_equal(sc, sc2)

    # E-step of Baum-Welch on a bounded slice of the same batch: Forward (stored) + Backward fused with
    # the posterior counts, then the path's one exchange step, the all-reduce of nTrans+1 doubles.
    from machineboss_b200 import shard
    n_em = min(P, args.em_pairs)
    em_batch = capi.Batch(x=x[: n_em * args.len], x_off=x_off[: n_em + 1], y=y[: n_em * args.len], y_off=y_off[: n_em + 1])
    capi.counts(mach, em_batch)
    barrier()
    t0 = time.perf_counter()
    cnt, cll = capi.counts(mach, em_batch)
    em_kernel_ms, em_launches = em_batch.last_kernel_ms()
    cnt, tot_ll = shard.allreduce_counts(cnt, float(cll.sum()), device="cuda")
    torch.cuda.synchronize()
    em_secs = time.perf_counter() - t0
    em_batch.close()

    if world > 1:
        t = torch.tensor([total, e2e_total, em_secs], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total, e2e_total, em_secs = float(t[0]), float(t[1]), float(t[2])

    if rank == 0:
        K = args.steps
        value = 2 * cells * world * K / total / 1e9
        e2e = 2 * cells * world * K / e2e_total / 1e9
        fwd_ms = statistics.mean(kernel_ms["forward"])
        vit_ms = statistics.mean(kernel_ms["viterbi"])
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": args.warmup,
            "ms_per_step": 1e3 * total / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": workload_config(args, P),
            "pairs_per_s": P * world * K / total,
            "forward_gcups": cells / fwd_ms / 1e6, "viterbi_gcups": cells / vit_ms / 1e6,
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": int(x.nbytes + y.nbytes + x_off.nbytes + y_off.nbytes),
                    "d2h_bytes_per_step": int(d2h), "ms_per_step": 1e3 * e2e_total / K},
            "gpu_launches": launches, "clocks": clocks, "engine": mach.engine,
            "check": {"forward_ll_pair0": float(ll[0]), "viterbi_pair0": float(sc[0]), "path_len_pair0": int(plen[0])},
        }
        em_cells = float(args.len + 1) * float(args.len + 1) * S * n_em
        out["em"] = {"what": "E-step (MachineCounts over the list): stored Forward + fused Backward/posterior counts on %d pairs per GPU, then all-reduce of %d doubles" % (n_em, len(cnt) + 1),
                     "pairs_per_s": n_em * world / em_secs, "gcups_2_sweeps": 2 * em_cells * world / em_secs / 1e9,
                     "kernel_ms": em_kernel_ms, "launches": em_launches, "sum_counts": float(cnt.sum()), "loglike": tot_ll}
        out["roofline"] = roofline(mj, cells, fwd_ms, vit_ms)
        if not args.no_cpu_baseline and world == 1:
            threads = os.cpu_count() or 1
            n = max(8, threads)
            g, pps, kind, secs = cpu_reference(args.len, args.len, n, threads)
            out["cpu_baseline"] = {"value": g, "unit": UNIT, "cores": threads, "kind": kind, "seconds": secs,
                                   "sample": "%d of the same synthetic %dx%d pairs, Forward (rolling) + Viterbi with traceback, %d threads"
                                             % (n, args.len, args.len, threads)}
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def roofline(mj: dict, cells: float, fwd_ms: float, vit_ms: float) -> dict:
    """Roofline of the step's kernels.  DESIGN.md section 'Roofline'.

    The fills are compute-bound (a score sweep's algorithmic HBM traffic is the tokens plus 8 B per
    pair; the Viterbi back-pointers add 1 B per cell), so the roofline is the FP64 pipe's issue rate
    over the per-cell transition fan-in: a dnapsw cell has T_c = 13 transition groups over 8 states,
    5 of which are a state's second group.
      Forward (scaled linear domain): one FP64 FMA per group  -> peak = S * DFMA / T_c
      Viterbi (log domain, exact):    one FP64 add per group + one FP64 compare/select per second group
                                      -> peak = S / (T_c / DADD + n_2nd / DSETP_SEL)
    Pipe rates are measured on this pool's B200 by tools/pipe_peaks.cu (profiles/r01_pipe_peaks.json).
    The dominant kernel of the step (longest) is reported at top level.
    """
    peaks = {}
    p = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            peaks = json.load(f)
    pipe = {}
    pp = os.path.join(REPO, "profiles", "r01_pipe_peaks.json")
    if os.path.exists(pp):
        with open(pp) as f:
            pipe = {r["op"]: r["gops_per_s"] for r in json.load(f)["results"]}
    groups, seen = set(), {}
    for t in range(len(mj["src"])):
        a, b = int(mj["tin"][t]), int(mj["tout"][t])
        if a == 0 and b == 0 and mj["dst"][t] <= mj["src"][t]:
            continue
        kind = 0 if (a and b) else 1 if a else 2 if b else 3
        key = (int(mj["dst"][t]), kind, int(mj["src"][t]), (a, b))
        rank = seen.get(key, 0)
        seen[key] = rank + 1
        groups.add((int(mj["dst"][t]), kind, int(mj["src"][t]), rank))
    t_c = len(groups)
    n_2nd = t_c - len({g[0] for g in groups})
    S = mj["n_states"]
    dfma = pipe.get("dfma", 17895.0) * 1e9
    dadd = pipe.get("dadd", 17775.0) * 1e9
    dsel = pipe.get("dsetp_sel", 8475.0) * 1e9
    peak_fwd = S * dfma / t_c / 1e9
    peak_vit = S / (t_c / dadd + n_2nd / dsel) / 1e9
    ach_fwd, ach_vit = cells / fwd_ms / 1e6, cells / vit_ms / 1e6
    hbm = peaks.get("hbm_gbs", 6650.0)
    fwd = {"kernel": "mb_k_forward_lin", "bound": "issue (FP64 FMA pipe)", "achieved": ach_fwd, "peak": peak_fwd,
           "unit": "GCUPS", "frac": ach_fwd / peak_fwd, "ms_per_launch": fwd_ms}
    vit = {"kernel": "mb_k_viterbi", "bound": "issue (FP64 add + compare/select)", "achieved": ach_vit, "peak": peak_vit,
           "unit": "GCUPS", "frac": ach_vit / peak_vit, "ms_per_launch": vit_ms,
           "hbm": {"algorithmic_bytes_per_launch": cells / S * 1.0, "achieved_gbs": cells / S / vit_ms / 1e6,
                   "peak_gbs": hbm, "frac": cells / S / vit_ms / 1e6 / hbm,
                   "peak_source": "measured (MEASURED_PEAKS.json)" if peaks else "fallback"}}
    top = dict(vit if vit_ms >= fwd_ms else fwd)
    # DRAM bytes (read + write) of one launch of that kernel from the committed `ncu --set full` capture,
    # which profile_round.sh takes at this bench's default size (10 000 pairs of 1000 x 1000)
    top["traffic"] = None
    sp = os.path.join(REPO, "profiles", "r01_ncu_summary.json")
    if os.path.exists(sp) and abs(cells - 1001.0 * 1001.0 * 8 * 10000) < 1:
        with open(sp) as f:
            k = json.load(f).get("kernels", {}).get(top["kernel"], {})
        if k.get("dram_traffic_bytes") and k.get("pairs", 10000) == 10000:
            top["traffic"] = k["dram_traffic_bytes"]
            top["traffic_source"] = "ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum (profiles/r01_ncu_summary.json)"
    top["per_cell"] = {"transition_groups": t_c, "second_groups": n_2nd, "states": S}
    top["peak_source"] = "measured DFMA %.0f, DADD %.0f, DSETP+SEL %.0f Gop/s (profiles/r01_pipe_peaks.json)" % (dfma / 1e9, dadd / 1e9, dsel / 1e9)
    top["forward"] = fwd
    top["viterbi"] = vit
    return top


if __name__ == "__main__":
    main()
