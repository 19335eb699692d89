"""CPU test of bench.py's reference arm (`--impl reference`): the reference's own CPU implementation of the path (oracle/_ref/refdrv,
the reference compiled unmodified) timed on this box's host cores, printing ONE JSON line with the contract's keys -- and that the
product arm refuses to run without a device instead of falling back."""
import json
import os
import subprocess
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    if not os.path.exists(os.path.join(REPO, "oracle", "_ref", "refdrv")):
        pytest.skip("oracle/_ref/refdrv not built here")
    r = subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, cwd=REPO, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
                "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["unit"] == "GCUPS" and d["value"] > 0 and d["higher_is_better"] is True
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in d["config"]


def test_product_arm_needs_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r = subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), "--steps", "1", "--warmup", "0", "--no-configs", "--no-cpu-baseline"],
                       capture_output=True, text=True, cwd=REPO, timeout=600)
    assert r.returncode != 0      # no CPU fallback: the product path fails loudly without CUDA
