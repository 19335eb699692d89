"""CPU tests of the boundary: the library builds, loads, exports every declared symbol, and fails
loudly (no CPU fallback) when there is no CUDA device."""
import ctypes
import os
import re

import pytest

from helpers import REPO


def test_library_exports_every_declared_symbol():
    from machineboss_b200 import capi
    L = capi.lib()
    hdr = open(os.path.join(REPO, "include", "machineboss_b200.h")).read()
    declared = set(re.findall(r"\b(mb_[a-z_0-9]+)\s*\(", hdr))
    assert declared == set(capi.SYMBOLS)
    for s in declared:
        assert hasattr(L, s), s
    assert L.mb_version() >= 100


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from machineboss_b200 import capi
    with pytest.raises(capi.MachineBossError, match="CUDA"):
        capi.Machine(1, 0, 0, [], [], [], [], [])


def test_product_does_not_import_oracle():
    """The product path must not touch oracle/ (only tests, smoke and the CPU baseline may)."""
    pkg = os.path.join(REPO, "machineboss_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".h", ".cpp", ".cuh")):
                txt = open(os.path.join(root, f)).read()
                assert "oracle" not in txt.lower() or f == "README.md", os.path.join(root, f)
