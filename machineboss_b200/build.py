"""In-tree build of the CUDA shared library (libmachineboss_b200.so) with nvcc for sm_100a."""
from __future__ import annotations

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libmachineboss_b200.so")
SOURCES = ["mb_api.cu", "mb_generic.cu", "mb_jit.cu", "mb_wide.cu", "mb_lane.cu", "mb_big.cu", "mb_group.cu", "mb_col.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]


def nvcc_path() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the CUDA toolkit is required to build machineboss_b200")


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(os.path.dirname(HERE), "include", "machineboss_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every .cu under csrc/ into one shared library next to this file."""
    if not force and not needs_build():
        return LIB
    cmd = [nvcc_path()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) \
        + [os.path.join(CSRC, s) for s in SOURCES] + ["-o", LIB, "-ldl"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + r.stdout + r.stderr)
    if verbose:
        print(r.stderr)
    return LIB


CLI = os.path.join(HERE, "boss_b200")


def build_host(force: bool = False) -> str:
    """Compile the C++ host mirror's CLI (host/boss_b200_cli.cpp) against the shared library."""
    src = os.path.join(HERE, "host", "boss_b200_cli.cpp")
    deps = [src, os.path.join(HERE, "host", "boss_b200.h"), os.path.join(HERE, "host", "boss_b200_fit.h"), os.path.join(HERE, "host", "boss_b200_ingest.h"), os.path.join(HERE, "host", "mbjson.h"), LIB]
    if not force and os.path.exists(CLI) and all(os.path.getmtime(d) <= os.path.getmtime(CLI) for d in deps):
        return CLI
    build()
    cmd = ["g++", "-std=c++14", "-O2", "-Wall", src, "-o", CLI, "-L" + HERE, "-lmachineboss_b200", "-lz", "-Wl,-rpath," + HERE]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("g++ failed:\n" + r.stdout + r.stderr)
    return CLI


if __name__ == "__main__":
    print(build(force=True, verbose=True))
    print(build_host(force=True))
