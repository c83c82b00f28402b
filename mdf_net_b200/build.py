"""Build the sm_100a shared library (mdf_net_b200/libmdf_b200.so) in-tree with nvcc.

`python -m mdf_net_b200.build [--force] [--verbose] [--tuning]`.  nvcc cross-compiles without a GPU; the
built .so is git-ignored but travels to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
INCLUDE = os.path.join(os.path.dirname(PKG), "include")
LIB_PATH = os.path.join(PKG, "libmdf_b200.so")
TUNING_LIB_PATH = os.path.join(PKG, "libmdf_b200_tuning.so")   # -DMDF_TUNING: shape / diagnostic variants for tools/ (never the product)
SOURCES = ("mdf_cost_volume.cu", "mdf_head.cu", "mdf_backward.cu", "mdf_hypos.cu", "mdf_prob_head.cu", "mdf_filter.cu", "mdf_fpn.cu")
ARCH_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found (set NVCC=...)")


def sources() -> list:
    return [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]


def is_fresh(path: str = LIB_PATH) -> bool:
    if not os.path.exists(path):
        return False
    t = os.path.getmtime(path)
    deps = [os.path.join(d, f) for d, _, fs in os.walk(CSRC) for f in fs] + [os.path.join(INCLUDE, f) for f in os.listdir(INCLUDE)]
    return all(os.path.getmtime(d) <= t for d in deps)


def build_library(force: bool = False, verbose: bool = False, tuning: bool = False) -> str:
    out = TUNING_LIB_PATH if tuning else LIB_PATH
    if is_fresh(out) and not force:
        return out
    cmd = [_nvcc(), "-O3", "-std=c++17", "--threads", "0", *ARCH_FLAGS, "-lineinfo", "-shared", "-Xcompiler", "-fPIC,-fvisibility=hidden",
           "-I", INCLUDE, *(["-DMDF_TUNING"] if tuning else []), "-o", out + ".tmp", *sources()]
    if verbose:
        cmd[1:1] = ["-Xptxas", "-v"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose:
        print(r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr)
    os.replace(out + ".tmp", out)
    return out


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="--verbose" in sys.argv, tuning="--tuning" in sys.argv))
