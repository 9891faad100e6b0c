"""Builds romis_b200/libromis_gpu.so in-tree with nvcc for sm_100a (no JIT cache: the .so must travel).

    python -m romis_b200.build [--force]

One translation unit per pass kernel, compiled in parallel, then linked into one shared library.
Flags that are part of the parity contract: -fmad=false (no fused multiply-add on the
device) and -ffp-contract=off for host code; never --use_fast_math.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libromis_gpu.so")
SOURCES = ["romis_gpu.cu", "k_initial.cu", "k_temporal.cu", "k_spatial.cu", "k_rmis.cu", "k_misc.cu", "bvh.cpp"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-fmad=false",
         "-Xcompiler", "-fPIC,-ffp-contract=off", "-ccbin", "g++",
         "-I" + os.path.join(ROOT, "include"), "-I" + CSRC]
if os.environ.get("ROMIS_LEAF_MAX"):   # tuning knob, see bvh.cpp
    FLAGS.append("-DROMIS_LEAF_MAX=" + os.environ["ROMIS_LEAF_MAX"])
if os.environ.get("ROMIS_DEFS"):       # tuning knob: extra -D flags, space separated
    FLAGS += ["-D" + d for d in os.environ["ROMIS_DEFS"].split()]
if os.environ.get("ROMIS_LIB_OUT"):    # tuning builds go next to the objects, the product stays libromis_gpu.so
    LIB = os.environ["ROMIS_LIB_OUT"]


def _deps():
    d = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    d += [os.path.join(ROOT, "include", f) for f in os.listdir(os.path.join(ROOT, "include"))]
    return d


def _compile(src: str, verbose: bool) -> str:
    obj = os.path.join(OBJ, os.path.splitext(src)[0] + ".o")
    cmd = [NVCC] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", os.path.join(CSRC, src), "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed on {src}:\n{r.stdout}\n{r.stderr}")
    if verbose:
        sys.stderr.write(r.stderr)
    return obj


def build(force: bool = False, verbose: bool = False) -> str:
    newest = max(os.path.getmtime(p) for p in _deps())
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= newest:
        return LIB
    os.makedirs(OBJ, exist_ok=True)
    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        objs = list(ex.map(lambda s: _compile(s, verbose), SOURCES))
    cmd = [NVCC, "-shared", "-o", LIB, "-ccbin", "g++", "-gencode", "arch=compute_100a,code=sm_100a"] + objs + ["-lcudart"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
