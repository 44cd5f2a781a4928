"""Build libgaplac_b200.so in-tree with nvcc for sm_100a (no JIT cache: the .so travels with the repo).

    python -m gaplac_b200.build [--force]
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "..", "build", "obj")
LIB = os.path.join(HERE, "libgaplac_b200.so")

NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-Wno-format-truncation"]
CU_SOURCES = ["api.cu", "lml_batched.cu", "lml_lockstep.cu", "lml_grad_lockstep.cu", "mcmc.cu", "kbuild.cu", "big.cu", "trail_int8.cu", "predict.cu"]
CPP_SOURCES = ["program.cpp", "multi.cpp"]


def _newest_header() -> float:
    t = 0.0
    for root in (CSRC, os.path.join(HERE, "..", "include")):
        for f in os.listdir(root):
            if f.endswith((".h", ".cuh")):
                t = max(t, os.path.getmtime(os.path.join(root, f)))
    return t


def _compile(src: str, force: bool, hdr_time: float) -> tuple[str, bool]:
    path = os.path.join(CSRC, src)
    obj = os.path.join(OBJ, src + ".o")
    if not force and os.path.exists(obj) and os.path.getmtime(obj) > max(os.path.getmtime(path), hdr_time):
        return obj, False
    if src.endswith(".cu"):
        cmd = ["nvcc", *NVCC_FLAGS, "-c", path, "-o", obj]
    else:
        cmd = ["g++", "-O2", "-std=c++17", "-fPIC", "-c", path, "-o", obj]
    subprocess.check_call(cmd)
    return obj, True


def build_variant(tag: str, defs: list[str]) -> str:
    """Experiment builds: libgaplac_b200_<tag>.so with extra -D flags (select with GAPLAC_B200_LIB=<path>)."""
    obj_dir = os.path.join(OBJ, tag)
    os.makedirs(obj_dir, exist_ok=True)
    objs = []
    for src in CU_SOURCES + CPP_SOURCES:
        path, obj = os.path.join(CSRC, src), os.path.join(obj_dir, src + ".o")
        if src.endswith(".cu"):
            subprocess.check_call(["nvcc", *NVCC_FLAGS, *defs, "-c", path, "-o", obj])
        else:
            subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-c", path, "-o", obj])
        objs.append(obj)
    lib = os.path.join(HERE, f"libgaplac_b200_{tag}.so")
    subprocess.check_call(["nvcc", "-shared", "-o", lib, *objs, "-gencode", "arch=compute_100a,code=sm_100a"])
    return lib


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    hdr_time = _newest_header()
    with ThreadPoolExecutor(max_workers=10) as ex:
        res = list(ex.map(lambda s: _compile(s, force, hdr_time), CU_SOURCES + CPP_SOURCES))
    objs = [r[0] for r in res]
    if force or any(r[1] for r in res) or not os.path.exists(LIB):
        subprocess.check_call(["nvcc", "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a"])
        if verbose:
            print("linked", LIB)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
