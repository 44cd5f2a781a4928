"""Build tools/ozaki/libozaki.so in-tree with nvcc for sm_100a (microbenchmark library, not part of libgaplac_b200.so)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "ozaki_syrk.cu")
LIB = os.path.join(HERE, "libozaki.so")
HDR = os.path.join(os.path.dirname(os.path.dirname(HERE)), "gaplac_b200", "csrc")
FLAGS = ["-O3", "-std=c++17", "-lineinfo", "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-Xcompiler", "-fPIC", "-I", HDR]


def build(force: bool = False) -> str:
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < max(os.path.getmtime(SRC), os.path.getmtime(os.path.join(HDR, "int8_syrk.cuh"))):
        subprocess.check_call(["nvcc", *FLAGS, "-o", LIB, SRC])
    return LIB


if __name__ == "__main__":
    print(build(force=True))
