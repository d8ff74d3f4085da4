"""Build the sm_100a shared library in-tree (shapes_b200/lib/libshapes_b200.so).

nvcc cross-compiles without a GPU.  -fmad=false belongs to the contract of the
library (the reference never fuses multiply-add); the kernels additionally use
explicit __dmul_rn/__dadd_rn so the results do not depend on it.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "shapes_b200", "csrc", "shapes_b200.cu")
INC = os.path.join(ROOT, "include")
LIB_DIR = os.path.join(ROOT, "shapes_b200", "lib")
LIB = os.path.join(LIB_DIR, "libshapes_b200.so")


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def nvcc_command(out: str = LIB, extra: list[str] | None = None) -> list[str]:
    return [
        nvcc_path(), "-O3", "-std=c++17",
        "-gencode", "arch=compute_100a,code=sm_100a",
        "-lineinfo", "-fmad=false",
        "-Xcompiler", "-fPIC", "-Xcompiler", "-ffp-contract=off", "-shared",
        "-I", INC, "-I", os.path.dirname(SRC), "-o", out, SRC, "-ldl",
    ] + (extra or [])


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [SRC, os.path.join(os.path.dirname(SRC), "world_step.cuh"),
            os.path.join(INC, "shapes_b200.h"), os.path.join(INC, "shapes_sincos.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build_library(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(LIB_DIR, exist_ok=True)
    if force or needs_build():
        cmd = nvcc_command()
        if verbose:
            print(" ".join(cmd), file=sys.stderr)
        subprocess.run(cmd, check=True)
    return LIB


HOST_TEST_SRC = os.path.join(ROOT, "tests", "cpp", "host_mirror_test.cpp")
HOST_TEST_EXE = os.path.join(ROOT, "tests", "cpp", "host_mirror_test")


def build_host_test(force: bool = False) -> str:
    """g++ build of the C++ host-mirror test program against include/shapes_b200.hpp."""
    build_library()
    deps = [HOST_TEST_SRC, os.path.join(INC, "shapes_b200.hpp"), os.path.join(INC, "shapes_b200.h"), LIB]
    if force or not os.path.exists(HOST_TEST_EXE) or \
            any(os.path.getmtime(d) > os.path.getmtime(HOST_TEST_EXE) for d in deps):
        subprocess.run(["g++", "-std=c++17", "-O2", "-Wall", "-I", INC, HOST_TEST_SRC, "-o", HOST_TEST_EXE,
                        "-L", LIB_DIR, "-lshapes_b200", "-Wl,-rpath," + LIB_DIR,
                        "-Wl,-rpath,$ORIGIN/../../shapes_b200/lib"], check=True)
    return HOST_TEST_EXE


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose=True))
