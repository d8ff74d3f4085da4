"""The C++ host-side mirror (include/shapes_b200.hpp) over the C ABI.

CPU: it compiles against the header + library and refuses to run without a GPU (exit 3, no
fallback).  GPU: the reference's own bench fixtures (KAT-1, KAT-3, config 1) through
culledKeys / prepareFrame / constraintGen."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "host_mirror_test")


def build_exe(product_lib):
    from shapes_b200 import build
    return build.build_host_test()


def test_cpp_host_mirror_compiles_and_fails_loudly_without_gpu(product_lib):
    exe = build_exe(product_lib)
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu test")
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 3, (r.returncode, r.stdout, r.stderr)
    assert "shapes_b200 error -2" in r.stdout


@pytest.mark.gpu
def test_cpp_host_mirror_reference_fixtures(product_lib):
    exe = build_exe(product_lib)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, (r.returncode, r.stdout, r.stderr)
    assert "all checks passed" in r.stdout
