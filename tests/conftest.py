import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (test infrastructure; oracle/shapes_oracle.h)."""
    from oracle import binding
    binding.build()
    binding.lib()
    return binding


@pytest.fixture(scope="session")
def product_lib():
    """The product's C-ABI library, built in-tree (cross-compiles without a GPU)."""
    from shapes_b200 import _lib, build
    build.build_library()
    return _lib.load()


INT_COLS = ("pair_i", "pair_j", "key_i", "key_j", "feat_a", "feat_b", "flip")
F64_COLS = (("normal_x", "normal_y", "center_x", "center_y", "depth")
            + tuple(f"j_np{q}" for q in range(6)) + ("b_np", "ra_x", "ra_y", "rb_x", "rb_y", "rn_x", "rn_y")
            + tuple(f"j_f{q}" for q in range(6)) + ("b_f", "inv_eff_np", "inv_eff_f"))
REL_TOL = 1e-9   # north_star: contacts, normals, depths, Jacobians within 1e-9 relative


def assert_frames_match(got, want, exact=True, cols=None):
    """Pair set / keys bit-exact; reals bit-exact by default (every op is IEEE and un-fused on
    both sides), never looser than the 1e-9 relative bar."""
    for k in INT_COLS:
        if cols is not None and k not in cols:
            continue
        if k not in got:
            continue
        g, w = np.asarray(got[k]), np.asarray(want[k])
        assert g.shape == w.shape, f"{k}: {g.shape} vs {w.shape}"
        assert np.array_equal(g, w), f"{k} differs at {np.nonzero(g != w)[0][:5]}"
    for k in F64_COLS:
        if cols is not None and k not in cols:
            continue
        if k not in got:
            continue
        g, w = np.asarray(got[k]), np.asarray(want[k])
        assert g.shape == w.shape, f"{k}: {g.shape} vs {w.shape}"
        same = (g == w) | (np.isnan(g) & np.isnan(w))
        if exact:
            assert same.all(), f"{k}: {np.count_nonzero(~same)} values differ, first at {np.nonzero(~same)[0][:5]}"
        else:
            err = np.abs(g - w) / np.maximum(np.abs(w), 1e-300)
            err = np.where(same, 0.0, err)
            assert np.all(err <= REL_TOL), f"{k}: max rel err {err.max()}"
