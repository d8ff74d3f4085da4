"""BASELINE.json configs 4 (4M mixed boxes/polygons) and 5 (1M-polygon Gaussian blob) at FULL size.
The oracle cannot walk such worlds pair by pair in test time, so the frame is checked through
size-independent properties plus the oracle on bounded samples:
  * pairs: the whole list, set and order, equals the oracle's Grid.culledKeys restatement (linear in N); also i > j,
    strictly descending, never static/static, and for sampled shapes the partner set equals a brute-force overlap
    test of the device's own AABBs against ALL other shapes (an oracle-independent completeness check);
  * AABBs of sampled shapes equal the oracle's (moveShapes + toAabb);
  * contact rows: keys follow the pair list in order, unit normals, the exact sign relations between the
    Jacobian halves / restitution normal / flip, b_f = 0;
  * the oracle's prepareFrame + constraintGen on a random SUBSET of the device's pairs reproduces those
    pairs' rows bit for bit;
  * a second frame on the same inputs is identical (determinism).
(Config 3 at full size is compared row by row with the oracle in test_gpu_parity.py.)"""
import numpy as np
import pytest

from shapes_b200 import scenes

pytestmark = pytest.mark.gpu

ROW_F64 = (("normal_x", "normal_y", "center_x", "center_y", "depth")
           + tuple(f"j_np{q}" for q in range(6)) + ("b_np", "ra_x", "ra_y", "rb_x", "rb_y", "rn_x", "rn_y")
           + tuple(f"j_f{q}" for q in range(6)) + ("inv_eff_np", "inv_eff_f"))


def check_full_size(oracle, w, max_pairs, max_contacts, seed):
    from shapes_b200.engine import Engine
    rng = np.random.default_rng(seed)
    c, s = oracle.cos_sin(w.rot)
    n = w.n_slots
    want = ("pairs", "contacts", "constraints", "aabb")
    with Engine(w, max_pairs=max_pairs, max_contacts=max_contacts) as eng:
        fr = eng.frame(cos_sin=(c, s), want=want)
        cols = {k: np.array(fr[k]) for k in fr.cols}
        fr2 = eng.frame(cos_sin=(c, s), want=want)
        for k in cols:                                                   # determinism
            a, b = cols[k], np.asarray(fr2[k])
            assert a.shape == b.shape and ((a == b) | ((a != a) & (b != b))).all(), k
    pi, pj = cols["pair_i"].astype(np.int64), cols["pair_j"].astype(np.int64)
    P = len(pi)
    # ---- pairs
    assert (pi > pj).all()
    key = pi * (n + 1) + pj
    assert (np.diff(key) < 0).all()                                          # strictly descending (i, j)
    static = (w.inv_lin == 0.0) & (w.inv_rot == 0.0)
    assert not (static[pi] & static[pj]).any()
    # ---- AABBs of sampled shapes against the oracle, completeness of their partner sets
    wx, wy, nx, ny = oracle.move_shapes(w, c, s)
    boxes = oracle.aabbs(w, wx, wy)
    for got, ref in zip(("aabb_min_x", "aabb_max_x", "aabb_min_y", "aabb_max_y"), boxes):
        assert np.array_equal(cols[got], ref), got
    # ---- the WHOLE pair list, set and order, against the oracle's restatement of Grid.culledKeys (Grid.hs:67-100),
    # which is linear in the number of shapes
    gi, gj = oracle.culled_keys_grid(w, boxes, oracle.is_static(w))
    assert len(gi) == P and np.array_equal(gi, cols["pair_i"]) and np.array_equal(gj, cols["pair_j"])
    x0, x1, y0, y1 = boxes
    starts = np.searchsorted(-pi, -np.arange(n, -1, -1))                     # rows of shape i: pairs are grouped by i, descending
    for i in rng.integers(1, n, 60):
        overlap = ~((x0[:i] > x1[i]) | (x1[:i] < x0[i])) & ~((y0[:i] > y1[i]) | (y1[:i] < y0[i]))   # aabbCheck (Aabb.hs:69-78)
        if static[i]:
            overlap &= ~static[:i]
        expect = np.nonzero(overlap)[0][::-1]
        lo, hi = starts[n - i], starts[n - i + 1]
        assert (pi[lo:hi] == i).all() and np.array_equal(pj[lo:hi], expect), i
    # ---- contact rows
    ki, kj = cols["key_i"].astype(np.int64), cols["key_j"].astype(np.int64)
    C = len(ki)
    rk = ki * (n + 1) + kj
    assert (np.diff(rk) <= 0).all()                                          # descending (i, j), <= 2 rows per pair
    row_pair = np.searchsorted(-key, -rk)                                    # every row belongs to a broadphase pair
    assert (key[row_pair] == rk).all()
    assert np.bincount(row_pair, minlength=P).max() <= 2
    flip = cols["flip"].astype(bool)
    nrm = np.hypot(cols["normal_x"], cols["normal_y"])
    assert np.abs(nrm - 1.0).max() < 1e-12 and np.isfinite(cols["depth"]).all()
    for q in (0, 1):                                                         # J = (ja, jb) with jb's linear part = -ja's, exactly
        assert np.array_equal(cols[f"j_np{q + 3}"], -cols[f"j_np{q}"]) and np.array_equal(cols[f"j_f{q + 3}"], -cols[f"j_f{q}"])
    sign = np.where(flip, -1.0, 1.0)
    assert np.array_equal(cols["j_np3"], sign * cols["normal_x"]) and np.array_equal(cols["j_np4"], sign * cols["normal_y"])
    assert np.array_equal(cols["rn_x"], sign * cols["normal_x"]) and np.array_equal(cols["rn_y"], sign * cols["normal_y"])
    assert np.array_equal(cols["j_f3"], sign * cols["normal_y"]) and np.array_equal(cols["j_f4"], -(sign * cols["normal_x"]))
    assert np.array_equal(cols["ra_x"], cols["center_x"] - w.pos_x[ki]) and np.array_equal(cols["rb_y"], cols["center_y"] - w.pos_y[kj])
    # ---- the oracle on a random subset of the pairs: those pairs' rows, bit for bit
    pick = np.sort(rng.choice(P, size=min(P, 40_000), replace=False))
    emin, emax = oracle.hull_extents(w)
    ref = oracle.contacts(w, pi[pick].astype(np.int32), pj[pick].astype(np.int32), wx, wy, nx, ny, emin, emax, 0.01, 0.01, 0.02)
    rows = np.nonzero(np.isin(row_pair, pick))[0]
    assert len(rows) == len(ref["key_i"]) and len(rows) > 1000
    for k in ("key_i", "key_j", "feat_a", "feat_b", "flip") + ROW_F64:
        a, b = cols[k][rows], ref[k]
        assert ((a == b) | ((a != a) & (b != b))).all(), k
    return P, C


def test_config4_mixed_4m(oracle):
    w = scenes.mixed_polygons(4_000_000)
    P, C = check_full_size(oracle, w, max_pairs=4_400_000, max_contacts=5_200_000, seed=4)
    assert P == 3_962_437 and C == 4_619_073


def test_config5_gaussian_blob_1m(oracle):
    w = scenes.gaussian_blob(1_000_000)
    P, C = check_full_size(oracle, w, max_pairs=2_000_000, max_contacts=2_400_000, seed=5)
    assert P > 1_500_000 and C > 1_800_000
