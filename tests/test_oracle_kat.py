"""The oracle against every vector and property that can be pinned for this path
(tests/golden/README.md).  CPU only."""
import ctypes as C
import json
import os

import numpy as np
import pytest

from shapes_b200 import scenes

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def kat():
    with open(os.path.join(GOLDEN, "kat.json")) as f:
        return json.load(f)


def test_kat1_test_opt_boxes(oracle, kat):
    """S.contact testOptBoxes (shapes/bench/Physics/Contact/Benchmark.hs:16-27)."""
    r = oracle.frame(scenes.test_opt_boxes())
    want = kat["kat1_testOptBoxes"]
    assert list(zip(r["pair_i"], r["pair_j"])) == [tuple(p) for p in want["pairs"]]
    assert len(r["key_i"]) == len(want["contacts"])
    for k, c in enumerate(want["contacts"]):
        assert [r["key_i"][k], r["key_j"][k]] == c["key"]
        assert [r["feat_a"][k], r["feat_b"][k]] == c["feat"]
        assert r["flip"][k] == c["flip"]
        assert [r["normal_x"][k], r["normal_y"][k]] == c["normal"]
        assert [r["center_x"][k], r["center_y"][k]] == c["center"]
        assert r["depth"][k] == c["depth"]


def test_kat2_solve_constraint(oracle, kat):
    """solveConstraint testConstraint testObjPair (bench/Physics/Constraint/Benchmark.hs:11-35)."""
    k = kat["kat2_solveConstraint"]
    # the Jacobian itself, through the NonPenetration generator: n=(0,1), xa=(0,0), xb=(0,4), p=(0,2)
    j = np.array(k["j"]); im = np.array(k["inv_mass6"]); v = np.array(k["vel6_in"])
    f = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    oracle.lib().orc_solve_constraint(f(j), C.c_double(k["b"]), f(im), f(v))
    assert v.tolist() == k["vel6_out"]


def test_kat2_jacobian_through_generator(oracle):
    """testConstraint's Jacobian (Benchmark.hs:11-19) from NonPenetration.jacobian: the same
    geometry (penetrated body at (0,0), penetrator at (0,4), n=(0,1), contact on y=2) through
    the generators.  Equal depths make the pair a Flip (SAT.hs:248), so the penetrated body is
    the lower-key one and the J halves come back swapped (Constraint.hs:96-98)."""
    from shapes_b200.world import World, rectangle_vertices
    w = World.from_objects([(rectangle_vertices(4, 4), (0.0, 0.0), 0.0, (1.0, 0.5)),
                            (rectangle_vertices(4, 4), (0.0, 4.0), 0.0, (1.0, 0.5))])
    r = oracle.frame(w)
    assert len(r["key_i"]) == 2 and set(r["flip"].tolist()) == {1}
    xa, xb = (0.0, 0.0), (0.0, 4.0)
    for k in range(2):
        n = (r["normal_x"][k], r["normal_y"][k])
        assert n == (0.0, 1.0)
        p = (r["center_x"][k], r["center_y"][k])
        assert p[1] == 2.0 and r["depth"][k] == 0.0
        ja = [-n[0], -n[1], (xa[0] - p[0]) * n[1] - (xa[1] - p[1]) * n[0]]
        jb = [n[0], n[1], (p[0] - xb[0]) * n[1] - (p[1] - xb[1]) * n[0]]
        assert [r[f"j_np{q}"][k] for q in range(6)] == jb + ja
        assert r["b_np"][k] == 0.0 and r["b_f"][k] == 0.0
        # Restitution: radii from the unflipped pair (i = upper box), normal negated for Flip
        assert (r["ra_x"][k], r["ra_y"][k]) == (p[0] - xb[0], p[1] - xb[1])
        assert (r["rb_x"][k], r["rb_y"][k]) == (p[0] - xa[0], p[1] - xa[1])
        assert (r["rn_x"][k], r["rn_y"][k]) == (-0.0, -1.0)


KAT678 = [("kat6_triangle_into_box_same", False, 0.0), ("kat7_triangle_into_box_flip", True, 0.0),
          ("kat8_triangle_lifted_single_point", False, 1.0)]


def assert_kat_rows(r, want, where):
    """Every pinned column of a KAT-6/7/8 entry, signed zeros included."""
    def bits(x):
        return np.float64(x).tobytes()
    assert list(zip(r["pair_i"].tolist(), r["pair_j"].tolist())) == [tuple(p) for p in want["pairs"]], where
    assert len(r["key_i"]) == len(want["contacts"]), where
    for k, c in enumerate(want["contacts"]):
        assert [int(r["key_i"][k]), int(r["key_j"][k])] == c["key"] and [int(r["feat_a"][k]), int(r["feat_b"][k])] == c["feat"], (where, k)
        assert int(r["flip"][k]) == c["flip"], (where, k)
        pairs = [("normal_x", c["normal"][0]), ("normal_y", c["normal"][1]), ("center_x", c["center"][0]),
                 ("center_y", c["center"][1]), ("depth", c["depth"]), ("b_np", c["b_np"]),
                 ("ra_x", c["ra"][0]), ("ra_y", c["ra"][1]), ("rb_x", c["rb"][0]), ("rb_y", c["rb"][1]),
                 ("rn_x", c["rn"][0]), ("rn_y", c["rn"][1]), ("inv_eff_np", c["inv_eff_np"]), ("inv_eff_f", c["inv_eff_f"])]
        pairs += [(f"j_np{q}", c["j_np"][q]) for q in range(6)] + [(f"j_f{q}", c["j_f"][q]) for q in range(6)]
        for col, val in pairs:
            assert bits(r[col][k]) == bits(val), (where, k, col, r[col][k], val)


@pytest.mark.parametrize("name,triangle_is_a,lift", KAT678)
def test_kat678_triangle_on_box(oracle, kat, name, triangle_is_a, lift):
    """KAT-6 (Same, ClipLeft, two points, active Baumgarte, non-trivial radii), KAT-7 (the same pair with the keys
    swapped: Flip, Jacobian halves swapped, restitution normal negated), KAT-8 (the clipped point lies above the
    face: the third clip removes it).  Hand traces: tests/golden/README.md."""
    want = kat[name]
    w, c, s = scenes.kat_triangle_on_box(triangle_is_a, lift)
    r = oracle.frame(w, c, s, broadphase="aabb", **want["behaviour"])
    if name.startswith("kat6"):
        t, b = want["world_vertices"]["triangle"], want["world_vertices"]["box"]
        assert list(zip(r["world_x"].tolist(), r["world_y"].tolist())) == [tuple(v) for v in t + b]
    assert_kat_rows(r, want, name)


def test_kat9_box_on_hexagon(oracle, kat):
    """KAT-9: a six-vertex hull with 3-4-5 edges under a box -- depth tie -> Flip, ClipRight + ClipLeft + the NaN-driven
    ClipNone, two points, live Baumgarte term.  Hand trace: tests/golden/README.md."""
    want = kat["kat9_box_on_hexagon"]
    w, c, s = scenes.kat_box_on_hexagon()
    r = oracle.frame(w, c, s, broadphase="aabb", **want["behaviour"])
    assert_kat_rows(r, want, "kat9")


def test_kat3_broadphase_knife_edge(oracle, kat):
    """Aabb.culledKeys testWorld (bench/Physics/Broadphase/Benchmark.hs:50-52)."""
    for name, spacing in (("spacing0", 0.0), ("spacing1", 1.0)):
        w = scenes.broadphase_bench_world(spacing=spacing)
        r = oracle.frame(w, broadphase="aabb")
        want = kat["kat3_broadphase"][name]
        assert len(r["pair_i"]) == want["n_pairs"]
        assert [r["pair_i"][0], r["pair_j"][0]] == want["first"]


@pytest.mark.parametrize("n", list(range(0, 31)))
def test_unordered_pairs_property(oracle, n):
    """|unorderedPairs n| = n(n-1)/2 (shapes/test/Physics/Broadphase/AabbSpec.hs:8-11),
    in the order (n-1,n-2), (n-1,n-3), ..., (1,0) (Aabb.hs:155-163)."""
    xs, ys = oracle.unordered_pairs(n)
    assert len(xs) == n * (n - 1) // 2
    want = [(x, y) for x in range(n - 1, 0, -1) for y in range(x - 1, -1, -1)]
    assert list(zip(xs.tolist(), ys.tolist())) == want


def test_template_spec_properties(oracle):
    """dotV2 / mul2x2x2 == plain left-to-right arithmetic
    (shapes-math/test/Shapes/Linear/TemplateSpec.hs:25-35)."""
    rng = np.random.default_rng(7)
    for _ in range(200):
        a, b, c, d = (float(x) for x in rng.normal(size=4) * 10.0 ** rng.integers(-3, 4))
        assert oracle.lib().orc_dot_v2(a, b, c, d) == (a * c) + (b * d)
        m = rng.normal(size=4); k = rng.normal(size=4); out = np.zeros(4)
        f = lambda x: x.ctypes.data_as(C.POINTER(C.c_double))
        oracle.lib().orc_mul2x2x2(f(m), f(k), f(out))
        want = [(m[0] * k[0]) + (m[1] * k[2]), (m[0] * k[1]) + (m[1] * k[3]),
                (m[2] * k[0]) + (m[3] * k[2]), (m[2] * k[1]) + (m[3] * k[3])]
        assert out.tolist() == want


@pytest.mark.parametrize("seed,side", [(1, 8.0), (2, 19.0), (3, 60.0)])
def test_grid_equals_aabb_culled_keys(oracle, seed, side):
    """Grid.culledKeys == Aabb.culledKeys, set and order, inside and outside the 20x20 grid
    (updateWorld calls the grid variant, Engine/Main.hs:75)."""
    w = scenes.random_polygons(400, density=400 / (side * side), static_frac=0.1, config=50 + seed)
    w.pos_x -= side / 2.0
    w.pos_y -= side / 2.0
    c, s = oracle.cos_sin(w.rot)
    wx, wy, _, _ = oracle.move_shapes(w, c, s)
    boxes = oracle.aabbs(w, wx, wy)
    st = oracle.is_static(w)
    a = oracle.culled_keys_aabb(w, boxes, st)
    g = oracle.culled_keys_grid(w, boxes, st)
    sw = oracle.culled_keys_sweep(w, boxes, st)
    assert len(a[0]) > 20
    assert np.array_equal(a[0], g[0]) and np.array_equal(a[1], g[1])
    assert np.array_equal(a[0], sw[0]) and np.array_equal(a[1], sw[1])


def test_sweep_equals_aabb_with_deletes_and_nan(oracle):
    w = scenes.random_polygons(600, density=3.0, static_frac=0.2, config=77)
    w.delete([5, 17, 300, 599])
    w.pos_x[40] = w.pos_y[40] = np.nan   # a NaN AABB "overlaps" everything (boundsOverlap, Aabb.hs:69-72)
    w.pos_x[42] = np.nan                 # NaN on one axis only: the other axis still filters
    w.pos_y[41] = np.inf
    r_a = oracle.frame(w, broadphase="aabb")
    r_s = oracle.frame(w, broadphase="sweep")
    assert np.array_equal(r_a["pair_i"], r_s["pair_i"]) and np.array_equal(r_a["pair_j"], r_s["pair_j"])
    n40 = (r_a["pair_i"] == 40).sum() + (r_a["pair_j"] == 40).sum()
    n_static = int(oracle.is_static(w)[w.alive == 1].sum())
    assert n40 == (w.alive.sum() - 1) - (n_static - 1 if oracle.is_static(w)[40] else 0)
    assert 0 < (r_a["pair_i"] == 42).sum() + (r_a["pair_j"] == 42).sum() < 200
    for dead in (5, 17, 300, 599):
        assert dead not in r_a["pair_i"] and dead not in r_a["pair_j"]


def test_stacks_scene_frame0(oracle):
    """Stacks.makeScene (30,30) 0 (BASELINE config 1): descending keys, contact invariants."""
    w = scenes.stacks_scene()
    r = oracle.frame(w, broadphase="aabb")
    key = r["pair_i"].astype(np.int64) << 32 | r["pair_j"]
    assert np.all(key[:-1] > key[1:])
    assert len(r["key_i"]) > 0
    nn = r["normal_x"] ** 2 + r["normal_y"] ** 2
    assert np.allclose(nn, 1.0)
    ck = (r["key_i"].astype(np.int64) << 32) | r["key_j"]
    assert np.all(ck[:-1] >= ck[1:])
    # the grid variant agrees on the reference's own world
    g = oracle.frame(w, broadphase="grid")
    assert np.array_equal(g["pair_i"], r["pair_i"]) and np.array_equal(g["pair_j"], r["pair_j"])


def test_golden_regression(oracle):
    """The oracle against its own committed output (tests/golden/make_golden.py)."""
    z = np.load(os.path.join(GOLDEN, "oracle_polygons64.npz"))
    w = scenes.random_polygons(64, density=2.0, static_frac=0.1, config=99)
    r = oracle.frame(w, z["cos"], z["sin"], broadphase="aabb")
    for k in z.files:
        if k in ("cos", "sin"):
            continue
        assert np.array_equal(r[k], z[k], equal_nan=True), k


def test_warm_join_desc_zip_vector_semantics(oracle):
    """descZipVector (Utils/Descending.hs:47-71) on a hand example: descending keys on both
    sides; equal key => cached Lagrangians (useCache), otherwise ContactLagrangian 0 0 (newCache)."""
    this = {"key_i": np.array([9, 9, 7, 5, 5, 2]), "key_j": np.array([3, 3, 1, 4, 4, 0]),
            "feat_a": np.array([2, 1, 0, 3, 3, 1]), "feat_b": np.array([0, 0, 2, 1, 0, 1])}
    that = {"key_i": np.array([9, 8, 7, 5, 1]), "key_j": np.array([3, 2, 1, 4, 0]),
            "feat_a": np.array([1, 0, 0, 3, 0]), "feat_b": np.array([0, 0, 2, 0, 0])}
    np_, f_, hit = oracle.warm_join(this, that, np.array([1., 2, 3, 4, 5]), np.array([10., 20, 30, 40, 50]))
    assert hit.tolist() == [0, 1, 1, 0, 1, 0]
    assert np_.tolist() == [0, 1, 3, 0, 4, 0] and f_.tolist() == [0, 10, 30, 0, 40, 0]
    # empty cache / empty frame
    empty = {k: np.zeros(0, np.int32) for k in this}
    assert oracle.warm_join(this, empty, np.zeros(0), np.zeros(0))[2].tolist() == [0] * 6
    assert len(oracle.warm_join(empty, that, np.ones(5), np.ones(5))[2]) == 0


def test_kat4_kat5_circles(oracle, kat):
    """Circle.contact and CircleVsHull/GJK on hand-derived vectors (tests/golden/README.md)."""
    from shapes_b200.world import World, rectangle_vertices
    k4 = kat["kat4_circle_circle"]
    w = World.from_objects([(k4["b"]["radius"], tuple(k4["b"]["center"]), 0.0, (1.0, 1.0)),
                            (k4["a"]["radius"], tuple(k4["a"]["center"]), 0.0, (1.0, 1.0))])
    r = oracle.frame(w)
    c = k4["contact"]
    assert len(r["key_i"]) == 1 and [r["feat_a"][0], r["feat_b"][0]] == c["feat"] and r["flip"][0] == c["flip"]
    assert [r["normal_x"][0], r["normal_y"][0]] == c["normal"] and [r["center_x"][0], r["center_y"][0]] == c["center"]
    assert r["depth"][0] == c["depth"]
    k5 = kat["kat5_circle_hull"]
    box = (rectangle_vertices(*k5["box"]["size"]), tuple(k5["box"]["center"]), 0.0, (1.0, 1.0))
    cir = (k5["circle"]["radius"], tuple(k5["circle"]["center"]), 0.0, (1.0, 1.0))
    for objs, name in (([box, cir], "circle_is_a"), ([cir, box], "hull_is_a")):
        r = oracle.frame(World.from_objects(objs))
        c = k5[name]
        assert len(r["key_i"]) == 1 and [r["feat_a"][0], r["feat_b"][0]] == c["feat"] and r["flip"][0] == c["flip"]
        assert [r["normal_x"][0], r["normal_y"][0]] == c["normal"] and [r["center_x"][0], r["center_y"][0]] == c["center"]
        assert r["depth"][0] == c["depth"]
        if "rn" in c:
            assert [r["rn_x"][0], r["rn_y"][0]] == c["rn"]


def test_circle_worlds_oracle_invariants(oracle):
    """Balls.makeScene and a random circle/polygon world: brute force == sweep == grid, unit normals."""
    for w in (scenes.balls_scene((8, 6), 0.5, 0.0), scenes.random_circles_and_polygons(1500, config=60)):
        a = oracle.frame(w, broadphase="aabb")
        s_ = oracle.frame(w, broadphase="sweep")
        assert np.array_equal(a["pair_i"], s_["pair_i"]) and np.array_equal(a["pair_j"], s_["pair_j"])
        assert len(a["key_i"]) > 0
        assert np.allclose(a["normal_x"] ** 2 + a["normal_y"] ** 2, 1.0)
        ck = (a["key_i"].astype(np.int64) << 32) | a["key_j"]
        assert np.all(ck[:-1] >= ck[1:])
    w = scenes.balls_scene((4, 3), 0.5, 0.0)
    assert w.radius is not None and (w.radius >= 0).sum() == 6 and w.n_slots == 13
