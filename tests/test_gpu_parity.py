"""Parity of the CUDA path (through the C ABI) with the CPU oracle.  Needs a GPU.

Pair sets and all integer columns must be bit-identical; reals are asserted bit-identical
too (every operation is IEEE and un-fused on both sides), which is stricter than the
north-star's 1e-9 relative bar.
"""
import json
import os

import numpy as np
import pytest

from conftest import assert_frames_match
from shapes_b200 import scenes
from shapes_b200.world import World, rectangle_vertices

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
WANT_ALL = ("pairs", "contacts", "constraints", "aabb", "world")


def gpu_frame(world, cos_sin, want=WANT_ALL, ext=None, cell=None, **kw):
    from shapes_b200.engine import Engine
    with Engine(world, ext=ext) as eng:
        if cell is not None:
            eng.set_cell_size(cell)
        fr = eng.frame_grow(cos_sin=cos_sin, want=want, **kw)
        got = {k: np.array(v) for k, v in fr.cols.items()}
        got["_n_big"] = fr.n_big
        got["_launches"] = eng.launch_count
    return got


def check_world(oracle, world, broadphase="auto", **kw):
    c, s = oracle.cos_sin(world.rot)
    want = oracle.frame(world, c, s, broadphase=broadphase,
                        **{k: v for k, v in kw.items() if k in ("dt", "baumgarte", "slop")})
    got = gpu_frame(world, (c, s), **kw)
    assert got["_launches"] > 0
    assert_frames_match(got, want)
    for k in ("aabb_min_x", "aabb_max_x", "aabb_min_y", "aabb_max_y"):
        live = world.alive == 1
        assert np.array_equal(got[k][live], want[k][live], equal_nan=True), k
    for k in ("world_x", "world_y"):
        live_v = np.repeat(world.alive == 1, np.diff(world.vert_offset))
        assert np.array_equal(got[k][live_v], want[k][live_v], equal_nan=True), k
    return got, want


def test_kat1_and_kat3_through_the_abi(oracle):
    with open(os.path.join(GOLDEN, "kat.json")) as f:
        kat = json.load(f)
    got, _ = check_world(oracle, scenes.test_opt_boxes())
    k1 = kat["kat1_testOptBoxes"]
    assert got["pair_i"].tolist() == [1] and got["pair_j"].tolist() == [0]
    for k, c in enumerate(k1["contacts"]):
        assert [got["feat_a"][k], got["feat_b"][k]] == c["feat"] and got["flip"][k] == c["flip"]
        assert [got["normal_x"][k], got["normal_y"][k]] == c["normal"]
        assert [got["center_x"][k], got["center_y"][k]] == c["center"] and got["depth"][k] == c["depth"]
    for name, spacing in (("spacing0", 0.0), ("spacing1", 1.0)):
        got, _ = check_world(oracle, scenes.broadphase_bench_world(spacing=spacing), broadphase="aabb")
        want = kat["kat3_broadphase"][name]
        assert len(got["pair_i"]) == want["n_pairs"]
        assert [got["pair_i"][0], got["pair_j"][0]] == want["first"]


@pytest.mark.parametrize("name,triangle_is_a,lift", [("kat6_triangle_into_box_same", False, 0.0),
                                                     ("kat7_triangle_into_box_flip", True, 0.0),
                                                     ("kat8_triangle_lifted_single_point", False, 1.0),
                                                     ("kat9_box_on_hexagon", None, 0.0)])
def test_kat678_through_the_abi(oracle, name, triangle_is_a, lift):
    """The hand-derived vectors KAT-6/7/8 (tests/golden/README.md) against the CUDA path itself, every pinned column
    bit for bit (signed zeros included): rotated non-box hull through minOverlap, Same and Flip, ClipLeft, the third
    clip removing a point, active Baumgarte term, Friction / Restitution radii, inverse effective masses."""
    from test_oracle_kat import assert_kat_rows
    with open(os.path.join(GOLDEN, "kat.json")) as f:
        want = json.load(f)[name]
    # (KAT-9's hexagon makes its world a "general polygon" world: cell-ordered work list + k_manifolds_coop)
    w, c, s = scenes.kat_box_on_hexagon() if triangle_is_a is None else scenes.kat_triangle_on_box(triangle_is_a, lift)
    got = gpu_frame(w, (c, s), **want["behaviour"])
    assert_kat_rows(got, want, name)
    assert_frames_match(got, oracle.frame(w, c, s, broadphase="aabb", **want["behaviour"]))


def test_golden_fixture(oracle):
    z = np.load(os.path.join(GOLDEN, "oracle_polygons64.npz"))
    w = scenes.random_polygons(64, density=2.0, static_frac=0.1, config=99)
    got = gpu_frame(w, (z["cos"], z["sin"]))
    assert_frames_match(got, {k: z[k] for k in z.files})


def test_config1_stacks_scene(oracle):
    """BASELINE config 1: Stacks.makeScene (30,30) 0, frame 0; floor = big-shape path."""
    got, want = check_world(oracle, scenes.stacks_scene(), broadphase="aabb")
    assert got["_n_big"] >= 1 and len(want["pair_i"]) > 3000


@pytest.mark.parametrize("dims,spacing", [((10, 10), 1.0), ((5, 40), 0.0)])
def test_stacks_variants(oracle, dims, spacing):
    check_world(oracle, scenes.stacks_scene(dims, spacing), broadphase="aabb")


def test_settled_stack_touches_floor(oracle):
    """A stack resting on (and slightly inside) the floor: floor contacts through the big path."""
    w = scenes.stacks_scene((12, 6), 0.0)
    w.pos_y[1:] -= 0.9 + 0.005
    got, want = check_world(oracle, w, broadphase="aabb")
    assert (want["key_j"] == 0).sum() >= 12


def test_config2_polygons_10k(oracle):
    got, want = check_world(oracle, scenes.random_polygons(10_000), broadphase="sweep")
    assert len(want["pair_i"]) > 5_000 and len(want["key_i"]) > 1_500


def test_polygons_dense_bruteforce_oracle(oracle):
    """Small enough for the Theta(n^2) restatement of Aabb.culledKeys itself."""
    check_world(oracle, scenes.random_polygons(2500, density=3.0, static_frac=0.15, config=21), broadphase="aabb")


def test_pile_with_floor(oracle):
    got, want = check_world(oracle, scenes.box_pile(120, 90))
    assert (want["key_j"] == 0).sum() > 100     # bottom row rests in the floor


def test_mixed_and_blob(oracle):
    check_world(oracle, scenes.mixed_polygons(30_000))
    check_world(oracle, scenes.gaussian_blob(30_000, density=1.0))


def test_tie_heavy_lattice_world(oracle):
    """Shapes from a small pool placed on a lattice of quarter units with rotations by multiples of 90 degrees
    (exact cos/sin): equal depths, touching edges, coincident vertices and parallel clip planes everywhere --
    every tie-break of the SAT fold, the incident-edge choice and the NaN-driven ClipNone at once, through both
    contact kernels (boxes only / hulls of up to 8 vertices)."""
    rng = np.random.default_rng(33)
    boxes = [rectangle_vertices(w, h) for w in (0.5, 1.0, 2.0) for h in (0.5, 1.0, 1.5)]
    polys = [[(0.0, 0.0), (1.0, 0.0), (0.0, 1.0)], [(0.5, 0.0), (0.0, 0.75), (-0.5, 0.0)],
             [(1.0, 0.0), (0.0, 1.0), (-1.0, 0.0), (0.0, -1.0)],
             [(1.0, 0.5), (0.0, 1.0), (-1.0, 0.5), (-1.0, -0.5), (0.0, -1.0), (1.0, -0.5)],
             [(1.0, 0.0), (0.75, 0.75), (0.0, 1.0), (-0.75, 0.75), (-1.0, 0.0), (-0.75, -0.75), (0.0, -1.0), (0.75, -0.75)]]
    quarter = [(1.0, 0.0), (0.0, 1.0), (-1.0, 0.0), (0.0, -1.0)]
    for pool in (boxes, boxes + polys):
        objs = []
        for k in range(4000):
            v = pool[rng.integers(len(pool))]
            pos = (float(rng.integers(-160, 161)) * 0.25, float(rng.integers(-160, 161)) * 0.25)
            objs.append((v, pos, 0.0, (1.0, 1.0) if rng.random() < 0.9 else (0.0, 0.0)))
        w = World.from_objects(objs)
        rot = rng.integers(0, 4, w.n_slots)
        cs = (np.array([quarter[r][0] for r in rot]), np.array([quarter[r][1] for r in rot]))
        want = oracle.frame(w, cs[0], cs[1], broadphase="sweep")
        got = gpu_frame(w, cs)
        assert_frames_match(got, want)
        assert len(want["key_i"]) > 3000 and want["flip"].sum() > 0     # axis-aligned boxes: depth ties everywhere, ties are Flip


def test_deleted_slots_and_all_static(oracle):
    w = scenes.random_polygons(3000, density=2.5, static_frac=0.3, config=31)
    w.delete(list(range(0, 3000, 7)))
    got, want = check_world(oracle, w, broadphase="aabb")
    assert not np.isin(got["pair_i"], np.arange(0, 3000, 7)).any()
    w2 = scenes.random_polygons(500, density=3.0, static_frac=1.1, config=32)
    got2, _ = check_world(oracle, w2, broadphase="aabb")
    assert len(got2["pair_i"]) == 0 and len(got2["key_i"]) == 0


def test_nonfinite_bounds_take_the_exact_path(oracle):
    """NaN / inf AABBs 'overlap' under boundsOverlap (Aabb.hs:69-72): identical pair set."""
    w = scenes.random_polygons(1500, density=2.0, static_frac=0.1, config=33)
    w.pos_x[40] = w.pos_y[40] = np.nan
    w.pos_x[42] = np.nan
    w.pos_y[41] = np.inf
    w.pos_x[1400] = -np.inf
    got, want = check_world(oracle, w, broadphase="aabb")
    assert got["_n_big"] >= 4 and (want["pair_i"] == 40).sum() + (want["pair_j"] == 40).sum() > 1000


def test_many_vertices_fallback_path(oracle):
    """Hulls with more than 8 vertices are not staged in shared memory."""
    rng = np.random.default_rng(5)
    objs = []
    for k in range(400):
        nv = int(rng.integers(3, 20))
        ang = np.sort(rng.uniform(0, 2 * np.pi, nv))
        r = rng.uniform(0.3, 0.6)
        verts = [(r * np.cos(a), r * np.sin(a)) for a in ang]
        objs.append((verts, (rng.uniform(0, 12), rng.uniform(0, 12)), rng.uniform(0, 6.28),
                     (0.0, 0.0) if k % 17 == 0 else (1.0, 1.0)))
    check_world(oracle, World.from_objects(objs), broadphase="aabb")


def test_host_supplied_extents_are_used(oracle):
    """_hullExtents is frozen at construction (ConvexHull.hs:159,165): the library must use the
    indices it is given, even 'wrong' ones, exactly like extentAlongSelf trusts the cache."""
    w = scenes.random_polygons(800, density=3.0, config=34)
    emin, emax = oracle.hull_extents(w)
    emin2 = emin.copy()
    nv = np.diff(w.vert_offset)
    starts = w.vert_offset[:-1]
    emin2[starts[::3]] = (emin2[starts[::3]] + 1) % nv[::3]     # perturb every third hull's first edge
    c, s = oracle.cos_sin(w.rot)
    want = oracle.frame(w, c, s, broadphase="aabb", ext=(emin2, emax))
    got = gpu_frame(w, (c, s), ext=(emin2, emax))
    assert_frames_match(got, want)
    base = oracle.frame(w, c, s, broadphase="aabb")
    assert len(base["key_i"]) != len(want["key_i"]) or not np.array_equal(base["depth"], want["depth"])


def test_out_of_range_extent_indices_are_rejected(oracle):
    """An extent index must name a vertex of its own hull; anything else is SHAPES_E_ARG, not an out-of-bounds read."""
    from shapes_b200.engine import Engine, ShapesError
    w = scenes.random_polygons(200, density=1.0, config=36)
    emin, emax = oracle.hull_extents(w)
    nv = np.diff(w.vert_offset)
    for which in (0, 1):
        bad = [emin.copy(), emax.copy()]
        bad[which][w.vert_offset[7]] = nv[7]              # one past the last vertex of hull 7
        with pytest.raises(ShapesError):
            Engine(w, ext=(bad[0], bad[1])).close()
    with Engine(w, ext=(emin, emax)) as eng:               # the exact table is accepted
        eng.frame_grow()


def test_cell_size_does_not_change_results(oracle):
    w = scenes.random_polygons(4000, density=2.0, config=35)
    c, s = oracle.cos_sin(w.rot)
    want = oracle.frame(w, c, s, broadphase="sweep")
    for cell in (0.25, 1.0, 3.7, 50.0, 1e6):
        got = gpu_frame(w, (c, s), want=("pairs", "contacts", "constraints"), cell=cell)
        assert_frames_match(got, want)


def test_behaviour_parameters(oracle):
    w = scenes.box_pile(40, 30)
    check_world(oracle, w, dt=1.0 / 60.0, baumgarte=0.2, slop=0.005)


def test_capacity_error_reports_required_sizes(oracle):
    from shapes_b200.engine import CapacityError, Engine
    w = scenes.random_polygons(3000, density=3.0, config=36)
    c, s = oracle.cos_sin(w.rot)
    want = oracle.frame(w, c, s, broadphase="sweep")
    with Engine(w, max_pairs=100, max_contacts=50) as eng:
        with pytest.raises(CapacityError) as ei:
            eng.frame(cos_sin=(c, s))
        assert ei.value.n_pairs == len(want["pair_i"])
    with Engine(w, max_pairs=len(want["pair_i"]), max_contacts=50) as eng:
        with pytest.raises(CapacityError) as ei:
            eng.frame(cos_sin=(c, s))
        assert ei.value.n_contacts == len(want["key_i"])
        fr = eng.frame_grow(cos_sin=(c, s))
        assert_frames_match(fr.cols, want)


def test_deterministic_and_reusable(oracle):
    from shapes_b200.engine import Engine
    w = scenes.random_polygons(20_000, density=1.5, config=37)
    c, s = oracle.cos_sin(w.rot)
    with Engine(w) as eng:
        a = {k: np.array(v) for k, v in eng.frame_grow(cos_sin=(c, s)).cols.items()}
        for _ in range(3):
            b = eng.frame(cos_sin=(c, s)).cols
            for k in a:
                assert np.array_equal(a[k], b[k], equal_nan=True), k
        # move everything and run again on the same ctx
        w.pos_x += 0.37
        w.rot += 0.1
        c2, s2 = oracle.cos_sin(w.rot)
        got = eng.frame_grow(cos_sin=(c2, s2)).cols
        assert_frames_match(got, oracle.frame(w, c2, s2, broadphase="sweep"))


def test_device_sincos_is_close_not_exact(oracle):
    """cos/sin NULL => device sincos: documented as not bit-exact; pair set and contacts must
    still agree to 1e-9 on a world away from knife edges."""
    w = scenes.random_polygons(3000, density=1.0, config=38)
    c, s = oracle.cos_sin(w.rot)
    want = oracle.frame(w, c, s, broadphase="sweep")
    got = gpu_frame(w, None, want=("pairs", "contacts", "constraints"))
    assert np.array_equal(got["pair_i"], want["pair_i"]) and np.array_equal(got["pair_j"], want["pair_j"])
    if len(got["key_i"]) == len(want["key_i"]):
        assert_frames_match(got, want, exact=False, cols=("normal_x", "normal_y", "depth"))


def test_empty_and_single_shape_worlds(oracle):
    w = World.from_objects([(rectangle_vertices(1, 1), (0.0, 0.0), 0.0, (1.0, 1.0))])
    got = gpu_frame(w, None, want=("pairs", "contacts"))
    assert len(got["pair_i"]) == 0 and len(got["key_i"]) == 0
    w0 = World.from_objects([])
    got0 = gpu_frame(w0, None, want=("pairs", "contacts"))
    assert len(got0["pair_i"]) == 0


def test_config3_full_size_pile(oracle):
    """BASELINE config 3 at full size (1M boxes + floor): bit-exact against the oracle, plus the
    size-independent properties (strictly descending unique pairs, keys subset of pairs)."""
    w = scenes.box_pile(1000, 1000)
    c, s = oracle.cos_sin(w.rot)
    got = gpu_frame(w, (c, s), want=("pairs", "contacts", "constraints"))
    key = (got["pair_i"].astype(np.int64) << 32) | got["pair_j"]
    assert np.all(key[:-1] > key[1:]) and np.all(got["pair_i"] > got["pair_j"])
    ck = (got["key_i"].astype(np.int64) << 32) | got["key_j"]
    assert np.all(ck[:-1] >= ck[1:]) and np.isin(ck, key).all()
    assert np.allclose(got["normal_x"] ** 2 + got["normal_y"] ** 2, 1.0, atol=1e-12)
    want = oracle.frame(w, c, s, broadphase="sweep")
    assert_frames_match(got, want)
    assert len(want["pair_i"]) > 3_500_000


# ---- cell-ordered hull records + SAT work list (general polygon worlds) -------------------------

def _polygon_world_with_big_shapes(seed=41, n=3000):
    """Random polygons (some with more than 8 vertices) plus big shapes of every kind: two static floors,
    a big dynamic polygon over many cells, a big 12-gon -- all on the big list, as queries and as partners."""
    rng = np.random.default_rng(seed)
    objs = []
    side = 40.0
    for k in range(n):
        nv = int(rng.integers(3, 9)) if k % 11 else int(rng.integers(9, 14))
        ang = np.sort(rng.uniform(0, 2 * np.pi, nv))
        r = rng.uniform(0.3, 0.5)
        verts = [(r * np.cos(a), r * np.sin(a)) for a in ang]
        objs.append((verts, (rng.uniform(0, side), rng.uniform(0, side)), rng.uniform(0, 6.28),
                     (0.0, 0.0) if k % 23 == 0 else (1.0, 1.0)))
    big = [(rectangle_vertices(side + 2.0, 1.0), (side / 2, 0.2), 0.0, (0.0, 0.0)),
           (rectangle_vertices(1.0, side), (side / 3, side / 2), 0.1, (0.0, 0.0)),
           ([(6.0, 0.0), (0.0, 5.0), (-6.0, 0.0), (0.0, -5.0)], (side / 2, side / 2), 0.3, (1.0, 1.0)),
           ([(7 * np.cos(a), 7 * np.sin(a)) for a in np.linspace(0, 2 * np.pi, 12, endpoint=False)],
            (side * 0.7, side * 0.6), 0.2, (1.0, 1.0))]
    # big shapes at both ends and in the middle of the key range
    objs = big[:1] + objs[: n // 2] + big[1:3] + objs[n // 2:] + big[3:]
    return World.from_objects(objs)


def test_sorted_mode_big_shapes_and_long_hulls(oracle):
    got, want = check_world(oracle, _polygon_world_with_big_shapes(), broadphase="aabb")
    assert got["_n_big"] >= 4 and len(want["key_i"]) > 500


def test_sorted_mode_crowded_cells(oracle):
    """More than 24 partners per query (the emit pass's global-memory sort) and more than 63 candidates (no
    hit-mask replay) with the work list in play."""
    w = scenes.random_polygons(1500, density=60.0, static_frac=0.05, config=43)
    got, want = check_world(oracle, w, broadphase="aabb")
    per_query = np.bincount(want["pair_i"], minlength=w.n_slots)
    assert per_query.max() > 40 and len(want["pair_i"]) > 30_000


@pytest.mark.parametrize("which", ["polygons", "big", "crowded", "nonfinite"])
def test_slot_order_path_still_matches(oracle, monkeypatch, which):
    """SHAPES_B200_NO_SORTED=1: hull records in slot order, SAT in the reference's pair order (the r1 path)."""
    monkeypatch.setenv("SHAPES_B200_NO_SORTED", "1")
    if which == "polygons":
        check_world(oracle, scenes.random_polygons(10_000), broadphase="sweep")
    elif which == "big":
        check_world(oracle, _polygon_world_with_big_shapes(seed=42), broadphase="aabb")
    elif which == "crowded":
        check_world(oracle, scenes.random_polygons(1200, density=60.0, static_frac=0.05, config=44), broadphase="aabb")
    else:
        w = scenes.random_polygons(1500, density=2.0, static_frac=0.1, config=33)
        w.pos_x[40] = w.pos_y[40] = np.nan
        w.pos_y[41] = np.inf
        check_world(oracle, w, broadphase="aabb")


def test_compact_wire_format_expands_to_the_full_rows(oracle):
    """shapes_frame with NULL pointers for the sixteen derived constraint columns (113 instead of 225 bytes per row
    over PCIe) + engine.expand_rows == the full fetch == the oracle, bit for bit."""
    from shapes_b200.engine import Engine
    for w in (scenes.random_polygons(20_000, density=2.0, config=45), scenes.box_pile(100, 60)):
        c, s = oracle.cos_sin(w.rot)
        want = oracle.frame(w, c, s, broadphase="sweep")
        with Engine(w) as eng:
            fr = eng.frame_grow(cos_sin=(c, s), compact=True)
            assert fr.d2h_bytes == 8 * fr.n_pairs + (17 + 12 * 8) * fr.n_contacts
            assert_frames_match(fr.cols, want)
            full = eng.frame(cos_sin=(c, s))
            assert full.d2h_bytes > fr.d2h_bytes + 12 * 8 * fr.n_contacts
            assert_frames_match(full.cols, want)


# ---- plan-ahead grid (the grid of a frame is planned from the previous frame's bounds) -----------------------

@pytest.mark.parametrize("kind", ["polygons", "pile"])
def test_plan_ahead_follows_a_moving_world_and_survives_teleports(oracle, kind):
    """Frames on one ctx: small drifts stay inside the planned grid's margin, a jump of 5000 units makes the plan
    stale (ERR_REPLAN: seeded again inside the same call), a jump of part of the world lands those shapes on the
    exact big-shape path.  Every frame bit-identical to the oracle, warm-start join included."""
    from shapes_b200.engine import Engine
    w = scenes.random_polygons(20_000, density=1.5, config=46) if kind == "polygons" else scenes.box_pile(150, 100)
    want = ("pairs", "contacts", "constraints", "warm")
    rng = np.random.default_rng(7)
    with Engine(w) as eng:
        prev = None
        for step, move in enumerate(["none", "drift", "drift", "teleport", "drift", "scatter_few", "drift"]):
            if move == "drift":
                w.pos_x += rng.uniform(-0.05, 0.05, w.n_slots); w.pos_y += rng.uniform(-0.05, 0.05, w.n_slots)
            elif move == "teleport":
                w.pos_x += 5000.0; w.pos_y -= 3000.0
            elif move == "scatter_few":
                far = rng.choice(w.n_slots, 40, replace=False)
                w.pos_x[far] += rng.uniform(50.0, 400.0, 40)
            c, s = oracle.cos_sin(w.rot)
            ref = oracle.frame(w, c, s, broadphase="sweep")
            if prev is not None:
                eng.set_lagrangian_cache(prev[1], prev[2])
            fr = eng.frame_grow(cos_sin=(c, s), want=want)
            assert_frames_match(fr.cols, ref)
            cur = {k: np.array(fr[k]) for k in ("key_i", "key_j", "feat_a", "feat_b")}
            if prev is not None:
                o_np, o_f, o_hit = oracle.warm_join(cur, prev[0], prev[1], prev[2])
                assert np.array_equal(fr["warm_hit"], o_hit) and np.array_equal(fr["warm_np"], o_np)
            if move == "scatter_few":
                assert fr.n_big >= 30
            prev = (cur, 0.5 + (cur["key_i"] % 97).astype(np.float64), 1.0 + (cur["feat_a"] % 5).astype(np.float64))


def test_plan_inside_the_frame_still_matches(oracle, monkeypatch):
    """SHAPES_B200_NO_PLAN_AHEAD=1: bounds -> plan -> keys inside the frame (the r1 path, what multi-rank ctxs run)."""
    monkeypatch.setenv("SHAPES_B200_NO_PLAN_AHEAD", "1")
    check_world(oracle, scenes.random_polygons(10_000), broadphase="sweep")
    check_world(oracle, scenes.box_pile(120, 90))
