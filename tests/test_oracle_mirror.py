"""Two restatements of the reference, written separately and structured differently, must agree bit for bit:
oracle/shapes_oracle.c (+ _step.c: index arithmetic over SoA columns) against oracle/hs_mirror.py (the
reference's own Neighborhood / Maybe / Either / foldl1 shapes, function by function).  This catches
transcription slips in either; it does not pin them to the Haskell program's outputs (no GHC here).  CPU only."""
import json
import os

import numpy as np
import pytest

from oracle import hs_mirror as hs
from shapes_b200 import scenes
from shapes_b200.world import Bodies, World, rectangle_vertices

COLS_J = [f"j_np{q}" for q in range(6)]
COLS_F = [f"j_f{q}" for q in range(6)]


def same(a, b):
    return a == b or (a != a and b != b)


def compare_frame(oracle, w, dt=0.01, beh=(0.01, 0.02)):
    c, s = oracle.cos_sin(w.rot)
    fr = oracle.frame(w, c, s, dt=dt, baumgarte=beh[0], slop=beh[1], broadphase="aabb")
    hulls = hs.hulls_of(w, c, s)
    # world vertices, normals, frozen extents
    for slot, h in enumerate(hulls):
        if h is None:
            continue
        o = int(w.vert_offset[slot])
        for k in range(h.count):
            assert same(h.vertices[k][0], fr["world_x"][o + k]) and same(h.vertices[k][1], fr["world_y"][o + k]), (slot, k)
            assert same(h.edge_normals[k][0], fr["normal_wx"][o + k]) and same(h.edge_normals[k][1], fr["normal_wy"][o + k])
            assert h.extents[k] == (int(fr["ext_min"][o + k]), int(fr["ext_max"][o + k])), (slot, k)
    pairs = list(zip(fr["pair_i"].tolist(), fr["pair_j"].tolist()))
    rows = hs.prepare_frame(w, hulls, pairs, beh, dt)
    assert len(rows) == len(fr["key_i"]), (len(rows), len(fr["key_i"]))
    for k, r in enumerate(rows):
        assert r["key"] == (fr["key_i"][k], fr["key_j"][k], fr["feat_a"][k], fr["feat_b"][k]), k
        assert r["flip"] == fr["flip"][k]
        ct = r["contact"]
        assert same(ct["normal"][0], fr["normal_x"][k]) and same(ct["normal"][1], fr["normal_y"][k]), k
        assert same(ct["center"][0], fr["center_x"][k]) and same(ct["center"][1], fr["center_y"][k]), k
        assert same(ct["depth"], fr["depth"][k]), k
        jn, bn = r["constraint"]["nonpen"]
        jf, bf = r["constraint"]["friction"]
        ra, rb, rn = r["constraint"]["restitution"]
        for q in range(6):
            assert same(jn[q], fr[COLS_J[q]][k]) and same(jf[q], fr[COLS_F[q]][k]), (k, q)
        assert same(bn, fr["b_np"][k]) and bf == 0.0
        assert same(ra[0], fr["ra_x"][k]) and same(ra[1], fr["ra_y"][k]) and same(rb[0], fr["rb_x"][k]) and same(rb[1], fr["rb_y"][k])
        assert same(rn[0], fr["rn_x"][k]) and same(rn[1], fr["rn_y"][k])
        i, j = r["key"][0], r["key"][1]
        a = {"inv": (w.inv_lin[i], w.inv_rot[i])}
        b = {"inv": (w.inv_lin[j], w.inv_rot[j])}
        assert same(hs.effMassM2(jn, a, b), fr["inv_eff_np"][k]) and same(hs.effMassM2(jf, a, b), fr["inv_eff_f"][k])
    return fr, rows


def test_kat1_through_the_mirror():
    """testOptBoxes (bench/Physics/Contact/Benchmark.hs:16-27), the hand trace of tests/golden/README.md."""
    with open(os.path.join(os.path.dirname(__file__), "golden", "kat.json")) as f:
        kat = json.load(f)["kat1_testOptBoxes"]
    a = hs.ConvexHull(rectangle_vertices(4, 4)).setHullTransform(lambda p: hs.afmul(hs.toTransform((0.0, 0.0), (1.0, 0.0)), p))
    b = hs.ConvexHull(rectangle_vertices(2, 2)).setHullTransform(lambda p: hs.afmul(hs.toTransform((1.0, 3.0), (1.0, 0.0)), p))
    got = hs.generateContacts(a, b)
    assert len(got) == len(kat["contacts"])
    for (feat, (flipping, c)), want in zip(got, kat["contacts"]):
        assert list(feat) == want["feat"] and (0 if flipping == "Same" else 1) == want["flip"]
        assert list(c["normal"]) == want["normal"] and list(c["center"]) == want["center"] and c["depth"] == want["depth"]


@pytest.mark.parametrize("name,triangle_is_a,lift", [("kat6_triangle_into_box_same", False, 0.0),
                                                     ("kat7_triangle_into_box_flip", True, 0.0),
                                                     ("kat8_triangle_lifted_single_point", False, 1.0),
                                                     ("kat9_box_on_hexagon", None, 0.0)])
def test_kat678_through_the_mirror(name, triangle_is_a, lift):
    """The hand traces of KAT-6/7/8 (tests/golden/README.md) through the second restatement: ClipLeft, Same / Flip on
    a non-box hull, the third clip removing a point, active Baumgarte term, Friction / Restitution with real radii."""
    with open(os.path.join(os.path.dirname(__file__), "golden", "kat.json")) as f:
        want = json.load(f)[name]
    w, c, s = scenes.kat_box_on_hexagon() if triangle_is_a is None else scenes.kat_triangle_on_box(triangle_is_a, lift)
    hulls = hs.hulls_of(w, c, s)
    beh = want["behaviour"]
    rows = hs.prepare_frame(w, hulls, [tuple(p) for p in want["pairs"]], (beh["baumgarte"], beh["slop"]), beh["dt"])
    assert len(rows) == len(want["contacts"])
    bits = lambda x: np.float64(x).tobytes()
    for r, c_ in zip(rows, want["contacts"]):
        assert r["key"] == tuple(c_["key"] + c_["feat"]) and r["flip"] == c_["flip"]
        ct = r["contact"]
        jn, bn = r["constraint"]["nonpen"]
        jf, bf = r["constraint"]["friction"]
        ra, rb, rn = r["constraint"]["restitution"]
        got = list(ct["normal"]) + list(ct["center"]) + [ct["depth"], bn] + list(jn) + list(jf) + list(ra) + list(rb) + list(rn)
        exp = c_["normal"] + c_["center"] + [c_["depth"], c_["b_np"]] + c_["j_np"] + c_["j_f"] + c_["ra"] + c_["rb"] + c_["rn"]
        assert [bits(x) for x in got] == [bits(x) for x in exp], (got, exp)
        assert bf == 0.0
        i, j = r["key"][0], r["key"][1]
        a = {"inv": (w.inv_lin[i], w.inv_rot[i])}
        b = {"inv": (w.inv_lin[j], w.inv_rot[j])}
        assert hs.effMassM2(jn, a, b) == c_["inv_eff_np"] and hs.effMassM2(jf, a, b) == c_["inv_eff_f"]


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_random_polygon_worlds(oracle, seed):
    w = scenes.random_polygons(260, density=2.5, config=300 + seed)
    fr, rows = compare_frame(oracle, w)
    assert len(rows) > 40 and {r["flip"] for r in rows} == {0, 1}


def test_stacks_pile_and_touching_boxes(oracle):
    """Axis-aligned boxes: parallel reference / incident edges (det = 0, NaN-driven ClipNone), exact ties."""
    w = scenes.stacks_scene((6, 5), 0.0)
    w.pos_y[1:] -= 0.905
    fr, rows = compare_frame(oracle, w)
    assert len(rows) > 50
    compare_frame(oracle, scenes.box_pile(12, 9))
    t = World.from_objects([(rectangle_vertices(4, 4), (0.0, 0.0), 0.0, (1.0, 1.0)),
                            (rectangle_vertices(2, 2), (1.0, 3.0), 0.0, (1.0, 1.0)),
                            (rectangle_vertices(2, 2), (3.0, 0.0), 0.0, (0.0, 0.0))])
    compare_frame(oracle, t, dt=0.02, beh=(0.2, 0.001))


def test_solver_sweeps_agree(oracle):
    """improveWorld (two sweeps) over a frame's contacts: the C loop against improveContactSln of the mirror."""
    w = scenes.random_polygons(160, density=3.0, config=310)
    fr, rows = compare_frame(oracle, w)
    n = w.n_slots
    rng = np.random.default_rng(3)
    b = Bodies(rng.uniform(-1, 1, n), rng.uniform(-1, 1, n), rng.uniform(-1, 1, n), rng.uniform(0, 0.6, n), rng.uniform(0, 0.5, n))
    objs = [{"vel": (b.vel_x[s], b.vel_y[s]), "rotvel": b.rot_vel[s], "pos": (w.pos_x[s], w.pos_y[s]),
             "inv": (w.inv_lin[s], w.inv_rot[s])} for s in range(n)]
    lam = [(0.0, 0.0)] * len(rows)
    lam_np, lam_f = np.zeros(len(rows)), np.zeros(len(rows))
    for sweep in range(2):
        oracle.improve_world(w, fr, b.mu, b.bounce, b.vel_x, b.vel_y, b.rot_vel, lam_np, lam_f)
        for k, r in enumerate(rows):
            i, j = r["key"][0], r["key"][1]
            (objs[i], objs[j]), lam[k] = hs.improveContactSln(r["constraint"], lam[k], (b.mu[i], b.mu[j]),
                                                                (b.bounce[i], b.bounce[j]), (objs[i], objs[j]))
        for s in range(n):
            assert same(objs[s]["vel"][0], b.vel_x[s]) and same(objs[s]["vel"][1], b.vel_y[s]) and same(objs[s]["rotvel"], b.rot_vel[s]), (sweep, s)
        for k in range(len(rows)):
            assert same(lam[k][0], lam_np[k]) and same(lam[k][1], lam_f[k]), (sweep, k)
    assert np.abs(lam_np).max() > 0.0


def mirror_world(w, c, s):
    """Shapes (hulls and circles), tagged AABBs in traversal order, as the reference's World holds them."""
    hulls = hs.hulls_of(w, c, s)
    shapes, tagged = {}, []
    for slot in range(w.n_slots):
        if not w.alive[slot]:
            continue
        if w.radius is not None and w.radius[slot] >= 0:
            m = hs.toTransform((float(w.pos_x[slot]), float(w.pos_y[slot])), (float(c[slot]), float(s[slot])))
            centre = hs.afmul(m, (0.0, 0.0))                         # setCircleTransform (Circle.hs:55-59)
            shapes[slot] = ("circle", centre, float(w.radius[slot]))
            box = hs.circleToAabb(centre, float(w.radius[slot]))
        else:
            shapes[slot] = ("hull", hulls[slot])
            box = hs.hullToAabb(hulls[slot])
        tagged.append((slot, box, w.inv_lin[slot] == 0.0 and w.inv_rot[slot] == 0.0))
    return shapes, tagged


@pytest.mark.parametrize("scene", ["balls", "mixed", "deleted"])
def test_circles_broadphase_and_dispatch(oracle, scene):
    """Aabb.culledKeys (unorderedPairs order, static/static dropped) and the four-way generateContacts
    dispatch (Circle.contact, GJK closestSimplex, CircleVsHull) against the C oracle."""
    if scene == "balls":
        w = scenes.balls_scene((7, 6), 0.5, 0.0)
        w.pos_y[1:] -= 0.55
    else:
        w = scenes.random_circles_and_polygons(240, config=320)
        if scene == "deleted":
            w.delete([3, 17, 100, 101])
    c, s = oracle.cos_sin(w.rot)
    fr = oracle.frame(w, c, s, broadphase="aabb")
    shapes, tagged = mirror_world(w, c, s)
    for slot, box, _ in tagged:
        assert box == ((fr["aabb_min_x"][slot], fr["aabb_max_x"][slot]), (fr["aabb_min_y"][slot], fr["aabb_max_y"][slot])), slot
    pairs = hs.culledKeys(tagged)
    assert pairs == list(zip(fr["pair_i"].tolist(), fr["pair_j"].tolist()))
    k = 0
    kinds = set()
    for (i, j) in pairs:
        for feat, (flipping, ct) in hs.generateContactsShapes(shapes[i], shapes[j]):
            assert (i, j) + feat == (fr["key_i"][k], fr["key_j"][k], fr["feat_a"][k], fr["feat_b"][k]), k
            assert (0 if flipping == "Same" else 1) == fr["flip"][k]
            assert same(ct["normal"][0], fr["normal_x"][k]) and same(ct["normal"][1], fr["normal_y"][k]), k
            assert same(ct["center"][0], fr["center_x"][k]) and same(ct["center"][1], fr["center_y"][k]), k
            assert same(ct["depth"], fr["depth"][k]), k
            kinds.add((shapes[i][0], shapes[j][0]))
            k += 1
    assert k == len(fr["key_i"]) and k > 10
    if scene == "mixed":
        assert len(kinds) == 4                                       # all four shape-pair kinds produced contacts


def test_desc_zip_vector_agrees(oracle):
    rng = np.random.default_rng(9)
    def keys(n):
        ks = sorted({(int(a), int(b), int(c), int(d)) for a, b, c, d in rng.integers(0, 6, (n, 4))}, reverse=True)
        return ks
    for trial in range(20):
        these, those = keys(40), keys(40)
        want = hs.descZipVector(these, those)
        cols = lambda ks: {n: np.array([k[q] for k in ks], np.int32) for q, n in enumerate(("key_i", "key_j", "feat_a", "feat_b"))}
        lam_np = np.arange(len(those), dtype=np.float64) + 1.0
        lam_f = -np.arange(len(those), dtype=np.float64) - 1.0
        o_np, o_f, hit = oracle.warm_join(cols(these), cols(those), lam_np, lam_f)
        for k, m in enumerate(want):
            assert bool(hit[k]) == (m is not None)
            assert (o_np[k], o_f[k]) == ((lam_np[m], lam_f[m]) if m is not None else (0.0, 0.0))


@pytest.mark.parametrize("seed", [0, 1])
def test_grid_culled_keys_agrees(oracle, seed):
    """Grid.toGrid / culledKeys (Grid.hs:67-141, what updateWorld calls) == Aabb.culledKeys == the C oracle's
    grid restatement, on worlds inside and partly outside the engine's 20x20 grid (gridAxes, Engine/Main.hs:39-40)."""
    w = scenes.random_polygons(220, density=0.7, config=330 + seed)
    w.pos_x -= 9.0; w.pos_y -= 9.0                                    # the square [0, 17.7]^2 -> [-9, 8.7]^2 (+ overhang)
    c, s = oracle.cos_sin(w.rot)
    shapes, tagged = mirror_world(w, c, s)
    axes = ((20, 1.0, -10.0), (20, 1.0, -10.0))
    got = hs.grid_culledKeys(axes, tagged)
    assert got == hs.culledKeys(tagged)
    wx, wy, _, _ = oracle.move_shapes(w, c, s)
    boxes = oracle.aabbs(w, wx, wy)
    pi, pj = oracle.culled_keys_grid(w, boxes, oracle.is_static(w))
    assert got == list(zip(pi.tolist(), pj.tolist())) and len(got) > 20


def test_differential_fuzz_on_tie_heavy_pairs(oracle):
    """Coordinates on a coarse lattice (multiples of 1/4, rotations by multiples of 90 degrees) make exact
    ties, touching edges, coincident vertices and parallel clip planes the norm rather than the exception;
    non-finite and huge positions ride along.  Both restatements must still agree on every row."""
    rng = np.random.default_rng(21)
    boxes = [rectangle_vertices(w, h) for w in (0.5, 1.0, 2.0) for h in (0.5, 1.0, 1.5)]
    tri = [[(0.0, 0.0), (1.0, 0.0), (0.0, 1.0)], [(0.5, 0.0), (0.0, 0.75), (-0.5, 0.0)], [(1.0, 0.0), (0.0, 1.0), (-1.0, 0.0), (0.0, -1.0)],
           [(1.0, 0.5), (0.0, 1.0), (-1.0, 0.5), (-1.0, -0.5), (0.0, -1.0), (1.0, -0.5)]]
    shapes_pool = boxes + tri
    quarter = [(1.0, 0.0), (0.0, 1.0), (-1.0, 0.0), (0.0, -1.0)]
    n_rows = n_flip = n_one = 0
    for trial in range(60):
        objs = []
        for k in range(12):
            v = shapes_pool[rng.integers(len(shapes_pool))]
            pos = (float(rng.integers(-6, 7)) * 0.25, float(rng.integers(-6, 7)) * 0.25)
            objs.append((v, pos, 0.0, (1.0, 1.0) if rng.random() < 0.8 else (0.0, 0.0)))
        w = World.from_objects(objs)
        rot = rng.integers(0, 4, w.n_slots)
        c = np.array([quarter[r][0] for r in rot]); s = np.array([quarter[r][1] for r in rot])
        if trial % 10 == 9:                                           # non-finite / huge bodies
            w.pos_x[3] = np.inf; w.pos_y[5] = np.nan; w.pos_x[7] = 1e300
        fr = oracle.frame(w, c, s, broadphase="aabb")
        hulls = hs.hulls_of(w, c, s)
        pairs = list(zip(fr["pair_i"].tolist(), fr["pair_j"].tolist()))
        shapes, tagged = mirror_world(w, c, s)
        assert hs.culledKeys(tagged) == pairs
        rows = hs.prepare_frame(w, hulls, pairs, (0.01, 0.02), 0.01)
        assert len(rows) == len(fr["key_i"]), trial
        for k, r in enumerate(rows):
            assert r["key"] == (fr["key_i"][k], fr["key_j"][k], fr["feat_a"][k], fr["feat_b"][k]) and r["flip"] == fr["flip"][k], (trial, k)
            ct = r["contact"]
            assert same(ct["normal"][0], fr["normal_x"][k]) and same(ct["normal"][1], fr["normal_y"][k])
            assert same(ct["center"][0], fr["center_x"][k]) and same(ct["center"][1], fr["center_y"][k]) and same(ct["depth"], fr["depth"][k])
            jn, bn = r["constraint"]["nonpen"]
            jf, _ = r["constraint"]["friction"]
            for q in range(6):
                assert same(jn[q], fr[COLS_J[q]][k]) and same(jf[q], fr[COLS_F[q]][k]), (trial, k, q)
            assert same(bn, fr["b_np"][k])
        n_rows += len(rows); n_flip += sum(r["flip"] for r in rows)
        counts = {}
        for r in rows:
            counts[r["key"][:2]] = counts.get(r["key"][:2], 0) + 1
        n_one += sum(1 for v in counts.values() if v == 1)
    assert n_rows > 1500 and 0 < n_flip < n_rows and n_one > 20      # both branches, one- and two-point manifolds


@pytest.mark.filterwarnings("ignore::RuntimeWarning")   # numpy scalars warn on the 1/0 of parallel clip planes; the NaN is the point
def test_update_world_assembly_agrees(oracle):
    """Whole frames of Physics.Engine.Main.updateWorld (Engine/Main.hs:71-86) -- culledKeys, applyExternal,
    prepareFrame, applyCachedSlns (join + applySln), improveWorld x2, advance, moveShapes -- assembled twice:
    the C oracle's update_world against the mirror's updateWorld, 25 frames of boxes landing on the floor."""
    import math
    w = scenes.stacks_scene((4, 3), 0.0)
    w.pos_y[1:] -= 0.85
    n = w.n_slots
    b = Bodies.at_rest(n, 0.2, 0.0)
    rng = np.random.default_rng(5)
    b.vel_x[1:] = rng.uniform(-0.3, 0.3, n - 1); b.rot_vel[1:] = rng.uniform(-0.5, 0.5, n - 1)
    libm = lambda r: (math.cos(r), math.sin(r))
    c, s = oracle.cos_sin(w.rot)
    objs = [{"vel": (b.vel_x[k], b.vel_y[k]), "rotvel": b.rot_vel[k], "pos": (w.pos_x[k], w.pos_y[k]), "rot": w.rot[k],
             "inv": (w.inv_lin[k], w.inv_rot[k]), "cs": (c[k], s[k])} for k in range(n)]
    local = [[(w.local_x[v], w.local_y[v]) for v in range(w.vert_offset[k], w.vert_offset[k + 1])] for k in range(n)]
    mats = [(b.mu[k], b.bounce[k]) for k in range(n)]
    cache_o, cache_m = None, []
    saw_contacts = saw_hits = False
    for f in range(25):
        fr, cache_o, c, s = oracle.update_world(w, b, cache_o, c, s, external=(oracle.EXT_ACCEL, 0.0, -2.0), broadphase="aabb")
        cache_m = hs.updateWorld(objs, local, mats, cache_m, hs.constantAccel((0.0, -2.0)), 0.01, (0.01, 0.02), libm)
        assert len(cache_m) == len(cache_o[1]), f
        for k, (key, lam) in enumerate(cache_m):
            assert key == (cache_o[0]["key_i"][k], cache_o[0]["key_j"][k], cache_o[0]["feat_a"][k], cache_o[0]["feat_b"][k]), (f, k)
            assert same(lam[0], cache_o[1][k]) and same(lam[1], cache_o[2][k]), (f, k)
        for k in range(n):
            o = objs[k]
            assert same(o["vel"][0], b.vel_x[k]) and same(o["vel"][1], b.vel_y[k]) and same(o["rotvel"], b.rot_vel[k]), (f, k)
            assert same(o["pos"][0], w.pos_x[k]) and same(o["pos"][1], w.pos_y[k]) and same(o["rot"], w.rot[k]), (f, k)
            assert same(o["cs"][0], c[k]) and same(o["cs"][1], s[k]), (f, k)
        saw_contacts |= len(cache_m) > 0
        saw_hits |= bool(fr["warm_hit"].any())
    assert saw_contacts and saw_hits
