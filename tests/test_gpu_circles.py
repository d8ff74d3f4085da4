"""Circle shapes through the C ABI: the full generateContacts dispatch
(shapes/src/Physics/Contact.hs:22-40) bit-exact against the oracle."""
import json
import os

import numpy as np
import pytest

from conftest import assert_frames_match
from shapes_b200 import scenes
from shapes_b200.world import World, rectangle_vertices

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
WANT = ("pairs", "contacts", "constraints", "aabb")


def run(oracle, w, broadphase="auto", **kw):
    from shapes_b200.engine import Engine
    c, s = oracle.cos_sin(w.rot)
    want = oracle.frame(w, c, s, broadphase=broadphase)
    with Engine(w) as eng:
        fr = eng.frame_grow(cos_sin=(c, s), want=WANT, **kw)
        got = {k: np.array(v) for k, v in fr.cols.items()}
    assert_frames_match(got, want)
    live = w.alive == 1
    for k in ("aabb_min_x", "aabb_max_x", "aabb_min_y", "aabb_max_y"):
        assert np.array_equal(got[k][live], want[k][live], equal_nan=True), k
    return got, want


def test_circle_kats_through_the_abi(oracle):
    with open(os.path.join(GOLDEN, "kat.json")) as f:
        kat = json.load(f)
    k4 = kat["kat4_circle_circle"]
    got, _ = run(oracle, World.from_objects([(k4["b"]["radius"], tuple(k4["b"]["center"]), 0.0, (1.0, 1.0)),
                                             (k4["a"]["radius"], tuple(k4["a"]["center"]), 0.0, (1.0, 1.0))]))
    c = k4["contact"]
    assert [got["feat_a"][0], got["feat_b"][0], got["flip"][0]] == c["feat"] + [c["flip"]]
    assert [got["normal_x"][0], got["normal_y"][0], got["center_x"][0], got["center_y"][0], got["depth"][0]] == \
        c["normal"] + c["center"] + [c["depth"]]
    k5 = kat["kat5_circle_hull"]
    box = (rectangle_vertices(*k5["box"]["size"]), tuple(k5["box"]["center"]), 0.0, (1.0, 1.0))
    cir = (k5["circle"]["radius"], tuple(k5["circle"]["center"]), 0.0, (1.0, 1.0))
    for objs, name in (([box, cir], "circle_is_a"), ([cir, box], "hull_is_a")):
        got, _ = run(oracle, World.from_objects(objs))
        c = k5[name]
        assert [got["feat_a"][0], got["feat_b"][0], got["flip"][0]] == c["feat"] + [c["flip"]]
        assert [got["normal_x"][0], got["normal_y"][0], got["center_x"][0], got["center_y"][0], got["depth"][0]] == \
            c["normal"] + c["center"] + [c["depth"]]


def test_balls_scene(oracle):
    """Balls.makeScene: alternating circle and box stacks on the floor, settled into contact."""
    w = scenes.balls_scene((12, 8), 0.5, -0.01)
    w.pos_y[1:] -= 0.9 + 0.002
    w.pos_x[1:] *= 0.985
    got, want = run(oracle, w, broadphase="aabb")
    kinds = set(zip((w.radius[want["key_i"]] >= 0).tolist(), (w.radius[want["key_j"]] >= 0).tolist()))
    assert kinds == {(True, True), (True, False), (False, True), (False, False)}, kinds


def test_random_circles_and_polygons(oracle):
    got, want = run(oracle, scenes.random_circles_and_polygons(40_000, density=2.0))
    assert len(want["key_i"]) > 10_000
    run(oracle, scenes.random_circles_and_polygons(2_000, circle_frac=1.0, density=3.0, config=61), broadphase="aabb")


def test_circles_with_deletes_static_and_deep_overlap(oracle):
    """Deep overlap (GJK simplex encloses the centre) yields NO contact in the reference
    (CircleVsHull.hs:29); static circles; deleted slots; a circle far larger than the rest (big path)."""
    w = scenes.random_circles_and_polygons(3_000, density=4.0, static_frac=0.3, config=62)
    w.delete(list(range(0, 3000, 11)))
    w.radius[5] = 25.0            # big circle, overlaps hundreds of shapes, swallows hull centres
    got, want = run(oracle, w, broadphase="aabb")
    assert (want["key_i"] == 5).sum() + (want["key_j"] == 5).sum() > 0
