"""shapes_create_multi / shapes_multi_frame: several GPUs driven from ONE process (the form the reference's
single-threaded ST host can use), rows mode -- sweep / SAT work split by grid rows, results delivered to the
slot-range homes.  The assembled frame must be bit-identical to the oracle.  Skipped on boxes with fewer than 2 GPUs."""
import numpy as np
import pytest

from conftest import assert_frames_match
from shapes_b200 import scenes
from shapes_b200.world import World, rectangle_vertices

pytestmark = pytest.mark.gpu


def _n_gpus():
    import torch
    return torch.cuda.device_count()


def _worlds():
    yield "polygons", scenes.random_polygons(40_000, density=1.5, config=71)
    yield "pile", scenes.box_pile(150, 120)
    yield "blob", scenes.gaussian_blob(30_000, density=1.0)
    yield "circles", scenes.random_circles_and_polygons(8000)
    rng = np.random.default_rng(3)
    objs = []
    for k in range(3000):      # hulls with more than 8 vertices (per-thread pass) + big shapes at both ends of the key range
        nv = int(rng.integers(3, 9)) if k % 9 else int(rng.integers(9, 14))
        ang = np.sort(rng.uniform(0, 2 * np.pi, nv))
        r = rng.uniform(0.3, 0.5)
        objs.append(([(r * np.cos(a), r * np.sin(a)) for a in ang], (rng.uniform(0, 40), rng.uniform(0, 40)), rng.uniform(0, 6.28),
                     (0.0, 0.0) if k % 19 == 0 else (1.0, 1.0)))
    big = [(rectangle_vertices(42.0, 1.0), (20.0, 0.2), 0.0, (0.0, 0.0)),
           ([(6.0, 0.0), (0.0, 5.0), (-6.0, 0.0), (0.0, -5.0)], (20.0, 20.0), 0.3, (1.0, 1.0))]
    yield "big_and_long", World.from_objects(big[:1] + objs + big[1:])


@pytest.mark.parametrize("n_gpus", [2, 4, 8])
def test_multi_frame_matches_the_oracle(oracle, n_gpus):
    if _n_gpus() < n_gpus:
        pytest.skip(f"needs {n_gpus} GPUs")
    from shapes_b200.engine import MultiEngine
    for name, w in _worlds():
        c, s = oracle.cos_sin(w.rot)
        want = oracle.frame(w, c, s, broadphase="sweep")
        with MultiEngine(w, n_gpus) as eng:
            fr = eng.frame(cos_sin=(c, s))
            assert_frames_match(fr.cols, want)
            assert sum(eng.rank_pairs()) == len(want["pair_i"]), name
            # frames 2..: the row cuts follow the measured row weights, frames replay from graphs; the world drifts
            rng = np.random.default_rng(5)
            for step in range(4):
                w.pos_x += rng.uniform(-0.04, 0.04, w.n_slots); w.pos_y += rng.uniform(-0.04, 0.04, w.n_slots)
                c, s = oracle.cos_sin(w.rot)
                want = oracle.frame(w, c, s, broadphase="sweep")
                fr = eng.frame(cos_sin=(c, s), compact=(step % 2 == 1))
                assert_frames_match(fr.cols, want)
            # a jump: stale plan, every GPU re-seeds inside the same call
            w.pos_x += 3000.0
            c, s = oracle.cos_sin(w.rot)
            fr = eng.frame(cos_sin=(c, s))
            assert_frames_match(fr.cols, oracle.frame(w, c, s, broadphase="sweep"))


def test_multi_with_one_gpu_is_plain_shapes_frame(oracle):
    from shapes_b200.engine import MultiEngine
    w = scenes.random_polygons(5000, density=1.5, config=72)
    c, s = oracle.cos_sin(w.rot)
    with MultiEngine(w, 1) as eng:
        assert_frames_match(eng.frame(cos_sin=(c, s)).cols, oracle.frame(w, c, s, broadphase="sweep"))
