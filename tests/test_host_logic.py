"""Scene generators and world storage (host logic).  CPU only."""
import numpy as np

from shapes_b200 import scenes
from shapes_b200.world import World, rectangle_vertices, to_inv_mass2


def test_splitmix64_reference_vector():
    # first outputs of Vigna's splitmix64.c for seed 1234567
    g = scenes.SplitMix64(1234567)
    got = [int(x) for x in g.u64(5)]
    assert got == [6457827717110365317, 3203168211198807973, 9817491932198370423,
                   4593380528125082431, 16408922859458223821]
    # the stream continues where it stopped
    g2 = scenes.SplitMix64(1234567)
    assert [int(x) for x in np.concatenate([g2.u64(2), g2.u64(3)])] == got


def test_rectangle_and_mass_helpers():
    assert rectangle_vertices(2, 4) == [(1, 2), (-1, 2), (-1, -2), (1, -2)]
    assert to_inv_mass2((2.0, 1.0)) == (0.5, 1.0) and to_inv_mass2((0.0, 0.0)) == (0.0, 0.0)


def test_stacks_scene_layout():
    """Stacks.makeScene (30,30) 0: floor first, then columns left to right, bottom to top,
    coordinates by repeated addition (Stacks.hs:34-56)."""
    w = scenes.stacks_scene()
    assert w.n_slots == 901 and w.n_verts == 3604
    assert (w.pos_x[0], w.pos_y[0]) == (0.0, -6.0) and w.inv_lin[0] == 0.0 and w.inv_rot[0] == 0.0
    left = 0.0 - (0.2 * 29.0 / 2.0)
    assert w.pos_x[1] == left and w.pos_y[1] == -4.5
    y = -4.5
    for k in range(1, 31):
        assert w.pos_y[k] == y
        y = y + (0.2 + 0.0)
    assert w.pos_x[31] == left + 0.2
    assert np.all(w.inv_lin[1:] == 0.5) and np.all(w.inv_rot[1:] == 1.0)


def test_generators_are_deterministic_and_convex():
    a = scenes.random_polygons(500)
    b = scenes.random_polygons(500)
    assert np.array_equal(a.local_x, b.local_x) and np.array_equal(a.pos_x, b.pos_x)
    nv = np.diff(a.vert_offset)
    assert nv.min() >= 3 and nv.max() <= 8
    # CCW and convex: every consecutive cross product is positive
    for s in range(0, 500, 7):
        o, n = a.vert_offset[s], nv[s]
        x, y = a.local_x[o:o + n], a.local_y[o:o + n]
        ex, ey = np.roll(x, -1) - x, np.roll(y, -1) - y
        cr = ex * np.roll(ey, -1) - ey * np.roll(ex, -1)
        assert np.all(cr > 0)


def test_pile_mixed_blob_shapes():
    p = scenes.box_pile(20, 10)
    assert p.n_slots == 201 and p.inv_lin[0] == 0.0 and np.all(np.diff(p.vert_offset) == 4)
    m = scenes.mixed_polygons(1000)
    nv = np.diff(m.vert_offset)
    assert m.n_slots == 1000 and (nv == 4).sum() >= 500 and m.n_verts == nv.sum()
    g = scenes.gaussian_blob(2000)
    assert abs(g.pos_x.mean()) < 3 * g.meta["sigma"] / np.sqrt(2000) * 2


def test_world_delete_keeps_keys_sparse():
    w = World.from_objects([(rectangle_vertices(1, 1), (float(k), 0.0), 0.0, (1.0, 1.0)) for k in range(4)])
    w.delete([1])
    assert w.alive.tolist() == [1, 0, 1, 1] and w.n_slots == 4


def test_expand_rows_rebuilds_the_derived_columns_bit_for_bit(oracle):
    """Compact wire format (engine.CONSTRAINT_COMPACT_F64): the sixteen derived constraint columns are rebuilt
    from the shipped ones exactly -- checked on the oracle's own rows (flips, signed zeros, clipped points)."""
    import numpy as np
    from shapes_b200 import engine, scenes
    for w in (scenes.random_polygons(3000, density=3.0, static_frac=0.1, config=51), scenes.box_pile(40, 30),
              scenes.stacks_scene((12, 8), 0.0)):
        c, s = oracle.cos_sin(w.rot)
        want = oracle.frame(w, c, s)
        assert len(want["key_i"]) > 100
        shipped = ("key_i", "key_j", "feat_a", "feat_b", "flip", "normal_x", "normal_y", "center_x", "center_y", "depth") \
            + engine.CONSTRAINT_COMPACT_F64
        cols = {k: np.array(want[k]) for k in shipped}
        engine.expand_rows(cols, w.pos_x, w.pos_y)
        for k in engine.CONSTRAINT_F64:
            a, b = cols[k], np.asarray(want[k])
            assert a.shape == b.shape and np.array_equal(a.view(np.uint64), b.view(np.uint64)), k
        if w.name.startswith("polygons"):
            assert want["flip"].any() and not want["flip"].all()
