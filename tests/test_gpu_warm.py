"""Warm-start cache join on the device against the oracle's
restatement of descZipVector (Utils/Descending.hs:47-71) as applyCachedSlns uses it."""
import numpy as np
import pytest

from shapes_b200 import scenes

pytestmark = pytest.mark.gpu


def _lagr(cols):
    """A deterministic stand-in for the host solver's ContactLagrangian cache."""
    k = (cols["key_i"].astype(np.int64) * 1000003 + cols["key_j"] * 7919 + cols["feat_a"] * 31 + cols["feat_b"])
    return 0.25 + (k % 1013).astype(np.float64), -3.0 + (k % 607).astype(np.float64) / 8.0


@pytest.mark.parametrize("kind", ["pile", "polygons"])
def test_warm_join_matches_desc_zip_vector(oracle, kind):
    from shapes_b200.engine import Engine
    w = scenes.box_pile(200, 150) if kind == "pile" else scenes.random_polygons(60_000, density=2.0, config=81)
    c, s = oracle.cos_sin(w.rot)
    want = ("pairs", "contacts", "warm")
    with Engine(w) as eng:
        f0 = eng.frame_grow(cos_sin=(c, s), want=want)
        prev = {k: np.array(f0[k]) for k in ("key_i", "key_j", "feat_a", "feat_b")}
        assert not f0["warm_hit"].any() and not f0["warm_np"].any()      # no cache yet: newCache everywhere
        lam_np, lam_f = _lagr(prev)
        for step in range(3):
            # the world moves: some contacts persist, some vanish, some appear
            rng = np.random.default_rng(step)
            w.pos_x += rng.uniform(-0.03, 0.03, w.n_slots)
            w.pos_y += rng.uniform(-0.03, 0.03, w.n_slots)
            w.rot += rng.uniform(-0.02, 0.02, w.n_slots)
            c, s = oracle.cos_sin(w.rot)
            eng.set_lagrangian_cache(lam_np, lam_f)
            f1 = eng.frame_grow(cos_sin=(c, s), want=want)
            cur = {k: np.array(f1[k]) for k in ("key_i", "key_j", "feat_a", "feat_b")}
            o_np, o_f, o_hit = oracle.warm_join(cur, prev, lam_np, lam_f)
            assert np.array_equal(f1["warm_hit"], o_hit)
            assert np.array_equal(f1["warm_np"], o_np) and np.array_equal(f1["warm_f"], o_f)
            assert 0 < o_hit.sum() < len(o_hit)
            # independent check of the join itself: dictionary lookup on the 4-tuple keys
            table = {t: (a, b) for t, a, b in zip(zip(prev["key_i"], prev["key_j"], prev["feat_a"], prev["feat_b"]), lam_np, lam_f)}
            for k in range(0, len(o_hit), max(1, len(o_hit) // 500)):
                t = (cur["key_i"][k], cur["key_j"][k], cur["feat_a"][k], cur["feat_b"][k])
                assert (t in table) == bool(o_hit[k])
                if t in table:
                    assert table[t] == (o_np[k], o_f[k])
            prev = cur
            lam_np, lam_f = _lagr(prev)
        # a frame without a fresh cache falls back to all-new
        f2 = eng.frame(cos_sin=(c, s), want=want)
        assert not f2["warm_hit"].any()


def test_cache_size_must_match_previous_frame(oracle):
    from shapes_b200.engine import Engine, ShapesError
    w = scenes.box_pile(30, 20)
    with Engine(w) as eng:
        with pytest.raises(ShapesError):
            eng.set_lagrangian_cache(np.zeros(3), np.zeros(3))      # no frame yet
        f0 = eng.frame_grow(want=("contacts",))
        with pytest.raises(ShapesError):
            eng.set_lagrangian_cache(np.zeros(f0.n_contacts + 1), np.zeros(f0.n_contacts + 1))
        eng.set_lagrangian_cache(np.ones(f0.n_contacts), np.ones(f0.n_contacts))
        f1 = eng.frame(want=("contacts", "warm"))
        assert f1["warm_hit"].all() and (f1["warm_np"] == 1.0).all()   # nothing moved: every key persists


def test_capacity_error_keeps_keys_and_cache_for_the_retry(oracle):
    """A frame that fails with SHAPES_E_CAPACITY leaves the previous frame's keys and the supplied cache in place;
    after Engine.grow (shapes_grow, in place) the retried frame joins against them as if nothing had happened."""
    from shapes_b200.engine import CapacityError, Engine, ShapesError
    w = scenes.random_polygons(20_000, density=2.0, config=83)
    half = np.arange(w.n_slots) % 2 == 1
    home_x = w.pos_x.copy()
    w.pos_x[half] += 1.0e4                                   # every second shape far away: few pairs
    c, s = oracle.cos_sin(w.rot)
    want = ("pairs", "contacts", "warm")
    small = oracle.frame(w, c, s, broadphase="aabb")
    with Engine(w, max_pairs=len(small["pair_i"]) + 8, max_contacts=len(small["key_i"]) + 8) as eng:
        f0 = eng.frame(cos_sin=(c, s), want=want)
        prev = {k: np.array(f0[k]) for k in ("key_i", "key_j", "feat_a", "feat_b")}
        lam_np, lam_f = _lagr(prev)
        eng.set_lagrangian_cache(lam_np, lam_f)
        w.pos_x[:] = home_x                                   # everybody comes back: the lists overflow
        with pytest.raises(CapacityError) as e:
            eng.frame(cos_sin=(c, s), want=want)
        assert e.value.n_pairs > eng.max_pairs
        with pytest.raises(ShapesError):
            eng.fetch(want=want)                              # the failed attempt left no results to fetch
        eng.grow(e.value.n_pairs, e.value.n_contacts)
        f1 = eng.frame(cos_sin=(c, s), want=want)             # no second set_lagrangian_cache: the first one is still valid
        cur = {k: np.array(f1[k]) for k in ("key_i", "key_j", "feat_a", "feat_b")}
        o_np, o_f, o_hit = oracle.warm_join(cur, prev, lam_np, lam_f)
        assert 0 < o_hit.sum() < len(o_hit)
        assert np.array_equal(f1["warm_hit"], o_hit)
        assert np.array_equal(f1["warm_np"], o_np) and np.array_equal(f1["warm_f"], o_f)
        full = oracle.frame(w, c, s, broadphase="aabb")
        assert np.array_equal(f1["pair_i"], full["pair_i"]) and np.array_equal(f1["key_j"], full["key_j"])
