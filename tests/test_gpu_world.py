"""Device-resident world step (applyExternal, applyCachedSlns,
improveWorld, advance on the GPU) against the oracle's sequential restatement of
Physics.Engine.Main.updateWorld.  Everything is compared BIT for bit after every step: body state,
contact rows, the cache join and the Lagrangians the solver left -- the device solver executes the
reference's sequential Gauss-Seidel order as a dependency graph, so no tolerance is needed.
Both sides rotate with shapes_sincos (include/shapes_sincos.h; host copy vs device copy)."""
import copy

import numpy as np
import pytest

from shapes_b200 import scenes
from shapes_b200.world import Bodies

pytestmark = pytest.mark.gpu

STATE = ("vel_x", "vel_y", "rot_vel", "pos_x", "pos_y", "rot", "cos_rot", "sin_rot")
ROW_I = ("key_i", "key_j", "feat_a", "feat_b", "flip")
ROW_F = ("normal_x", "normal_y", "center_x", "center_y", "depth", "b_np", "inv_eff_np", "inv_eff_f", "j_np2", "j_f5")


def same(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return a.shape == b.shape and bool(((a == b) | (np.isnan(a) & np.isnan(b))).all())


def random_bodies(n, seed, speed=0.5, spin=0.5, mu=(0.0, 0.6), bounce=(0.0, 0.5)):
    rng = np.random.default_rng(seed)
    return Bodies(rng.uniform(-speed, speed, n), rng.uniform(-speed, speed, n), rng.uniform(-spin, spin, n),
                  rng.uniform(*mu, n), rng.uniform(*bounce, n))


def run_both(oracle, w, bodies, steps, external=(0, 0.0, 0.0), iterations=2, warm_start=True, dt=0.01,
             baumgarte=0.01, slop=0.02, check_rows=True, max_pairs=None):
    """Steps the same world on the device and through the oracle; asserts equality after every step.
    Returns per-step stats."""
    from shapes_b200 import engine
    from shapes_b200.engine import Engine
    wo = copy.deepcopy(w)
    bo = bodies.copy()
    c, s = engine.sincos(wo.rot)
    cache = None
    all_stats = []
    with Engine(w, max_pairs=max_pairs) as eng:
        eng.world_upload(bodies)
        d = eng.world_download()
        assert same(d["cos_rot"], c) and same(d["sin_rot"], s), "device shapes_sincos differs from the host copy"
        for step in range(steps):
            st = eng.world_step(dt=dt, baumgarte=baumgarte, slop=slop, external=external, iterations=iterations,
                                warm_start=warm_start)
            fr, new_cache, c, s = oracle.update_world(wo, bo, cache if warm_start else None, c, s, dt=dt, baumgarte=baumgarte,
                                                      slop=slop, external=external, iterations=iterations,
                                                      sincos=engine.sincos)
            cache = new_cache
            assert st.n_pairs == len(fr["pair_i"]) and st.n_contacts == len(fr["key_i"]), (step, st.n_pairs, st.n_contacts)
            d = eng.world_download()
            want = dict(vel_x=bo.vel_x, vel_y=bo.vel_y, rot_vel=bo.rot_vel, pos_x=wo.pos_x, pos_y=wo.pos_y, rot=wo.rot,
                        cos_rot=c, sin_rot=s)
            live = wo.alive.astype(bool)
            for k in STATE:
                bad = ~((d[k] == want[k]) | (np.isnan(d[k]) & np.isnan(want[k]))) & live
                assert not bad.any(), f"step {step}: {k} differs on {int(bad.sum())} bodies, first {np.nonzero(bad)[0][:5]}"
            if check_rows:
                got = eng.fetch(want=("pairs", "contacts", "constraints", "warm"))
                for k in ROW_I + ROW_F:
                    assert same(got[k], fr[k]), f"step {step}: row column {k}"
                assert same(got["warm_hit"], fr["warm_hit"]), f"step {step}: warm_hit"
                assert same(got["warm_np"], cache[1]) and same(got["warm_f"], cache[2]), f"step {step}: lagrangians"
            all_stats.append(st)
    return all_stats


def test_stacks_scene_many_frames(oracle):
    """The reference's own scene (Stacks.makeScene, gravity (0,-2), mu 0.2, bounce 0): free fall, impact,
    rocking and settling, 150 frames, every frame bit-identical."""
    w = scenes.stacks_scene((12, 8), 0.0)
    w.pos_y[1:] -= 0.8                      # start close to the floor so that most frames have contacts
    stats = run_both(oracle, w, Bodies.at_rest(w.n_slots, 0.2, 0.0), 150, external=(1, 0.0, -2.0))
    assert stats[-1].n_contacts > 0 and stats[-1].warm == 1 and stats[0].warm == 0
    assert any(s.queue_pushes > 0 for s in stats)


def test_dense_pile_chains(oracle):
    """Config 3's lattice (long dependency chains through every row) with random velocities and materials."""
    w = scenes.box_pile(120, 90)
    stats = run_both(oracle, w, random_bodies(w.n_slots, 1), 4, external=(1, 0.0, -9.8))
    assert stats[0].solver_nodes > 0 and stats[1].warm == 1


def test_random_polygons_and_circles(oracle):
    w = scenes.random_polygons(10_000, density=2.0, config=91)
    run_both(oracle, w, random_bodies(w.n_slots, 2), 4, external=(2, 0.3, -1.0))     # constantForce
    w = scenes.random_circles_and_polygons(6_000, config=92)
    run_both(oracle, w, random_bodies(w.n_slots, 3, speed=1.0), 4, external=(1, 0.0, -2.0))


@pytest.mark.parametrize("iterations,warm", [(0, True), (1, True), (5, True), (12, True), (2, False)])
def test_solver_sweeps_and_cold_start(oracle, iterations, warm):
    w = scenes.box_pile(60, 40)
    run_both(oracle, w, random_bodies(w.n_slots, 4), 3, external=(1, 0.0, -2.0), iterations=iterations, warm_start=warm)


def test_deleted_slots_and_kinematic_bodies(oracle):
    """Empty EmptiesVector slots are skipped by applyExternal / advance; a static body with a velocity
    (a moving platform) advances but is never changed by the solver."""
    w = scenes.box_pile(50, 30)
    w.delete([7, 8, 400, 1200])
    b = random_bodies(w.n_slots, 5)
    static = (w.inv_lin == 0.0) & (w.inv_rot == 0.0)
    assert static.any()
    b.vel_x[static] = 0.25; b.vel_y[static] = 0.0; b.rot_vel[static] = 0.0
    run_both(oracle, w, b, 4, external=(1, 0.0, -2.0))


def test_capacity_error_leaves_the_world_untouched(oracle):
    from shapes_b200.engine import CapacityError, Engine
    w = scenes.box_pile(40, 30)
    b = random_bodies(w.n_slots, 6)
    with Engine(w, max_pairs=64, max_contacts=128) as eng:
        eng.world_upload(b)
        before = eng.world_download()
        with pytest.raises(CapacityError) as e:
            eng.world_step(external=(1, 0.0, -2.0))
        assert e.value.n_pairs > 64
        after = eng.world_download()
        for k in STATE:
            assert same(before[k], after[k]), k


def test_capacity_growth_mid_run_keeps_the_engine_cache(oracle):
    """growing after SHAPES_E_CAPACITY must not drop the EngineCache.  The stacks land on the floor while
    the capacities are those of the first frame: a later step overflows, shapes_grow enlarges the ctx in place and the
    retried step still applies the cached Lagrangians (warm == 1) -- every step stays bit-identical to the oracle."""
    from shapes_b200 import engine
    from shapes_b200.engine import CapacityError, Engine
    w = scenes.stacks_scene((12, 8), 0.0)
    w.pos_y[1:] -= 0.895                       # a hair above the floor: the floor contacts appear after a few frames
    bodies = Bodies.at_rest(w.n_slots, 0.2, 0.0)
    wo, bo = copy.deepcopy(w), bodies.copy()
    c, s = engine.sincos(wo.rot)
    first = oracle.frame(wo, c, s, broadphase="aabb")
    n_pairs0, n_contacts0 = len(first["pair_i"]), len(first["key_i"])
    cache, grown_warm = None, 0
    with Engine(w, max_pairs=n_pairs0 + 2, max_contacts=n_contacts0 + 2) as eng:
        eng.world_upload(bodies)
        for step in range(40):
            for attempt in range(4):
                try:
                    st = eng.world_step(external=(1, 0.0, -2.0))
                    break
                except CapacityError as e:
                    assert step > 0, "the first frame must fit (capacities were taken from it)"
                    before = eng.world_download()
                    eng.grow(e.n_pairs, e.n_contacts)
                    after = eng.world_download()
                    for k in STATE:
                        assert same(before[k], after[k]), k          # the uploaded world survives the growth
                    grown_warm += 1
            fr, cache, c, s = oracle.update_world(wo, bo, cache, c, s, external=(1, 0.0, -2.0), sincos=engine.sincos)
            assert st.n_pairs == len(fr["pair_i"]) and st.n_contacts == len(fr["key_i"]), step
            assert st.warm == (1 if step > 0 else 0), step
            d = eng.world_download()
            want = dict(vel_x=bo.vel_x, vel_y=bo.vel_y, rot_vel=bo.rot_vel, pos_x=wo.pos_x, pos_y=wo.pos_y, rot=wo.rot,
                        cos_rot=c, sin_rot=s)
            for k in STATE:
                assert same(d[k], want[k]), (step, k)
            got = eng.fetch(want=("contacts", "warm"))
            assert same(got["warm_hit"], fr["warm_hit"]), step
            assert same(got["warm_np"], cache[1]) and same(got["warm_f"], cache[2]), step
        assert grown_warm >= 1, "the scenario never overflowed: the test did not exercise shapes_grow"


def test_step_needs_an_uploaded_world():
    from shapes_b200.engine import Engine, ShapesError
    w = scenes.box_pile(10, 10)
    with Engine(w) as eng:
        with pytest.raises(ShapesError):
            eng.world_step()


def test_worlds_without_contacts_and_tiny_worlds(oracle):
    """No pairs at all (bodies far apart), a single body, a single touching pair: the solver kernels
    are skipped or run with one node; integration still happens."""
    from shapes_b200.world import World, rectangle_vertices
    far = World.from_objects([(rectangle_vertices(1, 1), (10.0 * k, 0.0), 0.1 * k, (1.0, 1.0)) for k in range(50)])
    st = run_both(oracle, far, random_bodies(50, 9), 3, external=(1, 0.0, -2.0))
    assert st[-1].n_pairs == 0 and st[-1].solver_nodes == 0
    one = World.from_objects([(rectangle_vertices(1, 1), (0.0, 0.0), 0.3, (2.0, 1.0))])
    run_both(oracle, one, random_bodies(1, 10), 3, external=(2, 1.0, 0.0))
    two = World.from_objects([(rectangle_vertices(4, 4), (0.0, 0.0), 0.0, (0.0, 0.0)),
                              (rectangle_vertices(2, 2), (0.5, 2.9), 0.05, (1.0, 0.5))])
    st = run_both(oracle, two, random_bodies(2, 11, mu=(0.3, 0.3)), 20, external=(1, 0.0, -2.0))
    assert st[-1].n_contacts >= 1 and st[-1].body_chains == 1


def test_hot_path_call_between_steps_resets_the_cache(oracle):
    """shapes_frame between two world steps replaces 'the last frame': the next step then starts cold
    (ContactLagrangian 0 0 everywhere), exactly like an engine whose cache was dropped."""
    from shapes_b200 import engine
    from shapes_b200.engine import Engine
    w = scenes.box_pile(40, 30)
    b = random_bodies(w.n_slots, 12)
    wo = copy.deepcopy(w); bo = b.copy()
    c, s = engine.sincos(wo.rot)
    with Engine(w) as eng:
        eng.world_upload(b)
        eng.world_step(external=(1, 0.0, -2.0))
        fr, cache, c, s = oracle.update_world(wo, bo, None, c, s, external=(1, 0.0, -2.0), sincos=engine.sincos)
        eng.frame_grow(cos_sin=engine.sincos(w.rot))           # an unrelated hot-path call on the host copy
        st = eng.world_step(external=(1, 0.0, -2.0))
        assert st.warm == 0
        fr, cache, c, s = oracle.update_world(wo, bo, None, c, s, external=(1, 0.0, -2.0), sincos=engine.sincos)
        d = eng.world_download()
        assert same(d["vel_x"], bo.vel_x) and same(d["rot_vel"], bo.rot_vel) and same(d["pos_y"], wo.pos_y)


def test_schedule_independence(oracle):
    """Two runs of the same world give the same bits although the dataflow schedule differs."""
    from shapes_b200.engine import Engine
    w = scenes.box_pile(150, 100)
    b = random_bodies(w.n_slots, 7)
    outs = []
    for _ in range(2):
        with Engine(w) as eng:
            eng.world_upload(b)
            for _ in range(5):
                eng.world_step(external=(1, 0.0, -2.0))
            outs.append(eng.world_download())
    for k in STATE:
        assert same(outs[0][k], outs[1][k]), k


def test_config3_full_size_one_step(oracle):
    """BASELINE config 3 at full size (1M boxes): one cold step and one warm step against the sequential oracle."""
    w = scenes.box_pile(1000, 1000)
    stats = run_both(oracle, w, random_bodies(w.n_slots, 8, speed=0.2, spin=0.2), 2, external=(1, 0.0, -2.0),
                     check_rows=False)
    assert stats[1].warm == 1 and stats[1].n_contacts > 7_000_000
