"""Oracle restatement of the rest of updateWorld (applyExternal, applyCachedSlns' velocity part,
improveWorld, advance) against the one known answer the
reference's fixtures hold (KAT-2, solveConstraint) and hand-evaluated steps.  CPU only."""
import json
import os

import numpy as np
import pytest

from shapes_b200 import scenes
from shapes_b200.world import Bodies, World, rectangle_vertices


@pytest.fixture(scope="module")
def kat():
    with open(os.path.join(os.path.dirname(__file__), "golden", "kat.json")) as f:
        return json.load(f)


def _one_row(j_np, j_f, b_np=0.0, ra=(0.0, 0.0), rb=(0.0, 0.0), rn=(0.0, 0.0)):
    fr = {"key_i": np.array([1], np.int32), "key_j": np.array([0], np.int32), "b_np": np.array([b_np]),
          "ra_x": np.array([ra[0]]), "ra_y": np.array([ra[1]]), "rb_x": np.array([rb[0]]), "rb_y": np.array([rb[1]]),
          "rn_x": np.array([rn[0]]), "rn_y": np.array([rn[1]])}
    for q in range(6):
        fr[f"j_np{q}"] = np.array([float(j_np[q])]); fr[f"j_f{q}"] = np.array([float(j_f[q])])
    return fr


def _two_bodies(inv_a, inv_b):
    """slot 1 = object a (the larger key comes first in a pair), slot 0 = object b"""
    w = World.from_objects([(rectangle_vertices(1, 1), (0.0, 4.0), 0.0, (1.0, 1.0)),
                            (rectangle_vertices(1, 1), (0.0, 0.0), 0.0, (1.0, 1.0))])
    w.inv_lin[:] = [inv_b[0], inv_a[0]]; w.inv_rot[:] = [inv_b[1], inv_a[1]]
    return w


def test_kat2_through_improve_world(oracle, kat):
    """solveConstraint testConstraint testObjPair (bench/Physics/Constraint/Benchmark.hs:11-35) is what
    improveContactSln does for a fresh contact with mu = 0 and bounce = 0: J.v = -2, effMass = 2,
    lambda = 1, both bodies end at (1, 0, 0)."""
    k = kat["kat2_solveConstraint"]
    j = k["j"]; im = k["inv_mass6"]; v = k["vel6_in"]
    w = _two_bodies((im[0], im[2]), (im[3], im[5]))
    b = Bodies(np.array([v[3], v[0]]), np.array([v[4], v[1]]), np.array([v[5], v[2]]), np.zeros(2), np.zeros(2))
    jf = [-1.0, 0.0, 2.0, 1.0, -0.0, 2.0]           # Friction.jacobian for n = (0,1), p = (0,2), xa = (0,0), xb = (0,4)
    fr = _one_row(j, jf, b_np=k["b"])
    lam_np, lam_f = np.zeros(1), np.zeros(1)
    oracle.improve_world(w, fr, b.mu, b.bounce, b.vel_x, b.vel_y, b.rot_vel, lam_np, lam_f)
    out = [b.vel_x[1], b.vel_y[1], b.rot_vel[1], b.vel_x[0], b.vel_y[0], b.rot_vel[0]]
    assert out == k["vel6_out"]
    assert lam_np[0] == 1.0 and lam_f[0] == 0.0      # clampAbs with maxThresh = 1 * 0


def test_solution_processors_by_hand(oracle):
    """positive (SolutionProcessors.hs:28-35) never lets the accumulated non-penetration impulse go
    negative; clampAbs (:37-53) keeps |friction| <= mu * non-penetration."""
    w = _two_bodies((1.0, 0.0), (1.0, 0.0))
    jn = [0.0, -1.0, 0.0, 0.0, 1.0, 0.0]
    jf = [-1.0, 0.0, 0.0, 1.0, 0.0, 0.0]
    # separating bodies with a cached impulse of 0.25: new = -(J.v)/2 = -1.5, apply = max(-1.5, -0.25) = -0.25
    b = Bodies(np.array([0.0, 0.0]), np.array([2.0, -1.0]), np.zeros(2), np.array([0.5, 0.5]), np.zeros(2))
    lam_np, lam_f = np.array([0.25]), np.array([0.0])
    oracle.improve_world(w, _one_row(jn, jf), b.mu, b.bounce, b.vel_x, b.vel_y, b.rot_vel, lam_np, lam_f)
    assert lam_np[0] == 0.0 and lam_f[0] == 0.0
    assert b.vel_y.tolist() == [2.0 - 0.25, -1.0 + 0.25]
    # approaching at 2 with sliding 3: lambda_np = 1, friction wants -(J_f.v)/2 = -1.5, clamped to -mu*1 = -0.5
    b = Bodies(np.array([3.0, 0.0]), np.array([-1.0, 1.0]), np.zeros(2), np.array([0.5, 0.5]), np.zeros(2))
    lam_np, lam_f = np.zeros(1), np.zeros(1)
    oracle.improve_world(w, _one_row(jn, jf), b.mu, b.bounce, b.vel_x, b.vel_y, b.rot_vel, lam_np, lam_f)
    assert lam_np[0] == 1.0 and lam_f[0] == -0.5
    assert b.vel_y.tolist() == [0.0, 0.0]
    assert b.vel_x.tolist() == [3.0 - 0.5, 0.0 + 0.5]


def test_restitution_bias_by_hand(oracle):
    """bounceB (Restitution.hs:34-47): min 0 (min(bounce_a, bounce_b) * (closingVelocity . n)); with
    closing speed -2 along n and bounciness 0.5 the bias is -1, so lambda = -(J.v + b)/mc = (2 + 1)/2."""
    w = _two_bodies((1.0, 0.0), (1.0, 0.0))
    jn = [0.0, -1.0, 0.0, 0.0, 1.0, 0.0]
    jf = [-1.0, 0.0, 0.0, 1.0, 0.0, 0.0]
    b = Bodies(np.zeros(2), np.array([-1.0, 1.0]), np.zeros(2), np.zeros(2), np.array([0.5, 0.75]))
    lam_np, lam_f = np.zeros(1), np.zeros(1)
    oracle.improve_world(w, _one_row(jn, jf, rn=(0.0, 1.0)), b.mu, b.bounce, b.vel_x, b.vel_y, b.rot_vel, lam_np, lam_f)
    assert lam_np[0] == 1.5
    assert b.vel_y.tolist() == [-1.0 + 1.5, 1.0 - 1.5]


def test_apply_cached_equals_apply_sln(oracle):
    w = _two_bodies((1.0, 2.0), (0.5, 0.0))
    jn = [0.0, -1.0, 0.25, 0.0, 1.0, -0.5]
    jf = [-1.0, 0.0, 2.0, 1.0, 0.0, 2.0]
    b = Bodies.at_rest(2)
    fr = _one_row(jn, jf)
    oracle.apply_cached(w, fr, np.array([0], np.uint8), np.array([3.0]), np.array([-1.0]), b.vel_x, b.vel_y, b.rot_vel)
    assert not b.vel_x.any() and not b.vel_y.any()        # newCache: nothing applied
    oracle.apply_cached(w, fr, np.array([1], np.uint8), np.array([3.0]), np.array([-1.0]), b.vel_x, b.vel_y, b.rot_vel)
    # a = slot 1 (inv 1, 2), b = slot 0 (inv 0.5, 0): v += (j*l)*im for non-penetration, then friction
    assert [b.vel_x[1], b.vel_y[1], b.rot_vel[1]] == [0.0 + (-1.0 * -1.0) * 1.0, (-1.0 * 3.0) * 1.0, (0.25 * 3.0) * 2.0 + (2.0 * -1.0) * 2.0]
    assert [b.vel_x[0], b.vel_y[0], b.rot_vel[0]] == [(1.0 * -1.0) * 0.5, (1.0 * 3.0) * 0.5, 0.0]


def test_external_and_advance(oracle):
    w = scenes.stacks_scene((3, 2), 0.0)
    n = w.n_slots
    b = Bodies.at_rest(n)
    b.vel_x[:] = 1.0
    oracle.apply_external(w, b.vel_x, b.vel_y, oracle.EXT_ACCEL, 0.5, -2.0, 0.01)
    static = w.inv_lin == 0.0
    assert static[0] and not static[1:].any()
    assert b.vel_y[0] == 0.0 and b.vel_x[0] == 1.0                       # isStaticLin: untouched
    assert np.all(b.vel_y[1:] == -2.0 * 0.01) and np.all(b.vel_x[1:] == 1.0 + 0.5 * 0.01)
    # constantForce as the reference parses it: (v + f*dt) * inv_lin
    b2 = Bodies.at_rest(n); b2.vel_x[:] = 1.0
    oracle.apply_external(w, b2.vel_x, b2.vel_y, oracle.EXT_FORCE, 0.0, -2.0, 0.01)
    assert np.array_equal(b2.vel_x, (1.0 + 0.0 * 0.01) * w.inv_lin) and np.array_equal(b2.vel_y, (0.0 + -2.0 * 0.01) * w.inv_lin)
    # dead slots are skipped (EmptiesVector traversal)
    w.delete([2])
    b.rot_vel[:] = 3.0
    px, py, rot = w.pos_x.copy(), w.pos_y.copy(), w.rot.copy()
    oracle.advance(w, b.vel_x, b.vel_y, b.rot_vel, 0.01)
    live = w.alive.astype(bool)
    assert np.array_equal(w.pos_x[live], b.vel_x[live] * 0.01 + px[live])
    assert np.array_equal(w.pos_y[live], b.vel_y[live] * 0.01 + py[live])
    assert np.array_equal(w.rot[live], 0.01 * b.rot_vel[live] + rot[live])
    assert w.pos_x[2] == px[2] and w.rot[2] == rot[2]


def test_update_world_box_lands_on_the_floor(oracle):
    """Behavioural check of the assembled frame loop on the reference's Stacks scene parameters
    (Stacks.hs: mu 0.2, bounce 0, gravity (0,-2), ContactBehavior 0.01 0.02): a dropped box lands on the
    floor, rocks (the sequential solver hands the whole impact to the first contact) and comes to rest;
    a 6x5 stack never sinks into the floor.  Warm starting carries the resting impulse."""
    w = scenes.stacks_scene((1, 1), 0.0)
    b = Bodies.at_rest(w.n_slots, 0.2, 0.0)
    c, s = oracle.cos_sin(w.rot)
    cache = None
    for f in range(400):
        fr, cache, c, s = oracle.update_world(w, b, cache, c, s, external=(oracle.EXT_ACCEL, 0.0, -2.0))
    assert abs(w.pos_y[1] - (-6.0 + 0.5 + 0.1)) < 0.01 and abs(w.rot[1]) < 0.05
    assert abs(b.vel_y[1]) < 0.02 and abs(b.vel_x[1]) < 0.02
    assert fr["warm_hit"].all() and 0.03 < cache[1].max() < 0.05      # ~ m g dt = 2 * 2 * 0.01 at rest
    assert w.pos_y[0] == -6.0 and b.vel_y[0] == 0.0                    # the static floor never moves
    w = scenes.stacks_scene((6, 5), 0.0)
    b = Bodies.at_rest(w.n_slots, 0.2, 0.0)
    c, s = oracle.cos_sin(w.rot)
    cache = None
    for f in range(250):
        fr, cache, c, s = oracle.update_world(w, b, cache, c, s, external=(oracle.EXT_ACCEL, 0.0, -2.0))
        assert w.pos_y[1:].min() > -5.45
    assert np.isfinite(b.vel_x).all() and np.abs(b.vel_y).max() < 3.0
