"""ctypes binding of oracle/libshapes_oracle.so (test infrastructure only).

PARITY UNPINNED by reference outputs (no GHC here); see shapes_oracle.h.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_DIR, "libshapes_oracle.so")

_i32p = C.POINTER(C.c_int32)
_u8p = C.POINTER(C.c_uint8)
_f64p = C.POINTER(C.c_double)

CONTACT_F64 = (["normal_x", "normal_y", "center_x", "center_y", "depth"]
               + [f"j_np{q}" for q in range(6)] + ["b_np", "ra_x", "ra_y", "rb_x", "rb_y", "rn_x", "rn_y"]
               + [f"j_f{q}" for q in range(6)] + ["b_f", "inv_eff_np", "inv_eff_f"])
CONTACT_I32 = ["key_i", "key_j", "feat_a", "feat_b"]


class ContactsOut(C.Structure):
    _fields_ = [
        ("cap", C.c_int64),
        ("key_i", _i32p), ("key_j", _i32p), ("feat_a", _i32p), ("feat_b", _i32p),
        ("flip", _u8p),
        ("normal_x", _f64p), ("normal_y", _f64p), ("center_x", _f64p), ("center_y", _f64p), ("depth", _f64p),
        ("j_np", _f64p * 6), ("b_np", _f64p),
        ("ra_x", _f64p), ("ra_y", _f64p), ("rb_x", _f64p), ("rb_y", _f64p), ("rn_x", _f64p), ("rn_y", _f64p),
        ("j_f", _f64p * 6), ("b_f", _f64p),
        ("inv_eff_np", _f64p), ("inv_eff_f", _f64p),
    ]


def build(force: bool = False) -> str:
    deps = [os.path.join(_DIR, f) for f in ("shapes_oracle.c", "shapes_oracle_step.c", "shapes_oracle.h", "Makefile")]
    if force or not os.path.exists(LIB_PATH) or \
            max(os.path.getmtime(d) for d in deps) > os.path.getmtime(LIB_PATH):
        subprocess.run(["make", "-C", _DIR, "-B", "libshapes_oracle.so"], check=True,
                       stdout=subprocess.DEVNULL)
    return LIB_PATH


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(LIB_PATH)
        _lib.orc_culled_keys_aabb.restype = C.c_int64
        _lib.orc_culled_keys_grid.restype = C.c_int64
        _lib.orc_culled_keys_sweep.restype = C.c_int64
        _lib.orc_unordered_pairs.restype = C.c_int64
        _lib.orc_contacts.restype = C.c_int64
        _lib.orc_dot_v2.restype = C.c_double
        _lib.orc_dot_v2.argtypes = [C.c_double] * 4
    return _lib


def _p(a, ty):
    return a.ctypes.data_as(ty) if a is not None else None


def _f(a):
    return _p(a, _f64p)


def cos_sin(rot: np.ndarray):
    c = np.empty_like(rot); s = np.empty_like(rot)
    lib().orc_cos_sin(C.c_int64(rot.shape[0]), _f(rot), _f(c), _f(s))
    return c, s


def hull_extents(world):
    emin = np.zeros(world.n_verts, np.int32); emax = np.zeros(world.n_verts, np.int32)
    lib().orc_hull_extents(C.c_int64(world.n_slots), _p(world.vert_offset, _i32p), _f(world.local_x),
                           _f(world.local_y), _p(emin, _i32p), _p(emax, _i32p))
    return emin, emax


def move_shapes(world, cos_rot, sin_rot):
    nv = world.n_verts
    wx = np.zeros(nv); wy = np.zeros(nv); nx = np.zeros(nv); ny = np.zeros(nv)
    lib().orc_move_shapes(C.c_int64(world.n_slots), _p(world.alive, _u8p), _p(world.vert_offset, _i32p),
                          _f(world.local_x), _f(world.local_y), _f(world.pos_x), _f(world.pos_y),
                          _f(cos_rot), _f(sin_rot), _f(wx), _f(wy), _f(nx), _f(ny))
    return wx, wy, nx, ny


def aabbs(world, wx, wy):
    n = world.n_slots
    out = [np.zeros(n) for _ in range(4)]
    lib().orc_aabbs(C.c_int64(n), _p(world.alive, _u8p), _p(world.vert_offset, _i32p), _f(wx), _f(wy),
                    *[_f(a) for a in out])
    return out


def is_static(world):
    st = np.zeros(world.n_slots, np.uint8)
    lib().orc_is_static(C.c_int64(world.n_slots), _f(world.inv_lin), _f(world.inv_rot), _p(st, _u8p))
    return st


def _culled(fn, world, boxes, static, extra=()):
    n = world.n_slots
    args = [C.c_int64(n), _p(world.alive, _u8p)] + [_f(b) for b in boxes] + [_p(static, _u8p)] + list(extra)
    cap = max(1024, 16 * n)
    while True:
        pi = np.zeros(cap, np.int32); pj = np.zeros(cap, np.int32)
        k = fn(*args, C.c_int64(cap), _p(pi, _i32p), _p(pj, _i32p))
        if k <= cap:
            return pi[:k].copy(), pj[:k].copy()
        cap = int(k)


def culled_keys_aabb(world, boxes, static):
    return _culled(lib().orc_culled_keys_aabb, world, boxes, static)


def culled_keys_sweep(world, boxes, static):
    return _culled(lib().orc_culled_keys_sweep, world, boxes, static)


def culled_keys_grid(world, boxes, static, x_axis=(20, 1.0, -10.0), y_axis=(20, 1.0, -10.0)):
    """Grid.culledKeys with gridAxes of Engine/Main.hs:40-41 by default."""
    extra = [C.c_int32(x_axis[0]), C.c_double(x_axis[1]), C.c_double(x_axis[2]),
             C.c_int32(y_axis[0]), C.c_double(y_axis[1]), C.c_double(y_axis[2])]
    return _culled(lib().orc_culled_keys_grid, world, boxes, static, extra)


def unordered_pairs(n: int):
    cap = max(1, n * (n - 1) // 2)
    xs = np.zeros(cap, np.int32); ys = np.zeros(cap, np.int32)
    k = lib().orc_unordered_pairs(C.c_int64(n), C.c_int64(cap), _p(xs, _i32p), _p(ys, _i32p))
    return xs[:k], ys[:k]


def move_circles(world, cos_rot, sin_rot):
    n = world.n_slots
    cx = np.zeros(n); cy = np.zeros(n)
    lib().orc_move_circles(C.c_int64(n), _p(world.alive, _u8p), _f(world.radius), _f(world.pos_x), _f(world.pos_y),
                           _f(cos_rot), _f(sin_rot), _f(cx), _f(cy))
    return cx, cy


def aabbs_circles(world, cx, cy, boxes):
    lib().orc_aabbs_circles(C.c_int64(world.n_slots), _p(world.alive, _u8p), _f(world.radius), _f(cx), _f(cy),
                            *[_f(b) for b in boxes])


def contacts(world, pair_i, pair_j, wx, wy, nx, ny, ext_min, ext_max, dt, baumgarte, slop, circles=None):
    """prepareFrame + constraintGen over the given pairs -> dict of columns.
    circles = (centre_x, centre_y) when the world has CircleShapes (full generateContacts dispatch)."""
    cap = max(16, 2 * int(pair_i.shape[0]))
    cols = {k: np.zeros(cap, np.int32) for k in CONTACT_I32}
    cols["flip"] = np.zeros(cap, np.uint8)
    for k in CONTACT_F64:
        cols[k] = np.zeros(cap, np.float64)
    out = ContactsOut()
    out.cap = cap
    for k in CONTACT_I32:
        setattr(out, k, _p(cols[k], _i32p))
    out.flip = _p(cols["flip"], _u8p)
    for k in CONTACT_F64:
        if k.startswith("j_np"):
            out.j_np[int(k[4:])] = _f(cols[k])
        elif k.startswith("j_f"):
            out.j_f[int(k[3:])] = _f(cols[k])
        else:
            setattr(out, k, _f(cols[k]))
    pi = np.ascontiguousarray(pair_i, np.int32); pj = np.ascontiguousarray(pair_j, np.int32)
    rad = world.radius if circles is not None else None
    ccx, ccy = circles if circles is not None else (None, None)
    lib().orc_contacts_shapes.restype = C.c_int64
    n = lib().orc_contacts_shapes(C.c_int64(pi.shape[0]), _p(pi, _i32p), _p(pj, _i32p), _p(world.vert_offset, _i32p),
                                  _f(wx), _f(wy), _f(nx), _f(ny), _p(ext_min, _i32p), _p(ext_max, _i32p),
                                  _f(rad), _f(ccx), _f(ccy),
                                  _f(world.pos_x), _f(world.pos_y), _f(world.inv_lin), _f(world.inv_rot),
                                  C.c_double(dt), C.c_double(baumgarte), C.c_double(slop), C.byref(out))
    assert n <= cap
    return {k: v[:n].copy() for k, v in cols.items()}


def frame(world, cos_rot=None, sin_rot=None, dt=0.01, baumgarte=0.01, slop=0.02, broadphase="auto",
          ext=None):
    """Whole hot path on the CPU: moveShapes -> culledKeys -> prepareFrame -> constraintGen."""
    if cos_rot is None:
        cos_rot, sin_rot = cos_sin(world.rot)
    emin, emax = ext if ext is not None else hull_extents(world)
    wx, wy, nx, ny = move_shapes(world, cos_rot, sin_rot)
    boxes = aabbs(world, wx, wy)
    circles = None
    if getattr(world, "radius", None) is not None:
        circles = move_circles(world, cos_rot, sin_rot)          # setCircleTransform
        aabbs_circles(world, circles[0], circles[1], boxes)      # circleToAabb
    static = is_static(world)
    if broadphase == "auto":
        broadphase = "aabb" if world.n_slots <= 3000 else "sweep"
    fn = {"aabb": culled_keys_aabb, "sweep": culled_keys_sweep, "grid": culled_keys_grid}[broadphase]
    pi, pj = fn(world, boxes, static)
    res = contacts(world, pi, pj, wx, wy, nx, ny, emin, emax, dt, baumgarte, slop, circles=circles)
    res.update(pair_i=pi, pair_j=pj, aabb_min_x=boxes[0], aabb_max_x=boxes[1], aabb_min_y=boxes[2],
               aabb_max_y=boxes[3], world_x=wx, world_y=wy, normal_wx=nx, normal_wy=ny,
               ext_min=emin, ext_max=emax, is_static=static)
    return res


def warm_join(this, that, that_np, that_f):
    """descZipVector join: `this`/`that` are dicts with key_i, key_j, feat_a, feat_b (descending)."""
    n = int(this["key_i"].shape[0]); m = int(that["key_i"].shape[0])
    out_np = np.zeros(n); out_f = np.zeros(n); hit = np.zeros(n, np.uint8)
    a = [np.ascontiguousarray(this[k], np.int32) for k in ("key_i", "key_j", "feat_a", "feat_b")]
    b = [np.ascontiguousarray(that[k], np.int32) for k in ("key_i", "key_j", "feat_a", "feat_b")]
    tn = np.ascontiguousarray(that_np, np.float64); tf = np.ascontiguousarray(that_f, np.float64)
    lib().orc_warm_join(C.c_int64(n), *[_p(x, _i32p) for x in a], C.c_int64(m), *[_p(x, _i32p) for x in b],
                        _f(tn), _f(tf), _f(out_np), _f(out_f), _p(hit, _u8p))
    return out_np, out_f, hit


# ---- the rest of updateWorld (shapes_oracle_step.c) ---------------------------------------------

EXT_NONE, EXT_ACCEL, EXT_FORCE = 0, 1, 2


def apply_external(world, vel_x, vel_y, kind, ex, ey, dt):
    """applyExternal (World.hs:156-158), in place on vel_x / vel_y."""
    lib().orc_apply_external(C.c_int64(world.n_slots), _p(world.alive, _u8p), C.c_int(kind), C.c_double(ex),
                             C.c_double(ey), C.c_double(dt), _f(world.inv_lin), _f(vel_x), _f(vel_y))


def advance(world, vel_x, vel_y, rot_vel, dt):
    """advance (World.hs:167-169), in place on world.pos_x / pos_y / rot."""
    lib().orc_advance(C.c_int64(world.n_slots), _p(world.alive, _u8p), C.c_double(dt), _f(vel_x), _f(vel_y),
                      _f(rot_vel), _f(world.pos_x), _f(world.pos_y), _f(world.rot))


def _cols6(fr, prefix):
    arr = (_f64p * 6)()
    keep = []
    for q in range(6):
        a = np.ascontiguousarray(fr[f"{prefix}{q}"], np.float64)
        keep.append(a)
        arr[q] = _f(a)
    return arr, keep


def apply_cached(world, fr, hit, lam_np, lam_f, vel_x, vel_y, rot_vel):
    """useCache's applySln for the rows the join hit (Solvers/Contact.hs:99-112)."""
    jn, k1 = _cols6(fr, "j_np"); jf, k2 = _cols6(fr, "j_f")
    lib().orc_apply_cached(C.c_int64(len(fr["key_i"])), _p(fr["key_i"], _i32p), _p(fr["key_j"], _i32p),
                           _p(hit, _u8p), _f(lam_np), _f(lam_f), jn, jf, _f(world.inv_lin), _f(world.inv_rot),
                           _f(vel_x), _f(vel_y), _f(rot_vel))


def improve_world(world, fr, mu, bounce, vel_x, vel_y, rot_vel, lam_np, lam_f):
    """One improveWorld solutionProcessor sweep (Solvers/Contact.hs:146-157), in place."""
    jn, k1 = _cols6(fr, "j_np"); jf, k2 = _cols6(fr, "j_f")
    lib().orc_improve_world(C.c_int64(len(fr["key_i"])), _p(fr["key_i"], _i32p), _p(fr["key_j"], _i32p),
                            jn, _f(fr["b_np"]), _f(fr["ra_x"]), _f(fr["ra_y"]), _f(fr["rb_x"]), _f(fr["rb_y"]),
                            _f(fr["rn_x"]), _f(fr["rn_y"]), jf, _f(mu), _f(bounce),
                            _f(world.inv_lin), _f(world.inv_rot), _f(vel_x), _f(vel_y), _f(rot_vel),
                            _f(lam_np), _f(lam_f))


def update_world(world, bodies, cache, cos_rot, sin_rot, dt=0.01, baumgarte=0.01, slop=0.02,
                 external=(EXT_NONE, 0.0, 0.0), iterations=2, sincos=None, broadphase="auto"):
    """Physics.Engine.Main.updateWorld (Engine/Main.hs:71-86) on the CPU, one frame, in place:
    culledKeys -> applyExternal -> prepareFrame -> applyCachedSlns -> improveWorld x iterations ->
    advance -> moveShapes.  `bodies` carries vel_x, vel_y, rot_vel, mu, bounce (numpy columns);
    `cache` is the EngineCache: None or (keys dict, lam_np, lam_f) from the previous frame;
    cos_rot / sin_rot are the rotation columns the shapes were last moved with; `sincos` maps the
    new rot column to (cos, sin) -- libm (the reference's rotate22) when None.
    Returns (frame dict, new cache, new cos, new sin)."""
    fr = frame(world, cos_rot, sin_rot, dt=dt, baumgarte=baumgarte, slop=slop, broadphase=broadphase)
    apply_external(world, bodies.vel_x, bodies.vel_y, external[0], external[1], external[2], dt)
    n = len(fr["key_i"])
    if cache is not None:
        lam_np, lam_f, hit = warm_join(fr, cache[0], cache[1], cache[2])
    else:
        lam_np, lam_f, hit = np.zeros(n), np.zeros(n), np.zeros(n, np.uint8)
    apply_cached(world, fr, hit, lam_np, lam_f, bodies.vel_x, bodies.vel_y, bodies.rot_vel)
    for _ in range(iterations):
        improve_world(world, fr, bodies.mu, bodies.bounce, bodies.vel_x, bodies.vel_y, bodies.rot_vel, lam_np, lam_f)
    advance(world, bodies.vel_x, bodies.vel_y, bodies.rot_vel, dt)
    c, s = (sincos or cos_sin)(world.rot)
    keys = {k: fr[k] for k in ("key_i", "key_j", "feat_a", "feat_b")}
    fr["warm_hit"] = hit
    return fr, (keys, lam_np, lam_f), c, s
