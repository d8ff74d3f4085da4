"""World generators: the reference's own scenes and the BASELINE.json configs.

Reference scenes are restated with the reference's floating-point evaluation
order (repeated addition), because whether neighbouring boxes' AABBs touch is
decided by last-bit rounding ("knife-edge inputs").
Synthetic configs use a counter-based splitmix64 stream seeded
0x5348415045530000 + config number.
"""
from __future__ import annotations

import math

import numpy as np

from .world import World, rectangle_vertices

SEED_BASE = 0x5348415045530000
_M64 = (1 << 64) - 1


class SplitMix64:
    """Vectorised splitmix64: element k of the stream is mix(seed + (k+1)*GOLDEN)."""
    GOLDEN = 0x9E3779B97F4A7C15

    def __init__(self, seed: int):
        self.state = seed & _M64

    def u64(self, n: int) -> np.ndarray:
        with np.errstate(over="ignore"):
            k = np.arange(1, n + 1, dtype=np.uint64)
            z = np.uint64(self.state) + k * np.uint64(self.GOLDEN)
            self.state = (self.state + n * self.GOLDEN) & _M64
            z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
            z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
            return z ^ (z >> np.uint64(31))

    def uniform(self, n: int, lo: float = 0.0, hi: float = 1.0) -> np.ndarray:
        u = (self.u64(n) >> np.uint64(11)).astype(np.float64) * (1.0 / (1 << 53))
        return lo + (hi - lo) * u

    def normal(self, n: int) -> np.ndarray:
        u1 = 1.0 - self.uniform(n)
        u2 = self.uniform(n)
        return np.sqrt(-2.0 * np.log(u1)) * np.cos(2.0 * math.pi * u2)


# ---------------------------------------------------------------------------
# reference scenes
# ---------------------------------------------------------------------------

_BOX_MASS = (2.0, 1.0)       # Stacks.box: makePhysicalObj ... (2, 1) (Stacks.hs:13-17)


def _box_stack(size, bottom, spacing, n):
    """boxStack (shapes/src/Physics/Scenes/Stacks.hs:34-44): y' = y + (h + spacing)."""
    w, h = size
    x, y = bottom
    out = []
    for _ in range(n):
        out.append((rectangle_vertices(w, h), (x, y), 0.0, _BOX_MASS))
        y = y + (h + spacing)
    return out


def _stacks(size, center_bottom, spacing, dims):
    """stacks (Stacks.hs:46-56): lefts = take n_w (iterate (+ w) leftmost)."""
    w, _ = size
    center, bottom = center_bottom
    n_w, n_h = dims
    left = center - (w * float(n_w - 1) / 2.0)
    out = []
    for _ in range(n_w):
        out.extend(_box_stack(size, (left, bottom), spacing, n_h))
        left = left + w
    return out


def _box_floor():
    """boxFloor' (Stacks.hs:19-32): static 18x1 rectangle at (0, -6)."""
    return (rectangle_vertices(18.0, 1.0), (0.0, -6.0), 0.0, (0.0, 0.0))


def stacks_scene(dims=(30, 30), spacing=0.0) -> World:
    """Stacks.makeScene dims spacing (Stacks.hs:110-113) -- BASELINE config 1 with (30,30) 0."""
    objs = [_box_floor()] + _stacks((0.2, 0.2), (0.0, -4.5), spacing, dims)
    w = World.from_objects(objs, name=f"stacks{dims[0]}x{dims[1]}")
    w.meta.update(dt=0.01, baumgarte=0.01, slop=0.02)   # contactBehavior (Stacks.hs:87-88)
    return w


def broadphase_bench_world(spacing=0.0, dims=(30, 30)) -> World:
    """testWorld of shapes/bench/Physics/Broadphase/Benchmark.hs:50-52 (no floor)."""
    return World.from_objects(_stacks((0.2, 0.2), (0.0, -4.5), spacing, dims), name="bp_bench")


def balls_scene(dims=(10, 10), diameter=0.5, spacing=0.0) -> World:
    """Balls.makeScene dims diameter spacing (shapes/src/Physics/Scenes/Balls.hs:69-72): the floor plus
    columns that alternate circle stacks and box stacks (stacks_ not, :36-57), repeated additions."""
    n_w, n_h = dims
    left = 0.0 - (diameter * float(n_w - 1) / 2.0)
    objs = [_box_floor()]
    is_circle = True
    for _ in range(n_w):
        y = -4.5
        for _ in range(n_h):
            shape = diameter / 2.0 if is_circle else rectangle_vertices(diameter, diameter)
            objs.append((shape, (left, y), 0.0, _BOX_MASS))
            y = y + (diameter + spacing)
        left = left + diameter
        is_circle = not is_circle
    w = World.from_objects(objs, name=f"balls{n_w}x{n_h}")
    w.meta.update(dt=0.01, baumgarte=0.01, slop=0.02)
    return w


def random_circles_and_polygons(n=10_000, circle_frac=0.5, density=1.5, static_frac=0.05, config=6) -> World:
    """Circles (radius U[0.2, 0.5]) mixed with convex polygons, uniform density."""
    rng = SplitMix64(SEED_BASE + config)
    is_c = rng.uniform(n) < circle_frac
    n_poly = int((~is_c).sum())
    off_p, lx, ly = _polygons(rng, n_poly)
    nv = np.zeros(n, np.int64)
    nv[~is_c] = np.diff(off_p)
    off = np.zeros(n + 1, np.int64)
    np.cumsum(nv, out=off[1:])
    radius = np.where(is_c, rng.uniform(n, 0.2, 0.5), -1.0)
    side = math.sqrt(n / density)
    px = rng.uniform(n, 0.0, side); py = rng.uniform(n, 0.0, side)
    rot = rng.uniform(n, 0.0, 2.0 * math.pi)
    static = rng.uniform(n) < static_frac
    w = _finish(f"circles_polygons{n}", off, lx, ly, px, py, rot, static, {"config": config})
    w.radius = np.ascontiguousarray(radius, np.float64)
    return w.validate()


def test_opt_boxes() -> World:
    """testOptBoxes (shapes/bench/Physics/Contact/Benchmark.hs:16-27): a 4x4 box at (0,0)
    and a 2x2 box at (1,3).  S.contact a b takes a = first, so the 4x4 box gets the larger
    key (slot 1) and the 2x2 box slot 0.  The fixture bakes the translation into the
    vertices; positions here carry it instead (identical world vertices)."""
    w = World.from_objects([
        (rectangle_vertices(2.0, 2.0), (1.0, 3.0), 0.0, (1.0, 1.0)),
        (rectangle_vertices(4.0, 4.0), (0.0, 0.0), 0.0, (1.0, 1.0)),
    ], name="testOptBoxes")
    w.meta.update(dt=0.01, baumgarte=0.01, slop=0.02)
    return w


def kat_triangle_on_box(triangle_is_a: bool = False, lift: float = 0.0):
    """KAT-6 / KAT-7 / KAT-8 (tests/golden/README.md): a 32x16 box at (1,-2) and a small right triangle (sides
    2.5 / 4.6875 / 5.3125, every edge a Pythagorean direction) turned by a quarter turn with cos = 0, sin = 1 given
    exactly, its tip poking through the box's top edge next to the top-left corner.  Every coordinate is a dyadic
    rational, so moveShapes, the clip points, depths, Jacobians and inverse effective masses are exact.
    Returns (world, cos, sin): the triangle is slot 0 (shape b) unless triangle_is_a; `lift` raises it."""
    tri = ([(-0.5, 1.5), (-2.0, -0.5), (1.75, -3.3125)], (-14.0, 6.0 + lift), 0.0, (0.5, 4.0))      # inverse masses (2, 0.25)
    box = (rectangle_vertices(32.0, 16.0), (1.0, -2.0), 0.0, (2.0, 8.0))                            # inverse masses (0.5, 0.125)
    objs = [box, tri] if triangle_is_a else [tri, box]
    w = World.from_objects(objs, name="kat_triangle_on_box")
    w.meta.update(dt=0.25, baumgarte=0.5, slop=0.25)
    t = 1 if triangle_is_a else 0
    c = np.ones(2); s = np.zeros(2)
    c[t], s[t] = 0.0, 1.0
    w.rot[t] = math.pi / 2        # informative only: the frame is evaluated with the exact (cos, sin) returned here
    return w, c, s


def kat_box_on_hexagon():
    """KAT-9 (tests/golden/README.md): a hexagon whose six edges are axis-parallel or 3-4-5 directions (slot 1 =
    shape a) and a 2x2 box (slot 0 = shape b) resting 0.25 deep in its top edge.  A world with a hull of more than four
    vertices: on the GPU this vector goes through the cell-ordered work list and k_manifolds_coop."""
    hexagon = ([(5.0, 0.0), (2.0, 4.0), (-2.0, 4.0), (-5.0, 0.0), (-2.0, -4.0), (2.0, -4.0)], (0.0, 0.0), 0.0, (1.0, 1.0))
    box = (rectangle_vertices(2.0, 2.0), (0.5, 4.75), 0.0, (1.0, 1.0))
    w = World.from_objects([box, hexagon], name="kat_box_on_hexagon")
    w.meta.update(dt=0.25, baumgarte=0.5, slop=0.125)
    return w, np.ones(2), np.zeros(2)


# ---------------------------------------------------------------------------
# BASELINE.json configs 2-5
# ---------------------------------------------------------------------------

def _polygons(rng: SplitMix64, n: int, vmin=3, vmax=8, rmin=0.3, rmax=0.5):
    """n convex polygons: V uniform in {vmin..vmax}, sorted random angles on a circle."""
    nv = (vmin + (rng.u64(n) % np.uint64(vmax - vmin + 1))).astype(np.int32)
    r = rng.uniform(n, rmin, rmax)
    ang = rng.uniform(n * vmax, 0.0, 2.0 * math.pi).reshape(n, vmax)
    ang[np.arange(vmax)[None, :] >= nv[:, None]] = np.inf
    ang.sort(axis=1)
    mask = np.isfinite(ang)
    off = np.zeros(n + 1, np.int32)
    np.cumsum(nv, out=off[1:])
    a = ang[mask]
    rr = np.repeat(r, nv)
    return off, rr * np.cos(a), rr * np.sin(a)


def _finish(name, off, lx, ly, px, py, rot, static_mask, meta=None) -> World:
    n = px.shape[0]
    il = np.where(static_mask, 0.0, 0.5)
    ir = np.where(static_mask, 0.0, 1.0)
    w = World(np.ones(n, np.uint8), off.astype(np.int32), np.ascontiguousarray(lx, np.float64),
              np.ascontiguousarray(ly, np.float64), np.ascontiguousarray(px, np.float64),
              np.ascontiguousarray(py, np.float64), np.ascontiguousarray(rot, np.float64),
              il.astype(np.float64), ir.astype(np.float64), name=name).validate()
    w.meta.update(dt=0.01, baumgarte=0.01, slop=0.02)
    if meta:
        w.meta.update(meta)
    return w


def random_polygons(n=10_000, density=1.0, static_frac=0.02, config=2) -> World:
    """Config 2: uniform density convex polygons in a square of side sqrt(n/density)."""
    rng = SplitMix64(SEED_BASE + config)
    off, lx, ly = _polygons(rng, n)
    side = math.sqrt(n / density)
    px = rng.uniform(n, 0.0, side)
    py = rng.uniform(n, 0.0, side)
    rot = rng.uniform(n, 0.0, 2.0 * math.pi)
    static = rng.uniform(n) < static_frac
    return _finish(f"polygons{n}", off, lx, ly, px, py, rot, static, {"config": config})


def box_pile(nx=1000, ny=1000, pitch=0.98, config=3) -> World:
    """Config 3: nx*ny unit boxes on a lattice of the given pitch (2 % interpenetration),
    centre jitter +-0.01, rotation jitter +-0.02, plus one static floor under the pile.
    Slot 0 is the floor; box (row, col) is slot 1 + row*nx + col."""
    rng = SplitMix64(SEED_BASE + config)
    n = nx * ny
    cols = np.tile(np.arange(nx, dtype=np.float64), ny)
    rows = np.repeat(np.arange(ny, dtype=np.float64), nx)
    px = cols * pitch + rng.uniform(n, -0.01, 0.01)
    py = rows * pitch + rng.uniform(n, -0.01, 0.01)
    rot = rng.uniform(n, -0.02, 0.02)
    bx = np.tile(np.array([0.5, -0.5, -0.5, 0.5]), n)
    by = np.tile(np.array([0.5, 0.5, -0.5, -0.5]), n)
    fw = nx * pitch + 2.0
    fv = rectangle_vertices(fw, 1.0)
    lx = np.concatenate([[v[0] for v in fv], bx])
    ly = np.concatenate([[v[1] for v in fv], by])
    off = np.arange(0, 4 * (n + 1) + 1, 4, dtype=np.int32)
    px = np.concatenate([[(nx - 1) * pitch / 2.0], px])
    py = np.concatenate([[-0.98], py])
    rot = np.concatenate([[0.0], rot])
    static = np.zeros(n + 1, bool)
    static[0] = True
    return _finish(f"pile{nx}x{ny}", off, lx, ly, px, py, rot, static, {"config": config})


def mixed_polygons(n=4_000_000, density=1.0, static_frac=0.02, config=4) -> World:
    """Config 4: 50 % boxes (sides U[0.4, 0.8]) and 50 % polygons as config 2."""
    rng = SplitMix64(SEED_BASE + config)
    n_box = n // 2
    n_poly = n - n_box
    off_p, lx_p, ly_p = _polygons(rng, n_poly)
    w = rng.uniform(n_box, 0.4, 0.8) / 2.0
    h = rng.uniform(n_box, 0.4, 0.8) / 2.0
    bx = np.stack([w, -w, -w, w], axis=1).reshape(-1)
    by = np.stack([h, h, -h, -h], axis=1).reshape(-1)
    # interleave kinds by a random permutation so rank ranges see the same mix
    kind_is_box = np.zeros(n, bool)
    kind_is_box[:n_box] = True
    perm = np.argsort(rng.u64(n), kind="stable")
    kind_is_box = kind_is_box[perm]
    nv = np.where(kind_is_box, 4, 0).astype(np.int64)
    nv[~kind_is_box] = np.diff(off_p)
    off = np.zeros(n + 1, np.int64)
    np.cumsum(nv, out=off[1:])
    lx = np.empty(off[-1]); ly = np.empty(off[-1])
    box_slots = np.nonzero(kind_is_box)[0]
    poly_slots = np.nonzero(~kind_is_box)[0]
    idx_b = (off[box_slots][:, None] + np.arange(4)[None, :]).reshape(-1)
    lx[idx_b] = bx; ly[idx_b] = by
    idx_p = np.repeat(off[poly_slots] - off_p[:-1], np.diff(off_p)) + np.arange(off_p[-1])
    lx[idx_p] = lx_p; ly[idx_p] = ly_p
    side = math.sqrt(n / density)
    px = rng.uniform(n, 0.0, side)
    py = rng.uniform(n, 0.0, side)
    rot = rng.uniform(n, 0.0, 2.0 * math.pi)
    static = rng.uniform(n) < static_frac
    return _finish(f"mixed{n}", off, lx, ly, px, py, rot, static, {"config": config})


def gaussian_blob(n=1_000_000, density=1.0, peak_factor=4.0, static_frac=0.02, config=5) -> World:
    """Config 5: polygons as config 2, centres ~ N(0, sigma^2 I) with peak density
    n / (2 pi sigma^2) = peak_factor * density (sigma ~ 199.5 at n = 1e6)."""
    rng = SplitMix64(SEED_BASE + config)
    off, lx, ly = _polygons(rng, n)
    sigma = math.sqrt(n / (2.0 * math.pi * peak_factor * density))
    px = sigma * rng.normal(n)
    py = sigma * rng.normal(n)
    rot = rng.uniform(n, 0.0, 2.0 * math.pi)
    static = rng.uniform(n) < static_frac
    return _finish(f"blob{n}", off, lx, ly, px, py, rot, static, {"config": config, "sigma": sigma})


def spatially_sorted(world: World, cell: float = 2.0) -> World:
    """The same objects with slot keys assigned in Morton (Z-curve) order of their positions.

    Keys are the host's choice (World.append hands them out in insertion order, World.hs:77-84); an engine
    that inserts objects region by region gets this layout.  It matters for performance only: partner hulls
    of a pair sit close in memory, and with several GPUs each rank's slot range is a compact region, so
    the multi-rank cell filter keeps only the neighbourhood of that region."""
    gx = np.floor((world.pos_x - np.nanmin(world.pos_x)) / cell).astype(np.uint64)
    gy = np.floor((world.pos_y - np.nanmin(world.pos_y)) / cell).astype(np.uint64)

    def spread(v):
        v = v & np.uint64(0xFFFFFFFF)
        v = (v | (v << np.uint64(16))) & np.uint64(0x0000FFFF0000FFFF)
        v = (v | (v << np.uint64(8))) & np.uint64(0x00FF00FF00FF00FF)
        v = (v | (v << np.uint64(4))) & np.uint64(0x0F0F0F0F0F0F0F0F)
        v = (v | (v << np.uint64(2))) & np.uint64(0x3333333333333333)
        v = (v | (v << np.uint64(1))) & np.uint64(0x5555555555555555)
        return v

    order = np.argsort(spread(gx) | (spread(gy) << np.uint64(1)), kind="stable")
    nv = np.diff(world.vert_offset).astype(np.int64)[order]
    off = np.zeros(world.n_slots + 1, np.int64)
    np.cumsum(nv, out=off[1:])
    src = np.repeat(world.vert_offset[:-1].astype(np.int64)[order] - off[:-1], nv) + np.arange(off[-1])
    out = World(world.alive[order].copy(), off.astype(np.int32), world.local_x[src].copy(), world.local_y[src].copy(),
                world.pos_x[order].copy(), world.pos_y[order].copy(), world.rot[order].copy(),
                world.inv_lin[order].copy(), world.inv_rot[order].copy(),
                radius=None if world.radius is None else world.radius[order].copy(), name=world.name + "_morton")
    out.meta.update(world.meta)
    out.meta["slot_order"] = "morton"
    return out.validate()
