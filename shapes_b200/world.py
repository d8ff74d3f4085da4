"""Host-side world storage in the flattened form the C ABI takes.

Mirrors the reference's World (shapes/src/Physics/World.hs:46-84): physical
objects are already structure-of-arrays there (`U.MVector PhysicalObj` unboxes
to one Double column per field, Constraint.hs:52-63); hull geometry, a boxed
vector of boxed arrays in the reference, is flattened once into CSR columns.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Iterable, Sequence

import numpy as np


def rectangle_vertices(w: float, h: float) -> list[tuple[float, float]]:
    """rectangleVertices (shapes/src/Physics/Contact/ConvexHull.hs:135-145): CCW from (+,+)."""
    w2, h2 = w / 2.0, h / 2.0
    return [(w2, h2), (-w2, h2), (-w2, -h2), (w2, -h2)]


def to_inv_mass2(mass: tuple[float, float]) -> tuple[float, float]:
    """toInvMass2 (shapes/src/Physics/Constraint.hs:79-83): mass 0 means infinite mass."""
    ml, mr = mass
    return (0.0 if ml == 0.0 else 1.0 / ml, 0.0 if mr == 0.0 else 1.0 / mr)


@dataclass
class Bodies:
    """Velocity and material columns of the world's objects: the part of PhysicalObj
    (shapes/src/Physics/Constraint.hs:52-63) and Material (World.hs:36-40) that only the solver and
    the integrator touch.  Used by the device-resident world step (Engine.world_upload / updateWorld)."""
    vel_x: np.ndarray          # f64 [n]  _physObjVel
    vel_y: np.ndarray
    rot_vel: np.ndarray        # f64 [n]  _physObjRotVel
    mu: np.ndarray             # f64 [n]  _mMu
    bounce: np.ndarray         # f64 [n]  _mBounce

    @staticmethod
    def at_rest(n: int, mu: float = 0.2, bounce: float = 0.2) -> "Bodies":
        return Bodies(np.zeros(n), np.zeros(n), np.zeros(n), np.full(n, float(mu)), np.full(n, float(bounce)))

    def copy(self) -> "Bodies":
        return Bodies(*(a.copy() for a in (self.vel_x, self.vel_y, self.rot_vel, self.mu, self.bounce)))


@dataclass
class World:
    """n_slots slots; slot s owns vertices [vert_offset[s], vert_offset[s+1])."""
    alive: np.ndarray          # uint8  [n]      EmptiesVector filled flags
    vert_offset: np.ndarray    # int32  [n+1]
    local_x: np.ndarray        # f64    [n_verts] _hullLocalVertices, CCW
    local_y: np.ndarray
    pos_x: np.ndarray          # f64    [n]      _physObjPos
    pos_y: np.ndarray
    rot: np.ndarray            # f64    [n]      _physObjRotPos
    inv_lin: np.ndarray        # f64    [n]      _physObjInvMass
    inv_rot: np.ndarray
    radius: "np.ndarray | None" = None   # f64 [n]: >= 0 marks a CircleShape of that radius (no vertices); < 0 / None = hull
    name: str = "world"
    meta: dict = field(default_factory=dict)

    @property
    def n_slots(self) -> int:
        return int(self.alive.shape[0])

    @property
    def n_verts(self) -> int:
        return int(self.vert_offset[-1]) if self.vert_offset.size else 0

    def validate(self) -> "World":
        n = self.n_slots
        assert self.vert_offset.shape == (n + 1,) and self.vert_offset.dtype == np.int32
        assert self.alive.dtype == np.uint8
        for a in (self.pos_x, self.pos_y, self.rot, self.inv_lin, self.inv_rot):
            assert a.shape == (n,) and a.dtype == np.float64
        for a in (self.local_x, self.local_y):
            assert a.shape == (self.n_verts,) and a.dtype == np.float64
        if self.radius is not None:
            assert self.radius.shape == (n,) and self.radius.dtype == np.float64
            assert np.all(np.diff(self.vert_offset)[self.radius >= 0] == 0), "a circle slot owns no vertices"
        return self

    def delete(self, slots: Iterable[int]) -> "World":
        """World.delete (World.hs:86-87): marks the slots empty; keys stay sparse."""
        for s in slots:
            self.alive[s] = 0
        return self

    @staticmethod
    def from_objects(objs: Sequence[tuple[Sequence[tuple[float, float]], tuple[float, float], float,
                                          tuple[float, float]]], name: str = "world") -> "World":
        """objs: (local CCW vertices | circle radius, position, rotation, (linear mass, rotational mass)),
        appended in order like World.fromList (World.hs:111-116)."""
        n = len(objs)
        off = np.zeros(n + 1, np.int32)
        lx: list[float] = []
        ly: list[float] = []
        px = np.zeros(n); py = np.zeros(n); rot = np.zeros(n)
        il = np.zeros(n); ir = np.zeros(n)
        radius = np.full(n, -1.0)
        for s, (verts, pos, r, mass) in enumerate(objs):
            if isinstance(verts, (int, float)):      # makeCircle radius (Engine.hs:53-54)
                radius[s] = float(verts)
                verts = ()
            for (x, y) in verts:
                lx.append(float(x)); ly.append(float(y))
            off[s + 1] = len(lx)
            px[s], py[s], rot[s] = pos[0], pos[1], r
            il[s], ir[s] = to_inv_mass2(mass)
        return World(np.ones(n, np.uint8), off, np.asarray(lx, np.float64), np.asarray(ly, np.float64),
                     px, py, rot, il, ir, radius=radius if (radius >= 0).any() else None, name=name).validate()
