"""Host-side mirror of the reference's per-frame contact path over the C ABI.

The three reference entry points this path replaces keep their names:

  culledKeys     Physics.Broadphase.Aabb.culledKeys / Grid.culledKeys
                 (shapes/src/Physics/Broadphase/Aabb.hs:168-183, Grid.hs:74-78)
  prepareFrame   Physics.Solvers.Contact.prepareFrame (Solvers/Contact.hs:40-52)
  constraintGen  Physics.Constraints.Contact.constraintGen (Constraints/Contact.hs:60-72)

One `Engine.frame` call produces all three on the GPU; the functions at the
bottom expose them with the reference's argument meaning.  Nothing here
computes on the CPU: without the CUDA library or a GPU every call raises.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional

import numpy as np

from . import _lib
from ._lib import FrameOut
from .world import World

PAIR_COLS = ("pair_i", "pair_j")
CONTACT_I32 = ("key_i", "key_j", "feat_a", "feat_b")
CONTACT_F64 = ("normal_x", "normal_y", "center_x", "center_y", "depth")
CONSTRAINT_F64 = (tuple(f"j_np{q}" for q in range(6)) + ("b_np", "ra_x", "ra_y", "rb_x", "rb_y", "rn_x", "rn_y")
                  + tuple(f"j_f{q}" for q in range(6)) + ("b_f", "inv_eff_np", "inv_eff_f"))
# Compact wire format of the constraint rows: the columns that carry arithmetic of their own.  The other sixteen
# are sign copies of the contact normal (j_np / j_f entries 0, 1, 3, 4; rn), c - pos (ra, rb) and a constant (b_f);
# `expand_rows` rebuilds them bit for bit (IEEE negation and subtraction are exact), so a host that asks for these
# only moves 113 instead of 225 bytes per row over PCIe.
CONSTRAINT_COMPACT_F64 = ("j_np2", "j_np5", "b_np", "j_f2", "j_f5", "inv_eff_np", "inv_eff_f")
WARM_F64 = ("warm_np", "warm_f")
AABB_COLS = ("aabb_min_x", "aabb_max_x", "aabb_min_y", "aabb_max_y")
WORLD_COLS = ("world_x", "world_y")


class ShapesError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"shapes_b200 error {code}: {msg}")
        self.code = code


class CapacityError(ShapesError):
    """SHAPES_E_CAPACITY: required sizes are in .n_pairs / .n_contacts."""
    def __init__(self, msg: str, n_pairs: int, n_contacts: int):
        super().__init__(_lib.E_CAPACITY, msg)
        self.n_pairs = n_pairs
        self.n_contacts = n_contacts


@dataclass
class ContactBehavior:
    """ContactBehavior (shapes/src/Physics/Contact/Types.hs:20-25)."""
    contactBaumgarte: float = 0.0
    contactPenetrationSlop: float = 0.0


def _ptr(a: Optional[np.ndarray]):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class _HostBuffers:
    """Output columns of one frame; pinned through shapes_host_alloc when asked."""

    def __init__(self, lib, max_pairs, max_contacts, n_slots, n_verts, want, pinned):
        self.lib = lib
        self.cols: dict[str, np.ndarray] = {}
        self._pinned_ptrs: list[int] = []
        self.pinned = pinned
        def alloc(n, dtype):
            n = max(int(n), 1)
            if not pinned:
                return np.zeros(n, dtype)
            nbytes = n * np.dtype(dtype).itemsize
            p = lib.shapes_host_alloc(nbytes)
            if not p:
                raise ShapesError(_lib.E_CUDA, "shapes_host_alloc failed")
            self._pinned_ptrs.append(p)
            buf = (C.c_char * nbytes).from_address(p)
            return np.frombuffer(buf, dtype=dtype, count=n)
        if "pairs" in want:
            for k in PAIR_COLS:
                self.cols[k] = alloc(max_pairs, np.int32)
        if "contacts" in want:
            for k in CONTACT_I32:
                self.cols[k] = alloc(max_contacts, np.int32)
            self.cols["flip"] = alloc(max_contacts, np.uint8)
            for k in CONTACT_F64:
                self.cols[k] = alloc(max_contacts, np.float64)
        if "constraints" in want:
            for k in CONSTRAINT_F64:
                self.cols[k] = alloc(max_contacts, np.float64)
        if "constraints_compact" in want:
            for k in CONSTRAINT_COMPACT_F64:
                self.cols[k] = alloc(max_contacts, np.float64)
        if "warm" in want:
            for k in WARM_F64:
                self.cols[k] = alloc(max_contacts, np.float64)
            self.cols["warm_hit"] = alloc(max_contacts, np.uint8)
        if "aabb" in want:
            for k in AABB_COLS:
                self.cols[k] = alloc(n_slots, np.float64)
        if "world" in want:
            for k in WORLD_COLS:
                self.cols[k] = alloc(n_verts, np.float64)

    def fill(self, out: FrameOut):
        for k, a in self.cols.items():
            if k.startswith("j_np"):
                out.j_np[int(k[4:])] = a.ctypes.data_as(C.POINTER(C.c_double))
            elif k.startswith("j_f"):
                out.j_f[int(k[3:])] = a.ctypes.data_as(C.POINTER(C.c_double))
            else:
                ty = dict(FrameOut._fields_)[k]
                setattr(out, k, a.ctypes.data_as(ty))

    def bytes_for(self, n_pairs, n_contacts, n_slots, n_verts) -> int:
        total = 0
        for k, a in self.cols.items():
            n = (n_pairs if k in PAIR_COLS else n_slots if k in AABB_COLS else n_verts if k in WORLD_COLS
                 else n_contacts)
            total += n * a.dtype.itemsize
        return total

    def free(self):
        self.cols.clear()
        for p in self._pinned_ptrs:
            self.lib.shapes_host_free(p)
        self._pinned_ptrs.clear()


class Frame:
    """Result of one frame: numpy views trimmed to the frame's counts."""

    def __init__(self, out: FrameOut, bufs: _HostBuffers, n_slots: int, n_verts: int):
        self.n_pairs = int(out.n_pairs)
        self.n_contacts = int(out.n_contacts)
        self.n_big = int(out.n_big)
        self.grid = (int(out.grid_w), int(out.grid_h), float(out.cell_size))
        self.device_ms = float(out.device_ms)
        self.total_ms = float(out.total_ms)
        self.cols = {}
        self.d2h_bytes = bufs.bytes_for(self.n_pairs, self.n_contacts, n_slots, n_verts)
        for k, a in bufs.cols.items():
            n = (self.n_pairs if k in PAIR_COLS else n_slots if k in AABB_COLS else n_verts if k in WORLD_COLS
                 else self.n_contacts)
            self.cols[k] = a[:n]

    def __getitem__(self, k):
        return self.cols[k]

    def __contains__(self, k):
        return k in self.cols

    @property
    def keys(self) -> np.ndarray:
        """Descending (i, j) pairs as an (n_pairs, 2) array -- culledKeys' result."""
        return np.stack([self.cols["pair_i"], self.cols["pair_j"]], axis=1)


class Engine:
    """Owns one shapes_ctx (one GPU)."""

    def __init__(self, world: World, max_pairs: Optional[int] = None, max_contacts: Optional[int] = None,
                 device: int = 0, rank: int = 0, world_size: int = 1, nccl_id: Optional[bytes] = None,
                 ext: Optional[tuple[np.ndarray, np.ndarray]] = None):
        self.lib = _lib.load()
        self.device, self.rank, self.world_size, self.nccl_id = device, rank, world_size, nccl_id
        n = world.n_slots
        self.max_pairs = int(max_pairs if max_pairs is not None else max(1024, 8 * n))
        self.max_contacts = int(max_contacts if max_contacts is not None else 2 * self.max_pairs)
        self.ctx = C.c_void_p()
        self._bufs: Optional[_HostBuffers] = None
        self._bufs_key = None
        self._create(world.n_slots, world.n_verts)
        self.world = None
        self.set_hulls(world, ext)

    # -- lifetime -------------------------------------------------------
    def _check(self, rc: int, out: Optional[FrameOut] = None):
        if rc == _lib.OK:
            return
        msg = self.lib.shapes_last_error(self.ctx if self.ctx else None)
        msg = msg.decode() if msg else ""
        if rc == _lib.E_CAPACITY and out is not None:
            raise CapacityError(msg, int(out.n_pairs), int(out.n_contacts))
        raise ShapesError(rc, msg)

    def _create(self, n_slots, n_verts):
        self.cap_slots, self.cap_verts = max(n_slots, 1), max(n_verts, 1)
        if self.world_size > 1:
            idbuf = C.create_string_buffer(self.nccl_id, _lib.NCCL_ID_BYTES)
            rc = self.lib.shapes_create_ranked(C.byref(self.ctx), self.device, self.rank, self.world_size, idbuf,
                                               self.cap_slots, self.cap_verts, self.max_pairs, self.max_contacts)
        else:
            rc = self.lib.shapes_create(C.byref(self.ctx), self.device, self.cap_slots, self.cap_verts,
                                        self.max_pairs, self.max_contacts)
        if rc != _lib.OK:
            msg = self.lib.shapes_last_error(None)
            self.ctx = C.c_void_p()
            raise ShapesError(rc, msg.decode() if msg else "")

    def close(self):
        if self._bufs is not None:
            self._bufs.free()
            self._bufs = None
        if self.ctx:
            self.lib.shapes_destroy(self.ctx)
            self.ctx = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- static geometry -----------------------------------------------------
    def set_hulls(self, world: World, ext=None):
        """World.fromList / append / delete (World.hs:77-116): register the hull geometry."""
        world.validate()
        if world.n_slots > self.cap_slots or world.n_verts > self.cap_verts:
            raise ShapesError(_lib.E_ARG, "world exceeds the ctx capacities")
        emin = emax = None
        if ext is not None:
            emin = np.ascontiguousarray(ext[0], np.int32); emax = np.ascontiguousarray(ext[1], np.int32)
        radius = getattr(world, "radius", None)
        self._check(self.lib.shapes_set_shapes(self.ctx, world.n_slots, _ptr(world.alive), _ptr(world.vert_offset),
                                               _ptr(world.local_x), _ptr(world.local_y), _ptr(emin), _ptr(emax),
                                               _ptr(radius)))
        self.world = world

    def set_lagrangian_cache(self, lambda_np: np.ndarray, lambda_f: np.ndarray):
        """The (key, ContactLagrangian) cache of the LAST frame (row k <-> that frame's contact k):
        the next frame(want=(..., "warm")) returns the descZipVector join against it."""
        a = np.ascontiguousarray(lambda_np, np.float64); b = np.ascontiguousarray(lambda_f, np.float64)
        assert a.shape == b.shape
        self._check(self.lib.shapes_set_lagrangian_cache(self.ctx, a.shape[0], _ptr(a), _ptr(b)))

    def set_lagrangian_cache_device(self, n_prev: int, lambda_np_ptr: int, lambda_f_ptr: int):
        """Same, with DEVICE addresses (a solver that keeps its cache in HBM)."""
        self._check(self.lib.shapes_set_lagrangian_cache_device(self.ctx, int(n_prev), lambda_np_ptr, lambda_f_ptr))

    def set_cell_size(self, cell: float):
        self._check(self.lib.shapes_set_cell_size(self.ctx, float(cell)))

    # -- frames -----------------------------------------------------------------
    def _buffers(self, want, pinned) -> _HostBuffers:
        key = (tuple(sorted(want)), pinned, self.max_pairs, self.max_contacts, self.world.n_slots, self.world.n_verts)
        if self._bufs is None or self._bufs_key != key:
            if self._bufs is not None:
                self._bufs.free()
            self._bufs = _HostBuffers(self.lib, self.max_pairs, self.max_contacts, self.world.n_slots,
                                      self.world.n_verts, want, pinned)
            self._bufs_key = key
        return self._bufs

    def frame(self, dt: float = 0.01, baumgarte: float = 0.01, slop: float = 0.02,
              cos_sin: Optional[tuple[np.ndarray, np.ndarray]] = None,
              want=("pairs", "contacts", "constraints"), pinned: bool = False,
              compact: bool = False, expand: bool = True) -> Frame:
        """One frame through shapes_frame with host buffers (H2D + kernels + D2H).

        compact: fetch the constraint rows in the compact wire format (CONSTRAINT_COMPACT_F64: NULL pointers for
        the sixteen derived columns, so the library copies 113 instead of 225 bytes per row); with `expand` the
        derived columns are rebuilt on the host (`expand_rows`), bit-identical to the full fetch."""
        w = self.world
        if compact and "constraints" in want:
            want = tuple(k for k in want if k != "constraints") + ("contacts", "constraints_compact")
            want = tuple(dict.fromkeys(want))
        else:
            compact = False
        bufs = self._buffers(want, pinned)
        out = FrameOut()
        bufs.fill(out)
        cos_rot = sin_rot = None
        if cos_sin is not None:
            cos_rot = np.ascontiguousarray(cos_sin[0], np.float64)
            sin_rot = np.ascontiguousarray(cos_sin[1], np.float64)
        rc = self.lib.shapes_frame(self.ctx, w.n_slots, _ptr(w.pos_x), _ptr(w.pos_y), _ptr(w.rot),
                                   _ptr(cos_rot), _ptr(sin_rot), _ptr(w.inv_lin), _ptr(w.inv_rot),
                                   dt, baumgarte, slop, C.byref(out))
        self._check(rc, out)
        fr = Frame(out, bufs, w.n_slots, w.n_verts)
        if compact and expand:
            expand_rows(fr.cols, w.pos_x, w.pos_y)
        return fr

    def frame_device(self, pos_x: int, pos_y: int, rot: int, cos_rot: int, sin_rot: int, inv_lin: int,
                     inv_rot: int, dt: float = 0.01, baumgarte: float = 0.01, slop: float = 0.02) -> FrameOut:
        """One frame through shapes_frame_device: arguments are DEVICE addresses (e.g.
        torch.Tensor.data_ptr()); results stay in HBM (see device_view)."""
        out = FrameOut()
        rc = self.lib.shapes_frame_device(self.ctx, self.world.n_slots, pos_x, pos_y, rot or None, cos_rot or None,
                                          sin_rot or None, inv_lin, inv_rot, dt, baumgarte, slop, C.byref(out))
        self._check(rc, out)
        return out

    def fetch(self, want=("pairs", "contacts", "constraints"), pinned: bool = False) -> Frame:
        """Copy the last frame's results out of HBM (shapes_fetch)."""
        bufs = self._buffers(want, pinned)
        out = FrameOut()
        bufs.fill(out)
        self._check(self.lib.shapes_fetch(self.ctx, C.byref(out)))
        view = self.device_view()
        out.n_pairs, out.n_contacts = view.n_pairs, view.n_contacts
        return Frame(out, bufs, self.world.n_slots, self.world.n_verts)

    def device_view(self) -> _lib.DeviceView:
        v = _lib.DeviceView()
        self._check(self.lib.shapes_device_view_get(self.ctx, C.byref(v)))
        return v

    def frame_info(self) -> _lib.FrameInfo:
        """Statistics of the last completed frame (shapes_last_frame_info)."""
        info = _lib.FrameInfo()
        self._check(self.lib.shapes_last_frame_info(self.ctx, C.byref(info)))
        return info

    def pairs_with_contacts(self) -> int:
        return int(self.frame_info().pairs_with_contacts)

    def sorted_mode(self) -> bool:
        return bool(self.frame_info().sorted_mode)

    def sat_kernel_name(self) -> str:
        return _lib.SAT_KERNEL_NAMES.get(int(self.frame_info().sat_kernel), "?")

    def ipc_export(self) -> bytes:
        """CUDA IPC handles of this rank's exchange buffers (shapes_ipc_export)."""
        buf = C.create_string_buffer(_lib.IPC_BYTES)
        self._check(self.lib.shapes_ipc_export(self.ctx, buf))
        return buf.raw

    def ipc_import(self, blobs: list[bytes]):
        """Blobs of ALL ranks in rank order: switches the AABB exchange to peer-to-peer stores."""
        assert len(blobs) == self.world_size
        buf = C.create_string_buffer(b"".join(blobs), _lib.IPC_BYTES * self.world_size)
        self._check(self.lib.shapes_ipc_import(self.ctx, buf))

    def rank_info(self):
        lo, hi = C.c_int64(), C.c_int64()
        pairs = (C.c_int64 * self.world_size)()
        contacts = (C.c_int64 * self.world_size)()
        self._check(self.lib.shapes_rank_info(self.ctx, C.byref(lo), C.byref(hi), pairs, contacts))
        return lo.value, hi.value, list(pairs), list(contacts)

    def rank_segments(self, rank: Optional[int] = None):
        """shapes_rank_segments: ((lo0, hi0, pairs0, contacts0), (lo1, hi1, pairs1, contacts1)) of a rank's two runs."""
        a = [(C.c_int64 * 2)() for _ in range(4)]
        self._check(self.lib.shapes_rank_segments(self.ctx, self.rank if rank is None else rank, *a))
        return tuple((int(a[0][q]), int(a[1][q]), int(a[2][q]), int(a[3][q])) for q in range(2))

    def grow(self, n_pairs: int, n_contacts: int):
        """The caller's answer to E_CAPACITY.  Single GPU: shapes_grow, in place -- the previous frame's key columns,
        the Lagrangian cache (EngineCache) and an uploaded world survive, so the retried frame warm-starts exactly as
        the failed attempt would have.  Multi-rank: the ctx is re-created (a collective decision of all ranks; the
        next frame starts cold)."""
        self.max_pairs = max(self.max_pairs, int(n_pairs * 1.25) + 1024)
        self.max_contacts = max(self.max_contacts, int(n_contacts * 1.25) + 1024)
        if self.world_size == 1:
            self._check(self.lib.shapes_grow(self.ctx, self.max_pairs, self.max_contacts))
            return          # host buffers are keyed on the capacities and follow on the next frame
        world = self.world
        max_pairs, max_contacts = self.max_pairs, self.max_contacts
        self.close()
        self.max_pairs, self.max_contacts = max_pairs, max_contacts
        self._create(world.n_slots, world.n_verts)
        self.set_hulls(world)

    def frame_grow(self, **kw) -> Frame:
        """frame(), growing capacities and retrying on SHAPES_E_CAPACITY."""
        for _ in range(4):
            try:
                return self.frame(**kw)
            except CapacityError as e:
                self.grow(e.n_pairs, e.n_contacts)
        return self.frame(**kw)

    # -- device-resident world ---------------------------------------------
    def world_upload(self, bodies, cos_sin: Optional[tuple[np.ndarray, np.ndarray]] = None):
        """shapes_world_upload: the world's PhysicalObj / Material columns move to HBM.
        `bodies` is a world.Bodies; positions and inverse masses come from self.world."""
        w = self.world
        cols = [np.ascontiguousarray(a, np.float64) for a in
                (bodies.vel_x, bodies.vel_y, bodies.rot_vel, w.pos_x, w.pos_y, w.rot)]
        cs = [None, None]
        if cos_sin is not None:
            cs = [np.ascontiguousarray(cos_sin[0], np.float64), np.ascontiguousarray(cos_sin[1], np.float64)]
        rest = [np.ascontiguousarray(a, np.float64) for a in (w.inv_lin, w.inv_rot, bodies.mu, bodies.bounce)]
        self._check(self.lib.shapes_world_upload(self.ctx, w.n_slots, *[_ptr(a) for a in cols], _ptr(cs[0]), _ptr(cs[1]),
                                                 *[_ptr(a) for a in rest]))

    def world_download(self) -> dict[str, np.ndarray]:
        """shapes_world_download: vel_x, vel_y, rot_vel, pos_x, pos_y, rot, cos_rot, sin_rot."""
        n = self.world.n_slots
        names = ("vel_x", "vel_y", "rot_vel", "pos_x", "pos_y", "rot", "cos_rot", "sin_rot")
        out = {k: np.empty(n) for k in names}
        self._check(self.lib.shapes_world_download(self.ctx, n, *[_ptr(out[k]) for k in names]))
        return out

    def world_step(self, dt: float = 0.01, baumgarte: float = 0.01, slop: float = 0.02,
                   external=(_lib.EXT_NONE, 0.0, 0.0), iterations: int = 2, warm_start: bool = True) -> _lib.StepStats:
        """shapes_world_step: one Physics.Engine.Main.updateWorld on the device."""
        cfg = _lib.StepConfig(dt, baumgarte, slop, int(external[0]), int(iterations), float(external[1]),
                              float(external[2]), 1 if warm_start else 0, 0)
        stats = _lib.StepStats()
        rc = self.lib.shapes_world_step(self.ctx, C.byref(cfg), C.byref(stats))
        if rc == _lib.E_CAPACITY:
            msg = self.lib.shapes_last_error(self.ctx)
            raise CapacityError(msg.decode() if msg else "", int(stats.n_pairs), int(stats.n_contacts))
        self._check(rc)
        return stats

    def set_profiling(self, on: bool = True):
        self._check(self.lib.shapes_set_profiling(self.ctx, 1 if on else 0))

    def stage_ms(self) -> dict[str, float]:
        """Device milliseconds of each stage of the last frame (profiling must be on)."""
        buf = (C.c_float * _lib.N_STAGES)()
        self._check(self.lib.shapes_stage_ms(self.ctx, buf))
        return {self.lib.shapes_stage_name(k).decode(): float(buf[k]) for k in range(_lib.N_STAGES)}

    @property
    def stream(self) -> int:
        return int(self.lib.shapes_stream(self.ctx) or 0)

    @property
    def launch_count(self) -> int:
        return int(self.lib.shapes_launch_count(self.ctx))


class MultiEngine:
    """Owns one shapes_multi: several GPUs of one box driven from THIS process (shapes_create_multi) -- the form a
    single-threaded host like the reference's ST engine can use.  frame() returns the whole frame, in the reference's
    global descending order, exactly like Engine.frame on one GPU."""

    def __init__(self, world: World, n_gpus: int, device_ids: Optional[list[int]] = None,
                 max_pairs_per_gpu: Optional[int] = None, max_contacts_per_gpu: Optional[int] = None,
                 ext: Optional[tuple[np.ndarray, np.ndarray]] = None):
        self.lib = _lib.load()
        self.n_gpus = int(n_gpus)
        n = world.n_slots
        own = (n + self.n_gpus - 1) // self.n_gpus
        self.max_pairs = int(max_pairs_per_gpu if max_pairs_per_gpu is not None else max(4096, 8 * own))
        self.max_contacts = int(max_contacts_per_gpu if max_contacts_per_gpu is not None else 2 * self.max_pairs)
        self.ctx = C.c_void_p()
        ids = None
        if device_ids is not None:
            ids = (C.c_int * self.n_gpus)(*device_ids)
        rc = self.lib.shapes_create_multi(C.byref(self.ctx), self.n_gpus, ids, max(n, 1), max(world.n_verts, 1),
                                          self.max_pairs, self.max_contacts)
        if rc != _lib.OK:
            msg = self.lib.shapes_multi_last_error(None)
            self.ctx = C.c_void_p()
            raise ShapesError(rc, msg.decode() if msg else "")
        self._bufs: Optional[_HostBuffers] = None
        self._bufs_key = None
        self.world = None
        self.set_hulls(world, ext)

    def _check(self, rc: int, out: Optional[FrameOut] = None):
        if rc == _lib.OK:
            return
        msg = self.lib.shapes_multi_last_error(self.ctx if self.ctx else None)
        msg = msg.decode() if msg else ""
        if rc == _lib.E_CAPACITY and out is not None:
            raise CapacityError(msg, int(out.n_pairs), int(out.n_contacts))
        raise ShapesError(rc, msg)

    def close(self):
        if self._bufs is not None:
            self._bufs.free()
            self._bufs = None
        if self.ctx:
            self.lib.shapes_multi_destroy(self.ctx)
            self.ctx = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def set_hulls(self, world: World, ext=None):
        world.validate()
        emin = emax = None
        if ext is not None:
            emin = np.ascontiguousarray(ext[0], np.int32); emax = np.ascontiguousarray(ext[1], np.int32)
        radius = getattr(world, "radius", None)
        self._check(self.lib.shapes_multi_set_shapes(self.ctx, world.n_slots, _ptr(world.alive), _ptr(world.vert_offset),
                                                     _ptr(world.local_x), _ptr(world.local_y), _ptr(emin), _ptr(emax),
                                                     _ptr(radius)))
        self.world = world

    def frame(self, dt: float = 0.01, baumgarte: float = 0.01, slop: float = 0.02,
              cos_sin: Optional[tuple[np.ndarray, np.ndarray]] = None,
              want=("pairs", "contacts", "constraints"), pinned: bool = False,
              compact: bool = False, expand: bool = True) -> Frame:
        """shapes_multi_frame: one frame on all GPUs, the whole result in host buffers."""
        w = self.world
        if compact and "constraints" in want:
            want = tuple(dict.fromkeys(tuple(k for k in want if k != "constraints") + ("contacts", "constraints_compact")))
        else:
            compact = False
        key = (tuple(sorted(want)), pinned, w.n_slots, w.n_verts)
        if self._bufs is None or self._bufs_key != key:
            if self._bufs is not None:
                self._bufs.free()
            self._bufs = _HostBuffers(self.lib, self.max_pairs * self.n_gpus, self.max_contacts * self.n_gpus,
                                      w.n_slots, w.n_verts, want, pinned)
            self._bufs_key = key
        out = FrameOut()
        self._bufs.fill(out)
        cos_rot = sin_rot = None
        if cos_sin is not None:
            cos_rot = np.ascontiguousarray(cos_sin[0], np.float64)
            sin_rot = np.ascontiguousarray(cos_sin[1], np.float64)
        rc = self.lib.shapes_multi_frame(self.ctx, w.n_slots, _ptr(w.pos_x), _ptr(w.pos_y), _ptr(w.rot),
                                         _ptr(cos_rot), _ptr(sin_rot), _ptr(w.inv_lin), _ptr(w.inv_rot),
                                         dt, baumgarte, slop, C.byref(out))
        self._check(rc, out)
        fr = Frame(out, self._bufs, w.n_slots, w.n_verts)
        if compact and expand:
            expand_rows(fr.cols, w.pos_x, w.pos_y)
        return fr

    def rank_pairs(self) -> list[int]:
        """Pairs each GPU holds as a home after the last frame."""
        n = self.n_gpus
        pairs = (C.c_int64 * n)(); contacts = (C.c_int64 * n)()
        lo, hi = C.c_int64(), C.c_int64()
        ctx0 = self.lib.shapes_multi_rank(self.ctx, 0)
        rc = self.lib.shapes_rank_info(ctx0, C.byref(lo), C.byref(hi), pairs, contacts)
        if rc != _lib.OK:
            raise ShapesError(rc, "shapes_rank_info")
        return list(pairs)


def nccl_unique_id() -> bytes:
    lib = _lib.load()
    buf = C.create_string_buffer(_lib.NCCL_ID_BYTES)
    rc = lib.shapes_nccl_unique_id(buf)
    if rc != _lib.OK:
        raise ShapesError(rc, (lib.shapes_last_error(None) or b"").decode())
    return buf.raw


# ---------------------------------------------------------------------------
# the reference's entry points, by name
# ---------------------------------------------------------------------------

def sincos(rot: np.ndarray) -> tuple[np.ndarray, np.ndarray]:
    """shapes_sincos on the host: the cos/sin the device-resident world uses (same bits as the GPU)."""
    r = np.ascontiguousarray(rot, np.float64)
    c = np.empty_like(r); s = np.empty_like(r)
    _lib.load().shapes_sincos(r.shape[0], _ptr(r), _ptr(c), _ptr(s))
    return c, s


def expand_rows(cols: dict, pos_x: np.ndarray, pos_y: np.ndarray) -> dict:
    """Rebuild the sixteen derived constraint columns from the compact wire format, in place and bit for bit.

    With n the contact normal, t = clockwiseV2 n = (n.y, -n.x), c the contact centre and flip the Flipping tag:
      j_np = (-n, x, n, x) and j_f = (-t, x, t, x) on (penetrated, penetrator), halves swapped back for Flip
             (NonPenetration.hs:34-43, Friction.hs:31-44, Utils.hs:175-177,212-215) -- so entries 0, 1, 3, 4 are
             +-n / +-t by the flip bit; the cross terms (entries 2, 5) are shipped;
      ra = c - pos_i, rb = c - pos_j, rn = n for Same / -n for Flip (Restitution.hs:21-31);  b_f = 0 (Friction.hs:26-29).
    Negation and one IEEE subtraction reproduce the device's bits exactly."""
    nx, ny = cols["normal_x"], cols["normal_y"]
    f = cols["flip"].astype(bool)
    sx = np.where(f, nx, -nx); sy = np.where(f, ny, -ny)          # -n for Same, n for Flip
    cols["j_np0"], cols["j_np1"], cols["j_np3"], cols["j_np4"] = sx, sy, -sx, -sy
    cols["j_f0"], cols["j_f1"], cols["j_f3"], cols["j_f4"] = sy, -sx, -sy, sx
    cols["rn_x"], cols["rn_y"] = -sx, -sy
    ki, kj = cols["key_i"], cols["key_j"]
    cols["ra_x"] = cols["center_x"] - pos_x[ki]; cols["ra_y"] = cols["center_y"] - pos_y[ki]
    cols["rb_x"] = cols["center_x"] - pos_x[kj]; cols["rb_y"] = cols["center_y"] - pos_y[kj]
    cols["b_f"] = np.zeros_like(nx)
    return cols


def updateWorld(engine: Engine, dt: float, beh: "ContactBehavior", external=(_lib.EXT_NONE, 0.0, 0.0)) -> _lib.StepStats:
    """Physics.Engine.Main.updateWorld (Engine/Main.hs:71-86) on a world uploaded with world_upload."""
    return engine.world_step(dt=dt, baumgarte=beh.contactBaumgarte, slop=beh.contactPenetrationSlop, external=external,
                             iterations=2)


def culledKeys(engine: Engine, cos_sin=None) -> np.ndarray:
    """Aabb.culledKeys world :: Descending (Int, Int) -- (n_pairs, 2), descending."""
    return engine.frame_grow(cos_sin=cos_sin, want=("pairs",)).keys.copy()


def prepareFrame(engine: Engine, cos_sin=None) -> Frame:
    """prepareFrame keys world :: Descending (ObjectFeatureKey Int, Flipping Contact).
    The broadphase keys are produced inside the same device pass."""
    return engine.frame_grow(cos_sin=cos_sin, want=("pairs", "contacts"))


def constraintGen(engine: Engine, beh: ContactBehavior, dt: float, cos_sin=None) -> Frame:
    """constraintGen beh dt fContact ab for every contact of the frame: row k of the
    constraint columns belongs to contact k, the order applyCachedSlns walks them in."""
    return engine.frame_grow(dt=dt, baumgarte=beh.contactBaumgarte, slop=beh.contactPenetrationSlop,
                             cos_sin=cos_sin, want=("pairs", "contacts", "constraints"))
