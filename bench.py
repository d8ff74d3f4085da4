#!/usr/bin/env python
"""Benchmark of the collision hot path (broadphase + SAT contacts + constraint generators).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload pile|polygons|mixed|blob|stacks]
  python bench.py --impl reference ...      # the CPU restatement of the reference on host cores

A "step" is one frame: AABB broadphase pair finding -> SAT contact generation -> per-contact
NonPenetration/Friction/Restitution generator evaluation, over one synthetic world.
Metric (BASELINE.json): pairs/s (and ms/frame as ms_per_step).

Headline workload (no --workload given):
 * N = 1 : BASELINE config 3, the dense pile of 1M unit boxes (+ floor) -- the configuration the metric
           is quoted on.  The same line carries `configs`: sub-records for config 1 (Stacks scene, whole
           updateWorld x 10 on the device next to the oracle's), config 2 (10k and 1M random polygons),
           config 4 (4M mixed) and config 5 (1M-polygon Gaussian blob), each with ms/frame, stage times
           and roofline fractions.
 * N > 1 : BASELINE config 4, ONE world of 4M mixed boxes/polygons in the generator's key order, sharded
           over the N ranks ("scaling": "strong"); `one_gpu_same_world` is the same world on rank 0 alone in
           the same run, `weak_config3` the weak-scaled pile (1M boxes per GPU in one world).

 * value   : whole-job broadphase pairs per second, inputs already resident in HBM
             (shapes_frame_device), timed between barrier+synchronize brackets, max over ranks.
 * e2e     : same metric through the public host-buffer call (shapes_frame): pinned host
             inputs are copied H2D and the result columns are copied D2H inside the timed region.
 * roofline: the longest kernel of the frame; algorithmic (compulsory) bytes / CUDA-event duration
             against the measured HBM copy bandwidth in MEASURED_PEAKS.json; `frame` = the whole frame
             by SURVEY.md section 8d's formula.  The SAT kernels are issue/FP64-bound, not HBM-bound:
             their entries say so and carry an FP64-pipe estimate next to the byte figure.
 * cpu_baseline: the C restatement of the reference (oracle/, "port": the Haskell reference
             cannot be built here) on a bounded sample of the same workload, 1 thread (the
             reference is single-threaded ST).
"""
from __future__ import annotations

import argparse
import copy
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "broadphase+SAT+constraint-generator pairs/s at 1M shapes per GPU"
UNIT = "pairs/s"
FP64_NONFUSED_PEAK = 18.2e12      # DMUL/DADD issue peak measured on this pool's B200 (profiles/micro/fp64_peak.cu)


def make_world(workload: str, shapes_per_gpu: int, n_gpus: int):
    from shapes_b200 import scenes
    n = shapes_per_gpu * n_gpus
    if workload == "pile":
        nx = min(1000, max(1, int(round(math.sqrt(shapes_per_gpu)))))
        ny = max(1, n // nx)
        return scenes.box_pile(nx, ny), f"config 3: {nx}x{ny} unit boxes, pitch 0.98, jitter, one static floor"
    if workload == "polygons":
        return scenes.random_polygons(n), f"config 2: {n} random convex polygons (3..8 vertices), density 1"
    if workload == "mixed":
        return scenes.mixed_polygons(n), f"config 4: {n} mixed boxes/polygons, density 1"
    if workload == "blob":
        return scenes.gaussian_blob(n), f"config 5: {n} polygons in a Gaussian blob (peak density 4)"
    if workload == "stacks":
        return scenes.stacks_scene(), "config 1: Stacks.makeScene (30,30) 0 (901 objects)"
    raise SystemExit(f"unknown workload {workload}")


def pair_capacity(workload: str, own_shapes: int) -> int:
    """max_pairs per rank: piles have ~4 pairs per box, uniform polygon worlds ~1, the blob ~1.6."""
    per = {"pile": 6.0, "stacks": 6.0, "blob": 3.0}.get(workload, 2.0)
    return int(own_shapes * per + 65536)


class ClockSampler:
    """nvidia-smi clocks and throttle reasons during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.rows: list[list[str]] = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50",
                 "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) < 7:
                continue
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except ValueError:
                continue
            for name, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def measured_peak_gbs():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured copy bandwidth)"
    except Exception:
        return 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md; MEASURED_PEAKS.json absent)"


# ---------------------------------------------------------------------------------------------
# algorithmic (compulsory) bytes: every datum read once and written once by the stage that needs it
# ---------------------------------------------------------------------------------------------

ROW_BYTES = 225.0   # keys 16 + flip 1 + contact 40 + constraint rows 152 + inverse effective masses 16


def rows_algorithmic_bytes(n_shapes, n_pairs, n_contacts, pairs_with_contacts) -> float:
    """k_rows (flatten + constraint generators): per row 4 B of row map read and 225 B written; per pair with
    contacts its 64 B manifold record and 8 B of keys; every shape's transform (32 B) and inverse masses (16 B)
    ONCE (the re-reads by the ~8 rows that share a shape are L2 hits, not compulsory traffic)."""
    return n_contacts * (ROW_BYTES + 4.0) + pairs_with_contacts * 72.0 + n_shapes * 48.0


def manifolds_algorithmic_bytes(n_shapes, n_pairs, pairs_with_contacts, vbar, sorted_mode) -> float:
    """SAT + clipping: every hull ONCE (world vertices + unit normals 32 B per vertex, 8 B packed extents, 8 B of
    offsets / counts), per pair 8 B of indices and a 4 B contact count (sorted mode: 12 B of work list as well),
    and one 64 B manifold record per pair with contacts."""
    return n_shapes * (32.0 * vbar + 16.0) + n_pairs * (24.0 if sorted_mode else 12.0) + pairs_with_contacts * 64.0


def manifolds_fp64_ops(n_pairs, pairs_with_contacts, vbar) -> float:
    """Non-fused FP64 operations of the SAT stage: both directions project V vertices (3 ops per dot product) on V
    axes plus the two cached extremes; clipping a pair with contacts costs ~160 ops (three line intersections with
    an IEEE division each, projections, the incident-edge choice)."""
    return n_pairs * 2.0 * vbar * (vbar + 2.0) * 3.0 + pairs_with_contacts * 160.0


def frame_algorithmic_bytes(n_shapes, n_pairs, n_contacts, vbar) -> float:
    """N (16 V + 150) + P (48 V + 96) + 217 C."""
    return n_shapes * (16.0 * vbar + 150.0) + n_pairs * (48.0 * vbar + 96.0) + 217.0 * n_contacts


def load_traffic(workload_key: str) -> dict:
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu capture of this workload."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            tj = json.load(f)
        return tj.get("workloads", {}).get(workload_key, {})
    except Exception:
        return {}


def kernel_rooflines(leg: dict, peak: float) -> list[dict]:
    """One entry per contact kernel of the leg, longest first."""
    n, p, c, pc, vbar = leg["shapes"], leg["pairs"], leg["contacts"], leg["pairs_with_contacts"], leg["mean_vertices"]
    st = leg["stage_ms"]
    traffic = load_traffic(leg["workload_key"])
    out = []
    ms = st.get("contact_rows", 0.0)
    nbytes = rows_algorithmic_bytes(n, p, c, pc)
    ach = nbytes / (ms * 1e-3) / 1e9 if ms > 0 else 0.0
    out.append({"bound": "hbm", "kernel": "k_rows", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                "traffic": traffic.get("k_rows"), "algorithmic_bytes_per_launch": nbytes, "avg_launch_ms": ms})
    ms = st.get("manifolds", 0.0)
    sorted_mode = bool(leg.get("sorted_mode"))
    nbytes = manifolds_algorithmic_bytes(n, p, pc, vbar, sorted_mode)
    ops = manifolds_fp64_ops(p, pc, vbar)
    ach = nbytes / (ms * 1e-3) / 1e9 if ms > 0 else 0.0
    out.append({"bound": "issue/fp64", "kernel": leg["sat_kernel"], "achieved": ach, "peak": peak, "unit": "GB/s",
                "frac": ach / peak, "traffic": traffic.get(leg["sat_kernel"]), "algorithmic_bytes_per_launch": nbytes,
                "avg_launch_ms": ms, "fp64_ops_per_launch": ops,
                "fp64_frac_of_nonfused_peak": (ops / (ms * 1e-3) / FP64_NONFUSED_PEAK) if ms > 0 else 0.0,
                "note": "latency / issue bound (ncu: DRAM < 25 %, issue slots 40-50 %): the HBM fraction is reported for "
                        "completeness, the FP64 figure is algorithmic DMUL/DADD ops against the measured 18.2 T op/s"})
    for k in out:
        k["share_of_device_time"] = k["avg_launch_ms"] / max(leg["device_ms_per_step"], 1e-9)
    out.sort(key=lambda k: -k["avg_launch_ms"])
    return out


def frame_roofline(leg: dict, peak: float) -> dict:
    nbytes = frame_algorithmic_bytes(leg["shapes"], leg["pairs"], leg["contacts"], leg["mean_vertices"])
    ms = leg["device_ms_per_step"]
    ach = nbytes / (ms * 1e-3) / 1e9 if ms > 0 else 0.0
    return {"bound": "hbm", "formula": "N(16V+150) + P(48V+96) + 217C (SURVEY.md 8d)", "algorithmic_bytes": nbytes,
            "ms": ms, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak}


# ---------------------------------------------------------------------------------------------
# CPU side: the oracle on a bounded sample
# ---------------------------------------------------------------------------------------------

def cpu_sample(workload: str, budget_frames: int = 3):
    """The oracle (C restatement, 1 thread) on a bounded sample of the same workload."""
    from oracle import binding as orc
    from shapes_b200 import scenes
    if workload == "pile":
        w = scenes.box_pile(1000, 200)
        desc = "pile 1000x200 (200k boxes + floor) of the same lattice"
    elif workload == "stacks":
        w = scenes.stacks_scene()
        desc = "the full 901-object scene"
    else:
        w, _ = make_world(workload, 200_000, 1)
        desc = f"{workload} with 200k shapes"
    orc.build()
    c, s = orc.cos_sin(w.rot)
    ext = orc.hull_extents(w)
    static = orc.is_static(w)
    times, pairs, contacts = [], 0, 0
    for _ in range(budget_frames):
        t0 = time.perf_counter()
        wx, wy, nx, ny = orc.move_shapes(w, c, s)                       # moveShapes
        boxes = orc.aabbs(w, wx, wy)                                    # toAabb
        bp = orc.culled_keys_aabb if w.n_slots <= 3000 else orc.culled_keys_grid
        pi, pj = bp(w, boxes, static)                                   # Aabb / Grid.culledKeys
        r = orc.contacts(w, pi, pj, wx, wy, nx, ny, ext[0], ext[1], 0.01, 0.01, 0.02)  # prepareFrame + constraintGen
        times.append(time.perf_counter() - t0)
        pairs, contacts = len(pi), len(r["key_i"])
    return {"seconds": times, "pairs": pairs, "contacts": contacts, "shapes": w.n_slots,
            "desc": f"{desc}; {budget_frames} frames; broadphase = "
                    f"{'Aabb.culledKeys (n^2)' if w.n_slots <= 3000 else 'Grid.culledKeys restatement, unit cells'}"}


def config1_leg():
    """BASELINE config 1, the reference's only published figure (shapes/bench/Main.hs:19-26: 10 x updateWorld on
    Stacks.makeScene (30,30) 0, "200ms"): shapes_world_step x 10 on the device next to the oracle's whole
    update_world x 10 (1 thread), same scene, same external force, warm start on."""
    from oracle import binding as orc
    from shapes_b200 import engine, scenes
    from shapes_b200.engine import Engine
    from shapes_b200.world import Bodies
    w = scenes.stacks_scene((30, 30), 0.0)
    ext = (1, 0.0, -2.0)                       # Stacks.externals: constantAccel (0, -2)
    res = {"workload": "config 1: Stacks.makeScene (30,30) 0 (901 objects), 10 x updateWorld, dt 0.01",
           "reference_published": "\"200ms\" for the 10 frames in a source comment (shapes/bench/Main.hs:24; author's machine, GHC)"}
    with Engine(w) as eng:
        eng.world_upload(Bodies.at_rest(w.n_slots, 0.2, 0.0))
        for _ in range(3):
            eng.world_step(external=ext)
        runs = []
        for _ in range(5):
            eng.world_upload(Bodies.at_rest(w.n_slots, 0.2, 0.0))
            t0 = time.perf_counter()
            dev = 0.0
            for _ in range(10):
                st = eng.world_step(external=ext)
                dev += st.total_ms
            runs.append(((time.perf_counter() - t0) * 1e3, dev))
        wall, dev = sorted(runs)[len(runs) // 2]
        res["device"] = {"api": "shapes_world_step", "ms_per_10_frames_wall": wall, "ms_per_10_frames_device": dev,
                         "pairs_last_frame": int(st.n_pairs), "contacts_last_frame": int(st.n_contacts)}
    runs = []
    for _ in range(3):
        wo, bo = copy.deepcopy(w), Bodies.at_rest(w.n_slots, 0.2, 0.0)
        c, s = engine.sincos(wo.rot)
        cache = None
        t0 = time.perf_counter()
        for _ in range(10):
            _, cache, c, s = orc.update_world(wo, bo, cache, c, s, external=ext, sincos=engine.sincos, broadphase="grid")
        runs.append((time.perf_counter() - t0) * 1e3)
    res["cpu"] = {"kind": "port", "cores": 1, "ms_per_10_frames": sorted(runs)[1],
                  "note": "oracle update_world: Grid.culledKeys restatement + SAT + generators + sequential solver + advance"}
    return res


# ---------------------------------------------------------------------------------------------
# one workload on the current set of ranks
# ---------------------------------------------------------------------------------------------

class Ranks:
    """torch.distributed plumbing of this process (None-safe for N = 1)."""

    def __init__(self):
        import torch
        self.torch = torch
        self.rank = int(os.environ.get("RANK", "0"))
        self.world_size = int(os.environ.get("WORLD_SIZE", "1"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.dist = None
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        self.host_binding = self.bind_near_gpu()
        if self.world_size > 1:
            import torch.distributed as dist
            dist.init_process_group("nccl", device_id=self.dev)
            self.dist = dist

    def bind_near_gpu(self) -> dict:
        """Multi-rank runs: pin this process to the cores of the GPU's NUMA node BEFORE any pinned buffer is allocated
        (first touch then places the staging buffers next to the GPU's PCIe root; torchrun starts every rank with the
        whole machine as its affinity mask, so without this every rank's buffers end up wherever the allocator ran)."""
        info = {"bound": False}
        if self.world_size <= 1 or os.environ.get("SHAPES_B200_NO_NUMA_BIND"):
            return info
        try:
            p = self.torch.cuda.get_device_properties(self.local_rank)
            bdf = "%04x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
            base = "/sys/bus/pci/devices/" + bdf
            with open(base + "/local_cpulist") as f:
                cpulist = f.read().strip()
            with open(base + "/numa_node") as f:
                info["numa_node"] = int(f.read().strip())
            cpus = set()
            for part in cpulist.split(","):
                if "-" in part:
                    a, b = part.split("-"); cpus.update(range(int(a), int(b) + 1))
                elif part:
                    cpus.add(int(part))
            allowed = os.sched_getaffinity(0)
            cpus &= allowed
            if cpus:
                os.sched_setaffinity(0, cpus)
                info.update(bound=True, pci=bdf, cpus=len(cpus))
        except Exception as e:      # containers without sysfs topology: keep the inherited mask
            info["error"] = repr(e)[:120]
        return info

    def bracket(self, group=True):
        self.torch.cuda.synchronize()
        if self.dist is not None and group:
            self.dist.barrier()
            self.torch.cuda.synchronize()

    def max_over_ranks(self, x: float) -> float:
        if self.dist is None:
            return x
        t = self.torch.tensor([x], device=self.dev, dtype=self.torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(self, xs):
        if self.dist is None:
            return [int(x) for x in xs]
        t = self.torch.tensor(list(xs), device=self.dev, dtype=self.torch.int64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return [int(v) for v in t.tolist()]

    def gather(self, x: float) -> list[float]:
        if self.dist is None:
            return [x]
        out = [None] * self.world_size
        self.dist.all_gather_object(out, x)
        return out


def run_leg(R: Ranks, workload: str, world, desc: str, steps: int, warmup: int, multi: bool, e2e: bool = False,
            warm_leg: bool = False, compact: bool = True) -> dict:
    """Build an engine for `world` on this rank (multi: one ctx per rank of the job; else rank-local single-GPU
    ctx), time `steps` frames of shapes_frame_device, then profile the stages.  Collective when `multi`."""
    torch = R.torch
    from shapes_b200.engine import Engine, nccl_unique_id
    ws = R.world_size if multi else 1
    rank = R.rank if multi else 0
    n = world.n_slots
    vbar = world.n_verts / max(n, 1)
    nccl_id = None
    if ws > 1:
        box = [nccl_unique_id() if R.rank == 0 else None]
        R.dist.broadcast_object_list(box, src=0)
        nccl_id = box[0]
    own = (n + ws - 1) // ws
    max_pairs = pair_capacity(workload, own)
    if ws > 1 and workload == "blob":
        # rows mode: the first frame of a geometry cuts the rows evenly (nothing measured yet), so the ranks that get
        # the blob's centre list several times their steady-state share before the pair-count balancing takes over
        max_pairs *= 4
    eng = Engine(world, max_pairs=max_pairs, max_contacts=2 * max_pairs, device=R.local_rank, rank=rank,
                 world_size=ws, nccl_id=nccl_id)
    exchange = "none"
    if ws > 1:
        exchange = "NCCL all-gather of the AABB records"
        if os.environ.get("SHAPES_B200_NO_P2P") is None:
            blobs = [None] * ws
            R.dist.all_gather_object(blobs, eng.ipc_export())
            eng.ipc_import(blobs)
            if os.environ.get("SHAPES_B200_NO_ROWS") is None:
                exchange = ("rows mode: home ranks push 4 B cell keys + 48 B body records only to the ranks whose grid rows need them; "
                            "the sweep / SAT work is split by grid rows (cuts balanced on the previous frame's per-row pair counts); "
                            "sweeping ranks re-transform the kept hulls, push per-slot pair counts, get the offsets back and store "
                            "every pair (keys, count, manifold) into its final place at the rank that owns its larger key "
                            "(CUDA IPC over NVLink, 5 flag barriers, CUDA-graph replay)")
            else:
                exchange = "4 B cell keys pushed to every peer + needed AABB / body records pulled through peer pointers (CUDA IPC over NVLink), flag barriers"
    cos_rot, sin_rot = np.cos(world.rot), np.sin(world.rot)
    cols = [world.pos_x, world.pos_y, world.rot, cos_rot, sin_rot, world.inv_lin, world.inv_rot]
    d_in = [torch.from_numpy(np.ascontiguousarray(a)).to(R.dev) for a in cols]
    ptrs = [t.data_ptr() for t in d_in]
    beh = dict(dt=world.meta.get("dt", 0.01), baumgarte=world.meta.get("baumgarte", 0.01), slop=world.meta.get("slop", 0.02))

    def step():
        return eng.frame_device(*ptrs, **beh)

    for _ in range(warmup):
        out = step()
    launches0 = eng.launch_count
    dev_ms = 0.0
    R.bracket(multi)
    t0 = time.perf_counter()
    for _ in range(steps):
        out = step()
        dev_ms += out.device_ms
    R.bracket(multi)
    elapsed = time.perf_counter() - t0
    launches = eng.launch_count - launches0
    # per-stage CUDA events right after the timed region, same process and buffers (the frame graph
    # is bypassed for these frames because events inside a captured graph cannot be timed)
    stage_acc: dict[str, float] = {}
    eng.set_profiling(True)
    prof_steps = max(3, min(steps, 10))
    step()
    for _ in range(prof_steps):
        step()
        for k, v in eng.stage_ms().items():
            stage_acc[k] = stage_acc.get(k, 0.0) + v
    eng.set_profiling(False)
    st_ms = {k: v / prof_steps for k, v in stage_acc.items()}
    n_pairs, n_contacts = int(out.n_pairs), int(out.n_contacts)
    pwc = eng.pairs_with_contacts()
    if multi:
        elapsed_all = R.gather(elapsed)
        elapsed = max(elapsed_all)
        tot_pairs, tot_contacts, tot_pwc = R.sum_over_ranks([n_pairs, n_contacts, pwc])
        per_rank_pairs = R.gather(n_pairs)
        per_rank_sat = R.gather(st_ms.get("manifolds", 0.0))
    else:
        tot_pairs, tot_contacts, tot_pwc = n_pairs, n_contacts, pwc
        per_rank_pairs, per_rank_sat = [n_pairs], [st_ms.get("manifolds", 0.0)]
    per_step = elapsed / steps
    leg = {"workload": desc, "workload_key": f"{workload}:{n}", "shapes": n, "pairs": tot_pairs, "contacts": tot_contacts,
           "pairs_with_contacts": tot_pwc, "mean_vertices": vbar, "n_gpus": ws, "steps": steps, "ms_per_step": per_step * 1e3,
           "value": tot_pairs / per_step, "contacts_per_s": tot_contacts / per_step,
           "device_ms_per_step": dev_ms / steps, "stage_ms": st_ms, "gpu_launches": int(launches),
           "grid": [int(out.grid_w), int(out.grid_h), float(out.cell_size)], "big_shapes": int(out.n_big),
           "sorted_mode": eng.sorted_mode(), "sat_kernel": eng.sat_kernel_name(), "exchange": exchange,
           "per_rank_pairs": per_rank_pairs, "per_rank_sat_ms": per_rank_sat}
    if warm_leg:
        # warm-start leg (not part of the headline metric): the same frame with
        # the previous frame's Lagrangian cache (device resident) joined against this frame's keys
        cache = torch.ones(2, max(int(out.n_contacts), 1), device=R.dev, dtype=torch.float64)
        warm_steps = max(3, min(steps, 10))
        for _ in range(2):
            eng.set_lagrangian_cache_device(int(out.n_contacts), cache[0].data_ptr(), cache[1].data_ptr())
            out = step()
        R.bracket(multi)
        tw = time.perf_counter()
        for _ in range(warm_steps):
            eng.set_lagrangian_cache_device(int(out.n_contacts), cache[0].data_ptr(), cache[1].data_ptr())
            out = step()
        R.bracket(multi)
        warm_ms = (time.perf_counter() - tw) / warm_steps * 1e3
        eng.set_profiling(True)
        eng.set_lagrangian_cache_device(int(out.n_contacts), cache[0].data_ptr(), cache[1].data_ptr())
        step()
        warm_join_ms = eng.stage_ms().get("warm_join", 0.0)
        eng.set_profiling(False)
        leg["warm_start"] = {"ms_per_step_with_cache_join": warm_ms, "join_kernel_ms": warm_join_ms, "steps": warm_steps,
                             "note": "descZipVector join of this frame's keys with the previous frame's Lagrangian cache "
                                     "(device resident); includes one 2 x 8 B x contacts D2D cache copy per step"}
    if e2e:
        # ---- end to end through the host-buffer API (H2D + kernels + D2H every step) -------------
        pin = {}
        for name, a in zip(("pos_x", "pos_y", "rot", "cos", "sin", "inv_lin", "inv_rot"), cols):
            pin[name] = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
        w_pinned = copy.copy(world)
        w_pinned.pos_x, w_pinned.pos_y, w_pinned.rot = pin["pos_x"].numpy(), pin["pos_y"].numpy(), pin["rot"].numpy()
        w_pinned.inv_lin, w_pinned.inv_rot = pin["inv_lin"].numpy(), pin["inv_rot"].numpy()
        eng.world = w_pinned
        cs = (pin["cos"].numpy(), pin["sin"].numpy())
        want = ("pairs", "contacts", "constraints")
        e_steps = max(3, min(steps, 10))
        res = {}
        for mode in (("compact", "full") if compact else ("full",)):
            kw = dict(cos_sin=cs, want=want, pinned=True, compact=(mode == "compact"), expand=False, **beh)
            for _ in range(2):
                fr = eng.frame(**kw)
            R.bracket(multi)
            t0 = time.perf_counter()
            for _ in range(e_steps):
                fr = eng.frame(**kw)
            R.bracket(multi)
            e_elapsed = R.max_over_ranks(time.perf_counter() - t0) if multi else time.perf_counter() - t0
            # 7 body columns; with the peer exchange each rank uploads only its own slot range
            n_up = (eng.rank_info()[1] - eng.rank_info()[0]) if not exchange.startswith("NCCL") and ws > 1 else n
            res[mode] = {"value": tot_pairs / (e_elapsed / e_steps), "unit": UNIT, "ms_per_step": e_elapsed / e_steps * 1e3,
                         "steps": e_steps, "h2d_bytes_per_step": int(7 * 8 * n_up), "d2h_bytes_per_step": int(fr.d2h_bytes),
                         "api": "shapes_frame (pinned host buffers, %s result columns fetched)" %
                                ("the independent" if mode == "compact" else "all")}
        leg["e2e"] = res
    leg["_engine"] = eng
    return leg


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path on host cores.  The
    Haskell reference cannot be compiled here (no GHC), so this is the oracle port; it is
    single-threaded like the reference's ST engine."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warm = args.steps, args.warmup
    frames = max(1, min(steps, 6))   # each frame of the 200k-shape sample is ~0.5 s of single-thread CPU work
    workload = args.workload or ("pile" if args.gpus <= 1 else "mixed")
    s = cpu_sample(workload, budget_frames=frames + min(warm, 1))
    secs = s["seconds"][min(warm, 1):]
    per = float(np.mean(secs))
    value = s["pairs"] / per
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": warm, "ms_per_step": per * 1e3, "higher_is_better": True,
        "scaling": "weak" if args.gpus <= 1 else "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": headline_desc(workload, args), "sample": s["desc"], "frames_timed": len(secs),
                   "note": "C restatement of the Haskell reference (GHC absent), 1 thread"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": "port", "sample": s["desc"]},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "contacts_per_s": s["contacts"] / per,
    }
    print(json.dumps(line))


def headline_desc(workload, args):
    if workload == "pile":
        return f"config 3: dense pile of unit boxes, {args.shapes_per_gpu} boxes per GPU + one static floor"
    if workload == "mixed" and args.gpus > 1:
        return f"config 4: one world of {args.strong_shapes} mixed boxes/polygons over {args.gpus} GPUs"
    return f"{workload}, {args.shapes_per_gpu} shapes per GPU"


def strip(leg: dict) -> dict:
    return {k: v for k, v in leg.items() if not k.startswith("_")}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=None, choices=["pile", "polygons", "mixed", "blob", "stacks"],
                    help="default: pile (config 3) on one GPU, mixed (config 4, one 4M world) on several")
    ap.add_argument("--shapes-per-gpu", type=int, default=1_000_000)
    ap.add_argument("--strong-shapes", type=int, default=4_000_000, help="size of the N>1 headline world (config 4)")
    ap.add_argument("--scaling", default=None, choices=["weak", "strong"],
                    help="with --workload on N>1 GPUs: weak = shapes-per-gpu x N shapes, strong = shapes-per-gpu shapes in total")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-world-step", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the sub-records of the other BASELINE configs")
    ap.add_argument("--slot-order", default="generator", choices=["generator", "morton"],
                    help="morton: slot keys assigned along a Z-curve of the positions (scenes.spatially_sorted)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
        return

    # Only the JSON line may reach stdout: libraries (e.g. NCCL's version banner) write to fd 1.
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(obj):
        os.write(real_stdout, (json.dumps(obj) + "\n").encode())

    import torch
    from shapes_b200 import build

    world_size = int(os.environ.get("WORLD_SIZE", "1"))
    if args.gpus != world_size:
        if world_size == 1 and args.gpus > 1:
            raise SystemExit("launch N>1 with: python -m torch.distributed.run --nproc-per-node N bench.py --gpus N ...")
        args.gpus = world_size
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback for the product path")
    R = Ranks()
    if R.rank == 0:
        build.build_library()
    R.bracket()

    default_run = args.workload is None
    workload = args.workload or ("pile" if world_size == 1 else "mixed")
    strong = world_size > 1 and (args.scaling == "strong" or (args.scaling is None and default_run))
    if strong:
        total = args.strong_shapes if default_run else args.shapes_per_gpu
        world, world_desc = make_world(workload, total, 1)
    else:
        world, world_desc = make_world(workload, args.shapes_per_gpu, world_size)
    if args.slot_order == "morton":
        from shapes_b200 import scenes as _scenes
        world = _scenes.spatially_sorted(world)
        world_desc += ", slot keys in Morton order of position"
    else:
        world_desc += ", slot keys in generator order"

    extras = {}
    if strong and R.rank == 0:
        # the same world on ONE GPU, in the same run (rank 0 alone; the other ranks wait at the next barrier)
        one = run_leg(R, workload, world, world_desc, max(10, min(args.steps, 30)), 5, multi=False)
        one.pop("_engine").close()
        extras["one_gpu_same_world"] = {k: one[k] for k in ("ms_per_step", "value", "device_ms_per_step", "stage_ms", "pairs", "contacts")}
    R.bracket()

    sampler = ClockSampler(R.local_rank)
    if R.rank == 0:
        sampler.start()
    head = run_leg(R, workload, world, world_desc, args.steps, args.warmup, multi=True, e2e=not args.no_e2e, warm_leg=True)
    clocks = sampler.stop() if R.rank == 0 else None
    eng = head.pop("_engine")

    # ---- whole updateWorld on the device (not part of the headline metric) ----
    world_step = None
    if world_size == 1 and not args.no_world_step:
        try:
            world_step = world_step_leg(eng, world, args, workload, R.bracket)
        except Exception as e:     # an extra leg must never cost the headline line
            world_step = {"error": f"{type(e).__name__}: {e}"}
    eng.close()

    peak, peak_src = measured_peak_gbs()
    configs = {}
    if default_run and not args.no_configs:
        if world_size == 1:
            subs = [("2_10k", "polygons", 10_000), ("2_1M", "polygons", 1_000_000), ("5", "blob", 1_000_000), ("4", "mixed", 4_000_000)]
            for key, wl, nshapes in subs:
                try:
                    w2, d2 = make_world(wl, nshapes, 1)
                    leg = run_leg(R, wl, w2, d2 + ", slot keys in generator order", max(10, min(args.steps, 30)), 5, multi=False)
                    leg.pop("_engine").close()
                    leg["kernels"] = kernel_rooflines(leg, peak)
                    leg["frame_roofline"] = frame_roofline(leg, peak)
                    configs[key] = strip(leg)
                    del w2
                except Exception as e:
                    configs[key] = {"error": f"{type(e).__name__}: {e}"}
            try:
                configs["1"] = config1_leg()
            except Exception as e:
                configs["1"] = {"error": f"{type(e).__name__}: {e}"}
        else:
            try:   # the weak-scaled pile: 1M boxes per GPU in one world
                w2, d2 = make_world("pile", args.shapes_per_gpu, world_size)
                leg = run_leg(R, "pile", w2, d2, max(10, min(args.steps, 100)), 5, multi=True)
                leg.pop("_engine").close()
                extras["weak_config3"] = strip(leg)
            except Exception as e:
                extras["weak_config3"] = {"error": f"{type(e).__name__}: {e}"}

    if R.rank != 0:
        if R.dist is not None:
            R.dist.destroy_process_group()
        return

    kernels = kernel_rooflines(head, peak)
    roofline = dict(kernels[0])
    roofline["peak_source"] = peak_src
    roofline["traffic_source"] = load_traffic("_source") or None
    roofline["other_kernels"] = kernels[1:]
    roofline["frame"] = frame_roofline(head, peak)

    cpu = None
    if not args.no_cpu_baseline and world_size == 1:
        s = cpu_sample(workload)
        per = float(np.mean(s["seconds"][1:])) if len(s["seconds"]) > 1 else s["seconds"][0]
        cpu = {"value": s["pairs"] / per, "unit": UNIT, "cores": 1, "kind": "port", "sample": s["desc"],
               "ms_per_frame_of_sample": per * 1e3, "host_cores_present": os.cpu_count()}

    e2e_all = head.pop("e2e", None)
    e2e = None
    if e2e_all:
        e2e = dict(e2e_all.get("compact") or e2e_all["full"])
        if "compact" in e2e_all:
            e2e["full_rows"] = e2e_all["full"]
    n = head["shapes"]
    k3_bytes = sum(k["algorithmic_bytes_per_launch"] for k in kernels)
    line = {
        "metric": METRIC, "value": head["value"], "unit": UNIT, "n_gpus": world_size, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": head["ms_per_step"], "higher_is_better": True,
        "scaling": "strong" if strong else "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": world_desc, "shapes": n, "pairs_per_step": head["pairs"], "contacts_per_step": head["contacts"],
                   "mean_vertices": head["mean_vertices"], "cos_sin": "host supplied (numpy)", "grid": head["grid"],
                   "big_shapes": head["big_shapes"],
                   "l2": "per-frame working set (~%.0f MB written + read) exceeds the 126 MB L2; no explicit flush" %
                         ((k3_bytes + n * 300.0) / 1e6),
                   "parallelism": "1 GPU" if world_size == 1 else
                                  f"{world_size} ranks, one world; exchange: {head['exchange']}"},
        "clocks": clocks, "e2e": e2e, "gpu_launches": head["gpu_launches"],
        "gpu_launches_note": "own kernels only; the three CUB scans launch 6 more per step",
        "device_ms_per_step": head["device_ms_per_step"], "stage_ms": head["stage_ms"],
        "stage_ms_note": "per-stage CUDA events over extra frames after the timed region (graph bypassed), rank 0",
        "contacts_per_s": head["contacts_per_s"], "per_rank_pairs": head["per_rank_pairs"],
        "per_rank_sat_ms": head["per_rank_sat_ms"],
        "warm_start": head.get("warm_start"),
        "roofline": roofline, "cpu_baseline": cpu, "world_step": world_step, "configs": configs,
        "host_binding": R.host_binding,
    }
    line.update(extras)
    emit(line)
    if R.dist is not None:
        R.dist.destroy_process_group()


def world_step_leg(eng, world, args, workload, bracket):
    """shapes_world_step: the whole Physics.Engine.Main.updateWorld on the device (body state resident in HBM,
    applyExternal, applyCachedSlns, 2 improveWorld sweeps executed as the sequential walk's dependency
    graph, advance), next to the oracle's sequential solver on a bounded sample."""
    from shapes_b200.world import Bodies
    n = world.n_slots
    rng = np.random.default_rng(11)
    bodies = Bodies(rng.uniform(-0.1, 0.1, n), rng.uniform(-0.1, 0.1, n), rng.uniform(-0.1, 0.1, n), np.full(n, 0.2), np.zeros(n))
    eng.world = world
    eng.world_upload(bodies)
    ext = (1, 0.0, -2.0)                       # Stacks.externals: constantAccel (0, -2)
    for _ in range(3):
        eng.world_step(external=ext)
    steps = max(3, min(args.steps, 10))
    rows = []
    bracket()
    t0 = time.perf_counter()
    for _ in range(steps):
        st = eng.world_step(external=ext)
        rows.append((st.frame_ms, st.chains_ms, st.solve_ms, st.integrate_ms, st.total_ms))
    bracket()
    wall_ms = (time.perf_counter() - t0) / steps * 1e3
    med = np.median(np.asarray(rows), axis=0)
    res = {"api": "shapes_world_step (no per-frame host<->device traffic)", "steps": steps, "ms_per_step": wall_ms,
           "device_ms": {"hot_path": float(med[0]), "chains": float(med[1]), "solver": float(med[2]), "integrate": float(med[3]),
                         "total": float(med[4])},
           "pairs": int(st.n_pairs), "contacts": int(st.n_contacts), "solver_nodes": int(st.solver_nodes),
           "solver_iterations": 2, "warm_start": bool(st.warm),
           "parity": "bit-identical to the sequential oracle (tests/test_gpu_world.py)"}
    if not args.no_cpu_baseline:
        from oracle import binding as orc
        from shapes_b200 import scenes
        ws = scenes.box_pile(1000, 100) if workload == "pile" else make_world(workload, 100_000, 1)[0]
        ns = ws.n_slots
        bs = Bodies(rng.uniform(-0.1, 0.1, ns), rng.uniform(-0.1, 0.1, ns), rng.uniform(-0.1, 0.1, ns), np.full(ns, 0.2), np.zeros(ns))
        c, s = orc.cos_sin(ws.rot)
        cache, ts, nrow = None, [], 0
        for _ in range(3):
            fr = orc.frame(ws, c, s, broadphase="sweep")
            t0 = time.perf_counter()
            orc.apply_external(ws, bs.vel_x, bs.vel_y, 1, 0.0, -2.0, 0.01)
            nrow = len(fr["key_i"])
            lam_np, lam_f, hit = (orc.warm_join(fr, *cache) if cache else (np.zeros(nrow), np.zeros(nrow), np.zeros(nrow, np.uint8)))
            orc.apply_cached(ws, fr, hit, lam_np, lam_f, bs.vel_x, bs.vel_y, bs.rot_vel)
            for _ in range(2):
                orc.improve_world(ws, fr, bs.mu, bs.bounce, bs.vel_x, bs.vel_y, bs.rot_vel, lam_np, lam_f)
            orc.advance(ws, bs.vel_x, bs.vel_y, bs.rot_vel, 0.01)
            c, s = orc.cos_sin(ws.rot)
            ts.append((time.perf_counter() - t0) * 1e3)
            cache = ({q: fr[q] for q in ("key_i", "key_j", "feat_a", "feat_b")}, lam_np, lam_f)
        per = float(np.median(ts))
        res["cpu_solver"] = {"kind": "port", "cores": 1, "sample_shapes": ns, "sample_contacts": nrow, "ms_per_step_of_sample": per,
                             "ns_per_contact": per * 1e6 / max(nrow, 1),
                             "extrapolated_ms_at_bench_contacts": per / max(nrow, 1) * int(st.n_contacts),
                             "note": "oracle: applyExternal + join + applyCachedSlns + 2 improveWorld sweeps + advance + libm "
                                     "cos/sin, sequential like the reference; contact generation excluded"}
    return res


if __name__ == "__main__":
    main()
