#!/usr/bin/env python
"""Benchmark of the collision hot path (broadphase + SAT contacts + constraint generators).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload pile|polygons|mixed|blob|stacks]
  python bench.py --impl reference ...      # the CPU restatement of the reference on host cores

A "step" is one frame: AABB broadphase pair finding -> SAT contact generation -> per-contact
NonPenetration/Friction/Restitution generator evaluation, over one synthetic world.
Metric (BASELINE.json): pairs/s (and ms/frame as ms_per_step) at 1M shapes per GPU.

 * value   : whole-job broadphase pairs per second, inputs already resident in HBM
             (shapes_frame_device), timed between barrier+synchronize brackets, max over ranks.
 * e2e     : same metric through the public host-buffer call (shapes_frame): pinned host
             inputs are copied H2D and every result column is copied D2H inside the timed region.
 * roofline: the dominant kernel (k_contacts), algorithmic bytes / CUDA-event duration against
             the measured HBM copy bandwidth in MEASURED_PEAKS.json.
 * cpu_baseline: the C restatement of the reference (oracle/, "port": the Haskell reference
             cannot be built here) on a bounded sample of the same workload, 1 thread (the
             reference is single-threaded ST).

N > 1 (torchrun, one rank per GPU): weak scaling, 1M shapes per GPU in one world; every rank
registers the whole world, owns the pairs whose larger key falls in its slot range, all-gathers
the AABB records with NCCL and keeps its slice of the (globally ordered) results in its own HBM.
"""
from __future__ import annotations

import argparse
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


ROW_BYTES = 225.0   # keys 16 + flip 1 + contact 40 + constraint rows 152 + inverse effective masses 16


def manifolds_algorithmic_bytes(n_pairs: int, n_contacts: int, vbar: float) -> float:
    """k_manifolds (SAT + clipping), compulsory traffic per launch: per pair 8 B (i, j) and per hull
    8 B CSR offsets + 16*V local vertices + 8 B packed extents + 32 B transform; writes 4 B count per
    pair and one 64 B manifold record per pair that has contacts (~contacts/2)."""
    return n_pairs * (8.0 + 2.0 * (8.0 + 16.0 * vbar + 8.0 + 32.0) + 4.0) + 0.5 * n_contacts * 64.0


def rows_algorithmic_bytes(n_pairs: int, n_contacts: int) -> float:
    """k_rows (flatten + constraint generators): per pair 8 B (count, offset); per pair with contacts
    the 64 B record, 8 B keys and 2 x (32 B transform + 16 B inverse mass); 225 B written per row."""
    return n_pairs * 8.0 + 0.5 * n_contacts * (64.0 + 8.0 + 96.0) + n_contacts * ROW_BYTES


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


def world_step_leg(eng, world, args, bracket):
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
        ws = scenes.box_pile(1000, 100) if args.workload == "pile" else make_world(args.workload, 100_000, 1)[0]
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


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path on host cores.  The
    Haskell reference cannot be compiled here (no GHC), so this is the oracle port; it is
    single-threaded like the reference's ST engine."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warm = args.steps, args.warmup
    frames = max(1, min(steps, 6))   # each frame of the 200k-shape sample is ~0.5 s of single-thread CPU work
    s = cpu_sample(args.workload, budget_frames=frames + min(warm, 1))
    secs = s["seconds"][min(warm, 1):]
    per = float(np.mean(secs))
    value = s["pairs"] / per
    world_desc = make_world_desc(args)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": warm, "ms_per_step": per * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": world_desc, "sample": s["desc"], "frames_timed": len(secs),
                   "note": "C restatement of the Haskell reference (GHC absent), 1 thread"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": "port", "sample": s["desc"]},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "contacts_per_s": s["contacts"] / per,
    }
    print(json.dumps(line))


def make_world_desc(args):
    if args.workload == "pile":
        return f"config 3: dense pile of unit boxes, {args.shapes_per_gpu} boxes per GPU + one static floor"
    return f"{args.workload}, {args.shapes_per_gpu} shapes per GPU"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="pile", choices=["pile", "polygons", "mixed", "blob", "stacks"])
    ap.add_argument("--shapes-per-gpu", type=int, default=1_000_000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-world-step", action="store_true")
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
    from shapes_b200.engine import Engine, nccl_unique_id

    rank = int(os.environ.get("RANK", "0"))
    world_size = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.gpus != world_size:
        if world_size == 1 and args.gpus > 1:
            raise SystemExit("launch N>1 with: python -m torch.distributed.run --nproc-per-node N bench.py --gpus N ...")
        args.gpus = world_size
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback for the product path")
    torch.cuda.set_device(local_rank)
    dist = None
    nccl_id = None
    if world_size > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        box = [nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        nccl_id = box[0]
    if rank == 0:
        build.build_library()
    if dist is not None:
        dist.barrier()

    world, world_desc = make_world(args.workload, args.shapes_per_gpu, world_size)
    if args.slot_order == "morton":
        from shapes_b200 import scenes as _scenes
        world = _scenes.spatially_sorted(world)
        world_desc += ", slot keys in Morton order of position"
    n = world.n_slots
    vbar = world.n_verts / max(n, 1)
    cos_rot, sin_rot = np.cos(world.rot), np.sin(world.rot)
    own = (n + world_size - 1) // world_size
    max_pairs = int(own * 6 + 4096) if args.workload in ("pile", "stacks") else int(own * 8 + 4096)
    eng = Engine(world, max_pairs=max_pairs, max_contacts=2 * max_pairs, device=local_rank, rank=rank,
                 world_size=world_size, nccl_id=nccl_id)
    dev = torch.device("cuda", local_rank)
    exchange = "none"
    if dist is not None:
        exchange = "NCCL all-gather"
        if os.environ.get("SHAPES_B200_NO_P2P") is None:
            blobs = [None] * world_size
            dist.all_gather_object(blobs, eng.ipc_export())
            eng.ipc_import(blobs)
            exchange = "peer-to-peer stores from K0 (CUDA IPC over NVLink), flag barrier"
    cols = [world.pos_x, world.pos_y, world.rot, cos_rot, sin_rot, world.inv_lin, world.inv_rot]
    d_in = [torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in cols]
    ptrs = [t.data_ptr() for t in d_in]
    beh = dict(dt=world.meta.get("dt", 0.01), baumgarte=world.meta.get("baumgarte", 0.01), slop=world.meta.get("slop", 0.02))

    def step():
        return eng.frame_device(*ptrs, **beh)

    def bracket():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(args.warmup):
        out = step()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = eng.launch_count
    stage_acc: dict[str, float] = {}
    dev_ms = 0.0
    bracket()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        out = step()
        dev_ms += out.device_ms
    bracket()
    elapsed = time.perf_counter() - t0
    launches = eng.launch_count - launches0
    # per-stage CUDA events right after the timed region, same process and buffers (the frame graph
    # is bypassed for these frames because events inside a captured graph cannot be timed)
    eng.set_profiling(True)
    prof_steps = max(3, min(args.steps, 10))
    step()
    for _ in range(prof_steps):
        step()
        for k, v in eng.stage_ms().items():
            stage_acc[k] = stage_acc.get(k, 0.0) + v
    eng.set_profiling(False)
    # warm-start leg (not part of the headline metric): the same frame with
    # the previous frame's Lagrangian cache (device resident) joined against this frame's keys
    cache = torch.ones(2, max(int(out.n_contacts), 1), device=dev, dtype=torch.float64)
    warm_steps = max(3, min(args.steps, 10))
    for _ in range(2):
        eng.set_lagrangian_cache_device(int(out.n_contacts), cache[0].data_ptr(), cache[1].data_ptr())
        out = step()
    bracket()
    tw = time.perf_counter()
    for _ in range(warm_steps):
        eng.set_lagrangian_cache_device(int(out.n_contacts), cache[0].data_ptr(), cache[1].data_ptr())
        out = step()
    bracket()
    warm_ms = (time.perf_counter() - tw) / warm_steps * 1e3
    eng.set_profiling(True)
    eng.set_lagrangian_cache_device(int(out.n_contacts), cache[0].data_ptr(), cache[1].data_ptr())
    step()
    warm_join_ms = eng.stage_ms().get("warm_join", 0.0)
    eng.set_profiling(False)
    bracket()
    clocks = sampler.stop() if rank == 0 else None

    n_pairs, n_contacts = int(out.n_pairs), int(out.n_contacts)
    if dist is not None:
        t = torch.tensor([elapsed], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed = float(t.item())
        cnt = torch.tensor([n_pairs, n_contacts], device=dev, dtype=torch.int64)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
        tot_pairs, tot_contacts = int(cnt[0].item()), int(cnt[1].item())
    else:
        tot_pairs, tot_contacts = n_pairs, n_contacts
    per_step = elapsed / args.steps
    value = tot_pairs / per_step

    # ---- end to end through the host-buffer API (H2D + kernels + D2H every step) -------------
    e2e = None
    if not args.no_e2e:
        pin = {}
        for name, a in zip(("pos_x", "pos_y", "rot", "cos", "sin", "inv_lin", "inv_rot"), cols):
            t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
            pin[name] = t
        import copy
        w_pinned = copy.copy(world)
        w_pinned.pos_x, w_pinned.pos_y, w_pinned.rot = pin["pos_x"].numpy(), pin["pos_y"].numpy(), pin["rot"].numpy()
        w_pinned.inv_lin, w_pinned.inv_rot = pin["inv_lin"].numpy(), pin["inv_rot"].numpy()
        eng.world = w_pinned
        cs = (pin["cos"].numpy(), pin["sin"].numpy())
        want = ("pairs", "contacts", "constraints")
        e_steps = max(3, min(args.steps, 10))
        for _ in range(2):
            fr = eng.frame(cos_sin=cs, want=want, pinned=True, **beh)
        bracket()
        t0 = time.perf_counter()
        for _ in range(e_steps):
            fr = eng.frame(cos_sin=cs, want=want, pinned=True, **beh)
        bracket()
        e_elapsed = time.perf_counter() - t0
        if dist is not None:
            t = torch.tensor([e_elapsed], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e_elapsed = float(t.item())
        # 7 body columns; with the peer exchange each rank uploads only its own slot range
        n_up = (eng.rank_info()[1] - eng.rank_info()[0]) if exchange.startswith("peer") else n
        h2d = 7 * 8 * n_up
        d2h = eng._bufs.bytes_for(fr.n_pairs, fr.n_contacts, n, world.n_verts)
        e2e = {"value": tot_pairs / (e_elapsed / e_steps), "unit": UNIT, "ms_per_step": e_elapsed / e_steps * 1e3,
               "steps": e_steps, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
               "api": "shapes_frame (pinned host buffers, all result columns fetched)"}

    # ---- whole updateWorld on the device (not part of the headline metric) ----
    world_step = None
    if world_size == 1 and not args.no_world_step:
        try:
            world_step = world_step_leg(eng, world, args, bracket)
        except Exception as e:     # an extra leg must never cost the headline line
            world_step = {"error": f"{type(e).__name__}: {e}"}

    if rank != 0:
        eng.close()
        if dist is not None:
            dist.destroy_process_group()
        return

    peak, peak_src = measured_peak_gbs()
    st_ms = {k: v / prof_steps for k, v in stage_acc.items()}
    traffic = {}
    try:   # dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu capture
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            tj = json.load(f)
        if tj.get("shapes") == n and tj.get("workload") == world_desc:
            traffic = tj["dram_bytes_per_launch"]
            traffic["_source"] = tj["source"]
    except Exception:
        pass
    kernels = []
    for kname, stage, nbytes in (("k_manifolds", "manifolds", manifolds_algorithmic_bytes(n_pairs, n_contacts, vbar)),
                                 ("k_rows", "contact_rows", rows_algorithmic_bytes(n_pairs, n_contacts))):
        ms = st_ms.get(stage, 0.0)
        ach = nbytes / (ms * 1e-3) / 1e9 if ms > 0 else 0.0
        kernels.append({"bound": "hbm", "kernel": kname, "achieved": ach, "peak": peak, "unit": "GB/s",
                        "frac": ach / peak, "traffic": traffic.get(kname), "algorithmic_bytes_per_launch": nbytes,
                        "avg_launch_ms": ms, "share_of_device_time": ms / max(dev_ms / args.steps, 1e-9)})
    roofline = dict(max(kernels, key=lambda k: k["avg_launch_ms"]))
    roofline["peak_source"] = peak_src
    roofline["traffic_source"] = traffic.get("_source")
    roofline["other_kernels"] = [k for k in kernels if k["kernel"] != roofline["kernel"]]
    k3_bytes = sum(k["algorithmic_bytes_per_launch"] for k in kernels)

    cpu = None
    if not args.no_cpu_baseline:
        s = cpu_sample(args.workload)
        per = float(np.mean(s["seconds"][1:])) if len(s["seconds"]) > 1 else s["seconds"][0]
        cpu = {"value": s["pairs"] / per, "unit": UNIT, "cores": 1, "kind": "port", "sample": s["desc"],
               "ms_per_frame_of_sample": per * 1e3, "host_cores_present": os.cpu_count()}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world_size, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": per_step * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": world_desc, "shapes": n, "pairs_per_step": tot_pairs, "contacts_per_step": tot_contacts,
                   "mean_vertices": vbar, "cos_sin": "host supplied (numpy)", "grid": [int(out.grid_w), int(out.grid_h), float(out.cell_size)],
                   "big_shapes": int(out.n_big),
                   "l2": "per-frame working set (~%.0f MB written + read) exceeds the 126 MB L2; no explicit flush" %
                         ((k3_bytes + n * 300.0) / 1e6),
                   "parallelism": "1 GPU" if world_size == 1 else f"slot-range ownership over {world_size} ranks; AABB exchange: {exchange}"},
        "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
        "gpu_launches_note": "own kernels only; CUB radix sort / scan launch ~9 more per step",
        "device_ms_per_step": dev_ms / args.steps, "stage_ms": st_ms,
        "stage_ms_note": f"per-stage CUDA events over {prof_steps} extra frames after the timed region (graph bypassed)",
        "contacts_per_s": tot_contacts / per_step,
        "warm_start": {"ms_per_step_with_cache_join": warm_ms, "join_kernel_ms": warm_join_ms, "steps": warm_steps,
                       "note": "descZipVector join of this frame's keys with the previous frame's Lagrangian cache "
                               "(device resident); includes one 2 x 8 B x contacts D2D cache copy per step"},
        "roofline": roofline, "cpu_baseline": cpu, "world_step": world_step,
    }
    emit(line)
    eng.close()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
