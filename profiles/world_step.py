"""Times shapes_world_step (the whole updateWorld on the device) on a named workload and prints one
JSON line: per-step device milliseconds split into the hot path / chain construction / solver /
integration, node counts, and the oracle's sequential CPU time for the same step on a bounded sample.

  python profiles/world_step.py --workload pile --nx 1000 --ny 1000 --steps 20 [--cpu-sample 200]
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from shapes_b200 import scenes  # noqa: E402
from shapes_b200.engine import Engine  # noqa: E402
from shapes_b200.world import Bodies  # noqa: E402


def make(args):
    if args.workload == "pile":
        return scenes.box_pile(args.nx, args.ny)
    if args.workload == "polygons":
        return scenes.random_polygons(args.nx * args.ny)
    if args.workload == "blob":
        return scenes.gaussian_blob(args.nx * args.ny)
    if args.workload == "stacks":
        return scenes.stacks_scene((args.nx, args.ny), 0.0)
    raise SystemExit("unknown workload")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="pile")
    ap.add_argument("--nx", type=int, default=1000)
    ap.add_argument("--ny", type=int, default=1000)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--iterations", type=int, default=2)
    ap.add_argument("--cpu-sample", type=int, default=0, help="rows of the pile (or thousands of polygons) the oracle steps for the CPU figure; 0 = skip")
    args = ap.parse_args()
    w = make(args)
    n = w.n_slots
    rng = np.random.default_rng(11)
    b = Bodies(rng.uniform(-0.1, 0.1, n), rng.uniform(-0.1, 0.1, n), rng.uniform(-0.1, 0.1, n), np.full(n, 0.2), np.zeros(n))
    rows = []
    with Engine(w) as eng:
        eng.world_upload(b)
        for k in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            st = eng.world_step(external=(1, 0.0, -2.0), iterations=args.iterations)
            wall = (time.perf_counter() - t0) * 1e3
            if k >= args.warmup:
                rows.append(dict(frame=st.frame_ms, chains=st.chains_ms, solve=st.solve_ms, integrate=st.integrate_ms,
                                 total=st.total_ms, wall=wall, pairs=st.n_pairs, contacts=st.n_contacts,
                                 nodes=st.solver_nodes, pushes=st.queue_pushes, chains_n=st.body_chains))
        launches = eng.launch_count
    med = {k: float(np.median([r[k] for r in rows])) for k in rows[0]}
    out = {"workload": f"{args.workload} {args.nx}x{args.ny}", "shapes": n, "steps": args.steps, "iterations": args.iterations,
           "ms_median": {k: round(med[k], 4) for k in ("frame", "chains", "solve", "integrate", "total", "wall")},
           "pairs": int(med["pairs"]), "contacts": int(med["contacts"]), "solver_nodes": int(med["nodes"]),
           "queue_pushes": int(med["pushes"]), "body_chains": int(med["chains_n"]),
           "first_step_ms": rows[0]["total"], "last_step_ms": rows[-1]["total"]}
    if args.cpu_sample:
        from oracle import binding as orc
        ws = scenes.box_pile(args.nx, args.cpu_sample) if args.workload == "pile" else scenes.random_polygons(args.cpu_sample * 1000)
        ns = ws.n_slots
        bs = Bodies(rng.uniform(-0.1, 0.1, ns), rng.uniform(-0.1, 0.1, ns), rng.uniform(-0.1, 0.1, ns), np.full(ns, 0.2), np.zeros(ns))
        c, s = orc.cos_sin(ws.rot)
        cache = None
        ts = []
        for k in range(3):
            fr = orc.frame(ws, c, s, broadphase="sweep")
            t0 = time.perf_counter()
            orc.apply_external(ws, bs.vel_x, bs.vel_y, 1, 0.0, -2.0, 0.01)
            nrow = len(fr["key_i"])
            lam_np, lam_f, hit = (orc.warm_join(fr, *cache) if cache else (np.zeros(nrow), np.zeros(nrow), np.zeros(nrow, np.uint8)))
            orc.apply_cached(ws, fr, hit, lam_np, lam_f, bs.vel_x, bs.vel_y, bs.rot_vel)
            for _ in range(args.iterations):
                orc.improve_world(ws, fr, bs.mu, bs.bounce, bs.vel_x, bs.vel_y, bs.rot_vel, lam_np, lam_f)
            orc.advance(ws, bs.vel_x, bs.vel_y, bs.rot_vel, 0.01)
            c, s = orc.cos_sin(ws.rot)
            ts.append((time.perf_counter() - t0) * 1e3)
            cache = ({q: fr[q] for q in ("key_i", "key_j", "feat_a", "feat_b")}, lam_np, lam_f)
        out["cpu_solver"] = {"kind": "port", "cores": 1, "sample_shapes": ns, "sample_contacts": nrow,
                             "ms_per_step_sample": round(float(np.median(ts)), 3),
                             "ns_per_contact": round(float(np.median(ts)) * 1e6 / max(nrow, 1), 2),
                             "note": "external + join + applyCachedSlns + improveWorld sweeps + advance + libm cos/sin; contact generation excluded"}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
