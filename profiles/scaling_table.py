"""Markdown scaling table from bench lines of N-GPU runs.  usage: python profiles/scaling_table.py file.json [...]"""
import json
import sys

rows = []
for path in sys.argv[1:]:
    try:
        d = json.loads([l for l in open(path).read().splitlines() if l.startswith("{")][-1])
    except Exception as e:
        print("skip", path, e, file=sys.stderr)
        continue
    one = d.get("one_gpu_same_world") or {}
    weak = d.get("weak_config3") or {}
    e2e = d.get("e2e") or {}
    rows.append((d["n_gpus"], d["ms_per_step"], one.get("ms_per_step"), d["value"], weak.get("ms_per_step"), weak.get("value"),
                 e2e.get("ms_per_step"), d.get("per_rank_pairs"), d.get("per_rank_sat_ms"), path))
rows.sort()
print("| GPUs | config 4 (4M mixed, generator keys) ms/frame | same world on 1 GPU, same run | speed-up | efficiency | G pairs/s | e2e ms/frame (host buffers) | weak config 3 (1M boxes per GPU) ms/frame | max/mean SAT time over ranks |")
print("|---|---|---|---|---|---|---|---|---|")
for n, ms, one, val, wms, wval, e, prp, sat, path in rows:
    sp = (one / ms) if one else float("nan")
    bal = (max(sat) / (sum(sat) / len(sat))) if sat else float("nan")
    print(f"| {n} | {ms:.3f} | {one:.3f} | {sp:.2f}× | {100 * sp / n:.0f} % | {val / 1e9:.2f} | " + (f"{e:.2f}" if e else "—") + " | " +
          (f"{wms:.3f}" if wms else "—") + f" | {bal:.3f} |")
