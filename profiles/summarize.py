"""Turn gpurun_out/{bench,launches,prof}_<tag> into profiles/<tag>_summary.md (run here, no GPU).

usage: python profiles/summarize.py <tag>
"""
import csv
import io
import json
import os
import subprocess
import sys
from collections import OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")

RAW_KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput %"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1/TEX throughput %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "FP64 pipe active %"),
    ("l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "global load requests"),
    ("l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "global load sectors"),
    ("l1tex__t_requests_pipe_lsu_mem_global_op_st.sum", "global store requests"),
    ("l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "global store sectors"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit %"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "active threads / warp instr"),
]


def launches_table(path):
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = OrderedDict()
    for r in rows[1:]:
        name = r[ki]
        name = name.split("(")[0].replace("<unnamed>::", "").replace("void ", "")
        if "cub::" in name:
            name = "cub::" + name.split("cub::")[1].split("<")[0]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += float(r[vi])
    tot = sum(v[1] for v in agg.values())
    lines = ["| kernel | launches | total us | share |", "|---|---|---|---|"]
    for k, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        lines.append(f"| `{k}` | {n} | {ns / 1e3:.1f} | {100 * ns / tot:.1f} % |")
    return "\n".join(lines), tot


def to_bytes(value, unit):
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    return float(value) * scale[unit]


def raw_tables(rep, traffic):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    out = []
    seen = set()
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        name = d["Kernel Name"].replace("<unnamed>::", "").replace("void ", "").split("(")[0]
        if name in seen:
            continue
        seen.add(name)
        traffic[name.split("<")[0]] = (to_bytes(d["dram__bytes_read.sum"], u["dram__bytes_read.sum"]) +
                                       to_bytes(d["dram__bytes_write.sum"], u["dram__bytes_write.sum"]))
        lines = [f"### `{name}`", "", "| metric | value |", "|---|---|"]
        for key, label in RAW_KEYS:
            if key in d:
                lines.append(f"| {label} (`{key}`) | {d[key]} {u[key]} |")
        stalls = [(float(d[h]), h) for h in hdr if "issue_stalled" in h and h.endswith("per_issue_active.ratio")]
        stalls.sort(reverse=True)
        lines.append("| top stall reasons (warps per issue) | " + ", ".join(
            f"{h.split('issue_stalled_')[1].split('_per_issue')[0]} {v:.2f}" for v, h in stalls[:5]) + " |")
        out.append("\n".join(lines))
    return "\n\n".join(out)


def main():
    tag = sys.argv[1]
    md = [f"# Profile summary {tag}", ""]
    bj = os.path.join(OUT, f"bench_{tag}.json")
    if os.path.exists(bj):
        d = json.loads([l for l in open(bj).read().splitlines() if l.startswith("{")][-1])
        md += ["## bench.py line (CUDA-event timing, not under a profiler)", "",
               f"* workload: {d['config']['workload']} — {d['config']['shapes']} shapes, "
               f"{d['config']['pairs_per_step']} pairs, {d['config']['contacts_per_step']} contacts per frame",
               f"* **{d['ms_per_step']:.3f} ms/frame** wall (device {d['device_ms_per_step']:.3f} ms), "
               f"{d['value'] / 1e9:.2f} G pairs/s, {d['contacts_per_s'] / 1e9:.2f} G contacts/s, n_gpus {d['n_gpus']}",
               f"* clocks: {d['clocks']}",
               "* stage times (ms, CUDA events on the ctx stream): " +
               ", ".join(f"{k} {v:.3f}" for k, v in d["stage_ms"].items()),
               f"* roofline: {json.dumps({k: v for k, v in d['roofline'].items() if k != 'other_kernels'})}",
               f"* other kernels: {json.dumps(d['roofline'].get('other_kernels'))}",
               f"* e2e: {json.dumps(d.get('e2e'))}",
               f"* cpu_baseline: {json.dumps(d.get('cpu_baseline'))}", ""]
    lc = os.path.join(OUT, f"launches_{tag}.csv")
    if os.path.exists(lc):
        table, tot = launches_table(lc)
        md += ["## ncu launch list (`--metrics gpu__time_duration.sum --clock-control none`, 60 launches ≈ 2 frames)", "",
               "Cold-cache, serialised per-launch times: compare SHARES with the stage times above, not absolutes.", "",
               table, ""]
    rep = os.path.join(OUT, f"prof_{tag}.ncu-rep")
    if os.path.exists(rep):
        traffic = {}
        md += ["## ncu --set full captures", "", raw_tables(rep, traffic), ""]
        with open(os.path.join(ROOT, "profiles", "traffic.json"), "w") as f:
            json.dump({"source": f"prof_{tag}.ncu-rep (ncu --set full --clock-control none)",
                       "workload": d["config"]["workload"] if os.path.exists(bj) else None,
                       "shapes": d["config"]["shapes"] if os.path.exists(bj) else None,
                       "dram_bytes_per_launch": traffic}, f, indent=1)
    # ---- device-resident world step (shapes_world_step)
    wj = [os.path.join(OUT, f"world_{tag}.json"), os.path.join(OUT, f"world_poly_{tag}.json")]
    if any(os.path.exists(x) for x in wj):
        md += ["## shapes_world_step (whole updateWorld on the device; profiles/world_step.py, CUDA events)", ""]
        for x in wj:
            if not os.path.exists(x):
                continue
            lines = [l for l in open(x).read().splitlines() if l.startswith("{")]
            if lines:
                w = json.loads(lines[-1])
                md += [f"* {w['workload']}: {json.dumps(w['ms_median'])} ms (median of {w['steps']} steps); {w['pairs']} pairs, "
                       f"{w['contacts']} contacts, {w['solver_nodes']} solver nodes, {w['queue_pushes']} queue pushes; "
                       f"cpu_solver: {json.dumps(w.get('cpu_solver'))}"]
        if os.path.exists(bj) and d.get("world_step"):
            md += [f"* bench.py `world_step` key: {json.dumps(d['world_step'])}"]
        md += [""]
    lw = os.path.join(OUT, f"launches_world_{tag}.csv")
    if os.path.exists(lw):
        table, tot = launches_table(lw)
        md += ["### ncu launch list of the step's own kernels (solver side)", "", table, ""]
    repw = os.path.join(OUT, f"prof_world_{tag}.ncu-rep")
    if os.path.exists(repw):
        md += ["### ncu --set full capture of `k_solve` (1M-box pile)", "", raw_tables(repw, {}), ""]
    repc = os.path.join(OUT, f"prof_coop_{tag}.ncu-rep")
    if os.path.exists(repc):
        md += ["### ncu --set full capture of `k_manifolds_coop` (config 2 at 1M polygons)", "", raw_tables(repc, {}), ""]
    path = os.path.join(ROOT, "profiles", f"{tag}_summary.md")
    with open(path, "w") as f:
        f.write("\n".join(md))
    print(path)


if __name__ == "__main__":
    main()
