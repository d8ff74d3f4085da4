"""Top stall-sample instructions of one kernel from an ncu report (source page, SASS view).
usage: python profiles/ncu_top.py <report.ncu-rep> <kernel-name-substring> [top N]"""
import csv, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kern],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
blocks, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "hdr": None, "rows": []}
        blocks.append(cur)
    elif cur is not None and cur["hdr"] is None:
        cur["hdr"] = r
    elif cur is not None and r:
        cur["rows"].append(r)
for b in blocks[:1]:
    h = b["hdr"]
    i_src, i_s, i_ex = h.index("Source"), h.index("Warp Stall Sampling (All Samples)"), h.index("Instructions Executed")
    tot = sum(int(r[i_s] or 0) for r in b["rows"])
    tot_ex = sum(int(r[i_ex] or 0) for r in b["rows"])
    print(b["name"], "instructions", len(b["rows"]), "samples", tot, "warp-instr executed", tot_ex)
    ranked = sorted(enumerate(b["rows"]), key=lambda t: -int(t[1][i_s] or 0))[:top]
    for k, r in ranked:
        print(f"{k:5d} {int(r[i_s] or 0):7d} {100.0 * int(r[i_s] or 0) / max(tot, 1):5.1f}%  ex {int(r[i_ex] or 0):9d}  {r[i_src].strip()}")
