"""Markdown table of the rows-mode per-kernel CUDA-event times (SHAPES_B200_KERNEL_TIMES=1) printed by every rank.
usage: python profiles/kernel_times.py label=gpurun_out/x.err [label=file ...]   (rank 0 of each file, one column per file)"""
import re
import sys
from collections import OrderedDict

cols = OrderedDict()
for arg in sys.argv[1:]:
    label, path = arg.rsplit("=", 1)
    line = next((l for l in open(path) if "[shapes_b200 rank 0]" in l and "rows-mode kernel ms" in l), None)
    if line is None:
        continue
    body = line.split("):", 1)[1].split()
    acc = OrderedDict()
    for name, val in zip(body[0::2], body[1::2]):
        name = {"k_rw_publish:KEYS": "barrier KEYS", "k_rw_wait:KEYS": "barrier KEYS", "k_rw_sync:KEYS": "barrier KEYS",
                "k_rw_publish:CNT": "barrier CNT", "k_rw_wait:CNT": "barrier CNT", "k_rw_sync:CNT": "barrier CNT",
                "k_rw_publish:OFF": "barrier OFF", "k_rw_wait:OFF": "barrier OFF", "k_rw_sync:OFF": "barrier OFF",
                "k_rw_publish:RESULTS": "barrier RESULTS", "k_rw_wait:RESULTS": "barrier RESULTS", "k_rw_sync:RESULTS": "barrier RESULTS",
                "k_rw_publish:COUNTS": "barrier COUNTS", "k_rw_wait:COUNTS": "barrier COUNTS", "k_rw_sync:COUNTS": "barrier COUNTS",
                "k_scan_cells_sums": "cell scan", "k_scan_cells_apply": "cell scan", "cub_scan": "cub scans (2)",
                "k_big": "k_big (2)"}.get(name, name)
        acc[name] = acc.get(name, 0.0) + float(val)
    cols[label] = acc
names = []
for acc in cols.values():
    for n in acc:
        if n not in names:
            names.append(n)
print("| kernel | " + " | ".join(cols) + " |")
print("|---|" + "---|" * len(cols))
for n in names:
    print(f"| `{n}` | " + " | ".join(("%.3f" % acc[n]) if n in acc else "—" for acc in cols.values()) + " |")
print("| **sum** | " + " | ".join("**%.3f**" % sum(acc.values()) for acc in cols.values()) + " |")
