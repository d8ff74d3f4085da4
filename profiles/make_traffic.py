"""profiles/traffic.json from full ncu captures: dram__bytes_read.sum + dram__bytes_write.sum per launch (mean over the
captured launches) of every kernel, per workload key as bench.py names it ("pile:1000001", "polygons:1000000").

usage: python profiles/make_traffic.py <workload_key>=<report.ncu-rep> [...]
"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def short(name: str) -> str:
    name = name.replace("<unnamed>::", "").replace("void ", "").split("(")[0]
    # k_manifolds<4, 0, 0> -> k_manifolds<4>; k_manifolds_coop<1> -> k_manifolds_coop; other templates keep the base name
    base = name.split("<")[0]
    if base == "k_manifolds" and "<" in name:
        first = name.split("<")[1].split(",")[0].split(">")[0].strip().replace("(int)", "")
        return f"k_manifolds<{first}>"
    return base


def main():
    out = {"source": {}, "note": "ncu --set full --clock-control none; bytes per launch, mean of the captured launches",
           "workloads": {}}
    for arg in sys.argv[1:]:
        key, rep = arg.split("=", 1)
        txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(txt)))
        hdr, units = rows[0], rows[1]
        ir, iw, ik = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("Kernel Name")
        acc = {}
        for r in rows[2:]:
            b = float(r[ir]) * SCALE[units[ir]] + float(r[iw]) * SCALE[units[iw]]
            a = acc.setdefault(short(r[ik]), [0, 0.0])
            a[0] += 1
            a[1] += b
        out["workloads"][key] = {k: v[1] / v[0] for k, v in acc.items()}
        out["source"][key] = os.path.basename(rep)
    with open(os.path.join(ROOT, "profiles", "traffic.json"), "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
