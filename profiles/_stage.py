import json, sys
d = json.loads([l for l in open(sys.argv[1]).read().splitlines() if l.startswith('{')][-1])
print(sys.argv[1], "ms/step %.3f dev %.3f" % (d["ms_per_step"], d["device_ms_per_step"]))
print("  " + "  ".join("%s %.3f" % (k, v) for k, v in d["stage_ms"].items()))
r = d["roofline"]
print("  roofline", r["kernel"], "%.0f GB/s frac %.3f" % (r["achieved"], r["frac"]),
      [(k["kernel"], round(k["frac"], 3)) for k in r.get("other_kernels", [])])
if d.get("e2e"):
    print("  e2e ms/step %.2f" % d["e2e"]["ms_per_step"])
