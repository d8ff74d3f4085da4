"""Build named variants of the library (extra nvcc -D flags) next to the default one, for A/B runs on the GPU box:
   python profiles/variants.py name1:-DFOO=1,-DBAR=2 name2:-DBAZ=3 ...   ->  shapes_b200/lib/var_<name>.so
Select one at run time with SHAPES_B200_LIB=shapes_b200/lib/var_<name>.so."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from shapes_b200 import build
procs = []
for spec in sys.argv[1:]:
    name, _, flags = spec.partition(":")
    out = os.path.join(build.LIB_DIR, f"var_{name}.so")
    cmd = build.nvcc_command(out, [f for f in flags.split(",") if f])
    procs.append((name, subprocess.Popen(cmd, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, text=True)))
for name, p in procs:
    _, err = p.communicate()
    print(name, "ok" if p.returncode == 0 else "FAILED\n" + err[-2000:])
