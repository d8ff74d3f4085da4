import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import binding as orc
from shapes_b200 import scenes
from shapes_b200.engine import MultiEngine
orc.build()
for name, w in (("polygons", scenes.random_polygons(4000, density=1.5, config=71)), ("pile", scenes.box_pile(60, 40))):
    c, s = orc.cos_sin(w.rot)
    want = orc.frame(w, c, s, broadphase="sweep")
    with MultiEngine(w, 2) as eng:
        fr = eng.frame(cos_sin=(c, s))
        print(name, "pairs", fr.n_pairs, len(want["pair_i"]), "contacts", fr.n_contacts, len(want["key_i"]))
        for k in fr.cols:
            if k in want and len(fr[k]) == len(want[k]):
                g, x = np.asarray(fr[k]), np.asarray(want[k])
                bad = ~((g == x) | ((g != g) & (x != x)))
                if bad.any():
                    idx = np.nonzero(bad)[0]
                    print("  ", k, "bad", bad.sum(), "of", len(g), "first", idx[:6], "got", g[idx[:4]], "want", x[idx[:4]])
