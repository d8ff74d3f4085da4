"""Regenerates oracle_polygons64.npz from the oracle (run from the repo root)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import binding as orc          # noqa: E402
from shapes_b200 import scenes             # noqa: E402

w = scenes.random_polygons(64, density=2.0, static_frac=0.1, config=99)
c, s = orc.cos_sin(w.rot)
r = orc.frame(w, c, s, broadphase="aabb")
keep = {k: v for k, v in r.items() if k not in ("world_x", "world_y", "normal_wx", "normal_wy")}
np.savez_compressed(os.path.join(os.path.dirname(__file__), "oracle_polygons64.npz"), cos=c, sin=s, **keep)
print({k: v.shape for k, v in keep.items()})
