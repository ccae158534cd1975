import os, sys
import numpy as np
REPO = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, "tests")); sys.path.insert(0, os.path.join(REPO, "oracle"))
from helpers import Case, gpu_columns
for tag, step in [("HD189", 100), ("HD189cho", 30), ("HD209S", 30), ("EarthS", 300), ("Jupiter", 30)]:
    c = Case(tag, step)
    col = gpu_columns(c, 2)
    y = np.repeat(c.y[None], 2, 0) * np.array([1.0, 1.0 + 1e-3])[:, None, None]
    dt = np.array([c.dt, 2 * c.dt])
    D0, u0, l0 = col.eval_lhs(y, dt)
    os.environ["VK_LHS_VIA_FUSED"] = "1"
    D1, u1, l1 = col.eval_lhs(y, dt)
    del os.environ["VK_LHS_VIA_FUSED"]
    for name, a, b in (("up", u0, u1), ("dn", l0, l1), ("D", D0, D1)):
        bad = np.argwhere(a != b)
        print(tag, step, name, "mismatches", len(bad), "first", bad[:4].tolist(), [(a[tuple(q)], b[tuple(q)]) for q in bad[:3]])
