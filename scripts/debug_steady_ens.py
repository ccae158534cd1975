import os, sys
import numpy as np
REPO = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, "tests")); sys.path.insert(0, os.path.join(REPO, "oracle"))
from helpers import Case, steady_ensemble_from_fixture
from vulcan_b200 import ensemble
c = Case("HD189", 0)
kz = np.array([0.1, 0.3, 1.0, 1.0, 3.0, 10.0]); met = np.array([1.0, 1.0, 1.0, 2.0, 1.0, 0.5]); co = np.array([0.55, 0.55, 0.55, 0.8, 0.3, 0.55])
y, atom_ini = ensemble.synthetic_columns(c.st["y_ini"], c.st["n_0"], c.st["compo"], c.cfg["atom_list"], kz, met, co)
r = steady_ensemble_from_fixture(c, y, atom_ini, kz)
print("hist", r.hist_cap, r.hist_stride)
print("state0", r.col.ens_get_state(False), r.col.ens_get_steady())
for it in range(3):
    left = r.col.ens_run_steady(4)
    st = r.col.ens_get_state(False)
    print("after", 4 * (it + 1), "left", left, st["n_accept"], st["n_reject"], st["t"], st["dt"], r.col.ens_get_steady()["end_case"])

out = r.run_to_steady_state(max_iterations=128)
print("run_to_steady_state(128):", out["n_accept"], out["end_case"], out["iterations"], out["columns_left"])
