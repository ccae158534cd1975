"""driver for ncu / timing: photolysis updates (flux_kernel + jrate_kernel) of an ncol-column steady-state ensemble.
python scripts/prof_photo.py [ncol] [n_updates]"""
import os, sys, time
import numpy as np
REPO = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, REPO)
from vulcan_b200.fixtures import Case, steady_ensemble_from_fixture
from vulcan_b200 import ensemble
ncol = int(sys.argv[1]) if len(sys.argv) > 1 else 64
nup = int(sys.argv[2]) if len(sys.argv) > 2 else 3
c = Case("HD189", 0)
kz, met, co = [a[:ncol] for a in ensemble.sweep_grid()]
y, atom_ini = ensemble.synthetic_columns(c.st["y_ini"], c.st["n_0"], c.st["compo"], c.cfg["atom_list"], kz, met, co)
se = steady_ensemble_from_fixture(c, y, atom_ini, kz, hist_cap=2, hist_stride=1)
ms = []
for _ in range(nup):
    se.col.ens_photo_update()
    ms.append(se.col.last_kernel_ms()[0])
print("photolysis update of %d columns: %s ms" % (ncol, ["%.3f" % m for m in ms]))
