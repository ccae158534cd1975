"""driver for ncu: N attempted steps of an ensemble built like the bench's (per-column Kzz => per-column atmosphere arrays), or of one column.
python scripts/prof_step.py [ncol] [n_steps]"""
import os, sys
import numpy as np
REPO = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, REPO)
from vulcan_b200.fixtures import Case
from vulcan_b200 import ensemble
ncol = int(sys.argv[1]) if len(sys.argv) > 1 else 592
nst = int(sys.argv[2]) if len(sys.argv) > 2 else 2
c = Case("HD189", 100)
kz, met, co = [a[:ncol] for a in ensemble.sweep_grid()]
y, atom_ini = ensemble.synthetic_columns(c.y, c.st["n_0"], c.st["compo"], c.cfg["atom_list"], kz, met, co)
kw = c.atm_kwargs()
kzz = kz[:, None] * np.asarray(kw["Kzz"])[None, :]
r = ensemble.EnsembleRunner(c.net, c.nz, y, np.full(ncol, c.dt), kw, kzz, c.k, c.cfg, c.st["compo"], atom_ini, c.st["n_0"])
r.run(nst)
print("ok", r.col.last_kernel_ms())
