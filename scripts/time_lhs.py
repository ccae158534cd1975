"""times the lhs path (which = 0) and the emitted Jacobian kernel alone (which = 7) on an ensemble built like the bench's.
python scripts/time_lhs.py [ncol]"""
import os, sys
import numpy as np
REPO = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, REPO)
from vulcan_b200.fixtures import Case
from vulcan_b200 import ensemble
ncol = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
c = Case("HD189", 100)
kz, met, co = [a[:ncol] for a in ensemble.sweep_grid()]
y, atom_ini = ensemble.synthetic_columns(c.y, c.st["n_0"], c.st["compo"], c.cfg["atom_list"], kz, met, co)
kw = c.atm_kwargs()
kzz = kz[:, None] * np.asarray(kw["Kzz"])[None, :]
r = ensemble.EnsembleRunner(c.net, c.nz, y, np.full(ncol, c.dt), kw, kzz, c.k, c.cfg, c.st["compo"], atom_ini, c.st["n_0"])
r.run(2)
print("TIMES", os.environ.get("VK_EMIT_JAC_TB"), os.environ.get("VK_EMIT_JAC_BLOCKS"), os.environ.get("VK_TAG"),
      "lhs %.3f ms" % r.col.time_kernel(0, 3), "negjac %.3f ms" % r.col.time_kernel(7, 3), "rhs %.3f ms" % r.col.time_kernel(1, 3),
      "step %.3f ms" % r.col.last_kernel_ms()[0])
