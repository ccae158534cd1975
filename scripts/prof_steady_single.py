"""driver for ncu: the device-resident loop of ONE HD189 column (vk_ens_run_steady), n iterations.  python scripts/prof_steady_single.py [n]"""
import os, sys, time
import numpy as np
REPO = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, REPO)
from vulcan_b200.fixtures import Case, steady_ensemble_from_fixture
n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
c = Case("HD189", 0)
y = c.st["y_ini"][None].copy()
atom_ini = np.einsum("cji,ia->ca", y, c.st["compo"])
se = steady_ensemble_from_fixture(c, y, atom_ini, np.ones(1))
se.col.ens_run_steady(n)          # warm-up (first launches, allocations)
t0 = time.time()
se.col.ens_run_steady(n)
print("steady loop, one column: %.3f ms per iteration (wall), device %.3f ms" % (1e3 * (time.time() - t0) / n, se.col.last_kernel_ms()[0] / n))
