import os, sys, time
import numpy as np
REPO = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, REPO)
from vulcan_b200.fixtures import Case, steady_ensemble_from_fixture
from vulcan_b200 import ensemble
def timeit(tag, f, col, n=100):
    f(n)
    t0 = time.time(); f(n); w = time.time() - t0
    print("%-58s wall %.3f ms / iteration, device %.3f" % (tag, 1e3 * w / n, col.last_kernel_ms()[0] / n), flush=True)
for step in (100, 0):
    c = Case("HD189", step)
    kw = c.atm_kwargs()
    y = (c.y if step else c.st["y_ini"])[None].copy()
    dt = np.array([c.dt if step else float(c.cfg["dttry"])])
    atom_ini = np.einsum("cji,ia->ca", y, c.st["compo"])
    one = ensemble.EnsembleRunner(c.net, c.nz, y, dt, dict(kw), np.asarray(kw["Kzz"])[None], c.k, c.cfg, c.st["compo"], atom_ini, c.st["n_0"])
    timeit("EnsembleRunner, state of step %d, dt %.1e" % (step, dt[0]), one.run, one.col)
    st = one.state(want_y=False)
    print("    accepted %d rejected %d dt now %.3e" % (st["n_accept"][0], st["n_reject"][0], st["dt"][0]))
c = Case("HD189", 0)
y = c.st["y_ini"][None].copy()
atom_ini = np.einsum("cji,ia->ca", y, c.st["compo"])
se = steady_ensemble_from_fixture(c, y, atom_ini, np.ones(1), photo=False)
timeit("SteadyEnsemble handle, plain vk_ens_run", se.col.ens_run, se.col)
timeit("SteadyEnsemble handle, vk_ens_run_steady", se.col.ens_run_steady, se.col)
