import os, sys, time, copy
import numpy as np
REPO = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, REPO)
from vulcan_b200.fixtures import Case, steady_ensemble_from_fixture
from vulcan_b200 import ensemble
n = 200
def run(tag, photo=True, update_frq=None, plain=False):
    c = Case("HD189", 0)
    if update_frq is not None:
        c.cfg["update_frq"] = update_frq
    y = c.st["y_ini"][None].copy()
    atom_ini = np.einsum("cji,ia->ca", y, c.st["compo"])
    se = steady_ensemble_from_fixture(c, y, atom_ini, np.ones(1), photo=photo)
    f = (lambda k: se.col.ens_run(k)) if plain else (lambda k: se.col.ens_run_steady(k))
    f(n)
    t0 = time.time(); f(n); w = time.time() - t0
    print("%-40s %.3f ms per iteration" % (tag, 1e3 * w / n), flush=True)
run("steady loop (photo, mu_dz)")
run("steady loop, no photolysis", photo=False)
run("steady loop, no photolysis, no mu_dz", photo=False, update_frq=0)
run("plain vk_ens_run on the same handle", photo=False, plain=True)
