"""how do the synthetic sweep columns behave under the step controller for different start dt?"""
import sys, os
import numpy as np
REPO = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, "tests"))
import bench
from vulcan_b200 import ensemble
case = bench.load_case()
cfg, st = case.cfg, case.st
ncol = 296
sel = np.linspace(0, 4095, ncol).astype(int)
kz, met, co = ensemble.sweep_grid()
y, atom_ini = ensemble.synthetic_columns(case.y, st["n_0"], st["compo"], cfg["atom_list"], kz[sel], met[sel], co[sel])
kw = case.atm_kwargs()
kzz = kz[sel][:, None] * np.asarray(kw["Kzz"])[None, :]
for dt0 in (case.dt, 1e-4, 1e-8):
    r = ensemble.EnsembleRunner(case.net, case.nz, y, np.full(ncol, dt0), dict(kw), kzz, case.k, cfg, st["compo"], atom_ini, st["n_0"])
    for n in (5, 10, 20, 40):
        r.run(n - (0 if n == 5 else {10: 5, 20: 10, 40: 20}[n]))
        s = r.state(want_y=False)
        acc, rej = s["n_accept"].sum(), s["n_reject"].sum()
        print("dt0 %.2e after %2d iterations: accepted %.3f  median dt %.2e  min dt %.2e max dt %.2e" % (dt0, n, acc / (acc + rej), np.median(s["dt"]), s["dt"].min(), s["dt"].max()), flush=True)
