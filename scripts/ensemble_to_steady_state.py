"""The BASELINE ensemble (configs[4]: HD189-like columns, Kzz x metallicity x C/O) run to PER-COLUMN convergence in the device-resident loop
(vk_ens_run_steady: stop / conv per column, photolysis at the reference cadence, finished columns frozen); prints the per-column statistics.
python scripts/ensemble_to_steady_state.py [ncol] [max_iterations] [out.json]"""
import json, os, sys, time
import numpy as np
REPO = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, REPO)
from vulcan_b200.fixtures import Case, steady_ensemble_from_fixture
from vulcan_b200 import ensemble
ncol = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
max_it = int(sys.argv[2]) if len(sys.argv) > 2 else 6000
c = Case("HD189", 0)
kz, met, co = [a[:ncol] for a in ensemble.sweep_grid()]
y, atom_ini = ensemble.synthetic_columns(c.st["y_ini"], c.st["n_0"], c.st["compo"], c.cfg["atom_list"], kz, met, co)
t0 = time.time()
runner = steady_ensemble_from_fixture(c, y, atom_ini, kz)
t_setup = time.time() - t0
out = runner.run_to_steady_state(max_iterations=max_it)
na, nr, ec, t = out["n_accept"], out["n_reject"], out["end_case"], out["t"]
res = dict(ncol=ncol, setup_s=t_setup, wall_s=out["wall_s"], iterations=int(out.get("iterations", 0)), converged=int((ec == 1).sum()),
           end_case_counts={int(k): int((ec == k).sum()) for k in np.unique(ec)}, accepted_min=int(na.min()), accepted_median=float(np.median(na)),
           accepted_max=int(na.max()), rejected_total=int(nr.sum()), attempted_total=int(na.sum() + nr.sum()), t_min=float(t.min()), t_max=float(t.max()),
           column_steps_per_s=float((na.sum() + nr.sum()) / out["wall_s"]))
print(json.dumps(res))
if len(sys.argv) > 3:
    json.dump(res, open(sys.argv[3], "w"))
