"""which final state is closer to a true steady state?  Run <cfg> to convergence on the GPU, then evaluate the tendency
f(y) = chemdf + diffdf (bit-identical to the reference's functions) at the GPU's final state and at the reference's final state
(tests/golden/<cfg>_full.npz), each with its own self-consistent photolysis rates, in the same final atmosphere.
python scripts/steady_residual.py HD209S"""
import os, sys
import numpy as np
REPO = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, "tests"))
from helpers import GOLD, run_config
tag = sys.argv[1]
case, var, atm, para, integ, wall = run_config(tag, refine=int(os.environ.get("REFINE", "0")))
solver = integ.odesolver
ref = np.load(os.path.join(GOLD, tag + "_full.npz"))
sp = list(case.net.species)
yg, yr = var.y.copy(), ref["y"].copy()
mg, mr = var.ymix.copy(), ref["ymix"].copy()


def tendency(y):
    var.y = y.copy(); var.ymix = y / np.vstack(np.sum(y, axis=1))
    for _ in range(4):                      # lagged diffuse flux: iterate the photolysis update to its fixed point
        solver.compute_tau(var, atm); solver.compute_flux(var, atm); solver.compute_J(var, atm)
    nz = y.shape[0]
    solver._sync_atm(atm, nz); solver._sync_k(var, nz)
    chem, diff = solver._col.eval_rhs(y)
    return chem[0] + diff[0]


fg, fr = tendency(yg), tendency(yr)
rel = np.abs(mg - mr) / np.maximum(mr, 1e-300)
print("GPU final t %.3e (%d steps, last dt %.3e), reference final t %.3e (last dt %.3e); max rel diff ymix>1e-4: %.3e, >1e-8: %.3e" % (var.t, para.count, var.dt, float(ref["t"]), float(ref["traj"][-1, 3]), rel[mr > 1e-4].max(), rel[mr > 1e-8].max()))
print("%-8s %5s %10s %10s | %12s %12s   (|f|/y = inverse e-folding time of the remaining drift, 1/s)" % ("species", "layer", "ymix ref", "rel diff", "|f|/y GPU", "|f|/y ref"))
big = np.argwhere((mr > 1e-8) & (rel > 0.02))
order = sorted(big, key=lambda ji: -rel[ji[0], ji[1]])[:14]
for j, i in order:
    print("%-8s %5d %10.3e %10.3e | %12.3e %12.3e" % (sp[i], j, mr[j, i], rel[j, i], abs(fg[j, i]) / yg[j, i], abs(fr[j, i]) / yr[j, i]))
m = mr > 1e-8
print("over all ymix > 1e-8: median |f|/y  GPU %.3e  reference %.3e ; 99th percentile GPU %.3e reference %.3e" % (
    np.median(np.abs(fg[m]) / yg[m]), np.median(np.abs(fr[m]) / yr[m]), np.percentile(np.abs(fg[m]) / yg[m], 99), np.percentile(np.abs(fr[m]) / yr[m], 99)))
