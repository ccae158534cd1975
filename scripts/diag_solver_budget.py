"""diagnostic: at selected steps of a single-column run compare the device block-tridiagonal solve (refine 0/1/2) with the 80-bit
truth on the SAME system: linear residual and the atom budget of k1.  python scripts/diag_solver_budget.py HD209S 150 138"""
import os, sys
import numpy as np
REPO = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, "tests")); sys.path.insert(0, os.path.join(REPO, "oracle"))
from helpers import run_config
from oracle import Oracle
from vulcan_b200 import ros2 as ros2_mod

tag = sys.argv[1]
count_max = int(sys.argv[2])
first = int(sys.argv[3])
orig = ros2_mod.Ros2.solver
state = {}


def traced(self, var, atm, para):
    c = para.count
    if c >= first:
        y, dt = np.array(var.y, copy=True), float(var.dt)
        nz = y.shape[0]
        self._sync_atm(atm, nz); self._sync_k(var, nz); self._sync_opts(var, atm, para, nz)
        col = self._col
        chem, diff = col.eval_rhs(y)
        rhs = chem[0] + diff[0]
        D, up, dn = col.eval_lhs(y, dt)
        D, up, dn = D[0], up[0], dn[0]
        o = state.setdefault("o", Oracle(self.network))
        xt = o.blocktri_truth(D, up, dn, rhs, 3)
        compo = self._compo
        tot = (y[:, :, None] * compo[None]).sum(axis=(0, 1))
        bud = lambda x: (x[:, :, None] * compo[None]).sum(axis=(0, 1)) / tot
        res = lambda x: np.abs(rhs - o.blocktri_matvec(D, up, dn, x)).max() / np.abs(rhs).max()
        line = "%4d dt %.3e | truth bud %s" % (c, dt, np.array2string(bud(xt), precision=2))
        for rf in (0, 1, 2):
            x, st = col.blocktri_solve(D, up, dn, rhs, refine=rf)
            line += " | r%d res %.1e bud-err %.1e st %d" % (rf, res(x[0]), np.abs(bud(x[0]) - bud(xt)).max(), st[0])
        print(line, flush=True)
    return orig(self, var, atm, para)


ros2_mod.Ros2.solver = traced
case, var, atm, para, integ, wall = run_config(tag, refine=int(os.environ.get("REFINE", "0")), count_max=count_max)
print("end", para.count, {a: "%.2e" % v for a, v in var.atom_loss.items()})
