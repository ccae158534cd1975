"""device block-tridiagonal solve vs oracle / truth on a fixture system: where do they differ?  python scripts/diag_fixture_solve.py HD209S 400"""
import os, sys
import numpy as np
REPO = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, "tests")); sys.path.insert(0, os.path.join(REPO, "oracle"))
from helpers import Case, gpu_columns
from oracle import Oracle
tag, step = sys.argv[1], int(sys.argv[2])
c = Case(tag, step); o = Oracle(c.net); atm = o.make_atm(**c.atm_kwargs())
col = gpu_columns(c)
D, up, dn = o.lhs(atm, c.y, c.k, c.dt)
rhs = c.fx["chemdf"] + c.fx["diffdf"]
xt = o.blocktri_truth(D, up, dn, rhs, 3)
xo = o.blocktri_solve(o.blocktri_factor(D, up, dn), up, dn, rhs)
compo = c.st["compo"]; tot = (c.y[:, :, None] * compo[None]).sum(axis=(0, 1))
bud = lambda v: (v[:, :, None] * compo[None]).sum(axis=(0, 1)) / tot
sp = list(c.net.species)
def report(name, x):
    r = rhs - o.blocktri_matvec(D, up, dn, x)
    be = bud(x) - bud(xt)
    # per-layer contribution to the worst atom's budget error
    a = int(np.argmax(np.abs(be)))
    contrib = ((x - xt) * compo[None, :, a]).sum(axis=1) / tot[a]
    jw = int(np.argmax(np.abs(contrib)))
    sw = int(np.argmax(np.abs((x - xt)[jw] * compo[:, a])))
    print("%-12s res %.1e  bud-err %s  worst atom %d: layer %d contributes %.1e (species %s, dx %.2e, x %.2e, y %.2e)" % (
        name, np.abs(r).max() / np.abs(rhs).max(), np.array2string(be, precision=1), a, jw, contrib[jw], sp[sw], (x - xt)[jw, sw], xt[jw, sw], c.y[jw, sw]))
report("LAPACK", c.fx["k1"]); report("oracle", xo)
for rf in (0, 1, 2):
    x, st = col.blocktri_solve(D, up, dn, rhs, refine=rf)
    report("gpu r%d" % rf, x[0])
