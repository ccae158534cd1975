import os, sys
import numpy as np
REPO = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, "tests")); sys.path.insert(0, os.path.join(REPO, "oracle"))
from helpers import Case, gpu_columns
from oracle import Oracle
tag, step = sys.argv[1], int(sys.argv[2])
c = Case(tag, step)
o = Oracle(c.net)
atm = o.make_atm(**c.atm_kwargs())
ncol = 40
rng = np.random.default_rng(11)
y = c.y[None] * (1.0 + 0.2 * rng.uniform(-1.0, 1.0, size=(ncol, 1, c.y.shape[1])))
dyn = np.asarray(c.net.tables()["dyn_k"])
kk = np.repeat(c.k[None], ncol, axis=0)
kk[:, :, dyn] *= rng.uniform(0.3, 3.0, size=(ncol, 1, len(dyn)))
kk[:, :, dyn] += 1e-12 * rng.uniform(0.0, 1.0, size=(ncol, c.nz, len(dyn)))
dt = np.full(ncol, min(c.dt, 1e2))
ymix = y / y.sum(axis=2, keepdims=True)
col = gpu_columns(c, ncol); col.set_k(kk, static_rows_shared=True)
s1, m1, d1, st1 = col.ros2_solve(y, ymix, dt)
tab = gpu_columns(c, ncol); tab.set_k(kk)
s2, m2, d2, st2 = tab.ros2_solve(y, ymix, dt)
os.environ["VK_EMIT_JAC"] = "0"
mid = gpu_columns(c, ncol); mid.set_k(kk, static_rows_shared=True)
s3, m3, d3, st3 = mid.ros2_solve(y, ymix, dt)
for q in (0, 13, 39):
    ref = o.ros2_solver(atm, y[q], ymix[q], kk[q], float(dt[q]), c.cfg["mtol"], c.cfg["atol"], refine=0)["sol"]
    m = np.abs(ref) > 1e-8 * np.abs(ref).max(axis=1, keepdims=True)
    e = lambda s: np.max(np.abs(s[q] - ref)[m] / np.abs(ref)[m])
    print("col %d vs oracle step: emitted %.2e | emitted chemdf + table J %.2e | table %.2e | status %d %d %d delta %.3e %.3e" % (q, e(s1), e(s3), e(s2), st1[q], st3[q], st2[q], d1[q], d2[q]))
