import os, sys
import numpy as np
REPO = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, "tests")); sys.path.insert(0, os.path.join(REPO, "oracle"))
from helpers import Case, gpu_columns, oracle_step
tag, step = sys.argv[1], int(sys.argv[2])
c = Case(tag, step)
ncol = 40
rng = np.random.default_rng(7)
y = c.y[None] * (1.0 + 0.3 * rng.uniform(-1.0, 1.0, size=(ncol,) + c.y.shape))
y[0] = c.y
y[1] = c.y
ymix = y / y.sum(axis=2, keepdims=True)
ymix[0] = c.ymix; ymix[1] = c.ymix
dt = np.full(ncol, min(c.dt, 1e2))
col = gpu_columns(c, ncol)
s1, m1, d1, st1 = col.ros2_solve(y, ymix, dt)
os.environ["VK_EMIT_JAC"] = "0"
tab = gpu_columns(c, ncol)
s2, m2, d2, st2 = tab.ros2_solve(y, ymix, dt)
ref = c.fx["sol"]
for q in (0, 1, 2, 17):
    m = np.abs(s2[q]) > 1e-8 * np.abs(s2[q]).max(axis=1, keepdims=True)
    rel = np.abs(s1[q] - s2[q]) / np.maximum(np.abs(s2[q]), 1e-300)
    jj, ii = np.unravel_index(np.argmax(np.where(m, rel, 0)), rel.shape)
    print("col %d: max rel diff emitted-J vs table-J %.2e at layer %d species %d (%s) values %.6e %.6e | delta %.6e %.6e" % (
        q, rel[m].max(), jj, ii, c.net.species[ii], s1[q][jj, ii], s2[q][jj, ii], d1[q], d2[q]))
mr = ref > 1e-30
print("col 0 vs the reference fixture sol: emitted-J %.2e, table-J %.2e" % (np.max(np.abs(s1[0] - ref)[mr] / ref[mr]), np.max(np.abs(s2[0] - ref)[mr] / ref[mr])))
