"""GPU: single-column kernel times (factor, solve) with the cyclic-reduction path (default for ncol = 1) and with block Thomas (VK_CR=0)."""
import os, sys
import numpy as np
REPO = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, "tests")); sys.path.insert(0, os.path.join(REPO, "oracle"))
from helpers import Case, gpu_columns
for tag, step in [("HD189", 100), ("HD209S", 150), ("HD189cho", 30)]:
    c = Case(tag, step)
    col = gpu_columns(c, 1, refine=0)
    for _ in range(3):
        sol, ymo, delta, st = col.ros2_solve(c.y, c.ymix, c.dt)
    import time
    t0 = time.time()
    for _ in range(50):
        col.ros2_solve(c.y, c.ymix, c.dt)
    el = (time.time() - t0) / 50
    print("%s-%d VK_CR=%s: lhs %.3f ms rhs %.3f factor %.3f solve %.3f | step (host call) %.3f ms  status %d delta %.6e" % (
        tag, step, os.environ.get("VK_CR", "1"), col.time_kernel(0, 10), col.time_kernel(1, 10), col.time_kernel(2, 10), col.time_kernel(4, 10), el * 1e3, st[0], delta[0]))
