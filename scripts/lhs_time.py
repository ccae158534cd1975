import sys, os, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from helpers import Case
from quick_time import make
c = Case("HD189", 100)
ncol = 592
col = make(c, ncol, 0)
y = np.repeat(c.y[None], ncol, 0); ym = np.repeat(c.ymix[None], ncol, 0); dt = np.full(ncol, c.dt)
for _ in range(3): col.ros2_solve(y, ym, dt)
tot = 0
for _ in range(5):
    col.ros2_solve(y, ym, dt); tot += col.last_kernel_ms()[0]
print("VK_LHS_DBG=%s  step %.3f ms" % (os.environ.get("VK_LHS_DBG", "0"), tot / 5))
