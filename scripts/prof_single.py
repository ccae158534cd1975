"""one single-column step (for ncu)."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from helpers import Case
from quick_time import make  # noqa
ncol = int(sys.argv[1]) if len(sys.argv) > 1 else 1
c = Case("HD189", 100)
col = make(c, ncol, 1)
y = np.repeat(c.y[None], ncol, 0); ym = np.repeat(c.ymix[None], ncol, 0); dt = np.full(ncol, c.dt)
for _ in range(int(sys.argv[2]) if len(sys.argv) > 2 else 2):
    col.ros2_solve(y, ym, dt)
print("ok", col.last_kernel_ms())
