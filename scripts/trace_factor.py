"""VK_TRACE=1 python scripts/trace_factor.py : per-pivot-step timeline (clock64) of layer 5 of a single-column factorisation."""
import os, sys, ctypes
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from helpers import Case
from quick_time import make
from vulcan_b200 import _abi
c = Case("HD189", 100)
col = make(c, 1, 0)
for _ in range(2):
    col.ros2_solve(c.y[None], c.ymix[None], c.dt)
buf = (ctypes.c_longlong * (128 * 8))()
_abi.load().vk_debug_trace(buf)
t = np.array(buf[:], dtype=np.int64).reshape(128, 8)
names = ["own:after_bar", "own:fma_done", "own:inv_ready", "own:published", "w8:step_done"]
print("step  " + "  ".join("%-16s" % n for n in names) + "  step_total")
for k in range(20, 44):
    base = t[k, 0]
    print("%4d  " % k + "  ".join("%-16d" % (t[k, e] - base) for e in range(5)) + "  %d" % (t[k + 1, 0] - t[k, 0]))
d = np.diff(t[8:64, 0]); print("mean cycles/step", d.mean(), "min", d.min(), "max", d.max())
