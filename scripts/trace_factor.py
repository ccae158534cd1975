"""VK_TRACE=1 python scripts/trace_factor.py : per-panel timeline (clock64) of layer 5 of a single-column factorisation."""
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
L = t[100]
print("layer: start 0 | schur done %d | P0 published %d | panels done %d | layer end %d" % tuple(L[1:5] - L[0]))
names = ["next:after_bar", "next:V+bfrag", "next:diag_upd", "next:gj_done", "next:published", "generic:done", "owner:done"]
print("panel " + "  ".join("%-15s" % n for n in names) + "  panel_total")
for k in range(9):
    base = t[k, 0]
    print("%4d  " % k + "  ".join("%-15d" % (t[k, e] - base) for e in range(7)) + "  %d" % ((t[k + 1, 0] if k < 8 else L[3]) - t[k, 0]))

f = t[64]
print("warp 4, panel 3: after_bar 0 | pa loaded %d | V done %d | nv frags %d | tile0 %d | tile1 %d | tile2 %d | tile8 %d" % tuple(f[1:8] - f[0]))
