"""per-kernel device time on the resident state of an ensemble (vk_debug_time_kernel): python scripts/kernel_times.py [ncol]"""
import sys, os, ctypes
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from helpers import Case
from quick_time import make
from vulcan_b200 import _abi
c = Case("HD189", 100)
ncol = int(sys.argv[1]) if len(sys.argv) > 1 else 592
col = make(c, ncol, 0)
y = np.repeat(c.y[None], ncol, 0); ym = np.repeat(c.ymix[None], ncol, 0); dt = np.full(ncol, c.dt)
for _ in range(2): col.ros2_solve(y, ym, dt)
lib = _abi.load()
lib.vk_debug_time_kernel.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_float)]
names = ["lhs", "rhs", "factor", "solve(fwd+bwd)"]
out = []
for w, n in enumerate(names):
    ms = ctypes.c_float(0)
    rc = lib.vk_debug_time_kernel(col.handle, w, 5, ctypes.byref(ms))
    out.append("%s %.3f" % (n, ms.value))
print("ncol %d VK_LHS_DBG=%s | " % (ncol, os.environ.get("VK_LHS_DBG", "0")) + " | ".join(out) + " ms")
