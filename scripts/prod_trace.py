"""GPU diagnostic: clock64 timeline of the fused kernel's producer warps (block 0, layer 7) - where the assembly of one layer spends
its time.  Build in the container: python -c "from vulcan_b200 import build; build.build(variant='prodtrace')"; run with
VK_LIB_PATH=vulcan_b200/_lib/libvulcan_b200_prodtrace.so python scripts/prod_trace.py [ncol]"""
import ctypes, os, sys
import numpy as np
REPO = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, "tests")); sys.path.insert(0, os.path.join(REPO, "oracle"))
from helpers import Case, gpu_columns
ncol = int(sys.argv[1]) if len(sys.argv) > 1 else 592
c = Case("HD189", 100)
col = gpu_columns(c, ncol)
y = np.repeat(c.y[None], ncol, 0); ym = np.repeat(c.ymix[None], ncol, 0)
for _ in range(2):
    col.ros2_solve(y, ym, np.full(ncol, c.dt))
out = (ctypes.c_longlong * 16)()
col.lib.vk_debug_prod_trace(out)
t = np.array(out[:10], dtype=np.int64)
names = ["wait prefetch", "layer sums", "transport", "products", "wait FREE", "gather", "split entries", "diagonal+fence", "end-of-layer barrier"]
for q in range(9):
    print("%-24s %7d clk" % (names[q], t[q + 1] - t[q]))
print("layer total %d clk" % (t[9] - t[0]))
