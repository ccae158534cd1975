"""TEST TOOL (not collected by pytest): a long run of the host protocol on the CPU - drop-in Ros2 object wired to the oracle-backed
stand-in of the C ABI - printed next to the reference's recorded trajectory (tests/golden/<cfg>_full.npz).  The oracle runs the same
block-tridiagonal algorithm as the GPU, so a production-dt problem of the solver shows up here without a GPU.
usage: python tests/trace_host.py Earth 1200 [every]"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from helpers import GOLD, run_config          # noqa: E402
from oracle_columns import oracle_backed_abi   # noqa: E402
from vulcan_b200 import ros2 as ros2_mod       # noqa: E402

tag = sys.argv[1]
count_max = int(sys.argv[2]) if len(sys.argv) > 2 else 300
every = int(sys.argv[3]) if len(sys.argv) > 3 else 50
ref = np.load(os.path.join(GOLD, tag + "_full.npz"))["traj"] if os.path.exists(os.path.join(GOLD, tag + "_full.npz")) else None
orig = ros2_mod.Ros2.one_step
worst = dict(t=0.0, where=0)


def traced(self, var, atm, para):
    c = para.count
    if ref is not None and c < len(ref) and ref[c, 1] > 0:      # column 1 of the recorded trajectory: model time BEFORE step c
        d = abs(var.t - ref[c, 1]) / ref[c, 1]
        if d > worst["t"]:
            worst["t"], worst["where"] = d, c
    var, para = orig(self, var, atm, para)
    if c % every == 0:
        r = "" if ref is None or c >= len(ref) else "   | ref dt %.3e delta %.3e rej %d" % (ref[c, 3], ref[c, 4], ref[c, 5])
        print("%5d t %.6e dt %.3e delta %.3e rej %d/%d/%d loss %.2e  max |t/t_ref-1| so far %.1e%s" % (
            c, var.t, var.dt, para.delta, para.delta_count, para.nega_count, para.loss_count,
            max(abs(v) for v in var.atom_loss.values()), worst["t"], r), flush=True)
    return var, para


ros2_mod.Ros2.one_step = traced
case, var, atm, para, integ, wall = run_config(tag, refine=0, count_max=count_max, abi=oracle_backed_abi(), max_wall_s=3000)
print("end: count %d end_case %d wall %.0f s, max relative deviation of t from the reference's trajectory %.2e (largest at step %d)" %
      (para.count, para.end_case, wall, worst["t"], worst["where"]))
