"""trace of a single-column run through the drop-in object + Integration mirror next to the reference's recorded trajectory:
python scripts/trace_config.py HD209S [count_max] [refine] [every]"""
import os, sys
import numpy as np
REPO = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, "tests"))
from helpers import GOLD, run_config
from vulcan_b200 import ros2 as ros2_mod

tag = sys.argv[1] if len(sys.argv) > 1 else "HD209S"
count_max = int(sys.argv[2]) if len(sys.argv) > 2 else 400
refine = int(sys.argv[3]) if len(sys.argv) > 3 else 0
every = int(sys.argv[4]) if len(sys.argv) > 4 else 10
ref = np.load(os.path.join(GOLD, tag + "_full.npz"))["traj"] if os.path.exists(os.path.join(GOLD, tag + "_full.npz")) else None
orig = ros2_mod.Ros2.one_step


def traced(self, var, atm, para):
    c, dt_try = para.count, var.dt
    n0 = (para.delta_count, para.nega_count, para.loss_count)
    var, para = orig(self, var, atm, para)
    if c % every == 0 or (para.delta_count, para.nega_count, para.loss_count) != n0:
        r = "" if ref is None or c >= len(ref) else "   | ref t %.4e dt %.3e delta %.3e rej %d" % (ref[c, 1], ref[c, 3], ref[c, 4], ref[c, 5])
        print("%5d t %.4e dt_try %.3e dt %.3e delta %.3e rej d/n/l %d/%d/%d loss %.2e%s" % (
            c, var.t, dt_try, var.dt, para.delta, para.delta_count, para.nega_count, para.loss_count,
            max(abs(v) for v in var.atom_loss.values()), r), flush=True)
    return var, para


ros2_mod.Ros2.one_step = traced
case, var, atm, para, integ, wall = run_config(tag, refine=refine, count_max=count_max)
print("end: count %d end_case %d wall %.1f" % (para.count, para.end_case, wall))
