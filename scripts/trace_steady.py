"""GPU diagnostic: a config to steady state through the drop-in class + Integration mirror with a per-step trace of dt, delta, element
losses, refinement passes kept / tried.   python scripts/trace_steady.py HD209S -1 [refine_dt_min]"""
import os, sys
import numpy as np
REPO = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, "tests")); sys.path.insert(0, os.path.join(REPO, "oracle"))
from helpers import run_config, GOLD
from vulcan_b200 import ros2 as ros2_mod, _abi

tag, refine = sys.argv[1], int(sys.argv[2])
if len(sys.argv) > 3:
    _abi.REFINE_DT_MIN = float(sys.argv[3])
every = int(os.environ.get("EVERY", "50"))
orig = ros2_mod.Ros2.one_step
rows = []

def traced(self, var, atm, para):
    v, p = orig(self, var, atm, para)
    c = p.count
    if c % every == 0:
        kept, tried = self._col.refine_stats()
        rows.append((c, v.t, v.dt, p.delta, kept[0], tried[0]) + tuple(v.atom_loss[a] for a in self.cfg.atom_list))
        print("%5d t %.3e dt %.3e delta %.2e refine kept/tried %d/%d loss %s rej %d" % (
            c, v.t, v.dt, p.delta, kept[0], tried[0], " ".join("%s %+.1e" % (a, v.atom_loss[a]) for a in self.cfg.atom_list),
            p.delta_count + p.nega_count + p.loss_count), flush=True)
    return v, p

ros2_mod.Ros2.one_step = traced
import contextlib, io
case, var, atm, para, integ, wall = run_config(tag, refine=refine)
ref = np.load("%s/%s_full.npz" % (GOLD, tag))
yr, ym = ref["ymix"], var.ymix
rel = np.abs(ym - yr) / np.maximum(yr, 1e-300)
print("END %s refine %d: %d steps (+%d rejected) t %.4e wall %.1f s end_case %s | vs reference final: >1e-4 %.2e  >1e-8 %.2e  >1e-12 %.2e median(>1e-20) %.2e | loss %s" % (
    tag, refine, para.count, para.delta_count + para.nega_count + para.loss_count, var.t, wall, para.end_case,
    rel[yr > 1e-4].max(), rel[yr > 1e-8].max(), rel[yr > 1e-12].max(), np.median(rel[yr > 1e-20]),
    " ".join("%s %+.1e" % (a, v) for a, v in var.atom_loss.items())))
