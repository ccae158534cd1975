#!/usr/bin/env python
"""TEST INFRASTRUCTURE - the CONVERGED fixed point of the unmodified reference ("next round" item 1c).

BASELINE's steady-state bar (mixing ratios within 1e-6 above 1e-20) cannot be checked at the reference's DEFAULT stopping rule
(yconv_cri = 0.01: the run stops while y still moves by up to 1 % per look-back window, so two hash seeds of the reference itself
differ by 1e-2).  The steady state f(y) = 0 does not depend on the path, so this script runs the unmodified reference (scratch copy
of oracle/stage_reference.py, its own op.Integration loop and op.Ros2 solver) with the stopping rule TIGHTENED - only the three cfg
numbers read live by Integration.conv (op.py:1023-1024, 1058) change:

    yconv_cri  0.01 -> --yconv     slope_cri 1e-4 -> 1.0 (off: longdy alone decides)     yconv_min 0.1 -> 0.0 (second clause off)

and records the final state plus (count, t, dt, longdy) per step and snapshots of ymix on the way, so that the floor the reference
reaches against ITSELF (another PYTHONHASHSEED = another summation order of tau / omega_0) can be measured.

usage:  python oracle/stage_reference.py --config HD189 --dest /tmp/vulcan_fp_HD189_s0
        PYTHONHASHSEED=0 python oracle/fixed_point_reference.py --config HD189 --refdir /tmp/vulcan_fp_HD189_s0 --tag HD189_fp_seed0
"""
import argparse
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
GOLD = os.path.join(REPO, "tests", "golden")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="HD189")
    ap.add_argument("--refdir", required=True)
    ap.add_argument("--tag", required=True)
    ap.add_argument("--yconv", type=float, default=1e-8)
    ap.add_argument("--max-steps", type=int, default=3000)
    ap.add_argument("--snap-every", type=int, default=100)
    ap.add_argument("--out-dir", default=GOLD)
    a = ap.parse_args()
    sys.path.insert(0, HERE)
    import ref_session
    s = ref_session.setup(a.refdir)
    cfg, var, atm, para, solver = s.cfg, s.var, s.atm, s.para, s.solver
    default = dict(yconv_cri=cfg.yconv_cri, slope_cri=cfg.slope_cri, yconv_min=cfg.yconv_min, count_max=cfg.count_max)
    cfg.yconv_cri, cfg.slope_cri, cfg.yconv_min, cfg.count_max = a.yconv, 1.0, 0.0, a.max_steps
    traj, snaps, snap_counts = [], [], []
    orig = solver.one_step
    t0 = time.time()

    def hooked(var_, atm_, para_):
        c = para_.count
        dt_try = var_.dt
        n0 = para_.delta_count + para_.nega_count + para_.loss_count
        v, p = orig(var_, atm_, para_)
        traj.append((c, v.t, dt_try, v.dt, p.delta, p.delta_count + p.nega_count + p.loss_count - n0,
                     float(getattr(v, "longdy", np.nan)), float(getattr(v, "aflux_change", np.nan))))
        if c % a.snap_every == 0 and c > 0:
            snaps.append(v.ymix.copy())
            snap_counts.append(c)
        if c % 100 == 0:
            print("count %d t %.3e dt %.3e longdy %.3e wall %.0f s" % (c, v.t, v.dt, getattr(v, "longdy", np.nan), time.time() - t0), flush=True)
        return v, p

    solver.one_step = hooked
    s.integ(var, atm, para, s.make_atm)
    wall = time.time() - t0
    os.makedirs(a.out_dir, exist_ok=True)
    np.savez_compressed(os.path.join(a.out_dir, "%s.npz" % a.tag), y=var.y, ymix=var.ymix, t=var.t, dt=var.dt, count=para.count,
                        delta_count=para.delta_count, nega_count=para.nega_count, loss_count=para.loss_count, end_case=para.end_case,
                        longdy=var.longdy, longdydt=var.longdydt, wall_s=wall, traj=np.array(traj),
                        traj_cols=np.array(["count", "t_before", "dt_try", "dt_used", "delta", "n_reject", "longdy", "aflux_change"]),
                        snaps=np.array(snaps), snap_counts=np.array(snap_counts), atom_loss=np.array([var.atom_loss[x] for x in cfg.atom_list]),
                        n_0=atm.n_0, seed=os.environ.get("PYTHONHASHSEED", "unset"),
                        tightened=json.dumps(dict(yconv_cri=a.yconv, slope_cri=1.0, yconv_min=0.0, count_max=a.max_steps, default=default)))
    print("fixed point run %s: %d steps, %d rejected, t = %.4e, longdy %.3e, end_case %s, wall %.0f s" % (
        a.tag, para.count, para.delta_count + para.nega_count + para.loss_count, var.t, var.longdy, para.end_case, wall), flush=True)


if __name__ == "__main__":
    main()
