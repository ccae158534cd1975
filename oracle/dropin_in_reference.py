#!/usr/bin/env python
"""TEST INFRASTRUCTURE - the drop-in of INTEGRATION.md §1 exercised inside the UNMODIFIED reference: the reference's own driver
sequence (vulcan.py:72-178 via oracle/ref_session.py) and its own `op.Integration` loop run with `vulcan_b200.ros2.Ros2` as the
solver object, for N steps from the reference's initial state; the state reached is compared with the fixture the reference
recorded with its own `op.Ros2` at the same step count (tests/golden/<cfg>_step00NN.npz).

Two modes:
  default   (build container, no GPU): the solver object is wired to the oracle-backed stand-in of the C ABI (tests/oracle_columns.py):
            what is checked is the BOUNDARY - that the class honours the protocol `op.Integration`, `vulcan.py` and the reference's
            condensation operators expect (attribute names, in-place mutation, call order);
  --cuda    (GPU box, staged copy oracle/_ref/<cfg> from oracle/build_ref.py): the real libvulcan_b200.so behind the same class -
            the seam vulcan.py:162-163 exercised end to end on the B200.  With --steady the run goes on to the reference's own
            stopping rule and the wall time is reported next to the reference's own (--reference runs the unmodified op.Ros2 on one
            host core of the same box: the north-star "reached N x faster than the reference on the box's host cores").
  --device-loop  additionally replaces `op.Integration` by vulcan_b200.steady.DeviceIntegration.

usage:  python oracle/stage_reference.py --config Earth
        PYTHONHASHSEED=0 python oracle/dropin_in_reference.py --config Earth --steps 30
        PYTHONHASHSEED=0 python oracle/dropin_in_reference.py --config HD189 --refdir oracle/_ref/HD189 --cuda --steady
"""
import time
import argparse
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
for p in (REPO, os.path.join(REPO, "tests"), HERE):
    if p not in sys.path:
        sys.path.insert(0, p)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="HD189")
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--refdir", default=None)
    ap.add_argument("--out", default=None, help="append a JSON line with the result")
    ap.add_argument("--ytol", type=float, default=1e-8, help="bound on y under the reference's mask (HD189cho: 5e-6, the reference's own "
                    "LAPACK solve is 2.8e-7 off the 80-bit solution there)")
    ap.add_argument("--cuda", action="store_true", help="the real CUDA library instead of the oracle-backed stand-in")
    ap.add_argument("--steady", action="store_true", help="run to the reference's own stopping rule instead of --steps")
    ap.add_argument("--reference", action="store_true", help="run the UNMODIFIED reference (op.Ros2) instead of the drop-in: the timing baseline")
    ap.add_argument("--device-loop", action="store_true", help="vulcan_b200.steady.DeviceIntegration instead of op.Integration")
    a = ap.parse_args()
    if a.out:
        a.out = os.path.abspath(a.out)           # ref_session.setup changes into the staged copy
    refdir = os.path.abspath(a.refdir or "/tmp/vulcan_ref_%s" % a.config)
    import ref_session
    from vulcan_b200 import ros2 as ros2_mod
    if not a.cuda and not a.reference:
        from oracle_columns import oracle_backed_abi
        ros2_mod._abi = oracle_backed_abi()      # no GPU in this container (see the module docstring)

    def factory(op, cfg, chem_funs):
        if a.cuda:
            return ros2_mod.Ros2(cfg=cfg, species=chem_funs.spec_list)           # product defaults; compo / charges read from cfg.com_file
        return ros2_mod.Ros2(cfg=cfg, species=chem_funs.spec_list, refine=1)

    t_setup = time.time()
    s = ref_session.setup(refdir, solver_factory=None if a.reference else factory)
    t_setup = time.time() - t_setup
    cfg = s.cfg
    if not a.steady:
        cfg.count_max = a.steps - 1              # Integration.stop: count > count_max (op.py:1080)
    t_run = time.time()
    if a.device_loop:
        from vulcan_b200.steady import DeviceIntegration
        DeviceIntegration(s.solver)(s.var, s.atm, s.para, s.make_atm)
    else:
        s.integ(s.var, s.atm, s.para, s.make_atm)    # the reference's own loop, condensation operators included
    t_run = time.time() - t_run
    var, para = s.var, s.para
    if a.steady:
        res = dict(config=a.config, mode="reference (unmodified op.Ros2, 1 host core)" if a.reference else ("drop-in, CUDA" if a.cuda else "drop-in, oracle stand-in"),
                   loop="vulcan_b200.steady.DeviceIntegration" if a.device_loop else type(s.integ).__module__ + "." + type(s.integ).__name__,
                   steps=int(para.count), rejected=int(para.delta_count + para.nega_count + para.loss_count), t=float(var.t), end_case=int(para.end_case),
                   longdy=float(var.longdy), wall_s=t_run, setup_s=t_setup, atom_loss={k: float(v) for k, v in var.atom_loss.items()})
        full = os.path.join(REPO, "tests", "golden", "%s_full.npz" % a.config)
        if os.path.exists(full):
            yr = np.load(full)["ymix"]
            rel = np.abs(var.ymix - yr) / np.maximum(yr, 1e-300)
            res.update(vs_fixture_gt_1e4=float(rel[yr > 1e-4].max()), vs_fixture_gt_1e12=float(rel[yr > 1e-12].max()), vs_fixture_median=float(np.median(rel[yr > 1e-20])))
        print(json.dumps(res))
        if a.out:
            with open(a.out, "a") as f:
                f.write(json.dumps(res) + "\n")
        return
    fx = np.load(os.path.join(REPO, "tests", "golden", "%s_step%04d.npz" % (a.config, a.steps)))
    assert para.count == a.steps, para.count
    yr, mr = fx["y"], fx["ymix"]
    m = (yr > cfg.atol) & (mr > cfg.mtol)
    res = dict(config=a.config, steps=a.steps, t=abs(var.t - float(fx["t"])) / float(fx["t"]), dt=abs(var.dt - float(fx["dt"])) / float(fx["dt"]),
               y=float(np.max(np.abs(var.y - yr)[m] / yr[m])), ymix=float(np.max(np.abs(var.ymix - mr)[m] / mr[m])),
               rejected=int(para.delta_count + para.nega_count + para.loss_count), solver=type(s.solver).__module__ + "." + type(s.solver).__name__,
               loop=type(s.integ).__module__ + "." + type(s.integ).__name__)
    print(json.dumps(res))
    if a.out:
        with open(a.out, "a") as f:
            f.write(json.dumps(res) + "\n")
    assert res["t"] < 1e-9 and res["dt"] < 1e-6 and res["y"] < a.ytol, res


if __name__ == "__main__":
    main()
