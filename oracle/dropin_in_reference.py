#!/usr/bin/env python
"""TEST INFRASTRUCTURE - the drop-in of INTEGRATION.md §1 exercised inside the UNMODIFIED reference: the reference's own driver
sequence (vulcan.py:72-178 via oracle/ref_session.py) and its own `op.Integration` loop run with `vulcan_b200.ros2.Ros2` as the
solver object, for N steps from the reference's initial state; the state reached is compared with the fixture the reference
recorded with its own `op.Ros2` at the same step count (tests/golden/<cfg>_step00NN.npz).

This container has no GPU, so the solver object is wired to the oracle-backed stand-in of the C ABI (tests/oracle_columns.py):
what is checked here is the BOUNDARY - that the class honours the protocol `op.Integration`, `vulcan.py` and the reference's
condensation operators expect (attribute names, in-place mutation, call order) - not the CUDA arithmetic, which the `-m gpu`
tests check against the same fixtures through the same class.

usage:  python oracle/stage_reference.py --config Earth
        PYTHONHASHSEED=0 python oracle/dropin_in_reference.py --config Earth --steps 30
"""
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
    a = ap.parse_args()
    refdir = a.refdir or "/tmp/vulcan_ref_%s" % a.config
    import ref_session
    from vulcan_b200 import ros2 as ros2_mod
    from oracle_columns import oracle_backed_abi
    ros2_mod._abi = oracle_backed_abi()          # no GPU in this container (see the module docstring)

    def factory(op, cfg, chem_funs):
        return ros2_mod.Ros2(cfg=cfg, species=chem_funs.spec_list, refine=1)     # compo / charges read from cfg.com_file

    s = ref_session.setup(refdir, solver_factory=factory)
    cfg = s.cfg
    cfg.count_max = a.steps - 1                  # Integration.stop: count > count_max (op.py:1080)
    s.integ(s.var, s.atm, s.para, s.make_atm)    # the reference's own loop, condensation operators included
    var, para = s.var, s.para
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
