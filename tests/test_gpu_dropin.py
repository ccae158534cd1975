"""The op.Ros2-protocol class (vulcan_b200/ros2.py) driven with stand-ins for the reference's store.Variables / AtmData /
Parameters containers (the GPU box has no /root/reference), checked against the reference fixtures."""
import json
from types import SimpleNamespace

import numpy as np
import pytest

from helpers import Case, GOLD, have, mock_objects

pytestmark = pytest.mark.gpu


def _mock(case):
    return mock_objects(case, with_photo=False)


@pytest.mark.parametrize("tag,step", [("HD189", 0), ("HD189", 100), ("HD189vm", 0), ("HD189vm", 30)])
def test_one_step_like_the_reference(tag, step):
    from vulcan_b200.ros2 import Ros2
    if not have(tag, "step%04d.npz" % step):
        pytest.skip("fixture missing")
    case = Case(tag, step)
    cfg, var, atm, para = _mock(case)
    solver = Ros2(cfg=cfg, species=list(case.st["species"]), compo=case.st["compo"], network=case.net, refine=1)
    solver.naming_solver(para)
    assert para.solver_str == "solver"
    var, para = solver.one_step(var, atm, para)
    fx = case.fx
    assert bool(fx["step_ok"])
    assert abs(para.delta - float(fx["delta"])) <= (1e-10 if case.dt <= 1e-6 else 1e-6) * float(fx["delta"])
    ref = fx["clip_y"]
    m = (ref > case.cfg["atol"]) & (fx["clip_ymix"] > case.cfg["mtol"])
    tol = 1e-10 if case.dt <= 1e-6 else 1e-6
    assert np.max(np.abs(var.y - ref)[m] / ref[m]) < tol
    for q, a in enumerate(cfg.atom_list):
        assert abs(var.atom_loss[a] - float(fx["atom_loss"][q])) <= 1e-9 * abs(float(fx["atom_loss"][q])) + 1e-13   # sums are rounding-level at step 0
    # step-size control against the reference trajectory (row count+1 holds the next dt_try)
    if tag == "HD189" and have("HD189", "full.npz"):
        tr = np.load("%s/HD189_full.npz" % GOLD)["traj"]
        var = solver.step_size(var, para)
        assert abs(var.dt - tr[step + 1, 2]) <= 1e-6 * tr[step + 1, 2]


def test_photolysis_methods():
    from vulcan_b200.ros2 import Ros2
    case = Case("HD189", 0)
    st = case.st
    cfg, var, atm, para = _mock(case)
    px = dict(np.load("%s/HD189_photo0000.npz" % GOLD))
    psp = [str(s) for s in st["photo_sp"]]
    var.photo_sp = set(psp)
    var.ion_sp = set()
    var.bins, var.sflux_top = st["bins"], st["sflux_top"]
    var.sflux_din12_indx, var.dbin1, var.dbin2 = int(st["sflux_din12_indx"]), float(st["dbin1"]), float(st["dbin2"])
    var.cross = {s: st["cross"][i] for i, s in enumerate(psp)}
    var.cross_scat = {s: st["cross_scat"][i] for i, s in enumerate(cfg.scat_sp)}
    var.n_branch, var.cross_J, var.pho_rate_index = {}, {}, {}
    for q in range(len(st["branch_sp"])):
        s, b = psp[int(st["branch_sp"][q])], int(st["branch_no"][q])
        var.n_branch[s] = max(var.n_branch.get(s, 0), b)
        var.cross_J[(s, b)] = st["cross_J"][q]
        var.pho_rate_index[(s, b)] = int(st["branch_rate_index"][q])
    var.y, var.ymix = px["y"], px["ymix"]
    atm.dz = px["dz"]
    solver = Ros2(cfg=cfg, compo=st["compo"], network=case.net)
    for it in (1, 2):
        solver.compute_tau(var, atm)
        solver.compute_flux(var, atm)
        solver.compute_J(var, atm)
        sel = px["bin_sel"]
        ref = px["aflux%d" % it]
        assert np.max(np.abs(var.aflux[:, sel] - ref) / np.maximum(np.abs(ref), 1e-30 * np.abs(ref).max())) < 1e-9
        assert abs(var.aflux_change - float(px["aflux_change%d" % it])) < 1e-9
        kref = px["kphoto%d" % it]
        for q in range(len(st["branch_sp"])):
            rid = int(st["branch_rate_index"][q])
            assert np.allclose(var.k[rid], kref[q], rtol=1e-9, atol=1e-12 * np.abs(kref[q]).max() + 1e-300)
