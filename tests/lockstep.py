"""Shared body of the lock-step tests: run the first N accepted steps of a config from the reference's initial state through the
reference-facing host code (drop-in Ros2 object + Integration mirror) and compare the state reached with the state the
UNMODIFIED reference had at the same step count (tests/golden/<cfg>_step00NN.npz: y, ymix, t, next dt, recorded at the top of
its loop iteration NN).  Between step 0 and step NN the reference ran: the photolysis update at count 0, update_mu_dz /
update_phi_esc at count 0, NN x (one_step incl. rejected attempts, condensation operators where enabled, hydrostatic rescale,
step_size) - so agreement checks the whole per-step protocol, not one solver call."""
import numpy as np

from helpers import Case, have, run_config

# (config, step count of the second fixture).  BASELINE.json configs 1-4 + the use_vm_mol variants + use_ion on the ion test network.
LOCKSTEP = [("HD189", 10), ("Jupiter", 30), ("Earth", 30), ("HD209S", 30), ("HD189vm", 30), ("JupiterVm", 30), ("EarthVm", 30),
            ("HD189ion", 30),
            # cfg-switch variants first run on the device in round 2: use_moldiff = False, vz != 0 in three stencil variants, thermochemistry
            # only, the smallest (ni = 41, block 48) and the largest (ni = 99, block 120) shipped networks
            ("HD189nomol", 30), ("HD189vz", 30), ("JupiterVz", 30), ("JupiterVmVz", 30), ("HD189thermo", 30), ("HD189cho", 30), ("EarthS", 30)]
LOCKSTEP = [p for p in LOCKSTEP if have(p[0], "step%04d.npz" % p[1]) and have(p[0], "step0000.npz")]


def lockstep(tag, nstep, abi=None, refine=-1):
    ref = Case(tag, nstep)
    case, var, atm, para, integ, wall = run_config(tag, refine=refine, count_max=nstep - 1, abi=abi)   # Integration.stop: count > count_max
    assert para.count == nstep
    fx, cfg = ref.fx, ref.cfg
    out = dict(t=abs(var.t - float(fx["t"])) / float(fx["t"]), dt=abs(var.dt - float(fx["dt"])) / float(fx["dt"]))
    yr, mr = fx["y"], fx["ymix"]
    m = (yr > cfg["atol"]) & (mr > cfg["mtol"])                      # the reference's own significance mask (op.py:2949-2950)
    out["y"] = float(np.max(np.abs(var.y - yr)[m] / yr[m]))
    m30 = yr > 1e-30
    out["y_all"] = float(np.max(np.abs(var.y - yr)[m30] / yr[m30]))
    out["ymix"] = float(np.max(np.abs(var.ymix - mr)[m] / mr[m]))
    out["rejected"] = para.delta_count + para.nega_count + para.loss_count
    out["wall"] = wall
    return out


def check_ion_photolysis(abi=None):
    """compute_Jion (op.py:2789-2820) through the drop-in object against the reference's own output on the ion test network"""
    from helpers import photolysis_via_dropin
    case, px, out = photolysis_via_dropin("HD189ion", 0, abi=abi)
    st = case.st
    isp = [str(s) for s in st["ion_sp"]]
    for it in (1, 2):
        o = out[it - 1]
        sel = px["bin_sel"]
        ref = px["aflux%d" % it]
        assert np.max(np.abs(o["aflux"][:, sel] - ref) / np.maximum(np.abs(ref), 1e-30 * np.abs(ref).max())) < 1e-9
        for q in range(len(st["ion_branch_sp"])):
            s, b = isp[int(st["ion_branch_sp"][q])], int(st["ion_branch_no"][q])
            jref, kref = px["Jion%d" % it][q], px["kion%d" % it][q]
            assert np.abs(jref).max() > 0
            assert np.allclose(o["Jion"][(s, b)], jref, rtol=1e-9, atol=1e-12 * np.abs(jref).max() + 1e-300), (s, b)
            assert np.allclose(o["k"][int(st["ion_branch_rate_index"][q])], kref, rtol=1e-9, atol=1e-12 * np.abs(kref).max() + 1e-300)
        for q in range(len(st["branch_sp"])):       # the photodissociation rows next to them are unaffected
            kref = px["kphoto%d" % it][q]
            assert np.allclose(o["k"][int(st["branch_rate_index"][q])], kref, rtol=1e-9, atol=1e-12 * np.abs(kref).max() + 1e-300)


def long_trajectory(tag, nstep, abi=None):
    """the host protocol over `nstep` steps next to the reference's recorded trajectory (tests/golden/<cfg>_full.npz, column 1 = model
    time before each step): returns the largest relative deviation of the model time and the two rejection counts."""
    import os
    from helpers import GOLD
    from vulcan_b200 import ros2 as ros2_mod
    ref = np.load(os.path.join(GOLD, tag + "_full.npz"))["traj"]
    orig = ros2_mod.Ros2.one_step
    worst = dict(t=0.0, where=0)

    def traced(self, var, atm, para):
        c = para.count
        if c < len(ref) and ref[c, 1] > 0:
            d = abs(var.t - ref[c, 1]) / ref[c, 1]
            if d > worst["t"]:
                worst["t"], worst["where"] = d, c
        return orig(self, var, atm, para)
    ros2_mod.Ros2.one_step = traced
    try:
        case, var, atm, para, integ, wall = run_config(tag, refine=-1, count_max=nstep - 1, abi=abi, max_wall_s=3000)
    finally:
        ros2_mod.Ros2.one_step = orig
    return dict(dev=worst["t"], where=worst["where"], count=para.count, rejected=para.delta_count + para.nega_count + para.loss_count,
                ref_rejected=int(ref[:nstep, 5].sum()), wall=wall)
