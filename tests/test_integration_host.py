"""Host mirror of the caller (tests/integration_mirror.py) against the reference's own operators: the condensation growth
rates `conden`, the relaxation operators `h2o_conden_evap_relax` / `nh3_conden_evap_relax` (op.py:1109-1421) and the
fix_species switch (op.py:862-893).  Inputs and outputs were recorded from the UNMODIFIED reference while it ran
(oracle/dump_fixtures.py --conden -> tests/golden/<cfg>_conden.npz); the mirror keeps the reference's evaluation order, so
the comparison is BIT-EXACT."""
import re

import numpy as np
import pytest

from helpers import Case, attach_conden, have, mock_objects

TAGS = [t for t in ("Jupiter", "JupiterFix", "Earth", "EarthS") if have(t, "conden.npz") and have(t, "step0000.npz" if t != "JupiterFix" else "static.npz")]


def _objects(tag):
    from integration_mirror import Integration
    step = 0
    if not have(tag, "step%04d.npz" % step):       # JupiterFix only has the post-switch step fixture
        import glob, os
        from helpers import GOLD
        step = int(re.search(r"step(\d+)", sorted(glob.glob(os.path.join(GOLD, tag + "_step*.npz")))[0]).group(1))
    case = Case(tag, step)
    cfg, var, atm, para = mock_objects(case, with_photo=False)
    cf = attach_conden(case, cfg, var, atm)
    integ = Integration(None, cfg, case.net.species)
    return case, cfg, var, atm, para, cf, integ


def _calls(cf):
    """(count, operator) pairs with recorded inputs/outputs"""
    out = []
    for key in cf:
        m = re.match(r"c(\d+)_(\w+?)_y_in$", key)
        if m:
            out.append((int(m.group(1)), m.group(2)))
    return sorted(out)


@pytest.mark.parametrize("tag", TAGS)
def test_condensation_operators_bit_exact(tag):
    case, cfg, var, atm, para, cf, integ = _objects(tag)
    calls = _calls(cf)
    assert calls, "no recorded calls"
    seen = set()
    for c, name in calls:
        pre = "c%05d_%s_" % (c, name)
        var.y, var.ymix, var.dt = cf[pre + "y_in"].copy(), cf[pre + "ymix_in"].copy(), float(cf[pre + "dt"])
        atm.Dzz = cf[pre + "Dzz"].copy()
        v = getattr(integ, name)(var, atm)
        assert np.array_equal(v.y, cf[pre + "y_out"]), "%s at count %d: y differs" % (name, c)
        assert np.array_equal(v.ymix, cf[pre + "ymix_out"]), "%s at count %d: ymix differs" % (name, c)
        if name == "conden":
            for q, r in enumerate(var.conden_re_list):
                for d in (0, 1):
                    assert np.array_equal(np.broadcast_to(np.asarray(v.k[r + d], dtype=float), (case.nz,)), cf[pre + "k_rows"][q, d])
        seen.add(name)
    # every operator the cfg enables was exercised
    # (the reference only has relaxation operators for H2O and NH3, op.py:896-902: EarthS lists H2SO4 in use_relax, which goes through conden)
    want = {"conden"} | {"%s_conden_evap_relax" % s.lower() for s in cfg.use_relax if s in ("H2O", "NH3")}
    assert want <= seen, (want, seen)


def test_relaxation_moves_mass_between_phases():
    """property: the relaxation operators only move molecules between vapour and condensate (op.py:1363-1372)"""
    tag = "Earth" if "Earth" in TAGS else (TAGS[0] if TAGS else None)
    if tag is None:
        pytest.skip("no condensation fixture")
    case, cfg, var, atm, para, cf, integ = _objects(tag)
    c, name = [x for x in _calls(cf) if x[1] == "h2o_conden_evap_relax"][-1]
    pre = "c%05d_%s_" % (c, name)
    var.y, var.ymix, var.dt = cf[pre + "y_in"].copy(), cf[pre + "ymix_in"].copy(), float(cf[pre + "dt"])
    atm.Dzz = cf[pre + "Dzz"].copy()
    sp = list(case.net.species)
    iw, il = sp.index("H2O"), sp.index("H2O_l_s")
    before = var.ymix[:, iw] + var.ymix[:, il]
    v = integ.h2o_conden_evap_relax(var, atm)
    after = v.ymix[:, iw] + v.ymix[:, il]
    assert np.allclose(before, after, rtol=1e-12, atol=0)
    other = [i for i in range(len(sp)) if i not in (iw, il)]
    assert np.array_equal(v.ymix[:, other], cf[pre + "ymix_in"][:, other])


def test_fix_species_switch_vs_reference():
    if "JupiterFix" not in TAGS:
        pytest.skip("fixture missing")
    case, cfg, var, atm, para, cf, integ = _objects("JupiterFix")
    assert bool(cf["fix_species_start"]) and "switch_y" in cf
    var.y, var.ymix, var.t = cf["switch_y"].copy(), cf["switch_ymix"].copy(), float(cf["switch_t"])
    atm.vs = cf["switch_vs_before"].copy()
    cfg.rtol = float(cf["switch_rtol_before"])
    para.fix_species_start = False
    integ.start_fix_species(var, atm, para)
    assert para.fix_species_start
    assert cfg.rtol == float(cf["rtol_after"]) == cfg.post_conden_rtol
    assert np.array_equal(atm.vs, cf["vs_after"]) and not atm.vs.any()
    fs = [str(s) for s in cf["static_fix_species"]]
    for q, s in enumerate(fs):
        assert np.array_equal(var.fix_y[s], cf["fix_y"][q]), s
        assert int(atm.conden_min_lev[s]) == int(cf["conden_min_lev"][q]), s


def test_adapt_rtol_follows_the_reference_policy():
    """`use_adapt_rtol` (op.py:835-853): every 10th step an element loss above `loss_criteria` doubles the criterion and cuts rtol by 25 %
    (floor rtol_min); every 1000th step a loss below 2e-4 raises it by 25 % (ceiling rtol_max).  The expected sequence below is the
    reference's block transcribed statement by statement; step_ok keeps the frozen rtol (default argument bound at import, op.py:2489)."""
    from types import SimpleNamespace
    from integration_mirror import Integration
    cfg = SimpleNamespace(ini_update_photo_frq=100, use_condense=False, rtol=0.25, rtol_min=0.02, rtol_max=2.5)
    integ = Integration(odesolver=None, cfg=cfg, species=["H"])
    integ.loss_criteria = 0.0005
    rng = np.random.default_rng(3)
    ref_rtol, ref_crit = 0.25, 0.0005
    for count in range(0, 4001):
        loss = {"H": float(10.0 ** rng.uniform(-5, -2.5)) * (-1) ** count, "O": 1e-5}
        var, para = SimpleNamespace(atom_loss=loss), SimpleNamespace(count=count)
        integ.adapt_rtol(var, para)
        # op.py:836-850
        if count % 10 == 0:
            if max([np.abs(v) for v in loss.values()]) >= ref_crit:
                ref_crit *= 2.
                ref_rtol *= 0.75
                ref_rtol = max(ref_rtol, cfg.rtol_min)
        if count % 1000 == 0 and count > 0:
            if max([np.abs(v) for v in loss.values()]) < 2e-4:
                ref_rtol *= 1.25
                ref_rtol = min(ref_rtol, cfg.rtol_max)
        assert cfg.rtol == ref_rtol and integ.loss_criteria == ref_crit
    assert cfg.rtol < 0.25                     # the policy did act on this sequence
