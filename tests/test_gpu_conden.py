"""Condensation operators on the device (vulcan_b200/csrc/vk_conden.cu: conden, h2o / nh3_conden_evap_relax, op.py:1109-1421) against the
inputs / outputs recorded from the UNMODIFIED reference's own operator calls (oracle/dump_fixtures.py --conden -> tests/golden/<cfg>_conden.npz):
rate coefficients of the growth reactions, y and ymix after the relaxation operators - BIT-EXACT (same expression order, -fmad=false)."""
import re

import numpy as np
import pytest

from helpers import Case, attach_conden, gpu_columns, have, mock_objects

pytestmark = pytest.mark.gpu
TAGS = [t for t in ("Jupiter", "JupiterFix", "Earth", "EarthS") if have(t, "conden.npz") and have(t, "static.npz")]


def _case(tag):
    import glob, os
    from helpers import GOLD
    step = int(re.search(r"step(\d+)", sorted(glob.glob(os.path.join(GOLD, tag + "_step*.npz")))[0]).group(1))
    return Case(tag, step)


@pytest.mark.parametrize("tag", TAGS)
def test_condensation_operators_bit_exact_on_the_device(tag):
    from vulcan_b200.steady import conden_tables
    case = _case(tag)
    cfg, var, atm, para = mock_objects(case, with_photo=False)
    cf = attach_conden(case, cfg, var, atm)
    col = gpu_columns(case, 1)
    col.set_k(np.ascontiguousarray(np.repeat(case.k[None], 1, 0)), shared=False)
    col.conden_setup(**conden_tables(cfg, var, atm, case.net.species))
    counts = sorted({int(m.group(1)) for m in (re.match(r"c(\d+)_(\w+?)_y_in$", k) for k in cf) if m})
    assert counts
    n_checked = 0
    for c in counts:
        ops = [n for n in ("conden", "h2o_conden_evap_relax", "nh3_conden_evap_relax") if "c%05d_%s_y_in" % (c, n) in cf]
        if "conden" not in ops:
            continue          # relaxation alone is never called by the reference loop (op.py:856-901): recorded after the switch by the hook only
        first, last = "c%05d_%s_" % (c, ops[0]), "c%05d_%s_" % (c, ops[-1])
        # the operators read atm.Dzz: the fixture stores the array of that call
        kw = case.atm_kwargs()
        kw["Dzz"] = cf[first + "Dzz"]
        col.set_atm(shared=True, **{a: kw[a] for a in ("Kzz", "vz", "dzi", "Dzz", "vs", "Tco", "g", "M", "Ti", "Hpi", "ms", "alpha", "top_flux", "bot_flux",
                                                          "bot_vdep", "use_moldiff", "use_settling", "use_topflux", "use_botflux", "gas_indx",
                                                          "gas_indx_lhs", "use_vm_mol", "vm", "diff_esc_idx")})
        y, ymix, kr = col.conden_apply(cf[first + "y_in"], cf[first + "ymix_in"], float(cf[first + "dt"]), case.st["n_0"])
        assert np.array_equal(y[0], cf[last + "y_out"]), "%s count %d: y differs (%s)" % (tag, c, ops)
        assert np.array_equal(ymix[0], cf[last + "ymix_out"]), "%s count %d: ymix differs" % (tag, c)
        want = cf["c%05d_conden_k_rows" % c]
        tab = conden_tables(cfg, var, atm, case.net.species)
        for q, re_id in enumerate(var.conden_re_list):
            if re_id in tab["re_idx"]:
                r = tab["re_idx"].index(re_id)
                assert np.array_equal(kr[0, r, 0], want[q, 0]) and np.array_equal(kr[0, r, 1], want[q, 1]), "%s count %d reaction %d" % (tag, c, re_id)
        n_checked += 1
    print("%s: %d recorded calls of the reference's condensation operators reproduced bit for bit on the device" % (tag, n_checked))
    assert n_checked > 0
