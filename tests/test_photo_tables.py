"""vulcan_b200/photo_tables.py (cross-section / stellar-flux readers) against the tables the UNMODIFIED reference built
from the same data files (`<cfg>_static.npz`: bins, cross, cross_J, cross_scat, cross_T, cross_J_T, cross_Jion, sflux_top).  Needs the
reference's data tree (thermo/photo_cross, atm/stellar_flux) - present in the build container, absent on the GPU box (skipped there)."""
import os
import runpy

import numpy as np
import pytest

from helpers import GOLD, REPO, have

REF = os.environ.get("VULCAN_REFERENCE", "/root/reference")
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "thermo", "photo_cross")), reason="reference data tree not present")

TAGS = [t for t in ("HD189", "Jupiter", "Earth", "HD209S", "HD189ion", "EarthS") if have(t, "static.npz")]


def _cfg(tag):
    """the cfg the fixture was recorded with: the reference's example cfg + the edits of oracle/stage_reference.py::CONFIGS"""
    import sys
    sys.path.insert(0, os.path.join(REPO, "oracle"))
    from stage_reference import CONFIGS
    c = CONFIGS[tag]
    cwd = os.getcwd()
    os.chdir(REF)
    try:
        ns = runpy.run_path(os.path.join(REF, c["src"]))
    finally:
        os.chdir(cwd)
    for k, v in c["edits"].items():
        ns[k] = eval(v, {}, {})
    return ns


@pytest.mark.parametrize("tag", TAGS)
def test_tables_match_the_reference(tag):
    import json
    from vulcan_b200 import photo_tables as pt
    st = dict(np.load(os.path.join(GOLD, tag + "_static.npz"), allow_pickle=False))
    cfgj = json.loads(str(st["cfg_json"]))
    cfg = _cfg(tag)
    psp = [str(s) for s in st["photo_sp"]]
    n_branch = {}
    for q in range(len(st["branch_sp"])):
        s = psp[int(st["branch_sp"][q])]
        n_branch[s] = max(n_branch.get(s, 0), int(st["branch_no"][q]))
    use_ion = bool(cfgj.get("use_ion"))
    isp, ion_branch = [], {}
    if use_ion:
        isp = [str(s) for s in st["ion_sp"]]
        for q in range(len(st["ion_branch_sp"])):
            s = isp[int(st["ion_branch_sp"][q])]
            ion_branch[s] = max(ion_branch.get(s, 0), int(st["ion_branch_no"][q]))
    tsp = [s for s in cfgj.get("T_cross_sp", []) if s in psp]
    sflux_file = os.path.join(REF, cfg["sflux_file"])
    lo, hi = pt.default_bin_range(sflux_file)
    T = pt.PhotoTables(os.path.join(REF, "thermo", "photo_cross") + "/", [str(s) for s in st["species"]], psp, n_branch, cfgj["scat_sp"], lo, hi,
                       cfgj["dbin1"], cfgj["dbin2"], cfgj["dbin_12trans"], T_cross_sp=tsp, Tco=st["Tco"], ion_sp=isp, ion_branch=ion_branch,
                       use_ion=use_ion)
    assert np.array_equal(T.bins, st["bins"])
    for i, s in enumerate(psp):
        assert np.array_equal(T.cross[s], st["cross"][i]), s
    for q in range(len(st["branch_sp"])):
        s, b = psp[int(st["branch_sp"][q])], int(st["branch_no"][q])
        assert np.array_equal(T.cross_J[(s, b)], st["cross_J"][q]), (s, b)
    for i, s in enumerate(cfgj["scat_sp"]):
        assert np.array_equal(T.cross_scat[s], st["cross_scat"][i]), s
    if use_ion:
        for i, s in enumerate(isp):
            assert np.array_equal(T.cross[s], st["ion_cross"][i]), s
        for q in range(len(st["ion_branch_sp"])):
            s, b = isp[int(st["ion_branch_sp"][q])], int(st["ion_branch_no"][q])
            assert np.array_equal(T.cross_Jion[(s, b)], st["cross_Jion"][q]), (s, b)
    if tsp:   # temperature-dependent tables: log10 / 10** evaluated on arrays here, per scalar in the reference -> last-ulp level
        for q, s in enumerate([str(x) for x in st["T_cross_sp"]]):
            ref = st["cross_T"][q]
            assert np.allclose(T.cross_T[s], ref, rtol=1e-13, atol=0), s
        for q, bq in enumerate(st["cross_J_T_branch"]):
            s, b = psp[int(st["branch_sp"][int(bq)])], int(st["branch_no"][int(bq)])
            assert np.allclose(T.cross_J_T[(s, b)], st["cross_J_T"][q], rtol=1e-13, atol=0), (s, b)
    top, i12 = pt.stellar_flux_top(sflux_file, T.bins, cfg["r_star"], cfg["orbit_radius"], cfgj["dbin_12trans"])
    assert i12 == int(st["sflux_din12_indx"])
    assert np.array_equal(top, st["sflux_top"])
