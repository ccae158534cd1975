"""`.vul` wire format: schema written by vulcan_b200.vulio.save_out = the reference's Output.save_out
(op.py:3216-3255, key lists store.py:83-90), and the `ini_mix='vulcan_ini'` warm start (build_atm.py:166-176)."""
import pickle

import numpy as np
import pytest

from helpers import Case, mock_objects
from vulcan_b200 import vulio


def _objects():
    case = Case("HD189", 10)
    cfg, var, atm, para = mock_objects(case)
    var.y_ini = case.st["y_ini"]
    var.atom_conden = {}
    var.Rf = {1: "OH + H2 -> H2O + H"}
    var.y_time = [var.y * f for f in (1.0, 1.1, 1.2, 1.3)]
    var.t_time = [0.0, 1.0, 2.0, 3.0]
    cfg.save_evolution, cfg.save_evo_frq = True, 2
    return case, cfg, var, atm, para


def test_vul_schema_and_round_trip(tmp_path):
    case, cfg, var, atm, para = _objects()
    path = str(tmp_path / "run.vul")
    missing = vulio.save_out(path, var, atm, para, case.net.species, case.nr, cfg)
    with open(path, "rb") as f:
        raw = pickle.load(f)                                   # what plot_py/*.py do
    assert set(raw) == {"variable", "atm", "parameter"}
    v = raw["variable"]
    assert v["species"] == list(case.net.species) and v["nr"] == case.nr
    for key in ("k", "y", "ymix", "y_ini", "t", "dt", "atom_ini", "atom_loss", "bins", "cross", "cross_J", "n_branch"):
        assert key in v, key
    assert set(missing) <= set(vulio.var_save_keys(cfg))
    assert np.array_equal(v["y"], var.y) and np.array_equal(v["ymix"], var.ymix)
    assert np.array_equal(v["y_time"], np.array(var.y_time)[::2]) and np.array_equal(v["t_time"], [0.0, 2.0])
    assert np.array_equal(raw["atm"]["Kzz"], atm.Kzz) and raw["parameter"]["count"] == para.count
    data = vulio.load_vul(path)
    assert np.array_equal(data["variable"]["k"][3], var.k[3])


def test_vulcan_ini_maps_species_by_name(tmp_path):
    case, cfg, var, atm, para = _objects()
    path = str(tmp_path / "prev.vul")
    vulio.save_out(path, var, atm, para, case.net.species, case.nr, cfg)
    sp = list(case.net.species)
    new_species = [sp[5], "XYZ_not_there", sp[0], sp[17]]
    base = np.full((case.nz, 4), 7.0)
    y, missing = vulio.ini_from_vul(path, new_species, case.nz, base)
    assert missing == ["XYZ_not_there"]
    assert np.array_equal(y[:, 0], var.y[:, 5]) and np.array_equal(y[:, 2], var.y[:, 0]) and np.array_equal(y[:, 3], var.y[:, 17])
    assert np.all(y[:, 1] == 7.0)
    with pytest.raises(ValueError):
        vulio.ini_from_vul(path, new_species, case.nz + 1)


def test_not_a_vul_file(tmp_path):
    p = tmp_path / "x.vul"
    with open(p, "wb") as f:
        pickle.dump({"variable": {}}, f)
    with pytest.raises(ValueError):
        vulio.load_vul(str(p))


def test_schema_matches_the_reference_output():
    """tests/golden/HD189_vul_schema.json is the key/type tree of a .vul the UNMODIFIED reference wrote (oracle/make_vul_schema.py)"""
    import json
    import os
    from helpers import GOLD
    p = os.path.join(GOLD, "HD189_vul_schema.json")
    if not os.path.exists(p):
        pytest.skip("schema fixture missing")
    ref = json.load(open(p))
    case, cfg, var, atm, para = _objects()
    cfg.save_evolution = False                        # the HD189 cfg does not save the evolution (vulcan_cfg_HD189.py)
    assert ref["_top"] == ["variable", "atm", "parameter"]
    # the key list our writer aims for is exactly what the reference wrote
    assert set(vulio.var_save_keys(cfg)) | {"species", "nr"} == set(ref["variable"])
    # value conventions of the entries the solver path produces
    assert ref["variable"]["k"]["key_type"] == "int" and ref["variable"]["k"]["n"] == case.nr
    assert ref["variable"]["k"]["first_value"]["shape"] == [case.nz]
    assert ref["variable"]["y"]["shape"] == [case.nz, case.ni] == list(var.y.shape)
    assert ref["variable"]["J_sp"]["key_type"] == "tuple"
    assert ref["variable"]["species"]["n"] == case.ni
    for key in ("Kzz", "Dzz", "dzi", "Hp", "mu", "g", "n_0", "top_flux", "vs", "zco", "pico"):      # atm fields the mirror updates
        assert key in ref["atm"] and hasattr(atm, key), key
    for key in ("count", "delta", "delta_count", "nega_count", "loss_count", "end_case", "fix_species_start", "small_y", "nega_y"):
        assert key in ref["parameter"] and hasattr(para, key), key
