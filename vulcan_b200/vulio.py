""".vul result files - the reference's output / warm-start wire format.

The reference writes ONE pickle (protocol 4) holding `{'variable': var_save, 'atm': vars(atm), 'parameter': vars(para)}`
(`Output.save_out`, op.py:3216-3255), where `var_save` = `{'species': [...], 'nr': nr}` + the attributes named in
`Variables.var_save` (store.py:83-88) + `y_time`, `t_time` thinned by `save_evo_frq` when `save_evolution` is on
(store.py:90, op.py:3241-3247).  `plot_py/*.py` read these files, and `ini_mix = 'vulcan_ini'` warm-starts a run from one
(build_atm.py:166-176: species matched BY NAME, so the two networks may differ).  This module writes and reads that schema
for the objects the drop-in solver and the Integration mirror hand back, so a GPU run can feed the reference's plotting
scripts or a later reference run, and a reference result can start a GPU run.
"""
import pickle

import numpy as np

VAR_SAVE = ['k', 'y', 'ymix', 'y_ini', 't', 'dt', 'longdy', 'longdydt', 'atom_ini', 'atom_sum', 'atom_loss', 'atom_conden',
            'aflux_change', 'Rf']                                                                    # store.py:83-84
VAR_SAVE_PHOTO = ['nbin', 'bins', 'dbin1', 'dbin2', 'tau', 'sflux', 'aflux', 'cross', 'cross_scat', 'cross_J', 'J_sp',
                  'n_branch']                                                                        # store.py:86
VAR_SAVE_TCROSS = ['cross_J', 'cross_T']                                                             # store.py:87
VAR_SAVE_ION = ['charge_list', 'ion_sp', 'cross_Jion', 'Jion_sp', 'ion_wavelen', 'ion_branch', 'ion_br_ratio']   # store.py:88
VAR_EVOL_SAVE = ['y_time', 't_time']                                                                 # store.py:90


def var_save_keys(cfg):
    keys = list(VAR_SAVE)
    if getattr(cfg, "use_photo", False):
        keys.extend(VAR_SAVE_PHOTO)
        if getattr(cfg, "T_cross_sp", None):
            keys.extend(VAR_SAVE_TCROSS)
        if getattr(cfg, "use_ion", False):
            keys.extend(VAR_SAVE_ION)
    return keys


def _vars(obj):
    return dict(obj) if isinstance(obj, dict) else dict(vars(obj))


def save_out(path, var, atm, para, species, nr, cfg, strict=False):
    """write `path` in the reference's .vul schema (op.py:3216-3255).  Attributes of `var` that a partial host object does not
    carry are skipped unless `strict`; returns the list of skipped keys."""
    var_save = {'species': list(species), 'nr': int(nr)}
    missing = []
    for key in var_save_keys(cfg):
        if hasattr(var, key):
            var_save[key] = getattr(var, key)
        else:
            missing.append(key)
    if strict and missing:
        raise KeyError("var lacks %s (store.py:83-88)" % missing)
    if getattr(cfg, "save_evolution", False):
        fq = int(getattr(cfg, "save_evo_frq", 1))
        for key in VAR_EVOL_SAVE:
            if hasattr(var, key):
                var_save[key] = np.array(getattr(var, key))[::fq]
    with open(path, 'wb') as f:
        pickle.dump({'variable': var_save, 'atm': _vars(atm), 'parameter': _vars(para)}, f, protocol=4)
    return missing


def load_vul(path):
    """-> the dict a reference .vul holds: keys 'variable', 'atm', 'parameter'"""
    with open(path, 'rb') as f:
        data = pickle.load(f)
    for k in ('variable', 'atm', 'parameter'):
        if k not in data:
            raise ValueError("%s: not a .vul file (no '%s' section)" % (path, k))
    if 'species' not in data['variable'] or 'y' not in data['variable']:
        raise ValueError("%s: 'variable' lacks species / y" % path)
    return data


def ini_from_vul(path, species, nz, y_ini=None):
    """`ini_mix = 'vulcan_ini'` (build_atm.py:166-176): number densities of a previous run mapped onto `species` by name;
    species the previous run did not have keep the values of `y_ini` (zeros by default).  Returns (y_ini, missing species)."""
    data = load_vul(path)
    prev_sp = list(data['variable']['species'])
    prev_y = np.asarray(data['variable']['y'])
    if prev_y.shape[0] != nz:
        raise ValueError("the previous run has %d layers, this one %d (vulcan_cfg.py: the T-P grids have to be the same)" % (prev_y.shape[0], nz))
    y = np.zeros((nz, len(species))) if y_ini is None else np.array(y_ini, dtype=float)
    missing = []
    for i, sp in enumerate(species):
        if sp in prev_sp:
            y[:, i] = prev_y[:, prev_sp.index(sp)]
        else:
            missing.append(sp)
    return y, missing
