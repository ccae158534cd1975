"""Loader of the state snapshots recorded from the unmodified reference (tests/golden/<cfg>_static.npz, <cfg>_stepNNNN.npz, written by
oracle/dump_fixtures.py): the synthetic inputs of bench.py and the inputs of the parity tests.  Plain file reading - nothing here touches
oracle/ or needs a GPU.  (The GPU box has no /root/reference: these files are how the reference's own states travel.)"""
import json
import os

import numpy as np

from .network import Network

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(REPO, "tests", "golden")


def have(tag, name):
    return os.path.exists(os.path.join(GOLD, "%s_%s" % (tag, name)))


# fixture variants that run a base config's network with different cfg switches (oracle/stage_reference.py::CONFIGS)
NETWORK_OF = {"JupiterFix": "Jupiter", "HD189vm": "HD189", "HD189nomol": "HD189", "HD189vz": "HD189", "JupiterVz": "Jupiter", "JupiterVmVz": "Jupiter", "JupiterFixAll": "Jupiter", "JupiterVm": "Jupiter", "EarthVm": "Earth"}   # HD189ion has its own


def load_network(tag):
    tag = NETWORK_OF.get(tag, tag)
    with open(os.path.join(GOLD, tag + "_network.json")) as f:
        return Network.from_json(f.read())


class Case(object):
    """one (config, step) fixture with everything needed to call the oracle / the CUDA path."""

    def __init__(self, tag, step):
        self.tag, self.step = tag, step
        self.st = dict(np.load(os.path.join(GOLD, "%s_static.npz" % tag), allow_pickle=False))   # eager: NpzFile is not thread-safe
        self.fx = dict(np.load(os.path.join(GOLD, "%s_step%04d.npz" % (tag, step)), allow_pickle=False))
        self.cfg = json.loads(str(self.st["cfg_json"]))
        self.net = load_network(tag)
        st, fx = self.st, self.fx
        self.ni, self.nr, self.nz = int(st["ni"]), int(st["nr"]), int(st["nz"])
        k = st["k"].copy()
        k[fx["k_rows_idx"]] = fx["k_rows"]
        self.k_rz = k                          # [nr+1, nz]  (reference dict order)
        self.k = np.ascontiguousarray(k.T)     # [nz, nr+1]  layer-major (product layout)
        self.y, self.ymix, self.dt = fx["y"], fx["ymix"], float(fx["dt"])
        self.gas_indx = st["gas_indx"]

    def atm_kwargs(self):
        st, fx, cfg = self.st, self.fx, self.cfg
        non_gas = bool(cfg.get("non_gas_sp"))
        return dict(
            nz=self.nz, Kzz=st["Kzz"], vz=st["vz"], dzi=fx["dzi"], Dzz=st["Dzz"], vs=fx["vs_dyn"], Tco=st["Tco"],
            g=fx["g"], Ti=fx["Ti"], Hpi=fx["Hpi"], ms=st["ms"], alpha=st["alpha"], top_flux=fx["top_flux_dyn"],
            bot_flux=st["bot_flux"], bot_vdep=st["bot_vdep"], M=st["M"],
            use_moldiff=bool(cfg["use_moldiff"]), use_settling=bool(cfg["use_settling"]),
            use_topflux=bool(cfg["use_topflux"]), use_botflux=bool(cfg["use_botflux"]),
            gas_indx=self.gas_indx if non_gas else None,
            # lhs_jac_tot keys the gas mask on use_condense (op.py:1981), the other variants on non_gas_sp
            gas_indx_lhs=self.gas_indx if (bool(cfg["use_condense"]) if (cfg["use_moldiff"] and not cfg["use_settling"]) else non_gas) else None,
            # use_vm_mol: the *_vm stencils (op.py:1599-1694, 1794-1898, 2044-2119, 2366-2444); diff_esc only enters their lhs
            use_vm_mol=bool(cfg.get("use_vm_mol", False)), vm=st["vm"],
            diff_esc_idx=[list(self.net.species).index(s) for s in cfg.get("diff_esc", [])],
        )


def photo_tables(st):
    """absorber / branch tables of a <cfg>_static.npz in the argument order of Oracle.compute_* (T-dependent cross sections,
    op.py:2588-2593, 2767-2773, are stored only for the species in T_cross_sp)."""
    psp = [str(x) for x in st["photo_sp"]]
    tsp = [str(x) for x in st["T_cross_sp"]] if "T_cross_sp" in st else []
    nz, nbin = int(st["nz"]), int(st["nbin"])
    abs_is_T = np.array([s in tsp for s in psp], dtype=np.uint8)
    cross = st["cross"].copy()
    cross_T = None
    if abs_is_T.any():
        cross_T = np.zeros((len(psp), nz, nbin))
        for q, s in enumerate(tsp):
            cross_T[psp.index(s)] = st["cross_T"][q]
    br_is_T = np.array([psp[b] in tsp for b in st["branch_sp"]], dtype=np.uint8)
    cross_J_T = None
    if br_is_T.any():
        cross_J_T = np.zeros((len(br_is_T), nz, nbin))
        for q, b in enumerate(st["cross_J_T_branch"]):
            cross_J_T[int(b)] = st["cross_J_T"][q]
    return dict(abs_is_T=abs_is_T if abs_is_T.any() else None, cross=cross, cross_T=cross_T,
                br_is_T=br_is_T if br_is_T.any() else None, cross_J_T=cross_J_T)


def steady_ensemble_from_fixture(case, y, atom_ini, kzz_scale, photo=True, **kw):
    """vulcan_b200.ensemble.SteadyEnsemble for columns derived from one fixture config (its atmosphere, rates, star): the inputs the
    reference would build per column, assembled from <cfg>_static.npz / <cfg>_step0000.npz"""
    from . import ensemble
    st, fx, cfg = case.st, case.fx, case.cfg
    akw = case.atm_kwargs()
    kzz = np.asarray(kzz_scale)[:, None] * np.asarray(akw["Kzz"])[None, :]
    grid = dict(pico=st["pico"], ms=st["ms"], zco=fx["zco"], Hp=fx["Hp"], dz=fx["dz"], pref_indx=int(st["pref_indx"]), gs=float(cfg["gs"]))
    ph = None
    if photo and cfg.get("use_photo"):
        pt = photo_tables(st)
        ph = dict(bins=st["bins"], sflux_top=st["sflux_top"], i12=int(st["sflux_din12_indx"]), dbin1=float(st["dbin1"]), dbin2=float(st["dbin2"]),
                  sl_angle=cfg["sl_angle"], edd=cfg["edd"], flux_atol=cfg["flux_atol"], f_diurnal=cfg["f_diurnal"], abs_idx=st["photo_sp_idx"],
                  cross_abs=pt["cross"], photo_idx=st["photo_sp_idx"], cross_photo=pt["cross"], scat_idx=st["scat_sp_idx"],
                  cross_scat=st["cross_scat"], cross_J=st["cross_J"], br_rate_index=st["branch_rate_index"], abs_is_T=pt["abs_is_T"],
                  cross_abs_T=pt["cross_T"], br_is_T=pt["br_is_T"], cross_J_T=pt["cross_J_T"])
    sp = list(case.net.species)
    return ensemble.SteadyEnsemble(case.net, case.nz, y, np.full(y.shape[0], float(cfg["dttry"])), akw, kzz, case.k, cfg, st["compo"], atom_ini,
                                   st["n_0"], grid, photo=ph, diff_esc_idx=[sp.index(s) for s in cfg.get("diff_esc", []) or []], **kw)
