"""Shared fixture loading for the parity tests (tests/ is the only place, besides smoke() and bench.py's CPU legs,
that may import oracle/)."""
import json
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(REPO, "tests", "golden")
for p in (REPO, os.path.join(REPO, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

from vulcan_b200.network import Network  # noqa: E402


def have(tag, name):
    return os.path.exists(os.path.join(GOLD, "%s_%s" % (tag, name)))


def load_network(tag):
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
        )


# every (config, step) fixture pair generated from the unmodified reference (oracle/dump_fixtures.py); BASELINE.json configs
CASES = [("HD189", 0), ("HD189", 10), ("HD189", 100), ("HD189", 300), ("Jupiter", 0), ("Jupiter", 30), ("Earth", 0), ("Earth", 30),
         ("HD209S", 0), ("HD209S", 30)]
PHOTO_CASES = [("HD189", 0), ("HD189", 300), ("Jupiter", 0), ("Jupiter", 30), ("Earth", 0), ("Earth", 30), ("HD209S", 0), ("HD209S", 30)]


def case_id(p):
    return "%s-%d" % p


def step_opts(case):
    """The vulcan_cfg / para state consulted inside Ros2.solver (op.py:2896-2970) as plain arrays: shared by the oracle call
    and the C-ABI call (vulcan_b200/ros2.py::_sync_opts builds the same from the live reference objects)."""
    cfg, sp, ni = case.cfg, list(case.net.species), case.ni
    fix_bot = cfg.get("use_fix_sp_bot") or {}
    fbi = [sp.index(s) for s in fix_bot.keys()]
    fbm = np.array([fix_bot[s] for s in fix_bot.keys()], dtype=float)
    dz = None
    if cfg.get("use_condense"):
        dz = np.zeros(ni, dtype=np.uint8)
        for s in list(cfg.get("non_gas_sp", [])) + list(cfg.get("condense_sp", [])):
            dz[sp.index(s)] = 1
    assert not bool(case.fx["fix_species_start"]), "fixtures are taken before fix_species starts"
    return dict(fix_bot_idx=fbi, fix_bot_mix=fbm, n0_bot=float(case.st["n_0"][0]),
                zero_delta_row0=bool(cfg.get("use_botflux") or fix_bot), delta_zero_sp=dz,
                gas_indx_mix=case.gas_indx if cfg.get("non_gas_sp") else None)


def oracle_step(case, oracle, atm, refine=0):
    o = step_opts(case)
    return oracle.ros2_solver(atm, case.y, case.ymix, case.k, case.dt, case.cfg["mtol"], case.cfg["atol"], refine=refine, **o)


def gpu_columns(case, ncol=1, refine=0):
    """a vk_column handle configured like the reference's solver object for this fixture (through the C ABI)."""
    from vulcan_b200 import _abi
    kw = case.atm_kwargs()
    net = _abi.DeviceNetwork(case.net)
    col = _abi.Columns(net, case.nz, ncol)
    col.set_atm(Kzz=kw["Kzz"], vz=kw["vz"], dzi=kw["dzi"], Dzz=kw["Dzz"], vs=kw["vs"], Tco=kw["Tco"], g=kw["g"], M=kw["M"],
                Ti=kw["Ti"], Hpi=kw["Hpi"], ms=kw["ms"], alpha=kw["alpha"], top_flux=kw["top_flux"], bot_flux=kw["bot_flux"],
                bot_vdep=kw["bot_vdep"], use_moldiff=kw["use_moldiff"], use_settling=kw["use_settling"],
                use_topflux=kw["use_topflux"], use_botflux=kw["use_botflux"], gas_indx=kw["gas_indx"],
                gas_indx_lhs=kw["gas_indx_lhs"], shared=True)
    col.set_k(case.k)
    o = step_opts(case)
    fbv = None
    if len(o["fix_bot_idx"]):
        fbv = np.repeat((o["fix_bot_mix"] * o["n0_bot"])[None], ncol, axis=0)
    col.set_step_opts(case.cfg["mtol"], case.cfg["atol"], refine=refine, zero_delta_row0=o["zero_delta_row0"],
                      fix_bot_idx=o["fix_bot_idx"], fix_bot_val=fbv, delta_zero_sp=o["delta_zero_sp"])
    return col


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


def ulp_diff(a, b):
    """max distance in units in the last place between two float64 arrays (same sign assumed where it matters)."""
    a = np.ascontiguousarray(a, dtype=np.float64)
    b = np.ascontiguousarray(b, dtype=np.float64)
    ia = a.view(np.int64).astype(np.float64)
    ib = b.view(np.int64).astype(np.float64)
    same = (np.sign(a) == np.sign(b)) | ((a == 0) & (b == 0))
    d = np.where(same, np.abs(ia - ib), np.inf)
    d = np.where((a == 0) & (b == 0), 0, d)
    return float(d.max())
