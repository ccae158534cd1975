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
