"""TEST INFRASTRUCTURE - a CPU stand-in for `vulcan_b200._abi.Columns` backed by the oracle (oracle/vk_oracle.c), used ONLY by
the `-m "not gpu"` host-logic tests: it lets the reference-facing host code (vulcan_b200/ros2.py, tests/integration_mirror.py)
run its whole per-step protocol on a box without a GPU, so that the control flow the GPU lock-step tests exercise
(tests/test_gpu_lockstep.py) is already checked against the reference's trajectory here.  The product never imports this
module and has no CPU path (tests/test_abi.py)."""
from types import SimpleNamespace

import numpy as np

from oracle import Oracle


class OracleNetwork(object):
    def __init__(self, network, device=0):
        self.network = network
        self.oracle = Oracle(network)
        self.ni, self.nr = network.ni, network.nr

    def close(self):
        pass


class OracleColumns(object):
    """same method names / argument meaning / return shapes as _abi.Columns for ncol = 1."""

    def __init__(self, devnet, nz, ncol=1):
        assert ncol == 1
        self.o, self.nz, self.ncol, self.ni, self.nr = devnet.oracle, nz, 1, devnet.ni, devnet.nr
        self.atm = self.k = self.opts = None
        self.photo = None

    def close(self):
        pass

    def set_atm(self, shared=True, **kw):
        self.gas_indx = kw.get("gas_indx")
        self.atm = self.o.make_atm(self.nz, **kw)

    def set_k(self, k, shared=None):
        self.k = np.array(k, dtype=np.float64).reshape(self.nz, self.nr + 1)

    def set_step_opts(self, mtol, atol, refine=0, zero_delta_row0=False, fix_bot_idx=(), fix_bot_val=None, delta_zero_sp=None,
                      fix_mask=None, fix_y=None, compo=None, refine_dt_min=1.0e3, rhs_order=0):
        self.opts = dict(mtol=mtol, atol=atol, refine=refine, zero_delta_row0=zero_delta_row0, fix_bot_idx=list(fix_bot_idx),
                         fix_bot_val=None if fix_bot_val is None else np.asarray(fix_bot_val, dtype=float).ravel(),
                         delta_zero_sp=delta_zero_sp, fix_mask=fix_mask, fix_y=fix_y, compo=compo, refine_dt_min=refine_dt_min)

    def ros2_solve(self, y, ymix, dt):
        o = self.opts
        fbv = o["fix_bot_val"]
        res = self.o.ros2_solver(self.atm, np.asarray(y).reshape(self.nz, self.ni), np.asarray(ymix).reshape(self.nz, self.ni), self.k,
                                 float(np.ravel(dt)[0]), o["mtol"], o["atol"], refine=o["refine"], fix_mask=o["fix_mask"], fix_y=o["fix_y"],
                                 fix_bot_idx=o["fix_bot_idx"], fix_bot_mix=fbv if fbv is not None else (), n0_bot=1.0,
                                 zero_delta_row0=o["zero_delta_row0"], delta_zero_sp=o["delta_zero_sp"], gas_indx_mix=self.gas_indx,
                                 compo=o["compo"], refine_dt_min=o["refine_dt_min"])
        return res["sol"][None], res["ymix"][None], np.array([res["delta"]]), np.zeros(1, dtype=np.int32)

    def clip_loss(self, y, ymix_in, compo, pos_cut, nega_cut, atom_sum=None, small_y=None, nega_y=None, atom_skip=None, mtol=None):
        res = self.o.clip_loss(np.asarray(y).reshape(self.nz, self.ni), np.asarray(ymix_in).reshape(self.nz, self.ni), compo, pos_cut,
                               nega_cut, mtol if mtol is not None else (self.opts["mtol"] if self.opts else 0.0), gas_indx=self.gas_indx, atom_skip=atom_skip)
        asum = res["atom_sum"]
        if atom_skip is not None and atom_sum is not None:
            asum = np.where(np.asarray(atom_skip, dtype=bool), np.ravel(atom_sum), asum)
        sm = res["small_y"] + (0.0 if small_y is None else float(np.ravel(small_y)[0]))
        ng = res["nega_y"] + (0.0 if nega_y is None else float(np.ravel(nega_y)[0]))
        return dict(y=res["y"][None], ymix=res["ymix"][None], atom_sum=asum[None], small_y=np.array([sm]), nega_y=np.array([ng]),
                    any_negative=np.array([int(np.any(res["y"] < 0))], dtype=np.int32))

    # ---- photolysis (compute_tau / compute_flux / compute_J, op.py:2580-2786)
    def photo_setup(self, bins, sflux_top, i12, dbin1, dbin2, sl_angle, edd, flux_atol, f_diurnal, abs_idx, cross_abs, photo_idx,
                    cross_photo, scat_idx, cross_scat, cross_J, br_rate_index, abs_is_T=None, cross_abs_T=None, br_is_T=None,
                    cross_J_T=None):
        nbin = len(bins)
        self.photo = SimpleNamespace(**{k: v for k, v in locals().items() if k != "self"})
        self.photo.dflux_u = np.zeros((self.nz + 1, nbin))
        self.photo.dflux_d = np.zeros((self.nz + 1, nbin))
        self.photo.aflux = np.zeros((self.nz, nbin))

    def photo_update(self, y, ymix, dz):
        p, o = self.photo, self.o
        y, ymix = np.asarray(y).reshape(self.nz, self.ni), np.asarray(ymix).reshape(self.nz, self.ni)
        p.tau = o.compute_tau(y, np.ravel(dz), p.abs_idx, p.cross_abs, p.scat_idx, p.cross_scat, p.abs_is_T, p.cross_abs_T)
        fl = o.compute_flux(ymix, p.tau, p.sflux_top, p.bins, p.photo_idx, p.cross_photo, p.scat_idx, p.cross_scat, p.sl_angle, p.edd,
                            p.flux_atol, p.dflux_u, p.dflux_d, p.aflux)
        p.sflux, p.dflux_u, p.dflux_d, p.aflux = fl["sflux"], fl["dflux_u"], fl["dflux_d"], fl["aflux"]
        J = o.compute_J(p.aflux, p.cross_J, p.i12, p.dbin1, p.dbin2, p.br_is_T, p.cross_J_T)
        for q, rid in enumerate(p.br_rate_index):          # the device path writes J f_diurnal into its copy of k
            if rid > 0 and self.k is not None:
                self.k[:, rid] = J[q] * p.f_diurnal
        return J[None], np.array([fl["aflux_change"]])

    def photo_read(self, names=("tau", "sflux", "dflux_u", "dflux_d", "aflux")):
        return {n: getattr(self.photo, n)[None] for n in names}


def oracle_backed_abi():
    """namespace with the two names vulcan_b200/ros2.py takes from _abi"""
    return SimpleNamespace(DeviceNetwork=OracleNetwork, Columns=OracleColumns)
