"""TEST INFRASTRUCTURE — ctypes front end of oracle/vk_oracle.c (the CPU restatement of the Ros2 path).

Imported only by tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference).
The product package vulcan_b200/ never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_build", "libvk_oracle.so")
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_bp = C.POINTER(C.c_ubyte)


def build(force=False):
    src = os.path.join(HERE, "vk_oracle.c")
    if not force and os.path.exists(LIB_PATH) and os.path.getmtime(LIB_PATH) >= os.path.getmtime(src):
        return LIB_PATH
    os.makedirs(os.path.dirname(LIB_PATH), exist_ok=True)
    subprocess.run(["gcc", "-O2", "-ffp-contract=off", "-fno-fast-math", "-shared", "-fPIC", "-o", LIB_PATH, src, "-lm"],
                   check=True)
    return LIB_PATH


class _Net(C.Structure):
    _fields_ = [("ni", C.c_int), ("nr", C.c_int), ("maxf", C.c_int), ("maxjf", C.c_int), ("n_ent", C.c_int),
                ("n_term", C.c_int), ("rate_fac", _ip), ("rate_pow", _ip), ("rhs_ptr", _ip), ("rhs_pair", _ip),
                ("rhs_coef", _dp), ("jac_ptr", _ip), ("jac_row", _ip), ("jac_col", _ip), ("jac_k", _ip),
                ("jac_coef", _dp), ("jac_fac", _ip)]


class _Atm(C.Structure):
    _fields_ = [("nz", C.c_int), ("ni", C.c_int), ("use_moldiff", C.c_int), ("use_settling", C.c_int),
                ("use_topflux", C.c_int), ("use_botflux", C.c_int), ("n_gas", C.c_int), ("gas_indx", _ip),
                ("n_gas_lhs", C.c_int), ("gas_indx_lhs", _ip), ("Kzz", _dp), ("vz", _dp), ("dzi", _dp), ("Dzz", _dp),
                ("vs", _dp), ("Tco", _dp), ("g", _dp), ("Ti", _dp), ("Hpi", _dp), ("ms", _dp), ("alpha", _dp),
                ("top_flux", _dp), ("bot_flux", _dp), ("bot_vdep", _dp), ("M", _dp), ("use_vm_mol", C.c_int), ("vm", _dp),
                ("n_diff_esc", C.c_int), ("diff_esc_idx", _ip)]


class _Opts(C.Structure):
    _fields_ = [("fix_mask", _bp), ("fix_y", _dp), ("n_fix_bot", C.c_int), ("fix_bot_idx", _ip), ("fix_bot_mix", _dp),
                ("n0_bot", C.c_double), ("zero_delta_row0", C.c_int), ("delta_zero_sp", _bp), ("n_gas_mix", C.c_int),
                ("gas_indx_mix", _ip), ("mtol", C.c_double), ("atol", C.c_double), ("refine", C.c_int),
                ("na", C.c_int), ("compo", _dp), ("refine_dt_min", C.c_double)]


def _d(a):
    return a.ctypes.data_as(_dp)


def _i(a):
    return a.ctypes.data_as(_ip)


def _b(a):
    return a.ctypes.data_as(_bp)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


class Oracle(object):
    """CPU oracle bound to one compiled network (vulcan_b200.network.Network)."""

    def __init__(self, network):
        self.lib = C.CDLL(build())
        self.lib.vko_np_sum.restype = C.c_double
        self.lib.vko_np_sum.argtypes = [_dp, C.c_long]
        self.net = network
        t = network.tables()
        self._keep = {k: (np.ascontiguousarray(v) if isinstance(v, np.ndarray) else v) for k, v in t.items()}
        k = self._keep
        self.cnet = _Net(t["ni"], t["nr"], t["maxf"], t["maxjf"], len(k["jac_row"]), len(k["jac_k"]),
                         _i(k["rate_fac"]), _i(k["rate_pow"]), _i(k["rhs_ptr"]), _i(k["rhs_pair"]), _d(k["rhs_coef"]),
                         _i(k["jac_ptr"]), _i(k["jac_row"]), _i(k["jac_col"]), _i(k["jac_k"]), _d(k["jac_coef"]),
                         _i(k["jac_fac"]))
        self.ni, self.nr = t["ni"], t["nr"]

    # -------------------------------------------------------------- helpers
    def make_atm(self, nz, Kzz, vz, dzi, Dzz, vs, Tco, g, Ti, Hpi, ms, alpha, top_flux, bot_flux, bot_vdep, M,
                 use_moldiff=True, use_settling=False, use_topflux=False, use_botflux=False, gas_indx=None,
                 gas_indx_lhs=None, use_vm_mol=False, vm=None, diff_esc_idx=None):
        ni = self.ni
        arrs = dict(Kzz=_f64(Kzz), vz=_f64(vz), dzi=_f64(dzi), Dzz=_f64(Dzz), vs=_f64(vs), Tco=_f64(Tco), g=_f64(g),
                    Ti=_f64(Ti), Hpi=_f64(Hpi), ms=_f64(ms), alpha=_f64(alpha), top_flux=_f64(top_flux),
                    bot_flux=_f64(bot_flux), bot_vdep=_f64(bot_vdep), M=_f64(M))
        gi = _i32(gas_indx) if gas_indx is not None and len(gas_indx) != ni else None
        gl = _i32(gas_indx_lhs) if gas_indx_lhs is not None and len(gas_indx_lhs) != ni else None
        arrs["gi"], arrs["gl"] = gi, gl
        arrs["vm"] = _f64(vm) if use_vm_mol else None
        de = _i32(diff_esc_idx) if (use_vm_mol and diff_esc_idx is not None and len(diff_esc_idx)) else None
        arrs["de"] = de
        a = _Atm(nz, ni, int(use_moldiff), int(use_settling), int(use_topflux), int(use_botflux),
                 0 if gi is None else len(gi), None if gi is None else _i(gi),
                 0 if gl is None else len(gl), None if gl is None else _i(gl),
                 _d(arrs["Kzz"]), _d(arrs["vz"]), _d(arrs["dzi"]), _d(arrs["Dzz"]), _d(arrs["vs"]), _d(arrs["Tco"]),
                 _d(arrs["g"]), _d(arrs["Ti"]), _d(arrs["Hpi"]), _d(arrs["ms"]), _d(arrs["alpha"]),
                 _d(arrs["top_flux"]), _d(arrs["bot_flux"]), _d(arrs["bot_vdep"]), _d(arrs["M"]),
                 int(bool(use_vm_mol)), None if arrs["vm"] is None else _d(arrs["vm"]),
                 0 if de is None else len(de), None if de is None else _i(de))
        a._keep = arrs
        return a

    def np_sum(self, a):
        a = _f64(a).ravel()
        return self.lib.vko_np_sum(_d(a), a.size)

    # -------------------------------------------------------------- components
    def chemdf(self, y, M, k):
        """y [nz,ni], M [nz], k [nz,nr+1] (layer-major)."""
        y, M, k = _f64(y), _f64(M), _f64(k)
        out = np.empty_like(y)
        self.lib.vko_chemdf(C.byref(self.cnet), y.shape[0], _d(y), _d(M), _d(k), _d(out))
        return out

    def chemjac(self, y, M, k):
        y, M, k = _f64(y), _f64(M), _f64(k)
        nz = y.shape[0]
        J = np.empty((nz, self.ni, self.ni))
        self.lib.vko_chemjac(C.byref(self.cnet), nz, _d(y), _d(M), _d(k), _d(J))
        return J

    def diffdf(self, atm, y):
        y = _f64(y)
        out = np.empty_like(y)
        self.lib.vko_diffdf(C.byref(atm), _d(y), _d(out))
        return out

    def lhs(self, atm, y, k, dt):
        y, k = _f64(y), _f64(k)
        nz = y.shape[0]
        D = np.empty((nz, self.ni, self.ni))
        up = np.empty((nz, self.ni))
        dn = np.empty((nz, self.ni))
        self.lib.vko_lhs(C.byref(self.cnet), C.byref(atm), _d(y), _d(k), C.c_double(dt), _d(D), _d(up), _d(dn))
        return D, up, dn

    def blocktri_factor(self, D, up, dn):
        D, up, dn = _f64(D), _f64(up), _f64(dn)
        nz, n = up.shape
        W = np.empty_like(D)
        rc = self.lib.vko_blocktri_factor_f64(nz, n, _d(D), _d(up), _d(dn), _d(W))
        if rc:
            raise FloatingPointError("singular block (code %d)" % rc)
        return W

    def blocktri_solve(self, W, up, dn, r):
        W, up, dn, r = _f64(W), _f64(up), _f64(dn), _f64(r)
        nz, n = up.shape
        x = np.empty_like(r)
        self.lib.vko_blocktri_solve_f64(nz, n, _d(W), _d(up), _d(dn), _d(r), _d(x))
        return x

    def blocktri_matvec(self, D, up, dn, x):
        D, up, dn, x = _f64(D), _f64(up), _f64(dn), _f64(x)
        nz, n = up.shape
        out = np.empty_like(x)
        self.lib.vko_blocktri_matvec_f64(nz, n, _d(D), _d(up), _d(dn), _d(x), _d(out))
        return out

    def blocktri_truth(self, D, up, dn, r, n_refine=3):
        """80-bit factorisation + 80-bit iterative refinement, rounded to double."""
        D, up, dn, r = _f64(D), _f64(up), _f64(dn), _f64(r)
        nz, n = up.shape
        x = np.empty_like(r)
        rc = self.lib.vko_blocktri_solve_truth(nz, n, _d(D), _d(up), _d(dn), _d(r), _d(x), n_refine)
        if rc:
            raise FloatingPointError("singular block (code %d)" % rc)
        return x

    # -------------------------------------------------------------- the step
    def ros2_solver(self, atm, y, ymix, k, dt, mtol, atol, refine=0, fix_mask=None, fix_y=None, fix_bot_idx=(),
                    fix_bot_mix=(), n0_bot=0.0, zero_delta_row0=False, delta_zero_sp=None, gas_indx_mix=None, compo=None,
                    refine_dt_min=1.0e3):
        y, ymix, k = _f64(y), _f64(ymix), _f64(k)
        nz, ni = y.shape
        fm = None if fix_mask is None else np.ascontiguousarray(fix_mask, dtype=np.uint8)
        fy = None if fix_y is None else _f64(fix_y)
        fbi, fbm = _i32(fix_bot_idx), _f64(fix_bot_mix)
        dz = None if delta_zero_sp is None else np.ascontiguousarray(delta_zero_sp, dtype=np.uint8)
        gm = _i32(gas_indx_mix) if gas_indx_mix is not None and len(gas_indx_mix) != ni else None
        cp = None if compo is None else _f64(compo)
        o = _Opts(None if fm is None else _b(fm), None if fy is None else _d(fy), len(fbi), _i(fbi), _d(fbm),
                  float(n0_bot), int(zero_delta_row0), None if dz is None else _b(dz),
                  0 if gm is None else len(gm), None if gm is None else _i(gm), float(mtol), float(atol), int(refine),
                  0 if cp is None else cp.shape[1], None if cp is None else _d(cp), float(refine_dt_min))
        sol, ymo, k1, k2 = np.empty_like(y), np.empty_like(y), np.empty_like(y), np.empty_like(y)
        delta = C.c_double(0)
        rc = self.lib.vko_ros2_solver(C.byref(self.cnet), C.byref(atm), C.byref(o), _d(y), _d(ymix), _d(k),
                                      C.c_double(dt), _d(sol), _d(ymo), C.byref(delta), _d(k1), _d(k2))
        if rc:
            raise FloatingPointError("singular block (code %d)" % rc)
        return dict(sol=sol, ymix=ymo, delta=delta.value, k1=k1, k2=k2)

    def clip_loss(self, y, ymix_in, compo, pos_cut, nega_cut, mtol, gas_indx=None, atom_skip=None):
        y = _f64(y).copy()
        ymix_in, compo = _f64(ymix_in), _f64(compo)
        nz, ni = y.shape
        na = compo.shape[1]
        ymo = np.empty_like(y)
        asum = np.zeros(na)
        small, nega = C.c_double(0), C.c_double(0)
        gi = _i32(gas_indx) if gas_indx is not None and len(gas_indx) != ni else None
        sk = None if atom_skip is None else np.ascontiguousarray(atom_skip, dtype=np.uint8)
        self.lib.vko_clip_loss(nz, ni, na, _d(y), _d(ymix_in), _d(ymo), _d(compo), None if sk is None else _b(sk),
                               _d(asum), C.c_double(pos_cut), C.c_double(nega_cut), C.c_double(mtol),
                               0 if gi is None else len(gi), None if gi is None else _i(gi), C.byref(small),
                               C.byref(nega))
        return dict(y=y, ymix=ymo, atom_sum=asum, small_y=small.value, nega_y=nega.value)

    # -------------------------------------------------------------- photolysis
    def compute_tau(self, y, dz, abs_idx, cross, scat_idx, cross_scat, abs_is_T=None, cross_T=None):
        y, dz, cross, cross_scat = _f64(y), _f64(dz), _f64(cross), _f64(cross_scat)
        nz, ni = y.shape
        nbin = cross.shape[1]
        ai, si = _i32(abs_idx), _i32(scat_idx)
        tau = np.empty((nz + 1, nbin))
        isT = None if abs_is_T is None else np.ascontiguousarray(abs_is_T, dtype=np.uint8)
        cT = None if cross_T is None else _f64(cross_T)
        self.lib.vko_compute_tau(nz, ni, nbin, _d(y), _d(dz), len(ai), _i(ai), _d(cross),
                                 None if isT is None else _b(isT), None if cT is None else _d(cT), len(si), _i(si),
                                 _d(cross_scat), _d(tau))
        return tau

    def compute_flux(self, ymix, tau, sflux_top, bins, photo_idx, cross, scat_idx, cross_scat, sl_angle, edd, flux_atol,
                     dflux_u, dflux_d, aflux):
        """returns dict; dflux_u/dflux_d/aflux are the PREVIOUS call's state (not modified)."""
        ymix, tau, sflux_top, bins, cross, cross_scat = map(_f64, (ymix, tau, sflux_top, bins, cross, cross_scat))
        nz, ni = ymix.shape
        nbin = bins.size
        pi, si = _i32(photo_idx), _i32(scat_idx)
        du, dd, af = _f64(dflux_u).copy(), _f64(dflux_d).copy(), _f64(aflux).copy()
        sflux = np.empty((nz + 1, nbin))
        prev = np.empty((nz, nbin))
        ch = C.c_double(0)
        self.lib.vko_compute_flux(nz, ni, nbin, _d(ymix), _d(tau), _d(sflux_top), _d(bins), len(pi), _i(pi), _d(cross),
                                  len(si), _i(si), _d(cross_scat), C.c_double(sl_angle), C.c_double(edd),
                                  C.c_double(flux_atol), _d(sflux), _d(du), _d(dd), _d(af), _d(prev), C.byref(ch))
        return dict(sflux=sflux, dflux_u=du, dflux_d=dd, aflux=af, prev_aflux=prev, aflux_change=ch.value)

    def compute_J(self, aflux, sigma, i12, dbin1, dbin2, br_is_T=None, sigma_T=None):
        aflux, sigma = _f64(aflux), _f64(sigma)
        nz, nbin = aflux.shape
        nbr = sigma.shape[0]
        J = np.empty((nbr, nz))
        isT = None if br_is_T is None else np.ascontiguousarray(br_is_T, dtype=np.uint8)
        sT = None if sigma_T is None else _f64(sigma_T)
        self.lib.vko_compute_J(nz, nbin, int(i12), C.c_double(dbin1), C.c_double(dbin2), _d(aflux), nbr, _d(sigma),
                               None if isT is None else _b(isT), None if sT is None else _d(sT), _d(J))
        return J
