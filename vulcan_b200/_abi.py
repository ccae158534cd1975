"""ctypes binding of include/vulcan_b200.h (libvulcan_b200.so, hand-written sm_100a CUDA).

There is NO CPU fallback: if the shared library is missing or no CUDA device is visible, every compute
entry point raises.  (Loading the library and listing its symbols works without a GPU.)
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("VK_LIB_PATH") or os.path.join(HERE, "_lib", "libvulcan_b200.so")     # VK_LIB_PATH: diagnostic build variants

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_bp = C.POINTER(C.c_ubyte)
_vp = C.c_void_p

VK_OK, VK_ERR_INVALID, VK_ERR_CUDA, VK_ERR_UNSUPPORTED, VK_ERR_SINGULAR = 0, -1, -2, -3, -4

# every symbol include/vulcan_b200.h declares (tests/test_abi.py checks the header against this list and the .so)
SYMBOLS = [
    "vk_abi_version", "vk_last_error", "vk_device_count", "vk_network_create", "vk_network_destroy",
    "vk_column_create", "vk_column_destroy", "vk_set_atm", "vk_set_k", "vk_set_k_rows", "vk_set_step_opts",
    "vk_ros2_solve", "vk_clip_loss", "vk_eval_rhs", "vk_eval_lhs", "vk_blocktri_solve", "vk_photo_setup",
    "vk_photo_update", "vk_photo_read", "vk_photo_reset", "vk_ens_setup", "vk_ens_set_state", "vk_ens_run",
    "vk_ens_get_state", "vk_last_kernel_ms", "vk_device_buffers", "vk_stream", "vk_rates_set", "vk_compute_k", "vk_get_k",
    "vk_debug_time_kernel", "vk_refine_stats", "vk_ens_setup_steady", "vk_ens_photo_update", "vk_ens_run_steady", "vk_ens_get_steady", "vk_ens_get_fix", "vk_conden_setup", "vk_conden_apply",
]


class VulcanB200Error(RuntimeError):
    def __init__(self, code, msg):
        RuntimeError.__init__(self, "vulcan_b200 error %d: %s" % (code, msg))
        self.code = code


class RateDesc(C.Structure):
    _fields_ = [("npair", C.c_int), ("kind", C.POINTER(C.c_int)), ("arrhenius", C.POINTER(C.c_double)), ("cap_kind", C.POINTER(C.c_int)),
                ("cap", C.POINTER(C.c_double)), ("reverse", C.POINTER(C.c_ubyte)), ("removed", C.POINTER(C.c_ubyte)),
                ("gibbs_ptr", C.POINTER(C.c_int)), ("gibbs_sp", C.POINTER(C.c_int)), ("gibbs_nu", C.POINTER(C.c_int)),
                ("dnu", C.POINTER(C.c_int)), ("nasa9", C.POINTER(C.c_double))]


class NetworkDesc(C.Structure):
    _fields_ = [("ni", C.c_int), ("nr", C.c_int), ("maxf", C.c_int), ("maxjf", C.c_int), ("n_rhs", C.c_int),
                ("n_ent", C.c_int), ("n_term", C.c_int), ("rate_fac", _ip), ("rate_pow", _ip), ("rhs_ptr", _ip),
                ("rhs_pair", _ip), ("rhs_coef", _dp), ("jac_ptr", _ip), ("jac_row", _ip), ("jac_col", _ip),
                ("jac_k", _ip), ("jac_coef", _dp), ("jac_fac", _ip)]


class AtmView(C.Structure):
    _fields_ = [("shared", C.c_int), ("use_moldiff", C.c_int), ("use_settling", C.c_int), ("use_topflux", C.c_int),
                ("use_botflux", C.c_int), ("n_gas", C.c_int), ("gas_indx", _ip), ("n_gas_lhs", C.c_int),
                ("gas_indx_lhs", _ip), ("Kzz", _dp), ("vz", _dp), ("dzi", _dp), ("Dzz", _dp), ("vs", _dp),
                ("Tco", _dp), ("g", _dp), ("M", _dp), ("Ti", _dp), ("Hpi", _dp), ("ms", _dp), ("alpha", _dp),
                ("top_flux", _dp), ("bot_flux", _dp), ("bot_vdep", _dp),
                ("use_vm_mol", C.c_int), ("vm", _dp), ("n_diff_esc", C.c_int), ("diff_esc_idx", _ip)]


class StepOpts(C.Structure):
    _fields_ = [("mtol", C.c_double), ("atol", C.c_double), ("refine", C.c_int), ("zero_delta_row0", C.c_int),
                ("n_fix_bot", C.c_int), ("fix_bot_idx", _ip), ("fix_bot_val", _dp), ("delta_zero_sp", _bp),
                ("fix_mask", _bp), ("fix_y", _dp), ("na", C.c_int), ("compo", _dp), ("refine_dt_min", C.c_double),
                ("rhs_order", C.c_int)]


class PhotoView(C.Structure):
    _fields_ = [("nbin", C.c_int), ("i12", C.c_int), ("dbin1", C.c_double), ("dbin2", C.c_double),
                ("sl_angle", C.c_double), ("edd", C.c_double), ("flux_atol", C.c_double), ("f_diurnal", C.c_double),
                ("bins", _dp), ("sflux_top", _dp),
                ("n_abs", C.c_int), ("abs_idx", _ip), ("cross_abs", _dp), ("abs_is_T", _bp), ("cross_abs_T", _dp),
                ("n_photo", C.c_int), ("photo_idx", _ip), ("cross_photo", _dp),
                ("n_scat", C.c_int), ("scat_idx", _ip), ("cross_scat", _dp),
                ("n_br", C.c_int), ("cross_J", _dp), ("br_rate_index", _ip), ("br_is_T", _bp), ("cross_J_T", _dp)]


class EnsOpts(C.Structure):
    _fields_ = [("rtol", C.c_double), ("loss_eps", C.c_double), ("dt_min", C.c_double), ("dt_max", C.c_double),
                ("dt_var_min", C.c_double), ("dt_var_max", C.c_double), ("pos_cut", C.c_double), ("nega_cut", C.c_double),
                ("na", C.c_int), ("compo", _dp), ("atom_ini", _dp), ("n_0", _dp)]


# refine = -1 (auto): columns are refined from this step size on.  Below it the element-budget error of one backward-stable solve,
# eps |A||x| r dt, is < 1e-9 per step on every fixture
REFINE_DT_MIN = 1.0e3

class SteadyOpts(C.Structure):
    _fields_ = [("st_factor", C.c_double), ("mtol_conv", C.c_double), ("atol", C.c_double), ("yconv_cri", C.c_double),
                ("slope_cri", C.c_double), ("yconv_min", C.c_double), ("flux_cri", C.c_double), ("trun_min", C.c_double),
                ("runtime", C.c_double), ("conv_step", C.c_int), ("count_min", C.c_int), ("count_max", C.c_int),
                ("conv_ignore_sp", _bp), ("use_photo", C.c_int), ("ini_update_photo_frq", C.c_int), ("final_update_photo_frq", C.c_int),
                ("update_frq", C.c_int), ("pref_indx", C.c_int), ("gs", C.c_double), ("Rp", C.c_double), ("max_flux", C.c_double),
                ("pico", _dp), ("ms", _dp), ("zco", _dp), ("Hp", _dp), ("dz", _dp), ("n_diff_esc", C.c_int), ("diff_esc_idx", _ip),
                ("hist_cap", C.c_int), ("hist_stride", C.c_int),
                ("use_condense", C.c_int), ("fix_species_switch", C.c_int), ("fix_from_coldtrap", C.c_int), ("n_fix", C.c_int),
                ("fix_sp", _ip), ("fix_whole_column", _bp), ("fix_sat_mix", _dp),
                ("start_conden_time", C.c_double), ("stop_conden_time", C.c_double), ("post_conden_rtol", C.c_double)]


class CondenDesc(C.Structure):
    _fields_ = [("n_re", C.c_int), ("re_idx", _ip), ("gas_idx", _ip), ("m", _dp), ("rho_p", _dp), ("r_p", _dp), ("sat", _dp), ("zero_rate", _bp),
                ("n_relax", C.c_int), ("relax_kind", _ip), ("relax_gas", _ip), ("relax_ice", _ip), ("relax_top", _ip), ("relax_m", _dp),
                ("relax_rho", _dp), ("relax_r", _dp), ("relax_sat", _dp), ("start_conden_time", C.c_double), ("stop_conden_time", C.c_double),
                ("post_conden_rtol", C.c_double)]


_lib = None


def load():
    """dlopen the in-tree library (built by vulcan_b200/build.py / __graft_entry__.build())."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise VulcanB200Error(VK_ERR_CUDA, "%s not built: run `python -m vulcan_b200.build` (no CPU fallback exists)" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    lib.vk_last_error.restype = C.c_char_p
    for name in SYMBOLS:
        getattr(lib, name)
    lib.vk_network_create.argtypes = [C.POINTER(NetworkDesc), C.c_int, C.POINTER(_vp)]
    lib.vk_network_destroy.argtypes = [_vp]
    lib.vk_network_destroy.restype = None
    lib.vk_column_create.argtypes = [_vp, C.c_int, C.c_int, C.POINTER(_vp)]
    lib.vk_column_destroy.argtypes = [_vp]
    lib.vk_column_destroy.restype = None
    lib.vk_set_atm.argtypes = [_vp, C.POINTER(AtmView)]
    lib.vk_set_k.argtypes = [_vp, _dp, C.c_int]
    lib.vk_set_k_rows.argtypes = [_vp, C.c_int, _ip, _dp]
    lib.vk_set_step_opts.argtypes = [_vp, C.POINTER(StepOpts)]
    lib.vk_ros2_solve.argtypes = [_vp, _dp, _dp, _dp, _dp, _dp, _dp, _ip]
    lib.vk_clip_loss.argtypes = [_vp, _dp, _dp, _dp, C.c_int, _dp, _bp, C.c_double, C.c_double, C.c_double, _dp, _dp, _dp, _ip]
    lib.vk_eval_rhs.argtypes = [_vp, _dp, _dp, _dp]
    lib.vk_eval_lhs.argtypes = [_vp, _dp, _dp, _dp, _dp, _dp]
    lib.vk_blocktri_solve.argtypes = [_vp, _dp, _dp, _dp, _dp, _dp, C.c_int, _ip]
    lib.vk_photo_setup.argtypes = [_vp, C.POINTER(PhotoView)]
    lib.vk_photo_update.argtypes = [_vp, _dp, _dp, _dp, _dp, _dp]
    lib.vk_photo_read.argtypes = [_vp, _dp, _dp, _dp, _dp, _dp]
    lib.vk_photo_reset.argtypes = [_vp]
    lib.vk_rates_set.argtypes = [_vp, C.POINTER(RateDesc)]
    lib.vk_compute_k.argtypes = [_vp, _dp, _dp, C.c_int]
    lib.vk_get_k.argtypes = [_vp, _dp]
    lib.vk_ens_setup.argtypes = [_vp, C.POINTER(EnsOpts)]
    lib.vk_ens_set_state.argtypes = [_vp, _dp, _dp]
    lib.vk_ens_run.argtypes = [_vp, C.c_int]
    lib.vk_ens_get_state.argtypes = [_vp, _dp, _dp, _dp, _ip, _ip]
    lib.vk_last_kernel_ms.argtypes = [_vp, C.POINTER(C.c_float), C.POINTER(C.c_float)]
    lib.vk_device_buffers.argtypes = [_vp, C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_vp)]
    lib.vk_stream.argtypes = [_vp, C.POINTER(_vp)]
    lib.vk_refine_stats.argtypes = [_vp, _ip, _ip]
    lib.vk_conden_setup.argtypes = [_vp, C.POINTER(CondenDesc)]
    lib.vk_conden_apply.argtypes = [_vp, _dp, _dp, _dp, _dp, _dp]
    lib.vk_ens_setup_steady.argtypes = [_vp, C.POINTER(SteadyOpts)]
    lib.vk_ens_photo_update.argtypes = [_vp]
    lib.vk_ens_run_steady.argtypes = [_vp, C.c_int, _ip]
    lib.vk_ens_get_steady.argtypes = [_vp, _ip, _dp, _dp, _dp, _dp, _dp]
    lib.vk_ens_get_fix.argtypes = [_vp, _ip, _bp, _dp]
    lib.vk_debug_time_kernel.argtypes = [_vp, C.c_int, C.c_int, C.POINTER(C.c_float)]
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        raise VulcanB200Error(rc, load().vk_last_error().decode("utf-8", "replace"))


def f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def u8(a):
    return np.ascontiguousarray(a, dtype=np.uint8)


def dptr(a):
    return None if a is None else a.ctypes.data_as(_dp)


def iptr(a):
    return None if a is None else a.ctypes.data_as(_ip)


def bptr(a):
    return None if a is None else a.ctypes.data_as(_bp)


class DeviceNetwork(object):
    """vk_network handle: the compiled reaction network resident on one GPU."""

    def __init__(self, network, device=0):
        self.lib = load()
        n = self.lib.vk_device_count()
        if n <= 0:
            raise VulcanB200Error(VK_ERR_CUDA, "no CUDA device visible (vulcan_b200 has no CPU fallback)")
        self.network = network
        t = network.tables()
        self._keep = {k: (np.ascontiguousarray(v) if isinstance(v, np.ndarray) else v) for k, v in t.items()}
        k = self._keep
        d = NetworkDesc(t["ni"], t["nr"], t["maxf"], t["maxjf"], len(k["rhs_pair"]), len(k["jac_row"]), len(k["jac_k"]),
                        iptr(k["rate_fac"]), iptr(k["rate_pow"]), iptr(k["rhs_ptr"]), iptr(k["rhs_pair"]),
                        dptr(k["rhs_coef"]), iptr(k["jac_ptr"]), iptr(k["jac_row"]), iptr(k["jac_col"]), iptr(k["jac_k"]),
                        dptr(k["jac_coef"]), iptr(k["jac_fac"]))
        h = _vp()
        check(self.lib.vk_network_create(C.byref(d), int(device), C.byref(h)))
        self.handle = h
        self.ni, self.nr, self.device = t["ni"], t["nr"], device

    def set_rates(self, rate_table):
        """upload a vulcan_b200.rates.RateTable (vk_rates_set): enables Columns.compute_k."""
        r = rate_table
        bp = lambda a: a.ctypes.data_as(C.POINTER(C.c_ubyte))
        self._rates_keep = r
        arr, cap = f64(r.arrhenius), f64(r.cap)
        self._rates_arrays = (arr, cap)
        d = RateDesc(int(r.npair), iptr(r.kind), dptr(arr), iptr(r.cap_kind), dptr(cap), bp(r.reverse), bp(r.removed),
                     iptr(r.gibbs_ptr), iptr(r.gibbs_sp), iptr(r.gibbs_nu), iptr(r.dnu), dptr(r.nasa9))
        check(self.lib.vk_rates_set(self.handle, C.byref(d)))

    def close(self):
        if getattr(self, "handle", None):
            self.lib.vk_network_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Columns(object):
    """vk_column handle: `ncol` independent columns of `nz` layers on the network's device."""

    def __init__(self, devnet, nz, ncol=1):
        self.lib = devnet.lib
        self.devnet = devnet
        self.nz, self.ncol, self.ni, self.nr = int(nz), int(ncol), devnet.ni, devnet.nr
        h = _vp()
        check(self.lib.vk_column_create(devnet.handle, self.nz, self.ncol, C.byref(h)))
        self.handle = h
        self._photo_nbr = 0

    def close(self):
        if getattr(self, "handle", None):
            self.lib.vk_column_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _shape(self, a, tail):
        a = f64(a)
        want = (self.ncol,) + tail
        if a.shape == tail and self.ncol == 1:
            a = a.reshape(want)
        if a.shape != want:
            raise ValueError("expected shape %s, got %s" % (want, a.shape))
        return a

    # ---------------------------------------------------------------- inputs
    def set_atm(self, Kzz, vz, dzi, Dzz, vs, Tco, g, M, Ti, Hpi, ms, alpha, top_flux, bot_flux, bot_vdep,
                use_moldiff=True, use_settling=False, use_topflux=False, use_botflux=False, gas_indx=None,
                gas_indx_lhs=None, use_vm_mol=False, vm=None, diff_esc_idx=None, shared=True):
        """use_vm_mol / vm / diff_esc_idx: the reference's *_vm stencil variants (op.py:2879-2888); vm is atm.vm [nz, ni]."""
        nz, ni = self.nz, self.ni
        rep = () if shared else (self.ncol,)

        def arr(a, tail):
            a = f64(a)
            if a.shape != rep + tail:
                raise ValueError("atm array: expected %s, got %s" % (rep + tail, a.shape))
            return a
        keep = dict(Kzz=arr(Kzz, (nz - 1,)), vz=arr(vz, (nz - 1,)), dzi=arr(dzi, (nz - 1,)), Dzz=arr(Dzz, (nz - 1, ni)),
                    vs=arr(vs, (nz - 1, ni)), Tco=arr(Tco, (nz,)), g=arr(g, (nz,)), M=arr(M, (nz,)), Ti=arr(Ti, (nz - 1,)),
                    Hpi=arr(Hpi, (nz - 1,)), ms=arr(ms, (ni,)), alpha=arr(alpha, (ni,)), top_flux=arr(top_flux, (ni,)),
                    bot_flux=arr(bot_flux, (ni,)), bot_vdep=arr(bot_vdep, (ni,)))
        gi = None if gas_indx is None or len(gas_indx) == ni else i32(gas_indx)
        gl = None if gas_indx_lhs is None or len(gas_indx_lhs) == ni else i32(gas_indx_lhs)
        use_vm_mol = bool(use_vm_mol) and bool(use_moldiff)
        if use_vm_mol:
            if vm is None:
                raise ValueError("use_vm_mol needs atm.vm")
            keep["vm"] = arr(vm, (nz, ni))
        de = i32(diff_esc_idx) if (use_vm_mol and diff_esc_idx is not None and len(diff_esc_idx)) else None
        v = AtmView(int(shared), int(use_moldiff), int(use_settling), int(use_topflux), int(use_botflux),
                    0 if gi is None else len(gi), iptr(gi), 0 if gl is None else len(gl), iptr(gl),
                    dptr(keep["Kzz"]), dptr(keep["vz"]), dptr(keep["dzi"]), dptr(keep["Dzz"]), dptr(keep["vs"]),
                    dptr(keep["Tco"]), dptr(keep["g"]), dptr(keep["M"]), dptr(keep["Ti"]), dptr(keep["Hpi"]),
                    dptr(keep["ms"]), dptr(keep["alpha"]), dptr(keep["top_flux"]), dptr(keep["bot_flux"]),
                    dptr(keep["bot_vdep"]), int(use_vm_mol), dptr(keep["vm"]) if use_vm_mol else None,
                    0 if de is None else len(de), iptr(de))
        check(self.lib.vk_set_atm(self.handle, C.byref(v)))

    def set_k(self, k, shared=None, static_rows_shared=False):
        """k: [nz, nr+1] (shared by all columns) or [ncol, nz, nr+1].  static_rows_shared (per-column k only): the rows outside the network's
        photolysis / ionisation / condensation sections are identical in every column (one T-P profile) - vk_set_k(shared = 2), the emitted
        chemistry kernels then apply to the batch."""
        k = f64(k)
        if shared is None:
            shared = (k.ndim == 2)
        want = (self.nz, self.nr + 1) if shared else (self.ncol, self.nz, self.nr + 1)
        if k.shape != want:
            raise ValueError("k: expected %s, got %s" % (want, k.shape))
        check(self.lib.vk_set_k(self.handle, dptr(k), 1 if shared else (2 if static_rows_shared else 0)))
        self._k_shared = bool(shared)

    def compute_k(self, Tco, M):
        """rate coefficients on the device (needs DeviceNetwork.set_rates): Tco, M [nz] (shared) or [ncol, nz]."""
        Tco, M = f64(Tco), f64(M)
        shared = Tco.ndim == 1
        want = (self.nz,) if shared else (self.ncol, self.nz)
        if Tco.shape != want or M.shape != want:
            raise ValueError("Tco / M: expected %s" % (want,))
        check(self.lib.vk_compute_k(self.handle, dptr(Tco), dptr(M), int(shared)))
        self._k_shared = bool(shared)

    def get_k(self):
        k = np.empty((self.nz, self.nr + 1) if getattr(self, "_k_shared", True) else (self.ncol, self.nz, self.nr + 1))
        check(self.lib.vk_get_k(self.handle, dptr(k)))
        return k

    def set_k_rows(self, rows, vals):
        rows, vals = i32(rows), f64(vals)
        check(self.lib.vk_set_k_rows(self.handle, len(rows), iptr(rows), dptr(vals)))

    def set_step_opts(self, mtol, atol, refine=0, zero_delta_row0=False, fix_bot_idx=(), fix_bot_val=None,
                      delta_zero_sp=None, fix_mask=None, fix_y=None, compo=None, refine_dt_min=REFINE_DT_MIN, rhs_order=0):
        """refine: passes of iterative refinement per solve (double-double residual); -1 = auto: one safeguarded pass on the columns
        whose dt >= refine_dt_min (needs compo [ni, na]).  rhs_order (chemdf): 0 (default) = the emitted straight-line kernel of the network
        when the library has one (vulcan_b200/emit.py; reference order, bit-identical to the generated chem_funs.py), else the table-driven
        kernel with the segmented summation; 1 = reference order (emitted, else table-driven); 2 = table-driven reference order; 3 =
        table-driven segmented."""
        fbi = i32(fix_bot_idx)
        fbv = None if len(fbi) == 0 else f64(fix_bot_val).reshape(self.ncol, len(fbi))
        dz = None if delta_zero_sp is None else u8(delta_zero_sp)
        fm = None if fix_mask is None else u8(fix_mask).reshape(self.ncol, self.nz, self.ni)
        fy = None if fix_y is None else f64(fix_y).reshape(self.ncol, self.nz, self.ni)
        cp = None if compo is None else f64(compo).reshape(self.ni, -1)
        o = StepOpts(float(mtol), float(atol), int(refine), int(zero_delta_row0), len(fbi), iptr(fbi) if len(fbi) else None,
                     dptr(fbv), bptr(dz), bptr(fm), dptr(fy), 0 if cp is None else cp.shape[1], dptr(cp), float(refine_dt_min), int(rhs_order))
        check(self.lib.vk_set_step_opts(self.handle, C.byref(o)))

    # ---------------------------------------------------------------- hot path
    def ros2_solve(self, y, ymix, dt):
        """≙ Ros2.solver (op.py:2860-3007) for every column.  Returns sol, ymix, delta[ncol], status[ncol]."""
        y = self._shape(y, (self.nz, self.ni))
        ymix = self._shape(ymix, (self.nz, self.ni))
        dt = f64(np.broadcast_to(np.asarray(dt, dtype=np.float64), (self.ncol,)))
        sol, ymo = np.empty_like(y), np.empty_like(y)
        delta = np.empty(self.ncol)
        status = np.zeros(self.ncol, dtype=np.int32)
        check(self.lib.vk_ros2_solve(self.handle, dptr(y), dptr(ymix), dptr(dt), dptr(sol), dptr(ymo), dptr(delta), iptr(status)))
        return sol, ymo, delta, status

    def ros2_solve_into(self, y, ymix, dt, sol, ymix_out, delta, status):
        """same call on caller-owned C-contiguous float64 arrays (e.g. numpy views of pinned torch tensors): no allocation,
        and page-locked buffers are DMA'd directly."""
        for a in (y, ymix, sol, ymix_out):
            assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"] and a.size == self.ncol * self.nz * self.ni
        check(self.lib.vk_ros2_solve(self.handle, dptr(y), dptr(ymix), dptr(dt), dptr(sol), dptr(ymix_out), dptr(delta), iptr(status)))

    def clip_loss(self, y, ymix_in, compo, pos_cut, nega_cut, atom_sum=None, small_y=None, nega_y=None, atom_skip=None, mtol=None):
        """mtol: vulcan_cfg.mtol of the clip mask (op.py:2459); None = the value last given to set_step_opts"""
        y = self._shape(y, (self.nz, self.ni)).copy()
        ymix_in = self._shape(ymix_in, (self.nz, self.ni))
        compo = f64(compo)
        na = compo.shape[1]
        ymo = np.empty_like(y)
        asum = np.zeros((self.ncol, na)) if atom_sum is None else f64(atom_sum).reshape(self.ncol, na).copy()
        sm = np.zeros(self.ncol) if small_y is None else f64(small_y).reshape(self.ncol).copy()
        ng = np.zeros(self.ncol) if nega_y is None else f64(nega_y).reshape(self.ncol).copy()
        anyneg = np.zeros(self.ncol, dtype=np.int32)
        sk = None if atom_skip is None else u8(atom_skip)
        check(self.lib.vk_clip_loss(self.handle, dptr(y), dptr(ymix_in), dptr(ymo), na, dptr(compo), bptr(sk),
                                    float(pos_cut), float(nega_cut), -1.0 if mtol is None else float(mtol), dptr(asum), dptr(sm),
                                    dptr(ng), iptr(anyneg)))
        return dict(y=y, ymix=ymo, atom_sum=asum, small_y=sm, nega_y=ng, any_negative=anyneg)

    # ---------------------------------------------------------------- components
    def eval_rhs(self, y):
        y = self._shape(y, (self.nz, self.ni))
        chem, diff = np.empty_like(y), np.empty_like(y)
        check(self.lib.vk_eval_rhs(self.handle, dptr(y), dptr(chem), dptr(diff)))
        return chem, diff

    def eval_lhs(self, y, dt):
        y = self._shape(y, (self.nz, self.ni))
        dt = f64(np.broadcast_to(np.asarray(dt, dtype=np.float64), (self.ncol,)))
        D = np.empty((self.ncol, self.nz, self.ni, self.ni))
        up, dn = np.empty_like(y), np.empty_like(y)
        check(self.lib.vk_eval_lhs(self.handle, dptr(y), dptr(dt), dptr(D), dptr(up), dptr(dn)))
        return D, up, dn

    def blocktri_solve(self, D, up, dn, rhs, refine=0):
        D = self._shape(D, (self.nz, self.ni, self.ni))
        up, dn, rhs = (self._shape(a, (self.nz, self.ni)) for a in (up, dn, rhs))
        x = np.empty_like(rhs)
        status = np.zeros(self.ncol, dtype=np.int32)
        check(self.lib.vk_blocktri_solve(self.handle, dptr(D), dptr(up), dptr(dn), dptr(rhs), dptr(x), int(refine), iptr(status)))
        return x, status

    # ---------------------------------------------------------------- photolysis
    def photo_setup(self, bins, sflux_top, i12, dbin1, dbin2, sl_angle, edd, flux_atol, f_diurnal, abs_idx, cross_abs,
                    photo_idx, cross_photo, scat_idx, cross_scat, cross_J, br_rate_index, abs_is_T=None, cross_abs_T=None,
                    br_is_T=None, cross_J_T=None):
        bins, sflux_top = f64(bins), f64(sflux_top)
        nbin = bins.size
        keep = [bins, sflux_top, i32(abs_idx), f64(cross_abs).reshape(-1, nbin), i32(photo_idx),
                f64(cross_photo).reshape(-1, nbin), i32(scat_idx), f64(cross_scat).reshape(-1, nbin),
                f64(cross_J).reshape(-1, nbin), i32(br_rate_index)]
        aT = None if abs_is_T is None else u8(abs_is_T)
        cT = None if cross_abs_T is None else f64(cross_abs_T)
        bT = None if br_is_T is None else u8(br_is_T)
        jT = None if cross_J_T is None else f64(cross_J_T)
        v = PhotoView(nbin, int(i12), float(dbin1), float(dbin2), float(sl_angle), float(edd), float(flux_atol),
                      float(f_diurnal), dptr(keep[0]), dptr(keep[1]),
                      len(keep[2]), iptr(keep[2]), dptr(keep[3]), bptr(aT), dptr(cT),
                      len(keep[4]), iptr(keep[4]), dptr(keep[5]),
                      len(keep[6]), iptr(keep[6]), dptr(keep[7]),
                      keep[8].shape[0], dptr(keep[8]), iptr(keep[9]), bptr(bT), dptr(jT))
        check(self.lib.vk_photo_setup(self.handle, C.byref(v)))
        self._photo_nbr, self._photo_nbin = keep[8].shape[0], nbin

    def photo_update(self, y, ymix, dz):
        y = self._shape(y, (self.nz, self.ni))
        ymix = self._shape(ymix, (self.nz, self.ni))
        dz = self._shape(dz, (self.nz,))
        J = np.empty((self.ncol, self._photo_nbr, self.nz))
        ch = np.empty(self.ncol)
        check(self.lib.vk_photo_update(self.handle, dptr(y), dptr(ymix), dptr(dz), dptr(J), dptr(ch)))
        return J, ch

    def photo_read(self, names=("tau", "sflux", "dflux_u", "dflux_d", "aflux")):
        nb = self._photo_nbin
        out = {}
        for n in names:
            out[n] = np.empty((self.ncol, self.nz + (0 if n == "aflux" else 1), nb))
        args = [dptr(out.get(n)) for n in ("tau", "sflux", "dflux_u", "dflux_d", "aflux")]
        check(self.lib.vk_photo_read(self.handle, *args))
        return out

    def photo_reset(self):
        check(self.lib.vk_photo_reset(self.handle))

    # ---------------------------------------------------------------- condensation operators (op.py:1109-1421)
    def conden_setup(self, re_idx, gas_idx, m, rho_p, r_p, sat, zero_rate, relax_kind=(), relax_gas=(), relax_ice=(), relax_top=(), relax_m=(),
                     relax_rho=(), relax_r=(), relax_sat=None, start_conden_time=0.0, stop_conden_time=1e300, post_conden_rtol=0.0):
        nz = self.nz
        re_idx, gas_idx = i32(re_idx), i32(gas_idx)
        n_re, n_rx = len(re_idx), len(relax_kind)
        keep = [re_idx, gas_idx, f64(m), f64(rho_p), f64(r_p), f64(sat).reshape(n_re, nz) if n_re else f64(np.zeros((0, nz))), u8(zero_rate),
                i32(relax_kind), i32(relax_gas), i32(relax_ice), i32(relax_top), f64(relax_m), f64(relax_rho), f64(relax_r),
                f64(relax_sat).reshape(n_rx, nz) if n_rx else f64(np.zeros((0, nz)))]
        d = CondenDesc(n_re, iptr(keep[0]), iptr(keep[1]), dptr(keep[2]), dptr(keep[3]), dptr(keep[4]), dptr(keep[5]), bptr(keep[6]), n_rx,
                       iptr(keep[7]), iptr(keep[8]), iptr(keep[9]), iptr(keep[10]), dptr(keep[11]), dptr(keep[12]), dptr(keep[13]), dptr(keep[14]),
                       float(start_conden_time), float(stop_conden_time), float(post_conden_rtol))
        check(self.lib.vk_conden_setup(self.handle, C.byref(d)))
        self._conden_nre = n_re

    def conden_apply(self, y, ymix, dt, n_0):
        """conden + the relaxation operators on the device: returns y, ymix, k_rows [ncol, n_re, 2, nz]"""
        y = self._shape(y, (self.nz, self.ni)).copy()
        ymix = self._shape(ymix, (self.nz, self.ni)).copy()
        dt = f64(np.broadcast_to(np.asarray(dt, dtype=np.float64), (self.ncol,)))
        n_0 = f64(np.broadcast_to(f64(n_0), (self.ncol, self.nz)))
        kr = np.zeros((self.ncol, self._conden_nre, 2, self.nz)) if self._conden_nre else None
        check(self.lib.vk_conden_apply(self.handle, dptr(y), dptr(ymix), dptr(dt), dptr(n_0), dptr(kr)))
        return y, ymix, kr

    def refine_stats(self):
        kept, tried = np.zeros(self.ncol, dtype=np.int32), np.zeros(self.ncol, dtype=np.int32)
        check(self.lib.vk_refine_stats(self.handle, iptr(kept), iptr(tried)))
        return kept, tried

    def last_kernel_ms(self):
        a, b = C.c_float(0), C.c_float(0)
        check(self.lib.vk_last_kernel_ms(self.handle, C.byref(a), C.byref(b)))
        return a.value, b.value

    def time_kernel(self, which, reps=3):
        """vk_debug_time_kernel: average ms of one kernel of the step on the resident state (0 lhs, 1 rhs, 2 factor, 4 solve)"""
        ms = C.c_float(0)
        check(self.lib.vk_debug_time_kernel(self.handle, int(which), int(reps), C.byref(ms)))
        return ms.value

    # ---------------------------------------------------------------- device-resident ensemble driver
    def ens_setup(self, rtol, loss_eps, dt_min, dt_max, dt_var_min, dt_var_max, pos_cut, nega_cut, compo, atom_ini, n_0):
        compo = f64(compo)
        na = compo.shape[1]
        atom_ini = f64(np.broadcast_to(f64(atom_ini), (self.ncol, na)))
        n_0 = f64(np.broadcast_to(f64(n_0), (self.ncol, self.nz)))
        o = EnsOpts(float(rtol), float(loss_eps), float(dt_min), float(dt_max), float(dt_var_min), float(dt_var_max),
                    float(pos_cut), float(nega_cut), na, dptr(compo), dptr(atom_ini), dptr(n_0))
        check(self.lib.vk_ens_setup(self.handle, C.byref(o)))

    def ens_set_state(self, y, dt):
        y = self._shape(y, (self.nz, self.ni))
        dt = f64(np.broadcast_to(np.asarray(dt, dtype=np.float64), (self.ncol,)))
        check(self.lib.vk_ens_set_state(self.handle, dptr(y), dptr(dt)))

    def ens_run(self, n_steps):
        check(self.lib.vk_ens_run(self.handle, int(n_steps)))

    def ens_setup_steady(self, cfg, pico, ms, zco, Hp, dz, pref_indx, gs, conv_ignore_sp=None, diff_esc_idx=(), hist_cap=None,
                         hist_stride=1, use_photo=None, condense=None):
        """device-resident run to steady state (vk_ens_setup_steady): cfg = mapping or object with the vulcan_cfg names read by
        Integration.stop / conv / __call__ (op.py:808-1087); zco / Hp / dz [ncol, nz(+1)] or one column's arrays (broadcast).
        condense (optional, after conden_setup): dict(fix_sp=[...], fix_whole=[...], fix_sat_mix=[n_fix, nz], from_coldtrap=bool,
        start_conden_time, stop_conden_time, post_conden_rtol) - conden / relaxation / the fix_species switch run inside the loop."""
        g = (lambda n, d=None: cfg.get(n, d)) if isinstance(cfg, dict) else (lambda n, d=None: getattr(cfg, n, d))
        nz, ncol = self.nz, self.ncol
        rep = lambda a, tail: f64(np.broadcast_to(f64(a), (ncol,) + tail))
        keep = dict(pico=f64(pico), ms=f64(ms), zco=rep(zco, (nz + 1,)), Hp=rep(Hp, (nz,)), dz=rep(dz, (nz,)))
        ig = None if conv_ignore_sp is None else u8(conv_ignore_sp)
        de = i32(diff_esc_idx) if len(diff_esc_idx) else None
        conv_step = int(g("conv_step"))
        cap = conv_step if hist_cap is None else int(hist_cap)
        up = bool(g("use_photo")) if use_photo is None else bool(use_photo)
        o = SteadyOpts(float(g("st_factor")), float(g("mtol_conv")), float(g("atol")), float(g("yconv_cri")), float(g("slope_cri")),
                       float(g("yconv_min")), float(g("flux_cri")), float(g("trun_min")), float(g("runtime")), conv_step,
                       int(g("count_min")), int(g("count_max")), bptr(ig), int(up), int(g("ini_update_photo_frq", 100) or 100),
                       int(g("final_update_photo_frq", 5) or 5), int(g("update_frq", 0) or 0), int(pref_indx), float(gs), float(g("Rp")),
                       float(g("max_flux", 1e13) or 1e13), dptr(keep["pico"]), dptr(keep["ms"]), dptr(keep["zco"]), dptr(keep["Hp"]),
                       dptr(keep["dz"]), 0 if de is None else len(de), iptr(de), cap, int(hist_stride))
        if condense is not None:
            n_fix = len(condense.get("fix_sp", ()))
            keep["fsp"] = i32(condense["fix_sp"]) if n_fix else None
            keep["fwh"] = u8(condense["fix_whole"]) if n_fix else None
            keep["fsm"] = f64(condense["fix_sat_mix"]).reshape(n_fix, nz) if n_fix else None
            o.use_condense, o.fix_species_switch, o.n_fix = 1, int(n_fix > 0), n_fix
            o.fix_from_coldtrap = int(bool(condense.get("from_coldtrap", False)))
            o.fix_sp, o.fix_whole_column, o.fix_sat_mix = iptr(keep["fsp"]), bptr(keep["fwh"]), dptr(keep["fsm"])
            o.start_conden_time = float(condense["start_conden_time"])
            o.stop_conden_time = float(condense["stop_conden_time"])
            o.post_conden_rtol = float(condense["post_conden_rtol"])
        check(self.lib.vk_ens_setup_steady(self.handle, C.byref(o)))

    def ens_photo_update(self):
        check(self.lib.vk_ens_photo_update(self.handle))

    def ens_run_steady(self, max_iterations):
        left = C.c_int(0)
        check(self.lib.vk_ens_run_steady(self.handle, int(max_iterations), C.byref(left)))
        return left.value

    def ens_get_steady(self, want_grid=False):
        ec = np.zeros(self.ncol, dtype=np.int32)
        ld, ldt, ch = np.empty(self.ncol), np.empty(self.ncol), np.empty(self.ncol)
        dz = np.empty((self.ncol, self.nz)) if want_grid else None
        zco = np.empty((self.ncol, self.nz + 1)) if want_grid else None
        check(self.lib.vk_ens_get_steady(self.handle, iptr(ec), dptr(ld), dptr(ldt), dptr(ch), dptr(dz), dptr(zco)))
        return dict(end_case=ec, longdy=ld, longdydt=ldt, aflux_change=ch, dz=dz, zco=zco)

    def ens_get_fix(self):
        fs = np.zeros(self.ncol, dtype=np.int32)
        fm = np.zeros((self.ncol, self.nz, self.ni), dtype=np.uint8)
        fy = np.zeros((self.ncol, self.nz, self.ni))
        check(self.lib.vk_ens_get_fix(self.handle, iptr(fs), bptr(fm), dptr(fy)))
        return dict(fix_started=fs, fix_mask=fm, fix_y=fy)

    def ens_get_state(self, want_y=True):
        y = np.empty((self.ncol, self.nz, self.ni)) if want_y else None
        t, dt = np.empty(self.ncol), np.empty(self.ncol)
        na, nrj = np.zeros(self.ncol, dtype=np.int32), np.zeros(self.ncol, dtype=np.int32)
        check(self.lib.vk_ens_get_state(self.handle, dptr(y), dptr(t), dptr(dt), iptr(na), iptr(nrj)))
        return dict(y=y, t=t, dt=dt, n_accept=na, n_reject=nrj)
