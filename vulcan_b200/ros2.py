"""Drop-in replacement of the reference's `op.Ros2` solver object (vulcan.py:162-163 looks the class up by name).

Same protocol as `op.ODESolver` / `op.Ros2` (op.py:1423-3125): `naming_solver, one_step, solver, step_size, clip, loss,
step_ok, step_reject, reset_y, compute_tau, compute_flux, compute_J, compute_Jion` with the same arguments, the same
in-place mutation of `var` / `para` and the same error convention (no exceptions for numerical failure: a bad step is
signalled through `step_ok`).  All arithmetic runs in the sm_100a CUDA library behind include/vulcan_b200.h; this file only
packs the reference's containers (store.Variables / AtmData / Parameters, store.py:21-209) into the C ABI's flat arrays.

There is no CPU fallback: constructing the object without the library or without a GPU raises.
"""
import numpy as np

from . import _abi
from .network import Network

_DYN_ATM = ("Kzz", "vz", "dzi", "Dzz", "vs", "Tco", "g", "M", "Ti", "Hpi", "ms", "alpha", "top_flux", "bot_flux", "bot_vdep")


class Ros2(object):
    def __init__(self, cfg=None, species=None, compo=None, device=0, refine=-1, network=None, charge=None, refine_dt_min=None):
        """cfg: the vulcan_cfg module (imported like the reference does when omitted); species: chem_funs.spec_list
        (only used to cross-check the network compiler's species order); compo: {species: {atom: n}} or an
        [ni][na] array for `loss` (read from cfg.com_file when omitted).
        refine: iterative refinement of the two linear solves of a step (double-double residual): 0 none, n > 0 passes, -1 (default)
        one pass per solve once dt >= refine_dt_min (default _abi.REFINE_DT_MIN), kept only if it lowers the element-weighted residual."""
        if cfg is None:
            import vulcan_cfg as cfg          # reference module (op.py:30)
        self.cfg = cfg
        self.network = network if network is not None else Network.from_file(cfg.network)
        if species is not None and list(species) != self.network.species:
            raise ValueError("species order of the network compiler differs from chem_funs.spec_list")
        self.species = self.network.species
        self.ni, self.nr = self.network.ni, self.network.nr
        self.device, self.refine = device, refine
        self.refine_dt_min = getattr(_abi, "REFINE_DT_MIN", 1.0e3) if refine_dt_min is None else refine_dt_min
        self.mtol, self.atol = cfg.mtol, cfg.atol                           # op.py:1427-1428
        self.non_gas_sp = list(getattr(cfg, "non_gas_sp", []))
        sp = self.species
        if getattr(cfg, "use_condense", False):                             # op.py:1431-1433
            self.non_gas_sp_index = [sp.index(s) for s in self.non_gas_sp]
            self.condense_sp_index = [sp.index(s) for s in cfg.condense_sp]
        self.fix_sp_bot_index = [sp.index(s) for s in cfg.use_fix_sp_bot.keys()]           # op.py:1435-1436
        self.fix_sp_bot_mix = np.array([cfg.use_fix_sp_bot[s] for s in cfg.use_fix_sp_bot.keys()], dtype=float)
        # defaults frozen at construction like the reference's def-time defaults (op.py:2447, 2489, 2495, 2524, 3105)
        self._pos_cut, self._nega_cut = cfg.pos_cut, cfg.nega_cut
        self._loss_eps, self._rtol0 = cfg.loss_eps, cfg.rtol
        self._dt_var_min, self._dt_var_max, self._dt_min, self._dt_max = cfg.dt_var_min, cfg.dt_var_max, cfg.dt_min, cfg.dt_max
        self._compo = self._load_compo(compo)
        # use_ion: the 'e' column of all_compose.txt (build_atm.py:161), read by the charge balance (op.py:2998-3004)
        self._charge = self._load_charge(charge) if getattr(cfg, "use_ion", False) else None
        self._devnet = _abi.DeviceNetwork(self.network, device)
        self._col = None
        self._k_cache = None
        self._k_ids = None
        self._atm_cache = None
        self._opts_key = None
        self._photo_ready = False
        self._photo_cache = None

    # ------------------------------------------------------------------ helpers
    def _load_compo(self, compo):
        atoms = list(self.cfg.atom_list)
        if compo is None:
            with open(self.cfg.com_file) as f:
                cols = f.readline().split()
            tab = np.genfromtxt(self.cfg.com_file, names=True, dtype=["U20"] + ["int"] * (len(cols) - 2) + ["float"])
            rows = list(tab["species"])
            return np.array([[tab[rows.index(s)][a] for a in atoms] for s in self.species], dtype=float)
        if isinstance(compo, dict):
            return np.array([[compo[s][a] for a in atoms] for s in self.species], dtype=float)
        return np.asarray(compo, dtype=float)

    def _masses(self):
        """molar mass per species as mean_mass reads it (build_atm.py:511-520: the `mass` column of vulcan_cfg.com_file)"""
        with open(self.cfg.com_file) as f:
            cols = f.readline().split()
        tab = np.genfromtxt(self.cfg.com_file, names=True, dtype=["U20"] + ["int"] * (len(cols) - 2) + ["float"])
        rows = list(tab["species"])
        return np.array([tab[rows.index(sp)][cols[-1]] for sp in self.species], dtype=np.float64)

    def _load_charge(self, charge):
        if charge is None:
            with open(self.cfg.com_file) as f:
                cols = f.readline().split()
            tab = np.genfromtxt(self.cfg.com_file, names=True, dtype=["U20"] + ["int"] * (len(cols) - 2) + ["float"])
            rows = list(tab["species"])
            return np.array([float(tab[rows.index(s)]["e"]) for s in self.species])
        if isinstance(charge, dict):
            return np.array([float(charge[s]) for s in self.species])
        return np.asarray(charge, dtype=float)

    def _columns(self, nz):
        if self._col is None or self._col.nz != nz:
            self._col = _abi.Columns(self._devnet, nz, 1)
            self._k_cache = self._atm_cache = self._opts_key = None
            self._photo_ready = False
        return self._col

    def _flag(self, name):
        """optional vulcan_cfg switches are read the same way everywhere (older cfg files lack some of them)"""
        return bool(getattr(self.cfg, name, False))

    def _gas(self, atm):
        return np.asarray(atm.gas_indx, dtype=np.int32) if self.non_gas_sp else None

    def _sync_atm(self, atm, nz):
        cfg = self.cfg
        if self._flag("use_moldiff"):
            vals = {n: np.asarray(getattr(atm, n), dtype=np.float64) for n in _DYN_ATM}
        else:
            # use_moldiff = False: vulcan.py never calls mol_diff, so atm.Ti / atm.Hpi do not exist (build_atm.py:569-571) and atm.ms is
            # np.empty garbage (store.py:129); diffdf_no_mol / lhs_jac_no_mol (op.py:1438-1494, 2122-2166) read none of the molecular-
            # diffusion arrays -> zeros of the right shape go to the device (found by running the class inside the unmodified reference)
            ni = len(self.species)
            zeros = {"Ti": (nz - 1,), "Hpi": (nz - 1,), "ms": (ni,), "alpha": (ni,), "Dzz": (nz - 1, ni)}
            vals = {n: (np.zeros(zeros[n]) if n in zeros else np.asarray(getattr(atm, n), dtype=np.float64)) for n in _DYN_ATM}
        flags = tuple(self._flag(n) for n in ("use_moldiff", "use_settling", "use_topflux", "use_botflux", "use_vm_mol"))
        use_vm = flags[4] and flags[0]                               # Ros2.solver's dispatch (op.py:2869-2888)
        if use_vm:
            vals["vm"] = np.asarray(atm.vm, dtype=np.float64)        # build_atm.py:735-739
        c = self._atm_cache
        if c is not None and c[0] == flags and all(np.array_equal(c[1][n], vals[n]) for n in vals):
            return
        gas = self._gas(atm)
        if flags[0] and not flags[1]:
            gas_lhs = np.asarray(atm.gas_indx, dtype=np.int32) if self._flag("use_condense") else None     # op.py:1981-1984
        else:
            gas_lhs = gas
        self._columns(nz).set_atm(use_moldiff=flags[0], use_settling=flags[1], use_topflux=flags[2], use_botflux=flags[3],
                                  gas_indx=gas, gas_indx_lhs=gas_lhs, use_vm_mol=use_vm,
                                  diff_esc_idx=[self.species.index(sp) for sp in getattr(cfg, "diff_esc", [])] if use_vm else None,
                                  shared=True, **vals)
        self._atm_cache = (flags, {n: v.copy() for n, v in vals.items()})

    def _sync_k(self, var, nz):
        """var.k is a dict {1..nr -> (nz,) array} (store.py:25).  During integration the reference only REBINDS entries
        (compute_J op.py:2785, conden op.py:1122-1176), so the identity of the value objects is a sufficient change detector - the
        previous objects are kept ALIVE here and compared with `is` (an id() alone could be reused by a new array once the old one
        is freed); a full comparison by value is still made every 64 calls (in-place writes)."""
        refs = list(var.k.values())
        self._k_calls = getattr(self, "_k_calls", 0) + 1
        old = self._k_ids
        if self._k_cache is not None and old is not None and len(old) == len(refs) and all(a is b for a, b in zip(old, refs)) \
                and self._k_calls % 64:
            return
        k = np.zeros((nz, self.nr + 1))
        for i in range(1, self.nr + 1):
            k[:, i] = var.k[i]
        if self._k_cache is None or not np.array_equal(self._k_cache, k):
            self._columns(nz).set_k(k)
            self._k_cache = k
        self._k_ids = refs

    def _sync_opts(self, var, atm, para, nz, alloc_fix=False):
        """alloc_fix: pass zeroed fix_mask / fix_y even before the fix_species switch - the device-resident loop (steady.py) writes
        them itself when the switch happens on the device"""
        cfg, ni = self.cfg, self.ni
        fix_mask = fix_y = dz_sp = None
        if alloc_fix:
            fix_mask = np.zeros((nz, ni), dtype=np.uint8)
            fix_y = np.zeros((nz, ni))
        if self._flag("use_condense"):
            dz_sp = np.zeros(ni, dtype=np.uint8)
            dz_sp[self.non_gas_sp_index] = 1
            dz_sp[self.condense_sp_index] = 1
            if para.fix_species_start:                                        # op.py:2896-2906, 2960-2970
                fix_mask = np.zeros((nz, ni), dtype=np.uint8)
                fix_y = np.zeros((nz, ni))
                alloc_fix = False
                for s in cfg.fix_species:
                    i = self.species.index(s)
                    top = nz if not cfg.fix_species_from_coldtrap_lev else int(atm.conden_min_lev[s])
                    fix_mask[:top, i] = 1
                    fix_y[:top, i] = np.asarray(var.fix_y[s])[:top]
                    dz_sp[i] = 1
        if self._flag("use_ion"):
            # atm.fix_e_indx (store.py:157): the electron row of every layer is  1/(r h) e_i  with a zero right-hand side in both
            # stages (op.py:2908-2911, 2926), i.e. the solve leaves e untouched: sol[:, e] = y[:, e] and its delta is exactly 0
            ie = self.species.index("e")
            if fix_mask is None:
                fix_mask = np.zeros((nz, ni), dtype=np.uint8)
                fix_y = np.zeros((nz, ni))
            if dz_sp is None:
                dz_sp = np.zeros(ni, dtype=np.uint8)
            fix_mask[:, ie] = 1
            fix_y[:, ie] = np.asarray(var.y)[:, ie]
            dz_sp[ie] = 1
        fbi = list(self.fix_sp_bot_index)
        fbv = self.fix_sp_bot_mix * atm.n_0[0] if fbi else None               # op.py:2946
        zero0 = bool(self._flag("use_botflux") or cfg.use_fix_sp_bot)         # op.py:2953
        key = (zero0, tuple(fbi), None if fbv is None else fbv.tobytes(), None if dz_sp is None else dz_sp.tobytes(),
               None if fix_mask is None else fix_mask.tobytes(), None if fix_y is None else fix_y.tobytes(), self.mtol, self.atol,
               self.refine, self.refine_dt_min)
        if key != self._opts_key or alloc_fix:
            self._columns(nz).set_step_opts(self.mtol, self.atol, refine=self.refine, zero_delta_row0=zero0, fix_bot_idx=fbi,
                                            fix_bot_val=fbv, delta_zero_sp=dz_sp, fix_mask=fix_mask, fix_y=fix_y, compo=self._compo,
                                            refine_dt_min=self.refine_dt_min)
            self._opts_key = None if alloc_fix else key      # the device loop changes the arrays behind this cache

    # ------------------------------------------------------------------ the Ros2 protocol
    def naming_solver(self, para):                                           # op.py:3078-3088
        print("Include molecular diffusion." if self._flag("use_moldiff") else "No molecular diffusion.")
        para.solver_str = "solver"

    def solver(self, var, atm, para):
        """one attempted step (op.py:2860-3007) on the GPU; writes var.y, var.ymix, para.delta."""
        cfg = self.cfg
        y = np.ascontiguousarray(var.y, dtype=np.float64)
        nz = y.shape[0]
        if getattr(cfg, "use_fix_H2He", False) and "H2" not in cfg.use_fix_sp_bot and var.t > 1e6:      # op.py:2935-2941
            cfg.use_fix_sp_bot["H2"] = var.ymix[0, self.species.index("H2")]
            cfg.use_fix_sp_bot["He"] = var.ymix[0, self.species.index("He")]
            self.fix_sp_bot_index = [self.species.index(s) for s in cfg.use_fix_sp_bot.keys()]
            self.fix_sp_bot_mix = np.array([cfg.use_fix_sp_bot[s] for s in cfg.use_fix_sp_bot.keys()], dtype=float)
        self._sync_atm(atm, nz)
        self._sync_k(var, nz)
        self._sync_opts(var, atm, para, nz)
        sol, ymix, delta, status = self._col.ros2_solve(y, np.ascontiguousarray(var.ymix, dtype=np.float64), var.dt)
        var.y = sol[0]
        var.ymix = ymix[0]
        para.delta = float(delta[0]) if status[0] == 0 else float("nan")     # a singular block fails step_ok like a NaN would
        if self._flag("use_ion"):                                            # op.py:2998-3004: [e] from charge neutrality (ymix is not redone)
            ie = self.species.index("e")
            var.y[:, ie] = 0
            for sp in var.charge_list:
                i = self.species.index(sp)
                var.y[:, ie] -= self._charge[i] * var.y[:, i]
        return var, para

    def one_step(self, var, atm, para):                                      # op.py:3091-3103
        while True:
            var, para = getattr(self, para.solver_str)(var, atm, para)
            var, para = self.clip(var, para, atm)
            if self.step_ok(var, para):
                break
            elif self.step_reject(var, para):
                break
        return var, para

    def clip(self, var, para, atm, pos_cut=None, nega_cut=None):             # op.py:2447-2470
        pos_cut = self._pos_cut if pos_cut is None else pos_cut
        nega_cut = self._nega_cut if nega_cut is None else nega_cut
        nz = var.y.shape[0]
        self._sync_atm(atm, nz)
        skip = None
        loss_ex = getattr(self.cfg, "loss_ex", [])
        if loss_ex:
            skip = np.array([a in loss_ex for a in self.cfg.atom_list], dtype=np.uint8)
        prev = np.array([var.atom_sum.get(a, 0.0) for a in self.cfg.atom_list], dtype=float)
        res = self._columns(nz).clip_loss(var.y, var.ymix, self._compo, pos_cut, nega_cut, atom_sum=prev, atom_skip=skip,
                                          small_y=[para.small_y], nega_y=[para.nega_y], mtol=self.mtol)
        para.small_y, para.nega_y = float(res["small_y"][0]), float(res["nega_y"][0])
        for q, a in enumerate(self.cfg.atom_list):
            if a not in loss_ex:                                             # op.py:2482-2485
                var.atom_sum[a] = res["atom_sum"][0][q]
                var.atom_loss[a] = (var.atom_sum[a] - var.atom_ini[a]) / var.atom_ini[a]
        var.y, var.ymix = res["y"][0], res["ymix"][0]
        return var, para

    def loss(self, data_var):                                                # op.py:2472-2487 (host mirror; clip() already does it on the GPU)
        loss_ex = getattr(self.cfg, "loss_ex", [])
        for q, a in enumerate(self.cfg.atom_list):
            if a not in loss_ex:
                data_var.atom_sum[a] = float(np.sum(self._compo[:, q][None, :] * data_var.y))
                data_var.atom_loss[a] = (data_var.atom_sum[a] - data_var.atom_ini[a]) / data_var.atom_ini[a]
        return data_var

    def step_ok(self, var, para, loss_eps=None, rtol=None):                  # op.py:2489-2493
        loss_eps = self._loss_eps if loss_eps is None else loss_eps
        rtol = self._rtol0 if rtol is None else rtol
        dl = np.abs(np.fromiter(var.atom_loss.values(), float) - np.fromiter(var.atom_loss_prev.values(), float))
        return bool(np.all(var.y >= 0) and np.amax(dl) < loss_eps and para.delta <= rtol)

    def step_reject(self, var, para, loss_eps=None, rtol=None):              # op.py:2495-2522
        rtol = self._rtol0 if rtol is None else rtol
        if para.delta > rtol:
            para.delta_count += 1
        elif np.any(var.y < 0):
            para.nega_count += 1
        else:
            para.loss_count += 1
        var = self.reset_y(var)
        if var.dt < self.cfg.dt_min:
            var.dt = self.cfg.dt_min
            var.y[var.y < 0] = 0.
            print("Keep producing negative values! Clipping negative solutions and moving on!")
            return True
        return False

    def reset_y(self, var, dt_reduc=None):                                   # op.py:2524-2534 (ymix is NOT restored, as in the reference)
        var.y = var.y_prev
        var.dt *= self._dt_var_min if dt_reduc is None else dt_reduc
        return var

    def step_size(self, var, para, dt_var_min=None, dt_var_max=None, dt_min=None, dt_max=None):   # op.py:3105-3125
        dt_var_min = self._dt_var_min if dt_var_min is None else dt_var_min
        dt_var_max = self._dt_var_max if dt_var_max is None else dt_var_max
        dt_min = self._dt_min if dt_min is None else dt_min
        dt_max = self._dt_max if dt_max is None else dt_max
        h, delta, rtol = var.dt, para.delta, self.cfg.rtol                   # rtol is read live (op.py:3111)
        if delta == 0:
            delta = 0.01 * rtol
        h_factor = 0.9 * (rtol / delta) ** 0.5
        h_factor = np.maximum(h_factor, dt_var_min)
        h_factor = np.minimum(h_factor, dt_var_max)
        h *= h_factor
        h = np.maximum(h, dt_min)
        h = np.minimum(h, dt_max)
        var.dt = h
        return var

    # ------------------------------------------------------------------ photolysis (op.py:2580-2786)
    def _photo_setup(self, var, atm, nz):
        cfg, sp = self.cfg, self.species
        absp = sorted(set(var.photo_sp) | set(getattr(var, "ion_sp", set())))
        psp = sorted(var.photo_sp)
        tsp = list(getattr(cfg, "T_cross_sp", []))
        nbin = len(var.bins)
        br = [(s, b) for s in psp for b in range(1, var.n_branch[s] + 1)]
        abs_is_T = np.array([s in tsp for s in absp], dtype=np.uint8)
        cross_abs = np.array([np.zeros(nbin) if s in tsp else var.cross[s] for s in absp])
        cross_abs_T = None
        if abs_is_T.any():
            cross_abs_T = np.array([var.cross_T[s] if s in tsp else np.zeros((nz, nbin)) for s in absp])
        br_is_T = np.array([s in tsp for s, b in br], dtype=np.uint8)
        cross_J = np.array([np.zeros(nbin) if s in tsp else var.cross_J[(s, b)] for s, b in br])
        cross_J_T = None
        if br_is_T.any():
            cross_J_T = np.array([var.cross_J_T[(s, b)] if s in tsp else np.zeros((nz, nbin)) for s, b in br])
        rid = np.array([0 if var.pho_rate_index[b] in cfg.remove_list else var.pho_rate_index[b] for b in br], dtype=np.int32)
        self._branches = br
        self._ion_branches = []
        if self._flag("use_ion"):
            # compute_Jion (op.py:2789-2820) is the same trapezoid contraction as compute_J over the ion cross sections (no
            # temperature dependence): its branches ride in the same device table behind the photodissociation branches
            ibr = [(s, b) for s in sorted(var.ion_sp) for b in range(1, var.ion_branch[s] + 1)]
            self._ion_branches = ibr
            cross_J = np.concatenate([cross_J, np.array([var.cross_Jion[b] for b in ibr])])
            rid = np.concatenate([rid, np.array([0 if var.ion_rate_index[b] in cfg.remove_list else var.ion_rate_index[b]
                                                  for b in ibr], dtype=np.int32)])
            if br_is_T.any():
                br_is_T = np.concatenate([br_is_T, np.zeros(len(ibr), dtype=np.uint8)])
                cross_J_T = np.concatenate([cross_J_T, np.zeros((len(ibr), nz, nbin))])
        self._columns(nz).photo_setup(var.bins, var.sflux_top, var.sflux_din12_indx, var.dbin1, var.dbin2, cfg.sl_angle, cfg.edd,
                                      cfg.flux_atol, cfg.f_diurnal, [sp.index(s) for s in absp], cross_abs,
                                      [sp.index(s) for s in psp], np.array([var.cross[s] for s in psp]),
                                      [sp.index(s) for s in cfg.scat_sp], np.array([var.cross_scat[s] for s in cfg.scat_sp]),
                                      cross_J, rid, abs_is_T=abs_is_T if abs_is_T.any() else None, cross_abs_T=cross_abs_T,
                                      br_is_T=br_is_T if br_is_T.any() else None, cross_J_T=cross_J_T)
        # the diffuse-flux state of the previous call lives on the device (op.py:2692); seed it from var if present
        self._photo_ready = True

    def _photo_run(self, var, atm):
        nz = var.y.shape[0]
        self._sync_k(var, nz)
        if not self._photo_ready:
            self._photo_setup(var, atm, nz)
        J, ch = self._col.photo_update(var.y, var.ymix, atm.dz)
        f = self._col.photo_read()
        self._photo_cache = dict(J=J[0], change=float(ch[0]), prev_aflux=np.copy(getattr(var, "aflux", np.zeros_like(f["aflux"][0]))),
                                 **{n: v[0] for n, v in f.items()})
        return self._photo_cache

    def compute_tau(self, var, atm):                                         # op.py:2580-2599
        c = self._photo_run(var, atm)
        var.tau = c["tau"]

    def compute_flux(self, var, atm):                                        # op.py:2603-2737
        c = self._photo_cache if self._photo_cache is not None else self._photo_run(var, atm)
        var.sflux, var.dflux_u, var.dflux_d = c["sflux"], c["dflux_u"], c["dflux_d"]
        var.prev_aflux = c["prev_aflux"]
        var.aflux = c["aflux"]
        var.aflux_change = c["change"]

    def compute_J(self, var, atm):                                           # op.py:2751-2786
        c = self._photo_cache if self._photo_cache is not None else self._photo_run(var, atm)
        nz = var.y.shape[0]
        var.J_sp = dict([((s, b), np.zeros(nz)) for s in var.photo_sp for b in range(var.n_branch[s] + 1)])
        for q, (s, b) in enumerate(self._branches):
            var.J_sp[(s, b)] = c["J"][q].copy()
            var.J_sp[(s, 0)] += var.J_sp[(s, b)]
            if var.pho_rate_index[(s, b)] not in self.cfg.remove_list:
                var.k[var.pho_rate_index[(s, b)]] = var.J_sp[(s, b)] * self.cfg.f_diurnal
        self._k_cache = None          # the device copy of k already holds the new J rows; re-verified on the next solver call
        self._k_ids = None
        self._jion_cache = c if self._flag("use_ion") else None     # Integration calls compute_Jion right after (op.py:828-829)
        self._photo_cache = None

    def compute_Jion(self, var, atm):                                        # op.py:2789-2820
        c = getattr(self, "_jion_cache", None)
        if c is None:
            c = self._photo_run(var, atm)
            self._photo_cache = None
        nz = var.y.shape[0]
        var.Jion_sp = dict([((s, b), np.zeros(nz)) for s in var.ion_sp for b in range(var.ion_branch[s] + 1)])
        n0 = len(self._branches)
        for q, (s, b) in enumerate(self._ion_branches):
            var.Jion_sp[(s, b)] = c["J"][n0 + q].copy()
            var.Jion_sp[(s, 0)] += var.Jion_sp[(s, b)]
            if var.ion_rate_index[(s, b)] not in self.cfg.remove_list:
                var.k[var.ion_rate_index[(s, b)]] = var.Jion_sp[(s, b)] * self.cfg.f_diurnal
        self._k_cache = None
        self._k_ids = None
        self._jion_cache = None
