"""Device-resident replacement of the reference's `op.Integration` loop (op.py:783-1105) for the configurations without condensation:
`DeviceIntegration(solver)(var, atm, para)` runs the whole time integration of ONE column on the GPU - attempted steps, accept / reject,
step size, hydrostatic rescale, stop / conv against the stored history, the photolysis cadence, update_mu_dz / update_phi_esc - with one
host round trip per `chunk` iterations instead of several per step, and writes the reference's containers back at the end
(var.y, ymix, t, dt, longdy, longdydt, aflux_change, atom_loss; para.count, end_case, reject counters; atm.dz, zco, ...).

The same C-ABI entry points (vk_ens_setup / vk_ens_setup_steady / vk_ens_run_steady) drive whole ensembles: `vulcan_b200.ensemble`.
Seam in the reference: `integ = op.Integration(solver, output)` (vulcan.py:166); INTEGRATION.md section 3.
"""
import time

import numpy as np


class DeviceIntegration(object):
    def __init__(self, odesolver, chunk=64, verbose=False):
        self.odesolver, self.chunk, self.verbose = odesolver, int(chunk), verbose
        self.wall = 0.0

    def __call__(self, var, atm, para, make_atm=None, max_wall_s=None):
        s = self.odesolver
        cfg = s.cfg
        if getattr(cfg, "use_condense", False):
            raise NotImplementedError("condensation / relaxation operators (op.py:1109-1421) run per accepted step on the host: use the "
                                      "reference's op.Integration with the drop-in solver object for this configuration")
        if getattr(cfg, "use_ion", False) or getattr(cfg, "use_adapt_rtol", False):
            raise NotImplementedError("use_ion / use_adapt_rtol are not part of the device-resident loop")
        y = np.ascontiguousarray(var.y, dtype=np.float64)
        nz = y.shape[0]
        s._sync_atm(atm, nz)
        s._sync_k(var, nz)
        s._sync_opts(var, atm, para, nz)
        use_photo = bool(getattr(cfg, "use_photo", False))
        if use_photo and not s._photo_ready:
            s._photo_setup(var, atm, nz)
        col = s._col
        atoms = list(cfg.atom_list)
        atom_ini = np.array([var.atom_ini[a] for a in atoms], dtype=float)
        col.ens_setup(cfg.rtol, cfg.loss_eps, cfg.dt_min, cfg.dt_max, cfg.dt_var_min, cfg.dt_var_max, cfg.pos_cut, cfg.nega_cut,
                      s._compo, atom_ini, np.asarray(atm.n_0, dtype=float))
        col.ens_set_state(y, float(var.dt))
        ignore = np.zeros(s.ni, dtype=np.uint8)
        for sp in getattr(cfg, "conver_ignore", []) or []:
            ignore[s.species.index(sp)] = 1
        ms = np.asarray(atm.ms, dtype=float) if getattr(cfg, "use_moldiff", True) else np.asarray(s._masses(), dtype=float)
        col.ens_setup_steady(cfg, atm.pico, ms, atm.zco, atm.Hp, atm.dz, int(atm.pref_indx), float(atm.gs), conv_ignore_sp=ignore,
                             diff_esc_idx=[s.species.index(sp) for sp in getattr(cfg, "diff_esc", []) or []], use_photo=use_photo)
        t0 = time.time()
        left, it = 1, 0
        while left:
            left = col.ens_run_steady(self.chunk)
            it += self.chunk
            if self.verbose:
                st = col.ens_get_state(want_y=False)
                print("iterations %d: accepted %d rejected %d t %.3e dt %.3e" % (it, st["n_accept"][0], st["n_reject"][0], st["t"][0], st["dt"][0]))
            if max_wall_s is not None and time.time() - t0 > max_wall_s:
                para.end_case = 4
                break
        self.wall = time.time() - t0
        st = col.ens_get_state(want_y=True)
        sd = col.ens_get_steady(want_grid=True)
        var.y = st["y"][0]
        gi = list(atm.gas_indx) if getattr(cfg, "non_gas_sp", None) else None
        var.ymix = var.y / np.vstack(np.sum(var.y[:, gi], axis=1)) if gi else var.y / np.vstack(np.sum(var.y, axis=1))
        var.t, var.dt = float(st["t"][0]), float(st["dt"][0])
        var.longdy, var.longdydt, var.aflux_change = float(sd["longdy"][0]), float(sd["longdydt"][0]), float(sd["aflux_change"][0])
        para.count = int(st["n_accept"][0])
        self.n_rejected = int(st["n_reject"][0])
        if left == 0:
            para.end_case = int(sd["end_case"][0])
        atm.dz, atm.zco = sd["dz"][0], sd["zco"][0]
        for q, a in enumerate(atoms):
            var.atom_sum[a] = float(np.sum(s._compo[:, q][None, :] * var.y))
            var.atom_loss[a] = (var.atom_sum[a] - var.atom_ini[a]) / var.atom_ini[a]
        s._k_cache = None           # the device copy of k holds newer photolysis rows than var.k
        s._k_ids = None
        s._atm_cache = None
        return var, atm, para
