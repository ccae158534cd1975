"""Device-resident replacement of the reference's `op.Integration` loop (op.py:783-1105), condensation included (conden, the H2O / NH3
relaxation operators and the fix_species switch, op.py:856-901):
`DeviceIntegration(solver)(var, atm, para)` runs the whole time integration of ONE column on the GPU - attempted steps, accept / reject,
step size, hydrostatic rescale, stop / conv against the stored history, the photolysis cadence, update_mu_dz / update_phi_esc - with one
host round trip per `chunk` iterations instead of several per step, and writes the reference's containers back at the end
(var.y, ymix, t, dt, longdy, longdydt, aflux_change, atom_loss; para.count, end_case, reject counters; atm.dz, zco, ...).

The same C-ABI entry points (vk_ens_setup / vk_ens_setup_steady / vk_ens_run_steady) drive whole ensembles: `vulcan_b200.ensemble`.
Seam in the reference: `integ = op.Integration(solver, output)` (vulcan.py:166); INTEGRATION.md section 3.
"""
import time

import numpy as np


KB = 1.38064852e-16      # phy_const.py:3
NAVO = 6.02214086e23     # phy_const.py:4
# growth reactions handled by Integration.conden: reaction string -> (gas species, particle key of atm.r_p / atm.rho_p, molecular mass in
# amu exactly as the reference writes it, op.py:1122-1296)
CONDEN_REACTIONS = {'H2O -> H2O_l_s': ('H2O', 'H2O_l_s', 18.), 'NH3 -> NH3_l': ('NH3', 'NH3_l_s', 17.),
                    'H2SO4 -> H2SO4_l': ('H2SO4', 'H2SO4_l', 98.022), 'S2 -> S2_l_s': ('S2', 'S2_l_s', 45.019),
                    'S4 -> S4_l_s': ('S4', 'S4_l_s', 32.06 * 4), 'S8 -> S8_l_s': ('S8', 'S8_l_s', 360.152), 'C -> C_s': ('C', 'C_s', 12.011)}


def conden_tables(cfg, var, atm, species):
    """keyword arguments of _abi.Columns.conden_setup from the reference's containers: what Integration.conden and the two relaxation
    operators read (op.py:1109-1421) - var.conden_re_list / var.Rf, atm.sat_p, atm.r_p, atm.rho_p, atm.Tco, atm.n_0, vulcan_cfg.humidity,
    condense_sp, use_relax."""
    sp = list(species)
    nz = len(atm.Tco)
    Tco = np.asarray(atm.Tco, dtype=float)
    use_relax = list(getattr(cfg, "use_relax", []) or [])
    re_idx, gas_idx, m, rho_p, r_p, sat, zero = [], [], [], [], [], [], []
    for re in var.conden_re_list:
        ent = CONDEN_REACTIONS.get(var.Rf[re])
        if ent is None or ent[0] not in cfg.condense_sp:
            continue
        gas, part, amu = ent
        re_idx.append(int(re)); gas_idx.append(sp.index(gas))
        if gas in ('H2O', 'NH3') and use_relax:                 # relaxation replaces the growth reaction (op.py:1124-1126)
            zero.append(1); m.append(0.0); rho_p.append(1.0); r_p.append(1.0); sat.append(np.zeros(nz))
            continue
        zero.append(0)
        m.append(amu / NAVO); rho_p.append(float(atm.rho_p[part])); r_p.append(float(atm.r_p[part]))
        s_ = np.asarray(atm.sat_p[gas], dtype=float) / KB / Tco
        if gas == 'H2O':
            s_ = s_ * cfg.humidity
        sat.append(s_)
    kind, rgas, rice, rtop, rm, rrho, rr, rsat = [], [], [], [], [], [], [], []
    if 'H2O' in use_relax:                                      # op.py:1340-1376
        kind.append(1); rgas.append(sp.index('H2O')); rice.append(sp.index('H2O_l_s')); rtop.append(nz)
        rm.append(18. / NAVO); rrho.append(float(atm.rho_p['H2O_l_s'])); rr.append(float(atm.r_p['H2O_l_s']))
        rsat.append(np.asarray(atm.sat_p['H2O'], dtype=float) / KB / Tco * cfg.humidity)
    if 'NH3' in use_relax:                                      # op.py:1378-1421
        satp = np.asarray(atm.sat_p['NH3'], dtype=float) / KB / Tco
        kind.append(2); rgas.append(sp.index('NH3')); rice.append(sp.index('NH3_l_s')); rtop.append(int(np.argmin(satp / np.asarray(atm.n_0, dtype=float))))
        rm.append(17. / NAVO); rrho.append(float(atm.rho_p['NH3_l_s'])); rr.append(float(atm.r_p['NH3_l_s']))
        rsat.append(satp)
    return dict(re_idx=re_idx, gas_idx=gas_idx, m=m, rho_p=rho_p, r_p=r_p, sat=np.array(sat).reshape(len(re_idx), nz), zero_rate=zero,
                relax_kind=kind, relax_gas=rgas, relax_ice=rice, relax_top=rtop, relax_m=rm, relax_rho=rrho, relax_r=rr,
                relax_sat=np.array(rsat).reshape(len(kind), nz), start_conden_time=float(getattr(cfg, "start_conden_time", 0.0)),
                stop_conden_time=float(getattr(cfg, "stop_conden_time", 1e300)), post_conden_rtol=float(getattr(cfg, "post_conden_rtol", 0.0)))


class DeviceIntegration(object):
    def __init__(self, odesolver, chunk=64, verbose=False):
        self.odesolver, self.chunk, self.verbose = odesolver, int(chunk), verbose
        self.wall = 0.0

    def __call__(self, var, atm, para, make_atm=None, max_wall_s=None):
        s = self.odesolver
        cfg = s.cfg
        if getattr(cfg, "use_ion", False) or getattr(cfg, "use_adapt_rtol", False):
            raise NotImplementedError("use_ion / use_adapt_rtol are not part of the device-resident loop")
        if getattr(cfg, "use_fix_H2He", False) and "H2" not in cfg.use_fix_sp_bot:
            raise NotImplementedError("use_fix_H2He changes the bottom boundary at t > 1e6 s on the host (op.py:2935-2941)")
        condensing = bool(getattr(cfg, "use_condense", False)) and not para.fix_species_start
        fix_species = list(getattr(cfg, "fix_species", []) or []) if condensing else []
        if condensing and not set(fix_species) <= set(cfg.condense_sp) | set(cfg.non_gas_sp):
            raise NotImplementedError("fix_species outside condense_sp / non_gas_sp changes delta_zero_sp at the switch")
        y = np.ascontiguousarray(var.y, dtype=np.float64)
        nz = y.shape[0]
        s._sync_atm(atm, nz)
        s._sync_k(var, nz)
        s._sync_opts(var, atm, para, nz, alloc_fix=bool(fix_species))
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
        condense = None
        if condensing:
            tabs = conden_tables(cfg, var, atm, s.species)
            col.conden_setup(**tabs)
            whole = [sp in ('H2O_l_s', 'H2SO4_l', 'NH3_l_s', 'S8_l_s') for sp in fix_species]                  # op.py:878
            sat_mix = [np.zeros(nz) if w else np.asarray(atm.sat_mix[sp], dtype=float) for sp, w in zip(fix_species, whole)]
            condense = dict(fix_sp=[s.species.index(sp) for sp in fix_species], fix_whole=whole, fix_sat_mix=np.array(sat_mix).reshape(len(fix_species), nz),
                            from_coldtrap=bool(getattr(cfg, "fix_species_from_coldtrap_lev", False)),
                            start_conden_time=tabs["start_conden_time"], stop_conden_time=tabs["stop_conden_time"],
                            post_conden_rtol=tabs["post_conden_rtol"])
        col.ens_setup_steady(cfg, atm.pico, ms, atm.zco, atm.Hp, atm.dz, int(atm.pref_indx), float(atm.gs), conv_ignore_sp=ignore,
                             diff_esc_idx=[s.species.index(sp) for sp in getattr(cfg, "diff_esc", []) or []], use_photo=use_photo,
                             condense=condense)
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
        if fix_species:
            fx = col.ens_get_fix()
            if fx["fix_started"][0]:                                                                          # op.py:866-892
                para.fix_species_start = True
                cfg.rtol = cfg.post_conden_rtol
                atm.vs = atm.vs * 0
                var.fix_y = {}
                for sp in fix_species:
                    i = s.species.index(sp)
                    var.fix_y[sp] = fx["fix_y"][0][:, i].copy()
                    if getattr(cfg, "fix_species_from_coldtrap_lev", False):
                        atm.conden_min_lev[sp] = int(fx["fix_mask"][0][:, i].sum())
        s._opts_key = None
        for q, a in enumerate(atoms):
            var.atom_sum[a] = float(np.sum(s._compo[:, q][None, :] * var.y))
            var.atom_loss[a] = (var.atom_sum[a] - var.atom_ini[a]) / var.atom_ini[a]
        s._k_cache = None           # the device copy of k holds newer photolysis rows than var.k
        s._k_ids = None
        s._atm_cache = None
        return var, atm, para
