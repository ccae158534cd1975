"""TEST INFRASTRUCTURE (moved out of the product package in round 2: it restates host code of the reference that is outside the hot path,
SURVEY.md section 2 row 10; the product's own loop is the device-resident vulcan_b200/steady.py).  Host mirror of the reference's `op.Integration` loop (op.py:808-1105) for ONE column, driving the GPU solver object
(`vulcan_b200.ros2.Ros2`) through the same protocol the reference's `vulcan.py:172-182` uses: this is what "run to steady
state" means for the single-column metrics (steps/s, time-to-steady-state) on a box that has no reference checkout.

Mirrored (same order of operations, same criteria):  __call__ (photo cadence switch op.py:818-829, one_step, update_mu_dz /
update_phi_esc every update_frq steps op.py:904-906, hydrostatic rescale op.py:909-914, f_dy, save_step, step_size), backup
(op.py:937-941), update_mu_dz (op.py:944-984), update_phi_esc (op.py:986-999), f_dy (op.py:1003-1015), conv (op.py:1018-1065),
stop (op.py:1067-1087), save_step (op.py:1089-1105).
Condensation: `conden` (op.py:1109-1300, growth rates of the `X -> X_l_s` reactions written into var.k),
`h2o_conden_evap_relax` / `nh3_conden_evap_relax` (op.py:1340-1421, implicit-Euler relaxation to saturation) and the
fix_species switch with its cold-trap levels (op.py:859-893) are mirrored expression by expression (same evaluation order, so
they are bit-identical to the reference on the same inputs: tests/test_integration_host.py against fixtures recorded from the
reference's own operators).  They act once per ACCEPTED step on nz-vectors of one or two species - bookkeeping of the caller,
not part of the Ros2 stage arithmetic - and stay on the host next to conv()/save_step, which need var.y on the host anyway.
`use_adapt_rtol` (op.py:835-853; no shipped cfg defines the switch, SURVEY 8c shim 2) is mirrored in `adapt_rtol`.

All arithmetic of the step itself runs on the GPU behind the C ABI; what is left here is the reference's own per-step host
bookkeeping (numpy), kept on the host because `conv()` compares against the stored history `y_time`.
"""
import time

import numpy as np

KB = 1.38064852e-16      # phy_const.py:3
NAVO = 6.02214086e23     # phy_const.py:4


class Integration(object):
    def __init__(self, odesolver, cfg, species, verbose=False, mass=None):
        self.odesolver, self.cfg, self.species, self.verbose = odesolver, cfg, list(species), verbose
        self._mass = None if mass is None else np.asarray(mass, dtype=np.float64)     # molar masses, see _mol_mass
        self.update_photo_frq = cfg.ini_update_photo_frq                      # op.py:800
        if getattr(cfg, "use_condense", False):
            self.non_gas_sp_index = [self.species.index(sp) for sp in cfg.non_gas_sp]
        self.n_photo_updates = 0
        self.t_photo = 0.0

    # ------------------------------------------------------------------------------------------------ op.py:808-935
    def __call__(self, var, atm, para, max_wall_s=None):
        cfg = self.cfg
        self.loss_criteria = 0.0005                                                                  # op.py:811
        t0 = time.time()
        while not self.stop(var, para, atm):
            var = self.backup(var)
            if cfg.use_photo and var.longdy < cfg.yconv_min * 10. and var.longdydt < 1.e-6:          # op.py:818-822
                self.update_photo_frq = cfg.final_update_photo_frq
                para.switch_final_photo_frq = True
            if cfg.use_photo and para.count % self.update_photo_frq == 0:                            # op.py:824-829
                tp = time.time()
                self.odesolver.compute_tau(var, atm)
                self.odesolver.compute_flux(var, atm)
                self.odesolver.compute_J(var, atm)
                if getattr(cfg, "use_ion", False):
                    self.odesolver.compute_Jion(var, atm)
                self.n_photo_updates += 1
                self.t_photo += time.time() - tp
            var, para = self.odesolver.one_step(var, atm, para)                                      # op.py:832
            if getattr(cfg, "use_adapt_rtol", False):                                                # op.py:835-853 ("TEST 2025")
                self.adapt_rtol(var, para)
            if getattr(cfg, "use_condense", False) and var.t >= cfg.start_conden_time and not para.fix_species_start:   # op.py:859-902
                var = self.conden(var, atm)
                if cfg.fix_species and var.t > cfg.stop_conden_time:
                    self.start_fix_species(var, atm, para)
                if cfg.use_relax:
                    if 'H2O' in cfg.use_relax:
                        var = self.h2o_conden_evap_relax(var, atm)
                    if 'NH3' in cfg.use_relax:
                        var = self.nh3_conden_evap_relax(var, atm)
            if para.count % cfg.update_frq == 0:                                                     # op.py:904-906
                atm = self.update_mu_dz(var, atm)
                atm = self.update_phi_esc(var, atm)
            if getattr(cfg, "use_condense", False):                                                  # op.py:909-914
                var.y[:, atm.gas_indx] = np.vstack(atm.n_0) * var.ymix[:, atm.gas_indx]
            else:
                var.y = np.vstack(atm.n_0) * var.ymix
            var = self.f_dy(var, para)
            var, para = self.save_step(var, para)
            var = self.odesolver.step_size(var, para)
            if max_wall_s is not None and time.time() - t0 > max_wall_s:
                para.end_case = 4
                break
        return var, atm, para

    # ------------------------------------------------------------------------------------------------ op.py:835-853
    def adapt_rtol(self, var, para):
        """`use_adapt_rtol`: element loss steers cfg.rtol.  Only step_size sees the new value (it reads vulcan_cfg.rtol live,
        op.py:3111); step_ok / step_reject keep the rtol their default arguments froze at import (op.py:2489, 2495) - the mirror's
        `_rtol0` - exactly as in the reference."""
        cfg = self.cfg
        worst = max(abs(v) for v in var.atom_loss.values())
        if para.count % 10 == 0 and worst >= self.loss_criteria:
            self.loss_criteria *= 2.
            cfg.rtol = max(cfg.rtol * 0.75, cfg.rtol_min)
        if para.count % 1000 == 0 and para.count > 0 and worst < 2e-4:
            cfg.rtol = min(cfg.rtol * 1.25, cfg.rtol_max)

    # ------------------------------------------------------------------------------------------------ op.py:862-893
    def start_fix_species(self, var, atm, para):
        """first step after stop_conden_time: freeze the condensable species, record their cold-trap levels, switch rtol,
        turn the settling velocity off (op.py:862-893)."""
        cfg, sp_index = self.cfg, self.species.index
        nz = var.y.shape[0]
        para.fix_species_start = True
        cfg.rtol = cfg.post_conden_rtol
        atm.vs *= 0
        var.fix_y = {}
        if not hasattr(atm, "conden_min_lev"):
            atm.conden_min_lev = {}
        for sp in cfg.fix_species:
            var.fix_y[sp] = np.copy(var.y[:, sp_index(sp)])
            if cfg.fix_species_from_coldtrap_lev:
                if sp in ('H2O_l_s', 'H2SO4_l', 'NH3_l_s', 'S8_l_s'):
                    atm.conden_min_lev[sp] = nz - 1
                else:
                    sat_rho = atm.n_0 * atm.sat_mix[sp]
                    conden_status = var.y[:, sp_index(sp)] >= sat_rho
                    atm.conden_status = conden_status
                    if list(var.y[conden_status, sp_index(sp)]):
                        min_sat = np.amin(atm.sat_mix[sp][conden_status])
                        atm.min_sat = min_sat
                        atm.conden_min_lev[sp] = np.where(atm.sat_mix[sp] == min_sat)[0].item()
                    else:
                        atm.conden_min_lev[sp] = 0

    # growth-rate reactions handled by conden: Rf string -> (gas species, particle key of r_p / rho_p, molecular mass in amu as the
    # reference writes it, op.py:1122-1296), humidity factor applies to water only
    CONDEN = {'H2O -> H2O_l_s': ('H2O', 'H2O_l_s', 18.), 'NH3 -> NH3_l': ('NH3', 'NH3_l_s', 17.),
              'H2SO4 -> H2SO4_l': ('H2SO4', 'H2SO4_l', 98.022), 'S2 -> S2_l_s': ('S2', 'S2_l_s', 45.019),
              'S4 -> S4_l_s': ('S4', 'S4_l_s', 32.06 * 4), 'S8 -> S8_l_s': ('S8', 'S8_l_s', 360.152), 'C -> C_s': ('C', 'C_s', 12.011)}

    def conden(self, var, atm):                                                                      # op.py:1109-1300
        cfg, sp_index = self.cfg, self.species.index
        nz = var.y.shape[0]
        for re in var.conden_re_list:
            ent = self.CONDEN.get(var.Rf[re])
            if ent is None or ent[0] not in cfg.condense_sp:
                continue
            gas, part, amu = ent
            if gas in ('H2O', 'NH3') and cfg.use_relax:                        # relaxation replaces the growth reaction (op.py:1124-1126)
                var.k[re] = np.repeat(0., nz)
                var.k[re + 1] = np.repeat(0., nz)
                continue
            m = amu / NAVO
            rho_p, r_p = atm.rho_p[part], atm.r_p[part]
            sat = atm.sat_p[gas] / KB / atm.Tco
            if gas == 'H2O':
                sat = sat * cfg.humidity
            Dg = np.insert(atm.Dzz[:, sp_index(gas)], 0, atm.Dzz[0, sp_index(gas)])
            rate = Dg * m / rho_p / r_p ** 2 * (var.y[:, sp_index(gas)] - sat)
            var.k[re] = np.maximum(rate, 0)                                    # positive: condensation
            var.k[re + 1] = np.abs(np.minimum(rate, 0))                        # negative: evaporation
        return var

    def h2o_conden_evap_relax(self, var, atm):                                                       # op.py:1340-1376
        cfg, sp_index = self.cfg, self.species.index
        iw, il = sp_index('H2O'), sp_index('H2O_l_s')
        m = 18. / NAVO
        rho_p, r_p = atm.rho_p['H2O_l_s'], atm.r_p['H2O_l_s']
        sat_humidity = atm.sat_p['H2O'] / KB / atm.Tco * cfg.humidity
        Dg = np.insert(atm.Dzz[:, iw], 0, atm.Dzz[0, iw])
        with np.errstate(divide='ignore', invalid='ignore'):
            tau = 1. / (Dg * m / rho_p / r_p ** 2 * (var.y[:, iw] - sat_humidity))
            conden_indx = np.where(tau > 0)
            evap_indx = np.where(tau < 0)
            sat_mix = sat_humidity / atm.n_0
            y_conden = (var.ymix[:, iw] + var.dt / tau * sat_mix) / (1. + var.dt / tau)
            ice_loss = (var.y[:, iw] - sat_humidity) * var.dt / tau
        ice_loss = np.minimum(var.y[:, il], ice_loss)
        var.ymix[conden_indx, il] += (var.ymix[conden_indx, iw] - y_conden[conden_indx])
        var.ymix[conden_indx, iw] = y_conden[conden_indx]
        var.ymix[evap_indx, iw] += ice_loss[evap_indx] / atm.n_0[evap_indx]
        var.ymix[evap_indx, il] -= ice_loss[evap_indx] / atm.n_0[evap_indx]
        var.y = var.ymix * np.vstack(np.sum(var.y[:, atm.gas_indx], axis=1))
        return var

    def nh3_conden_evap_relax(self, var, atm):                                                       # op.py:1378-1421
        sp_index = self.species.index
        ig, il = sp_index('NH3'), sp_index('NH3_l_s')
        m = 17. / NAVO
        rho_p, r_p = atm.rho_p['NH3_l_s'], atm.r_p['NH3_l_s']
        sat_p = atm.sat_p['NH3'] / KB / atm.Tco
        sat_mix = sat_p / atm.n_0
        conden_top = np.argmin(sat_mix)
        Dg = np.insert(atm.Dzz[:, ig], 0, atm.Dzz[0, ig])
        with np.errstate(divide='ignore', invalid='ignore'):
            tau = 1. / (Dg * m / rho_p / r_p ** 2 * (var.y[:, ig] - sat_p))
            conden_indx = np.where(tau > 0)[0]
            evap_indx = np.where(tau < 0)[0]
            conden_indx = [i for i in conden_indx if i <= conden_top]         # no condensation above the top of the condensation zone
            y_conden = (var.ymix[:, ig] + var.dt / tau * sat_mix) / (1. + var.dt / tau)
            ice_loss = (var.y[:, ig] - sat_p) * var.dt / tau
        ice_loss = np.minimum(var.y[:, il], ice_loss)
        var.ymix[conden_indx, il] += (var.ymix[conden_indx, ig] - y_conden[conden_indx])
        var.ymix[conden_indx, ig] = y_conden[conden_indx]
        var.ymix[evap_indx, ig] += ice_loss[evap_indx] / atm.n_0[evap_indx]
        var.ymix[evap_indx, il] -= ice_loss[evap_indx] / atm.n_0[evap_indx]
        var.ymix[:, il] = np.maximum(var.ymix[:, il], 0)
        var.y = var.ymix * np.vstack(np.sum(var.y[:, atm.gas_indx], axis=1))
        return var

    def backup(self, var):                                                                           # op.py:937-941
        var.y_prev = np.copy(var.y)
        var.dy_prev = np.copy(var.dy)
        var.atom_loss_prev = var.atom_loss.copy()
        return var

    def _mol_mass(self, atm):
        """molar mass per species as mean_mass reads it (build_atm.py:511-520: the `mass` column of vulcan_cfg.com_file).  atm.ms holds the
        same numbers once mol_diff has run (build_atm.py:675) - but with use_moldiff = False it is np.empty garbage (store.py:129), so
        then the masses come from the `mass=` argument or the compose file."""
        if self._mass is None:
            if self.cfg.use_moldiff:
                return atm.ms
            with open(self.cfg.com_file) as f:
                cols = f.readline().split()
            tab = np.genfromtxt(self.cfg.com_file, names=True, dtype=["U20"] + ["int"] * (len(cols) - 2) + ["float"])
            rows = list(tab["species"])
            self._mass = np.array([tab[rows.index(sp)][cols[-1]] for sp in self.species], dtype=np.float64)
        return self._mass

    def update_mu_dz(self, var, atm):                                                                # op.py:944-984
        cfg = self.cfg
        nz = var.y.shape[0]
        pref_indx = int(atm.pref_indx)
        Tco, pico = atm.Tco.copy(), atm.pico.copy()
        mu = np.zeros(nz)
        ms = self._mol_mass(atm)
        for i in range(len(self.species)):                                                           # build_atm.py:515-520
            mu += ms[i] * var.ymix[:, i]
        atm.mu = mu
        Hp = atm.Hp
        for i in range(pref_indx, nz):
            if i == pref_indx:
                atm.g[i] = atm.gs
                Hp[i] = KB * Tco[i] / (atm.mu[i] / NAVO * atm.gs)
            else:
                atm.g[i] = atm.gs * (cfg.Rp / (cfg.Rp + atm.zco[i])) ** 2
                Hp[i] = KB * Tco[i] / (atm.mu[i] / NAVO * atm.g[i])
            atm.dz[i] = Hp[i] * np.log(pico[i] / pico[i + 1])
            atm.zco[i + 1] = atm.zco[i] + atm.dz[i]
        if not pref_indx == 0:
            for i in range(pref_indx - 1, -1, -1):
                atm.g[i] = atm.gs * (cfg.Rp / (cfg.Rp + atm.zco[i + 1])) ** 2
                Hp[i] = KB * Tco[i] / (atm.mu[i] / NAVO * atm.g[i])
                atm.dz[i] = Hp[i] * np.log(pico[i] / pico[i + 1])
                atm.zco[i] = atm.zco[i + 1] - atm.dz[i]
        zmco = 0.5 * (atm.zco + np.roll(atm.zco, -1))
        atm.zmco = zmco[:-1]
        dzi = 0.5 * (atm.dz + np.roll(atm.dz, 1))
        atm.dzi = dzi[1:]
        if cfg.use_moldiff:
            Ti = 0.5 * (Tco + np.roll(Tco, -1))
            atm.Ti = Ti[:-1]
            Hpi = 0.5 * (Hp + np.roll(Hp, -1))
            atm.Hpi = Hpi[:-1]
        return atm

    def update_phi_esc(self, var, atm):                                                              # op.py:986-999
        cfg = self.cfg
        for sp in getattr(cfg, "diff_esc", []):
            i = self.species.index(sp)
            atm.top_flux[i] = -atm.Dzz[-1, i] * var.y[-1, i] * (1. / atm.Hp[-1] - atm.ms[i] * atm.g[-1] / (NAVO * KB * atm.Tco[-1]))
            atm.top_flux[i] = max(atm.top_flux[i], cfg.max_flux * (-1))
        return atm

    def f_dy(self, var, para):                                                                       # op.py:1003-1015
        if para.count == 0:
            var.dy, var.dydt = 1., 1.
            return var
        y, ymix, y_prev = var.y, var.ymix, var.y_prev
        dy = np.abs(y - y_prev)
        dy[ymix < self.cfg.mtol] = 0
        dy[y < self.cfg.atol] = 0
        dy = np.amax(dy[y > 0] / y[y > 0])
        var.dy, var.dydt = dy, dy / var.dt
        return var

    def conv(self, var, para, atm):                                                                  # op.py:1018-1065
        cfg = self.cfg
        st_factor, mtol_conv, atol, yconv_cri, slope_cri, yconv_min = \
            cfg.st_factor, cfg.mtol_conv, cfg.atol, cfg.yconv_cri, cfg.slope_cri, cfg.yconv_min
        y, ymix, y_time, t_time = var.y.copy(), var.ymix.copy(), var.y_time, var.t_time
        count = para.count
        slope_min = min(np.amin(atm.Kzz / (0.1 * atm.Hp[:-1]) ** 2), 1.e-8)
        slope_min = max(slope_min, 1.e-10)
        indx = np.abs(np.asarray(t_time) - var.t * st_factor).argmin()
        if indx == para.count - 1:
            indx -= 1
        indx = max(para.count - cfg.conv_step, indx)
        longdy = np.abs((y_time[count - 1] - y_time[indx]) / np.vstack(atm.n_0))
        longdy[ymix < mtol_conv] = 0
        longdy[y < atol] = 0
        for sp in getattr(cfg, "conver_ignore", []):
            longdy[:, self.species.index(sp)] = 0
        if getattr(cfg, "use_condense", False):
            longdy[:, self.non_gas_sp_index] = 0
        longdy = np.amax(longdy[ymix > 0] / ymix[ymix > 0])
        longdydt = longdy / (t_time[-1] - t_time[indx])
        var.longdy, var.longdydt = longdy, longdydt
        if (longdy < yconv_cri and longdydt < slope_cri or longdy < yconv_min and longdydt < slope_min) and var.aflux_change < cfg.flux_cri:
            return True
        return False

    def stop(self, var, para, atm):                                                                  # op.py:1067-1087
        cfg = self.cfg
        if var.t > cfg.trun_min and para.count > cfg.count_min and self.conv(var, para, atm):
            para.end_case = 1
            return True
        elif var.t > cfg.runtime:
            para.end_case = 2
            return True
        elif para.count > cfg.count_max:
            para.end_case = 3
            return True
        return False

    def save_step(self, var, para):                                                                  # op.py:1089-1105
        var.t += var.dt
        para.count += 1
        var.y_time.append(var.y)
        var.t_time.append(var.t)
        var.atom_loss_time.append(list(var.atom_loss.values()))
        return var, para
