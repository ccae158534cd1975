"""Host mirror of the reference's `op.Integration` loop (op.py:808-1105) for ONE column, driving the GPU solver object
(`vulcan_b200.ros2.Ros2`) through the same protocol the reference's `vulcan.py:172-182` uses: this is what "run to steady
state" means for the single-column metrics (steps/s, time-to-steady-state) on a box that has no reference checkout.

Mirrored (same order of operations, same criteria):  __call__ (photo cadence switch op.py:818-829, one_step, update_mu_dz /
update_phi_esc every update_frq steps op.py:904-906, hydrostatic rescale op.py:909-914, f_dy, save_step, step_size), backup
(op.py:937-941), update_mu_dz (op.py:944-984), update_phi_esc (op.py:986-999), f_dy (op.py:1003-1015), conv (op.py:1018-1065),
stop (op.py:1067-1087), save_step (op.py:1089-1105).
Not built yet ("next"): the condensation operators `conden` / `*_conden_evap_relax` (op.py:1109-1421) and
`use_adapt_rtol`; configurations that enable them raise NotImplementedError instead of silently skipping physics.

All arithmetic of the step itself runs on the GPU behind the C ABI; what is left here is the reference's own per-step host
bookkeeping (numpy), kept on the host because `conv()` compares against the stored history `y_time`.
"""
import time

import numpy as np

KB = 1.38064852e-16      # phy_const.py:3
NAVO = 6.02214086e23     # phy_const.py:4


class Integration(object):
    def __init__(self, odesolver, cfg, species, verbose=False):
        self.odesolver, self.cfg, self.species, self.verbose = odesolver, cfg, list(species), verbose
        self.update_photo_frq = cfg.ini_update_photo_frq                      # op.py:800
        if getattr(cfg, "use_condense", False):
            self.non_gas_sp_index = [self.species.index(sp) for sp in cfg.non_gas_sp]
        self.n_photo_updates = 0
        self.t_photo = 0.0

    # ------------------------------------------------------------------------------------------------ op.py:808-935
    def __call__(self, var, atm, para, max_wall_s=None):
        cfg = self.cfg
        if getattr(cfg, "use_adapt_rtol", False):
            raise NotImplementedError("use_adapt_rtol (op.py:836-856)")
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
            if getattr(cfg, "use_condense", False) and var.t >= cfg.start_conden_time and not para.fix_species_start:
                raise NotImplementedError("condensation operators conden / relax (op.py:859-902, 1109-1421)")
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

    def backup(self, var):                                                                           # op.py:937-941
        var.y_prev = np.copy(var.y)
        var.dy_prev = np.copy(var.dy)
        var.atom_loss_prev = var.atom_loss.copy()
        return var

    def update_mu_dz(self, var, atm):                                                                # op.py:944-984
        cfg = self.cfg
        nz = var.y.shape[0]
        pref_indx = int(atm.pref_indx)
        Tco, pico = atm.Tco.copy(), atm.pico.copy()
        mu = np.zeros(nz)
        for i in range(len(self.species)):                                                           # build_atm.py:515-520
            mu += atm.ms[i] * var.ymix[:, i]
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
