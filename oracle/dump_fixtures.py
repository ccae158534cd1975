#!/usr/bin/env python
"""TEST INFRASTRUCTURE — generates the golden fixtures under tests/golden/ by RUNNING THE
UNMODIFIED REFERENCE (a scratch copy made by oracle/stage_reference.py) with
PYTHONHASHSEED=0.  The reference ships no tests or golden vectors, so every
fixture is the output of the reference's own functions on the reference's own state.

Fixture levels:
  L0  chem_funs.chemdf, chem_funs.neg_symjac blocks, ODESolver.diffdf*, lhs_jac_* (op.py:1496-2444)
  L1  Ros2.solver output (sol, ymix, delta) + the two stage vectors k1, k2 (op.py:2860-3007)
  L2  clip / loss / step_ok (op.py:2447-2493)
  L4  compute_tau / compute_flux / compute_J (op.py:2580-2786)
  L5  full run trajectory (written by --full): per-step (t, dt, delta), final y
  L6  condensation operators of the caller (written by --conden): inputs / outputs of Integration.conden and
      *_conden_evap_relax (op.py:1109-1421) at selected counts + the fix_species switch (op.py:862-893)

usage (from the repo root):
  python oracle/stage_reference.py --config HD189
  PYTHONHASHSEED=0 python oracle/dump_fixtures.py --config HD189 --steps 0,10,100,300
  PYTHONHASHSEED=0 python oracle/dump_fixtures.py --config HD189 --full
"""
import argparse
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
GOLD = os.path.join(REPO, "tests", "golden")

JAC_LAYERS_FRAC = (0.0, 0.007, 0.3, 0.5, 0.8, 1.0)   # layers whose dense blocks are stored


def _sel_layers(nz):
    return sorted(set(min(nz - 1, int(round(f * (nz - 1)))) for f in JAC_LAYERS_FRAC))


def cfg_scalars(cfg):
    names = ["sl_angle", "edd", "f_diurnal", "flux_atol", "mtol", "atol", "rtol", "loss_eps", "pos_cut",
             "nega_cut", "dttry", "dt_min", "dt_max", "dt_var_min", "dt_var_max", "dbin1", "dbin2",
             "dbin_12trans", "use_moldiff", "use_settling", "use_vm_mol", "use_condense", "use_topflux",
             "use_botflux", "use_ion", "use_photo", "update_frq", "ini_update_photo_frq",
             "final_update_photo_frq", "nz", "yconv_cri", "slope_cri", "yconv_min", "flux_cri",
             "mtol_conv", "st_factor", "conv_step", "trun_min", "count_min", "count_max", "runtime",
             "gs", "Rp", "post_conden_rtol", "start_conden_time", "stop_conden_time", "max_flux"]
    d = {n: getattr(cfg, n) for n in names if hasattr(cfg, n)}
    for n in ["atom_list", "scat_sp", "T_cross_sp", "remove_list", "non_gas_sp", "condense_sp", "fix_species",
              "diff_esc", "use_relax", "atm_base", "network", "use_fix_sp_bot", "ode_solver"]:
        if hasattr(cfg, n):
            d[n] = getattr(cfg, n)
    return d


def compo_matrix(s):
    cfg = s.cfg
    with open(cfg.com_file) as f:
        cols = f.readline().split()
    num_ele = len(cols) - 2
    types = ["U20"] + ["int"] * num_ele + ["float"]
    compo = np.genfromtxt(cfg.com_file, names=True, dtype=types)
    rows = list(compo["species"])
    m = np.zeros((s.ni, len(cfg.atom_list)))
    for i, sp in enumerate(s.species):
        for a, atom in enumerate(cfg.atom_list):
            m[i, a] = compo[rows.index(sp)][atom]
    return m


def static_dict(s):
    from ref_session import pack_k
    var, atm, cfg = s.var, s.atm, s.cfg
    ni, nr, nz = s.ni, s.nr, s.nz
    d = dict(
        species=np.array(s.species), ni=ni, nr=nr, nz=nz,
        k=pack_k(var, nr, nz),
        M=atm.M, n_0=atm.n_0, Tco=atm.Tco, pco=atm.pco, pico=atm.pico, Kzz=atm.Kzz, vz=atm.vz,
        Dzz=atm.Dzz, vs=atm.vs, vm=atm.vm, ms=atm.ms, alpha=atm.alpha, top_flux=atm.top_flux, bot_flux=atm.bot_flux,
        bot_vdep=atm.bot_vdep, gas_indx=np.array(atm.gas_indx), pref_indx=atm.pref_indx,
        compo=compo_matrix(s), atom_ini=np.array([var.atom_ini[a] for a in cfg.atom_list]),
        y_ini=var.y_ini, cfg_json=json.dumps(cfg_scalars(cfg), default=str),
        photo_indx=getattr(var, "photo_indx", -1), stop_rev_indx=getattr(var, "stop_rev_indx", -1),
        conden_indx=getattr(var, "conden_indx", -1),
    )
    if cfg.use_photo:
        psp = sorted(var.photo_sp)          # canonical (sorted) order for the fixture tables
        br = [(sp, b) for sp in psp for b in range(1, var.n_branch[sp] + 1)]
        d.update(
            photo_sp=np.array(psp), photo_sp_idx=np.array([s.species.index(x) for x in psp]),
            scat_sp_idx=np.array([s.species.index(x) for x in cfg.scat_sp]),
            bins=var.bins, nbin=var.nbin, dbin1=var.dbin1, dbin2=var.dbin2,
            sflux_din12_indx=var.sflux_din12_indx, sflux_top=var.sflux_top,
            cross=np.array([var.cross[sp] for sp in psp]),
            cross_scat=np.array([var.cross_scat[sp] for sp in cfg.scat_sp]),
            cross_J=np.array([var.cross_J[b] for b in br]),
            branch_sp=np.array([psp.index(sp) for sp, b in br]), branch_no=np.array([b for sp, b in br]),
            branch_rate_index=np.array([var.pho_rate_index[b] for b in br]),
        )
        if cfg.use_ion:                     # photo-ionisation tables (op.py:253-269, 617-618, 751-768) and the charge bookkeeping
            isp = sorted(var.ion_sp)
            ibr = [(sp, b) for sp in isp for b in range(1, var.ion_branch[sp] + 1)]
            import build_atm
            d.update(
                ion_sp=np.array(isp), ion_sp_idx=np.array([s.species.index(x) for x in isp]),
                ion_cross=np.array([var.cross[sp] for sp in isp]),          # total absorption of the ionising species (compute_tau)
                cross_Jion=np.array([var.cross_Jion[b] for b in ibr]),
                ion_branch_sp=np.array([isp.index(sp) for sp, b in ibr]), ion_branch_no=np.array([b for sp, b in ibr]),
                ion_branch_rate_index=np.array([var.ion_rate_index[b] for b in ibr]),
                charge_list=np.array(list(var.charge_list)),
                charge=np.array([float(build_atm.compo[build_atm.compo_row.index(sp)]["e"]) for sp in s.species]),
            )
        if cfg.T_cross_sp:
            tsp = [sp for sp in psp if sp in cfg.T_cross_sp]
            d.update(T_cross_sp=np.array(tsp),
                     cross_T=np.array([var.cross_T[sp] for sp in tsp]),
                     cross_J_T=np.array([var.cross_J_T[b] for b in br if b[0] in tsp]),
                     cross_J_T_branch=np.array([br.index(b) for b in br if b[0] in tsp]))
    return d


def dyn_atm(atm):
    # use_moldiff = False: build_atm never calls mol_diff, so the interface arrays Hpi / Ti do not exist (and the _no_mol stencils
    # never read them): stored as zeros of the interface shape
    zi = np.zeros(len(atm.dzi))
    return dict(dz=atm.dz.copy(), dzi=atm.dzi.copy(), mu=atm.mu.copy(), Hp=atm.Hp.copy(), Hpi=np.array(getattr(atm, "Hpi", zi), copy=True),
                Ti=np.array(getattr(atm, "Ti", zi), copy=True), g=atm.g.copy(), top_flux_dyn=atm.top_flux.copy(), vs_dyn=atm.vs.copy(),
                zco=atm.zco.copy())


def capture_step(s, tag, count):
    """dump L0-L2 around ONE call of the reference's own one_step at the current state."""
    import copy
    import scipy.linalg
    from ref_session import pack_k
    var, atm, para, solver, cfg, op = s.var, s.atm, s.para, s.solver, s.cfg, s.op
    ni, nr, nz = s.ni, s.nr, s.nz
    kk = pack_k(var, nr, nz)
    rows = np.array([i for i in range(nr + 1) if not np.array_equal(kk[i], s.k_static[i])], dtype=int)
    out = dict(count=count, t=var.t, dt=var.dt, y=var.y.copy(), ymix=var.ymix.copy(),
               k_rows_idx=rows, k_rows=kk[rows],   # rows of k that differ from <cfg>_static.npz['k']
               atom_loss_prev=np.array([var.atom_loss_prev.get(a, 0.0) for a in cfg.atom_list]),
               atom_loss_in=np.array([var.atom_loss.get(a, 0.0) for a in cfg.atom_list]),
               fix_species_start=para.fix_species_start)
    out.update(dyn_atm(atm))
    if cfg.use_condense and para.fix_species_start:       # the state Ros2.solver's fixed-species rows read (op.py:2896-2906, 2960-2970)
        out["fix_species"] = np.array(list(cfg.fix_species))
        out["fix_y"] = np.array([var.fix_y[sp] for sp in cfg.fix_species])
        # fix_species_from_coldtrap_lev = False: the whole column is fixed and conden_min_lev is only filled for the condensing gases
        out["conden_min_lev"] = np.array([int(atm.conden_min_lev.get(sp, s.nz)) for sp in cfg.fix_species])
        out["fix_from_coldtrap"] = bool(cfg.fix_species_from_coldtrap_lev)
    # ---- L0: components, evaluated by the reference functions on a copy of the state
    y = var.y.copy()
    if not cfg.use_vm_mol and cfg.use_moldiff and not cfg.use_settling:
        diffdf, jac = solver.diffdf, solver.lhs_jac_tot
    elif not cfg.use_vm_mol and cfg.use_moldiff and cfg.use_settling:
        diffdf, jac = solver.diffdf_settling, solver.lhs_jac_settling
    elif cfg.use_vm_mol and cfg.use_moldiff and not cfg.use_settling:
        diffdf, jac = solver.diffdf_vm, solver.lhs_jac_tot_vm
    elif cfg.use_vm_mol and cfg.use_moldiff and cfg.use_settling:
        diffdf, jac = solver.diffdf_settling_vm, solver.lhs_jac_settling_vm
    else:
        diffdf, jac = solver.diffdf_no_mol, solver.lhs_jac_no_mol
    out["chemdf"] = op.chemdf(y, atm.M, var.k)
    out["diffdf"] = diffdf(y, atm)
    lay = _sel_layers(nz)
    out["layers"] = np.array(lay)
    nj = op.neg_achemjac(y, atm.M, var.k)
    out["negjac_blocks"] = np.array([nj[j * ni:(j + 1) * ni, j * ni:(j + 1) * ni] for j in lay])
    del nj
    lhs = jac(var, atm)
    out["lhs_blocks"] = np.array([lhs[j * ni:(j + 1) * ni, j * ni:(j + 1) * ni] for j in lay])
    idx = np.arange(ni)
    out["lhs_diag"] = np.array([lhs[j * ni + idx, j * ni + idx] for j in range(nz)])
    up = np.zeros((nz, ni)); dn = np.zeros((nz, ni))
    for j in range(nz - 1):
        up[j] = lhs[j * ni + idx, (j + 1) * ni + idx]
        dn[j + 1] = lhs[(j + 1) * ni + idx, j * ni + idx]
    out["lhs_up"], out["lhs_dn"] = up, dn
    # the off-diagonal blocks must be exactly diagonal - assert on the reference output
    j = lay[len(lay) // 2]
    if 0 < j < nz - 1:
        blk = lhs[j * ni:(j + 1) * ni, (j + 1) * ni:(j + 2) * ni].copy()
        blk[idx, idx] = 0
        assert not blk.any()
    # ---- stage vectors with the reference's own band storage + LAPACK (op.py:2913-2929), only when no row surgery
    r = 1. + 1. / 2. ** 0.5
    if not (cfg.use_condense and para.fix_species_start) and not cfg.use_ion:
        df = out["chemdf"].flatten() + out["diffdf"].flatten()
        lhs_b, bw = solver.store_bandM(lhs, ni, nz)
        k1 = scipy.linalg.solve_banded((bw, bw), lhs_b, df)
        yk2 = y + k1.reshape(y.shape) / r
        df2 = op.chemdf(yk2, atm.M, var.k).flatten() + diffdf(yk2, atm).flatten()
        rhs = df2 - 2. / (r * var.dt) * k1
        k2 = scipy.linalg.solve_banded((bw, bw), lhs_b, rhs)
        out["k1"], out["k2"], out["rhs2"] = k1.reshape(y.shape), k2.reshape(y.shape), rhs.reshape(y.shape)
        del lhs_b
    del lhs
    # ---- L1: the reference's solver() itself (on the live objects, then restored)
    y_save, ymix_save, delta_save = var.y.copy(), var.ymix.copy(), para.delta
    aloss_save, asum_save = dict(var.atom_loss), dict(var.atom_sum)
    small_save, nega_save = para.small_y, para.nega_y
    var, para = solver.solver(var, atm, para)
    out["sol"], out["sol_ymix"], out["delta"] = var.y.copy(), var.ymix.copy(), float(para.delta)
    # ---- L2: clip + loss + step_ok
    para.small_y, para.nega_y = 0., 0.
    var, para = solver.clip(var, para, atm)
    out["clip_y"], out["clip_ymix"] = var.y.copy(), var.ymix.copy()
    out["clip_small_y"], out["clip_nega_y"] = float(para.small_y), float(para.nega_y)
    out["atom_sum"] = np.array([var.atom_sum[a] for a in cfg.atom_list])
    out["atom_loss"] = np.array([var.atom_loss[a] for a in cfg.atom_list])
    out["step_ok"] = bool(solver.step_ok(var, para))
    # restore
    var.y, var.ymix, para.delta = y_save, ymix_save, delta_save
    var.atom_loss, var.atom_sum = aloss_save, asum_save
    para.small_y, para.nega_y = small_save, nega_save
    s.var, s.para = var, para
    path = os.path.join(GOLD, "%s_step%04d.npz" % (tag, count))
    np.savez_compressed(path, **out)
    print("wrote", path, "dt=%.3e delta=%.3e ok=%s" % (out["dt"], out["delta"], out["step_ok"]), flush=True)


def capture_photo(s, tag, count, nsub=16):
    """L4: two consecutive photolysis updates from a zeroed diffuse-flux state (the scheme is a lagged
    fixed point in var.dflux_u, op.py:2692), at the current y.  2-D outputs are stored on every
    `nsub`-th wavelength bin (the computation is independent per bin); J rates are stored in full."""
    var, atm, solver, cfg = s.var, s.atm, s.solver, s.cfg
    nz = s.nz
    keep = {n: getattr(var, n).copy() for n in ("tau", "sflux", "dflux_u", "dflux_d", "aflux", "prev_aflux") if hasattr(var, n)}
    keep_k = {i: np.copy(v) for i, v in var.k.items()}
    keep_change = var.aflux_change
    sel = np.arange(0, var.nbin, nsub)
    psp = sorted(var.photo_sp)
    br = [(sp, b) for sp in psp for b in range(1, var.n_branch[sp] + 1)]
    out = dict(count=count, y=var.y.copy(), ymix=var.ymix.copy(), dz=atm.dz.copy(), bin_sel=sel)
    var.dflux_u = np.zeros((nz + 1, var.nbin)); var.dflux_d = np.zeros((nz + 1, var.nbin))
    var.aflux = np.zeros((nz, var.nbin)); var.prev_aflux = np.zeros((nz, var.nbin))
    for it in (1, 2):
        solver.compute_tau(var, atm)
        solver.compute_flux(var, atm)
        solver.compute_J(var, atm)
        out["tau%d" % it] = var.tau[:, sel].copy()
        out["sflux%d" % it] = var.sflux[:, sel].copy()
        out["dflux_u%d" % it] = var.dflux_u[:, sel].copy()
        out["dflux_d%d" % it] = var.dflux_d[:, sel].copy()
        out["aflux%d" % it] = var.aflux[:, sel].copy()
        out["aflux_change%d" % it] = float(var.aflux_change)
        out["J%d" % it] = np.array([var.J_sp[b] for b in br])
        out["kphoto%d" % it] = np.array([np.asarray(var.k[var.pho_rate_index[b]]) for b in br])
        if cfg.use_ion:                      # compute_Jion (op.py:2789-2820)
            solver.compute_Jion(var, atm)
            ibr = [(sp, b) for sp in sorted(var.ion_sp) for b in range(1, var.ion_branch[sp] + 1)]
            out["Jion%d" % it] = np.array([var.Jion_sp[b] for b in ibr])
            out["kion%d" % it] = np.array([np.asarray(var.k[var.ion_rate_index[b]]) for b in ibr])
    for n, v in keep.items():
        setattr(var, n, v)
    var.k.update(keep_k)
    var.aflux_change = keep_change
    path = os.path.join(GOLD, "%s_photo%04d.npz" % (tag, count))
    np.savez_compressed(path, **out)
    print("wrote", path, "aflux_change=%.3e,%.3e" % (out["aflux_change1"], out["aflux_change2"]), flush=True)


def hook_conden(s, tag, counts, store):
    """L6: wrap the reference's own condensation operators (bound methods of op.Integration) and record what they read and
    write at the counts in `counts`, plus every call's (count, name) in order.  Nothing numerical is altered."""
    var0, atm, cfg, integ = s.var, s.atm, s.cfg, s.integ
    sp_idx = s.species.index
    st = store
    st["calls"] = []
    st["static"] = dict(
        condense_sp=np.array(list(cfg.condense_sp)), non_gas_sp=np.array(list(cfg.non_gas_sp)),
        use_relax=np.array(list(cfg.use_relax) if cfg.use_relax else [], dtype="U8"), humidity=float(getattr(cfg, "humidity", 1.0)),
        conden_re_list=np.array(list(var0.conden_re_list), dtype=int),
        conden_Rf=np.array([var0.Rf[re] for re in var0.conden_re_list]),
        sat_p=np.array([atm.sat_p[sp] for sp in cfg.condense_sp]), sat_mix=np.array([atm.sat_mix[sp] for sp in cfg.condense_sp]),
        r_p_keys=np.array(list(atm.r_p.keys())), r_p=np.array([atm.r_p[k] for k in atm.r_p.keys()]),
        rho_p_keys=np.array(list(atm.rho_p.keys())), rho_p=np.array([atm.rho_p[k] for k in atm.rho_p.keys()]),
        fix_species=np.array(list(cfg.fix_species)), pco=atm.pco.copy(),
        fix_species_from_coldtrap_lev=bool(getattr(cfg, "fix_species_from_coldtrap_lev", False)),
    )

    def wrap(name):
        orig = getattr(integ, name)

        def f(var, atm_):
            c = s.para.count
            nprev = sum(1 for _, n in st["calls"] if n == name)
            st["calls"].append((c, name))
            cap = c in counts or nprev < 2 or (c % 400 == 0 and nprev < 4000) or st.get("switch_count") == c
            # the first two calls, every 400th count, and the iteration of the fix_species switch
            if cap:
                pre = "c%05d_%s_" % (c, name)
                st[pre + "y_in"], st[pre + "ymix_in"], st[pre + "dt"], st[pre + "t"] = var.y.copy(), var.ymix.copy(), var.dt, var.t
                st[pre + "Dzz"] = atm_.Dzz.copy()
            v = orig(var, atm_)
            if name == "conden" and cfg.fix_species and v.t > cfg.stop_conden_time and not s.para.fix_species_start:
                # Integration.__call__ performs the fix_species switch right after this call (op.py:862-893): keep its input
                st["switch_count"], st["switch_t"], st["switch_y"], st["switch_ymix"] = c, v.t, v.y.copy(), v.ymix.copy()
                st["switch_vs_before"], st["switch_rtol_before"] = atm_.vs.copy(), float(cfg.rtol)
            if cap:
                st[pre + "y_out"], st[pre + "ymix_out"] = v.y.copy(), v.ymix.copy()
                if name == "conden":
                    st[pre + "k_rows"] = np.array([[np.broadcast_to(np.asarray(v.k[re + d], dtype=float), (s.nz,)) for d in (0, 1)]
                                                    for re in v.conden_re_list])
            return v
        setattr(integ, name, f)

    for n in ("conden", "h2o_conden_evap_relax", "nh3_conden_evap_relax"):
        wrap(n)


def finish_conden(s, tag, store, traj):
    """after the run: the fix_species switch as the reference left it (op.py:862-893)."""
    var, atm, para, cfg = s.var, s.atm, s.para, s.cfg
    out = {k: v for k, v in store.items() if k not in ("calls", "static")}
    out.update({"static_" + k: v for k, v in store["static"].items()})
    out["calls_count"] = np.array([c for c, n in store["calls"]], dtype=int)
    out["calls_name"] = np.array([n for c, n in store["calls"]])
    out["fix_species_start"] = bool(para.fix_species_start)
    if para.fix_species_start:
        out["fix_y"] = np.array([var.fix_y[sp] for sp in cfg.fix_species])
        # fix_species_from_coldtrap_lev = False: the whole column is fixed and conden_min_lev is only filled for the condensing gases
        out["conden_min_lev"] = np.array([int(atm.conden_min_lev.get(sp, s.nz)) for sp in cfg.fix_species])
        out["vs_after"] = atm.vs.copy()
        out["rtol_after"] = float(cfg.rtol)
    path = os.path.join(GOLD, "%s_conden.npz" % tag)
    np.savez_compressed(path, **out)
    print("wrote", path, "calls=%d fix_species_start=%s" % (len(store["calls"]), para.fix_species_start), flush=True)


def run(config, refdir, steps, full, max_steps=None, conden=None, after_switch=None, tag=None):
    sys.path.insert(0, HERE)
    import ref_session
    # fixtures are pinned to hash seed 0; `--tag` runs measure the reference's own seed-to-seed spread under another name
    assert os.environ.get("PYTHONHASHSEED") == "0" or tag is not None, "run with PYTHONHASHSEED=0 (SURVEY.md §8c)"
    s = ref_session.setup(refdir)
    os.makedirs(GOLD, exist_ok=True)
    tag = tag or config
    var, atm, para, integ, solver, cfg = s.var, s.atm, s.para, s.integ, s.solver, s.cfg
    sd = static_dict(s)
    s.k_static = sd["k"]
    np.savez_compressed(os.path.join(GOLD, "%s_static.npz" % tag), **sd)
    print("wrote static", flush=True)
    steps = sorted(steps)
    traj = []
    orig_one_step = solver.one_step
    state = dict(t0=time.time())

    def hooked(var_, atm_, para_):
        c = para_.count
        if after_switch is not None and para_.fix_species_start:
            state.setdefault("switch", c)                  # first one_step with the fixed-species rows active
            if c == state["switch"] + after_switch:
                s.var, s.atm, s.para = var_, atm_, para_
                capture_step(s, tag, c)
                var_, para_ = s.var, s.para
                print("captured step %d (fix_species switch seen at %d)" % (c, state["switch"]), flush=True)
                cfg.count_max = c + 1          # nothing more to record: let Integration.stop end the run (op.py:1080)
        if c in steps:
            s.var, s.atm, s.para = var_, atm_, para_
            capture_step(s, tag, c)
            if cfg.use_photo and c in (steps[0], steps[-1]):
                capture_photo(s, tag, c)
            var_, para_ = s.var, s.para
        n0 = para_.delta_count + para_.nega_count + para_.loss_count
        dt_try = var_.dt
        v, p = orig_one_step(var_, atm_, para_)
        traj.append((c, v.t, dt_try, v.dt, p.delta, p.delta_count + p.nega_count + p.loss_count - n0))
        return v, p

    solver.one_step = hooked
    cstore = {}
    if conden is not None:
        hook_conden(s, tag, set(conden), cstore)
    last = max(steps) if steps else 0
    if not full:
        cfg.count_max = last          # Integration.stop: count > count_max (op.py:1080)
    elif max_steps:
        cfg.count_max = max_steps
    integ(var, atm, para, s.make_atm)
    wall = time.time() - state["t0"]
    if conden is not None:
        finish_conden(s, tag, cstore, traj)
    if full:
        tr = np.array(traj)
        np.savez_compressed(os.path.join(GOLD, "%s_full.npz" % tag), traj=tr, y=var.y, ymix=var.ymix, t=var.t,
                            count=para.count, delta_count=para.delta_count, nega_count=para.nega_count,
                            loss_count=para.loss_count, end_case=para.end_case, wall_s=wall,
                            longdy=var.longdy, longdydt=var.longdydt,
                            traj_cols=np.array(["count", "t_before", "dt_try", "dt_used", "delta", "n_reject"]))
        print("full run: %d steps, %d rejected, t=%.4e, wall %.1f s" % (
            para.count, para.delta_count + para.nega_count + para.loss_count, var.t, wall), flush=True)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="HD189")
    ap.add_argument("--refdir", default=None)
    ap.add_argument("--steps", default="0,10,100,300")
    ap.add_argument("--full", action="store_true")
    ap.add_argument("--max-steps", type=int, default=None)
    ap.add_argument("--conden", default=None, help="comma list of counts at which the condensation operators are captured (L6)")
    ap.add_argument("--after-switch", type=int, default=None, help="capture an L0-L2 step fixture this many steps after fix_species starts")
    ap.add_argument("--tag", default=None, help="output name instead of the config name (self-spread runs with another PYTHONHASHSEED)")
    a = ap.parse_args()
    steps = [int(x) for x in a.steps.split(",") if x != ""]
    conden = None if a.conden is None else [int(x) for x in a.conden.split(",") if x != ""]
    run(a.config, a.refdir or "/tmp/vulcan_ref_%s" % a.config, steps, a.full, a.max_steps, conden, a.after_switch, a.tag)
