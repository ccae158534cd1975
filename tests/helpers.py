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


from vulcan_b200.fixtures import Case, NETWORK_OF, have, load_network, photo_tables, steady_ensemble_from_fixture  # noqa: E402,F401


# every (config, step) fixture pair generated from the unmodified reference (oracle/dump_fixtures.py); BASELINE.json configs
CASES = [("HD189", 0), ("HD189", 10), ("HD189", 100), ("HD189", 300), ("Jupiter", 0), ("Jupiter", 30), ("Earth", 0), ("Earth", 30),
         ("HD209S", 0), ("HD209S", 30)]
CASES += [p for p in [("HD209S", 150), ("HD209S", 400)] if have(p[0], "step%04d.npz" % p[1])]     # production dt (1.8e4 s, 2.4e5 s)
# use_vm_mol variants: diffdf_vm / lhs_jac_tot_vm (HD189vm), diffdf_settling_vm / lhs_jac_settling_vm (JupiterVm),
# the latter with the diffusion-limited escape term (EarthVm)
VM_CASES = [p for p in [("HD189vm", 0), ("HD189vm", 30), ("JupiterVm", 0), ("JupiterVm", 30), ("EarthVm", 0), ("EarthVm", 30)]
            if have(p[0], "step%04d.npz" % p[1])]
# use_ion on the ion test network (oracle/stage_reference.py::write_ion_test_network): electron rows, charge balance, compute_Jion
ION_CASES = [p for p in [("HD189ion", 0), ("HD189ion", 30)] if have(p[0], "step%04d.npz" % p[1])]
# Jupiter after the fix_species switch (fixture-only cfg variant JupiterFix): fixed-species rows inside Ros2.solver
import glob as _glob
FIX_CASES = [("JupiterFix", int(os.path.basename(f)[len("JupiterFix_step"):-4])) for f in sorted(_glob.glob(os.path.join(GOLD, "JupiterFix_step*.npz")))]
CASES = CASES + VM_CASES + ION_CASES + FIX_CASES
# use_moldiff = False (diffdf_no_mol / lhs_jac_no_mol, op.py:1438-1494, 2122-2166).  The fixture sets below were recorded at the end of
# round 1 in a GPU-less session and pinned the oracle only; since round 2 they are part of CASES / LOCKSTEP, i.e. they run on the device
# in the -m gpu tests as well (cfg-switch branches _no_mol / vz != 0, padded block sizes 48 and 120, fix_species over whole columns).
NOMOL_CASES = [p for p in [("HD189nomol", 0), ("HD189nomol", 30)] if have(p[0], "step%04d.npz" % p[1])]
# the Earth cfg with the network it actually names (cfg_examples/vulcan_cfg_Earth.py:11: SNCHO_full_photo_network.txt, ni = 99, nr = 1284;
# H2SO4 condensation, sulphur boundary fluxes): the largest shipped cfg, padded block size 120.
# use_vz = True with a sign-changing vertical wind (fixture-only atm file, oracle/stage_reference.py::write_vz_test_atm): the upwind advection
# terms of every stencil variant, which no shipped cfg switches on.
NOMOL_CASES += [p for p in [("HD189vz", 0), ("HD189vz", 30), ("JupiterVz", 0), ("JupiterVz", 30), ("JupiterVmVz", 0), ("JupiterVmVz", 30)]
                if have(p[0], "step%04d.npz" % p[1])]         # diffdf / _settling / _settling_vm with vz != 0
# fix_species with fix_species_from_coldtrap_lev = False: whole columns of the fixed species are replaced rows (op.py:2898-2899, 2962-2963)
NOMOL_CASES += [p for p in [("JupiterFixAll", 153)] if have(p[0], "step%04d.npz" % p[1])]
# the smallest shipped photochemical network (CHO_photo_network.txt, ni = 41: padded block size 48)
NOMOL_CASES += [p for p in [("HD189cho", 0), ("HD189cho", 30)] if have(p[0], "step%04d.npz" % p[1])]
# thermochemistry only (NCHO_thermo_network.txt: no photo section, use_photo = False)
NOMOL_CASES += [p for p in [("HD189thermo", 0), ("HD189thermo", 30)] if have(p[0], "step%04d.npz" % p[1])]
NOMOL_CASES += [p for p in [("EarthS", 0), ("EarthS", 30), ("EarthS", 100), ("EarthS", 300)] if have(p[0], "step%04d.npz" % p[1])]   # dt 1e-10 ... 1e4 s
CASES = CASES + NOMOL_CASES
PHOTO_CASES = [("HD189", 0), ("HD189", 300), ("Jupiter", 0), ("Jupiter", 30), ("Earth", 0), ("Earth", 30), ("HD209S", 0), ("HD209S", 30)]
PHOTO_CASES += [p for p in [("EarthS", 0), ("EarthS", 30)] if have(p[0], "photo%04d.npz" % p[1])]


def case_id(p):
    return "%s-%d" % p


def step_opts(case):
    """The vulcan_cfg / para state consulted inside Ros2.solver (op.py:2896-2970) as plain arrays: shared by the oracle call
    and the C-ABI call (vulcan_b200/ros2.py::_sync_opts builds the same from the live reference objects)."""
    cfg, sp, ni = case.cfg, list(case.net.species), case.ni
    fix_bot = cfg.get("use_fix_sp_bot") or {}
    fbi = [sp.index(s) for s in fix_bot.keys()]
    fbm = np.array([fix_bot[s] for s in fix_bot.keys()], dtype=float)
    dz = None
    if cfg.get("use_condense"):
        dz = np.zeros(ni, dtype=np.uint8)
        for s in list(cfg.get("non_gas_sp", [])) + list(cfg.get("condense_sp", [])):
            dz[sp.index(s)] = 1
    fix_mask = fix_y = None
    if bool(case.fx["fix_species_start"]):       # op.py:2896-2906, 2921-2924, 2960-2970: rows of the fixed species below their cold trap
        fix_mask = np.zeros((case.nz, ni), dtype=np.uint8)
        fix_y = np.zeros((case.nz, ni))
        for q, s in enumerate(case.fx["fix_species"]):
            i = sp.index(str(s))
            top = case.nz if not bool(case.fx["fix_from_coldtrap"]) else int(case.fx["conden_min_lev"][q])
            fix_mask[:top, i] = 1
            fix_y[:top, i] = case.fx["fix_y"][q][:top]
            dz[i] = 1
    if cfg.get("use_ion"):       # atm.fix_e_indx rows (store.py:157, op.py:2908-2911, 2926): the solve leaves the electrons untouched
        ie = sp.index("e")
        fix_mask = np.zeros((case.nz, ni), dtype=np.uint8)
        fix_y = np.zeros((case.nz, ni))
        fix_mask[:, ie] = 1
        fix_y[:, ie] = case.y[:, ie]
        dz = np.zeros(ni, dtype=np.uint8) if dz is None else dz
        dz[ie] = 1
    return dict(fix_bot_idx=fbi, fix_bot_mix=fbm, n0_bot=float(case.st["n_0"][0]),
                zero_delta_row0=bool(cfg.get("use_botflux") or fix_bot), delta_zero_sp=dz,
                gas_indx_mix=case.gas_indx if cfg.get("non_gas_sp") else None, fix_mask=fix_mask, fix_y=fix_y)


def charge_balance(case, sol):
    """op.py:2998-3004: [e] from charge neutrality after the step (host bookkeeping on one nz-vector)"""
    if not case.cfg.get("use_ion"):
        return sol
    sp = list(case.net.species)
    sol = sol.copy()
    ie = sp.index("e")
    sol[..., ie] = 0
    for s in case.st["charge_list"]:
        i = sp.index(str(s))
        sol[..., ie] -= case.st["charge"][i] * sol[..., i]
    return sol


def oracle_step(case, oracle, atm, refine=0):
    o = step_opts(case)
    res = oracle.ros2_solver(atm, case.y, case.ymix, case.k, case.dt, case.cfg["mtol"], case.cfg["atol"], refine=refine, compo=case.st["compo"], **o)
    res["sol"] = charge_balance(case, res["sol"])
    return res


def gpu_columns(case, ncol=1, refine=0, rhs_order=0):
    """a vk_column handle configured like the reference's solver object for this fixture (through the C ABI)."""
    from vulcan_b200 import _abi
    kw = case.atm_kwargs()
    net = _abi.DeviceNetwork(case.net)
    col = _abi.Columns(net, case.nz, ncol)
    col.set_atm(Kzz=kw["Kzz"], vz=kw["vz"], dzi=kw["dzi"], Dzz=kw["Dzz"], vs=kw["vs"], Tco=kw["Tco"], g=kw["g"], M=kw["M"],
                Ti=kw["Ti"], Hpi=kw["Hpi"], ms=kw["ms"], alpha=kw["alpha"], top_flux=kw["top_flux"], bot_flux=kw["bot_flux"],
                bot_vdep=kw["bot_vdep"], use_moldiff=kw["use_moldiff"], use_settling=kw["use_settling"],
                use_topflux=kw["use_topflux"], use_botflux=kw["use_botflux"], gas_indx=kw["gas_indx"],
                gas_indx_lhs=kw["gas_indx_lhs"], use_vm_mol=kw["use_vm_mol"], vm=kw["vm"], diff_esc_idx=kw["diff_esc_idx"],
                shared=True)
    col.set_k(case.k)
    o = step_opts(case)
    fbv = None
    if len(o["fix_bot_idx"]):
        fbv = np.repeat((o["fix_bot_mix"] * o["n0_bot"])[None], ncol, axis=0)
    rep = lambda a: None if a is None else np.repeat(a[None], ncol, axis=0)
    col.set_step_opts(case.cfg["mtol"], case.cfg["atol"], refine=refine, zero_delta_row0=o["zero_delta_row0"],
                      fix_bot_idx=o["fix_bot_idx"], fix_bot_val=fbv, delta_zero_sp=o["delta_zero_sp"],
                      fix_mask=rep(o["fix_mask"]), fix_y=rep(o["fix_y"]), compo=case.st["compo"], rhs_order=rhs_order)
    return col


def mock_objects(case, with_photo=True):
    """Stand-ins for the reference's vulcan_cfg module and store.Variables / AtmData / Parameters containers (store.py:21-209)
    filled from a fixture: the GPU box has no /root/reference, and the drop-in class only reads attributes."""
    from types import SimpleNamespace
    st, fx, cfgd = case.st, case.fx, case.cfg
    cfg = SimpleNamespace(**cfgd)
    cfg.use_fix_sp_bot = {} if not isinstance(cfgd.get("use_fix_sp_bot"), dict) else cfgd["use_fix_sp_bot"]
    for name, default in (("non_gas_sp", []), ("condense_sp", []), ("fix_species", []), ("remove_list", []), ("T_cross_sp", []),
                          ("diff_esc", []), ("use_relax", [])):
        if not isinstance(getattr(cfg, name, None), list):
            setattr(cfg, name, default)
    atoms = cfg.atom_list
    var = SimpleNamespace(y=case.y.copy(), ymix=case.ymix.copy(), dt=case.dt, t=float(fx["t"]), y_prev=case.y.copy(),
                          k={i: case.k_rz[i].copy() for i in range(1, case.nr + 1)},
                          atom_ini={a: float(st["atom_ini"][q]) for q, a in enumerate(atoms)}, atom_sum={},
                          atom_loss={a: float(fx["atom_loss_in"][q]) for q, a in enumerate(atoms)},
                          atom_loss_prev={a: float(fx["atom_loss_prev"][q]) for q, a in enumerate(atoms)},
                          dy=1., dydt=1., dy_prev=1., longdy=1., longdydt=1., aflux_change=0., y_time=[], t_time=[],
                          atom_loss_time=[])
    atm = SimpleNamespace(Kzz=st["Kzz"].copy(), vz=st["vz"].copy(), dzi=fx["dzi"].copy(), Dzz=st["Dzz"].copy(), vs=fx["vs_dyn"].copy(),
                          Tco=st["Tco"].copy(), g=fx["g"].copy(), M=st["M"].copy(), Ti=fx["Ti"].copy(), Hpi=fx["Hpi"].copy(),
                          ms=st["ms"].copy(), alpha=st["alpha"].copy(), top_flux=fx["top_flux_dyn"].copy(), bot_flux=st["bot_flux"].copy(),
                          bot_vdep=st["bot_vdep"].copy(), gas_indx=list(st["gas_indx"]), n_0=st["n_0"].copy(), dz=fx["dz"].copy(),
                          pico=st["pico"].copy(), pco=st["pco"].copy(), pref_indx=int(st["pref_indx"]), gs=float(cfgd["gs"]),
                          Hp=fx["Hp"].copy(), zco=fx["zco"].copy(), mu=fx["mu"].copy(), vm=st["vm"].copy())
    if not cfgd["use_moldiff"]:
        # what the reference's atm object looks like then (found by running the drop-in class inside the unmodified reference): vulcan.py
        # never calls mol_diff, so Ti / Hpi do not exist (build_atm.py:569-571) and ms is np.empty garbage (store.py:129)
        del atm.Ti, atm.Hpi
        atm.ms = np.full(case.ni, np.nan)
    para = SimpleNamespace(delta=0.0, small_y=0.0, nega_y=0.0, delta_count=0, nega_count=0, loss_count=0, count=int(fx["count"]),
                           fix_species_start=False, solver_str="", end_case=0, switch_final_photo_frq=False)
    if with_photo and cfgd.get("use_photo") and "photo_sp" in st:
        psp = [str(s) for s in st["photo_sp"]]
        var.photo_sp = set(psp)
        var.ion_sp = set()
        var.bins, var.sflux_top = st["bins"], st["sflux_top"]
        var.sflux_din12_indx, var.dbin1, var.dbin2 = int(st["sflux_din12_indx"]), float(st["dbin1"]), float(st["dbin2"])
        var.cross = {s: st["cross"][i] for i, s in enumerate(psp)}
        var.cross_scat = {s: st["cross_scat"][i] for i, s in enumerate(cfg.scat_sp)}
        var.n_branch, var.cross_J, var.pho_rate_index = {}, {}, {}
        for q in range(len(st["branch_sp"])):
            s, b = psp[int(st["branch_sp"][q])], int(st["branch_no"][q])
            var.n_branch[s] = max(var.n_branch.get(s, 0), b)
            var.cross_J[(s, b)] = st["cross_J"][q]
            var.pho_rate_index[(s, b)] = int(st["branch_rate_index"][q])
        if "ion_sp" in st:                      # use_ion: op.py:253-269, 617-618; build_atm.py:148-161, 271-277
            isp = [str(s) for s in st["ion_sp"]]
            var.ion_sp = set(isp)
            var.cross.update({s: st["ion_cross"][i] for i, s in enumerate(isp)})
            var.ion_branch, var.cross_Jion, var.ion_rate_index = {}, {}, {}
            for q in range(len(st["ion_branch_sp"])):
                s, b = isp[int(st["ion_branch_sp"][q])], int(st["ion_branch_no"][q])
                var.ion_branch[s] = max(var.ion_branch.get(s, 0), b)
                var.cross_Jion[(s, b)] = st["cross_Jion"][q]
                var.ion_rate_index[(s, b)] = int(st["ion_branch_rate_index"][q])
        if "T_cross_sp" in st:
            tsp = [str(x) for x in st["T_cross_sp"]]
            var.cross_T = {s: st["cross_T"][q] for q, s in enumerate(tsp)}
            var.cross_J_T = {}
            for q, b in enumerate(st["cross_J_T_branch"]):
                b = int(b)
                var.cross_J_T[(psp[int(st["branch_sp"][b])], int(st["branch_no"][b]))] = st["cross_J_T"][q]
    if "charge_list" in st:
        var.charge_list = [str(s) for s in st["charge_list"]]
    return cfg, var, atm, para


def attach_conden(case, cfg, var, atm):
    """add what the condensation operators of the caller read (op.py:1109-1421, 862-893) from <cfg>_conden.npz (recorded from
    the reference's live objects by oracle/dump_fixtures.py --conden): saturation pressures, particle sizes / densities, the
    condensation reaction list.  Returns the fixture dict (None when the config does not condense)."""
    path = os.path.join(GOLD, "%s_conden.npz" % case.tag)
    if not os.path.exists(path):       # cfg-switch variants share the base config's saturation curves / particle tables
        path = os.path.join(GOLD, "%s_conden.npz" % NETWORK_OF.get(case.tag, case.tag))
    if not os.path.exists(path):       # Jupiter: the tables were recorded with the JupiterFix timing variant of the same cfg
        path = os.path.join(GOLD, "%sFix_conden.npz" % case.tag)
    if not os.path.exists(path):
        return None
    cf = dict(np.load(path, allow_pickle=False))
    csp = [str(x) for x in cf["static_condense_sp"]]
    atm.sat_p = {s: cf["static_sat_p"][q] for q, s in enumerate(csp)}
    atm.sat_mix = {s: cf["static_sat_mix"][q] for q, s in enumerate(csp)}
    atm.r_p = {str(k): float(v) for k, v in zip(cf["static_r_p_keys"], cf["static_r_p"])}
    atm.rho_p = {str(k): float(v) for k, v in zip(cf["static_rho_p_keys"], cf["static_rho_p"])}
    atm.conden_min_lev = {}
    cfg.humidity = float(cf["static_humidity"])
    cfg.fix_species_from_coldtrap_lev = bool(cf["static_fix_species_from_coldtrap_lev"]) if "static_fix_species_from_coldtrap_lev" in cf \
        else True          # both shipped condensing cfgs set it (vulcan_cfg_Jupiter.py:121, vulcan_cfg_Earth.py:120)
    var.conden_re_list = [int(x) for x in cf["static_conden_re_list"]]
    var.Rf = {re: str(n) for re, n in zip(var.conden_re_list, cf["static_conden_Rf"])}
    return cf


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


def photolysis_via_dropin(tag, step, abi=None):
    """two consecutive photolysis updates (compute_tau / flux / J and, with use_ion, compute_Jion) through the drop-in solver object
    from a zeroed diffuse-flux state, as recorded in <cfg>_photo<step>.npz; returns (fixture, [var after update 1, 2] summaries)"""
    from vulcan_b200 import ros2 as ros2_mod
    from vulcan_b200.ros2 import Ros2
    real_abi = ros2_mod._abi
    if abi is not None:
        ros2_mod._abi = abi
    try:
        case = Case(tag, step)
        cfg, var, atm, para = mock_objects(case)
        px = dict(np.load(os.path.join(GOLD, "%s_photo%04d.npz" % (tag, step))))
        var.y, var.ymix = px["y"], px["ymix"]
        atm.dz = px["dz"]
        solver = Ros2(cfg=cfg, compo=case.st["compo"], network=case.net, charge=case.st["charge"] if "charge" in case.st else None)
        out = []
        for it in (1, 2):
            solver.compute_tau(var, atm)
            solver.compute_flux(var, atm)
            solver.compute_J(var, atm)
            if cfg.use_ion:
                solver.compute_Jion(var, atm)
            out.append(dict(aflux=var.aflux.copy(), aflux_change=var.aflux_change, k={i: np.array(v, dtype=float, copy=True) for i, v in var.k.items()},
                            Jion={b: v.copy() for b, v in getattr(var, "Jion_sp", {}).items()}))
        return case, px, out
    finally:
        ros2_mod._abi = real_abi


def run_config(tag, refine=-1, max_wall_s=600, count_max=None, abi=None, cfg_edit=None, device_loop=False, dt_schedule=None):
    """one BASELINE.json single-column config from the reference's initial state (fixture step 0) through the drop-in solver
    object and the Integration mirror until Integration.stop() says so (op.py:1067-1087).  `abi`: replaces the ctypes binding
    the solver object talks to (only the CPU host-logic tests pass the oracle-backed stand-in of tests/oracle_columns.py)."""
    import time
    from vulcan_b200 import ros2 as ros2_mod
    from integration_mirror import Integration
    from vulcan_b200.ros2 import Ros2
    real_abi = ros2_mod._abi
    if abi is not None:
        ros2_mod._abi = abi
    try:
        case = Case(tag, 0)
        cfg, var, atm, para = mock_objects(case)
        attach_conden(case, cfg, var, atm)
        if count_max is not None:
            cfg.count_max = count_max
        if cfg_edit:
            for name, val in cfg_edit.items():
                setattr(cfg, name, val)
        var.y = case.st["y_ini"].copy()
        if cfg.non_gas_sp:
            var.ymix = var.y / np.vstack(np.sum(var.y[:, atm.gas_indx], axis=1))
        else:
            var.ymix = var.y / np.vstack(np.sum(var.y, axis=1))
        solver = Ros2(cfg=cfg, species=list(case.st["species"]), compo=case.st["compo"], network=case.net, refine=refine,
                      charge=case.st["charge"] if "charge" in case.st else None)
        solver.naming_solver(para)
        if dt_schedule is not None:
            # REPLAY of the reference's own time grid: accepted step number c is taken with the step size the reference used for it
            # (traj column dt_used of <cfg>_full.npz), no accept / reject decision and no step-size control of our own - what remains
            # between the two final states is the arithmetic of the step alone, not the chaos of the controller
            sched = np.asarray(dt_schedule, dtype=float)

            def replay_step(var_, atm_, para_):
                var_.dt = float(sched[para_.count])
                var_, para_ = solver.solver(var_, atm_, para_)
                return solver.clip(var_, para_, atm_)
            solver.one_step = replay_step
            solver.step_size = lambda var_, para_: var_
            cfg.count_max = len(sched) - 1           # the run ends when the schedule does (Integration.stop: count > count_max) ...
            cfg.flux_cri = -1.0                      # ... and not earlier: conv() still evaluates longdy (the photolysis cadence switch reads
                                                     # it, op.py:818-822) but can never be satisfied (aflux_change < flux_cri, op.py:1058)
        # vulcan.py:170-176: one photolysis update at set-up, then the loop updates again at count 0
        if cfg.use_photo:
            solver.compute_tau(var, atm)
            solver.compute_flux(var, atm)
            solver.compute_J(var, atm)
        if device_loop:      # the device-resident loop (vulcan_b200/steady.py) instead of the host mirror of op.Integration
            from vulcan_b200.steady import DeviceIntegration
            integ = DeviceIntegration(solver)
            if not cfg.use_moldiff:
                solver._masses = lambda: case.st["ms"]
        else:
            integ = Integration(solver, cfg, case.net.species, mass=None if cfg.use_moldiff else case.st["ms"])
        t0 = time.time()
        var, atm, para = integ(var, atm, para, max_wall_s=max_wall_s)
        return case, var, atm, para, integ, time.time() - t0
    finally:
        ros2_mod._abi = real_abi           # never leak the stand-in into other tests of the same process
