"""TEST INFRASTRUCTURE — drives the UNMODIFIED reference through its own public API.

Runs inside a scratch copy produced by oracle/stage_reference.py (cwd must be that copy,
PYTHONPATH must contain its `_stubs`).  `setup()` performs the same call sequence as the
reference driver (vulcan.py:72-178) and returns the live objects so that fixtures can be
dumped (oracle/dump_fixtures.py) or the CPU path timed.  No reference source is copied:
every numerical routine executed here is the reference's own.
"""
import os
import sys
import time

import numpy as np


class Session(object):
    pass


def setup(refdir=None, solver_factory=None):
    """solver_factory(op, vulcan_cfg, chem_funs) -> solver object: replaces `getattr(op, vulcan_cfg.ode_solver)()` (vulcan.py:162-163),
    i.e. the ONE line INTEGRATION.md changes; everything else below is the reference's own call sequence."""
    if refdir is not None:
        os.chdir(refdir)
        sys.path.insert(0, refdir)
        sys.path.insert(0, os.path.join(refdir, "_stubs"))
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    import store, build_atm, op, chem_funs, vulcan_cfg  # noqa: E401  (reference modules)

    s = Session()
    s.store, s.build_atm, s.op, s.chem_funs, s.cfg = store, build_atm, op, chem_funs, vulcan_cfg
    s.species = chem_funs.spec_list
    s.ni, s.nr, s.nz = chem_funs.ni, chem_funs.nr, vulcan_cfg.nz

    var, atm, para = store.Variables(), store.AtmData(), store.Parameters()
    para.start_time = time.time()
    make_atm = build_atm.Atm()
    output = op.Output()
    atm = make_atm.f_pico(atm)
    atm = make_atm.load_TPK(atm)
    if vulcan_cfg.use_condense:
        make_atm.sp_sat(atm)
    rate = op.ReadRate()
    var = rate.read_rate(var, atm)
    if vulcan_cfg.use_lowT_limit_rates:
        var = rate.lim_lowT_rates(var, atm)
    var = rate.rev_rate(var, atm)
    var = rate.remove_rate(var)
    ini = build_atm.InitialAbun()
    var = ini.ini_y(var, atm)
    var = ini.ele_sum(var)
    atm = make_atm.f_mu_dz(var, atm, output)
    make_atm.mol_diff(atm)
    make_atm.BC_flux(atm)
    solver = getattr(op, vulcan_cfg.ode_solver)() if solver_factory is None else solver_factory(op, vulcan_cfg, chem_funs)
    if vulcan_cfg.use_photo:
        rate.make_bins_read_cross(var, atm)
        make_atm.read_sflux(var, atm)
        solver.compute_tau(var, atm)
        solver.compute_flux(var, atm)
        solver.compute_J(var, atm)
        var = rate.remove_rate(var)
    integ = op.Integration(solver, output)
    solver.naming_solver(para)
    s.var, s.atm, s.para = var, atm, para
    s.make_atm, s.output, s.rate, s.solver, s.integ = make_atm, output, rate, solver, integ
    return s


def pack_k(var, nr, nz):
    """var.k dict{1..nr -> (nz,) or scalar} -> dense [nr+1, nz] (row 0 unused)."""
    k = np.zeros((nr + 1, nz))
    for i in range(1, nr + 1):
        k[i] = np.broadcast_to(np.asarray(var.k[i], dtype=float), (nz,))
    return k
