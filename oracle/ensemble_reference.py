#!/usr/bin/env python
"""TEST INFRASTRUCTURE - sampled columns of the synthetic HD189-like ensemble (BASELINE.json configs[4]: the base state re-weighted in
metallicity and C/O, Kzz scaled) run to steady state by the UNMODIFIED reference: its own set-up sequence (vulcan.py:72-178 through
oracle/ref_session.py), its own op.Integration loop and op.Ros2 solver in oracle/_ref/HD189 (staged by oracle/build_ref.py).  Only the
INPUTS of the column change, exactly as vulcan_b200.ensemble.synthetic_columns / fixtures.steady_ensemble_from_fixture build them for the
device: var.y (re-weighted, renormalised to n_0), atm.Kzz (scaled); the grid (dz, zco, Hp, mu) starts from the base state like on the device.

    python oracle/ensemble_reference.py --index 3 --out /tmp/ens_ref/col3.npz      one column (one host core, ~15-25 min)
    python oracle/ensemble_reference.py --pack /tmp/ens_ref                         -> tests/golden/HD189_ens8_reference.npz

The fixture feeds tests/test_gpu_device_steady.py::test_ensemble_columns_against_reference_runs."""
import argparse
import contextlib
import glob
import io
import os
import sys
import time

import numpy as np

os.environ["OMP_NUM_THREADS"] = "1"
HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, REPO)

# (Kzz scale, metallicity scale, C/O) of the sampled columns; column 2 is the unmodified HD189 cfg.  Column 3 (C/O 0.8 at twice solar
# metallicity) had not converged after two hours of the reference and is not part of the fixture; column 8 was added in its place
PARAMS = np.array([[0.1, 1.0, 0.55], [0.3, 1.0, 0.55], [1.0, 1.0, 0.55], [1.0, 2.0, 0.8], [3.0, 1.0, 0.3], [10.0, 0.5, 0.55],
                   [0.5, 3.0, 1.0], [2.0, 0.3, 0.4], [5.0, 1.5, 0.45]])


def run_one(refdir, index, out):
    from ref_worker import private_view
    from vulcan_b200.ensemble import synthetic_columns
    view = private_view(os.path.abspath(refdir))
    import ref_session
    log = io.StringIO()
    devnull = os.open(os.devnull, os.O_WRONLY)
    keep = os.dup(1)
    os.dup2(devnull, 1)
    with contextlib.redirect_stdout(log):
        s = ref_session.setup(view)
        cfg, var, atm, para, solver = s.cfg, s.var, s.atm, s.para, s.solver
        species = list(s.species)
        compo = np.array([[s.build_atm.compo[s.build_atm.compo_row.index(sp)][a] for a in cfg.atom_list] for sp in species], dtype=float)
        kz, met, co = PARAMS[index]
        y, atom_ini = synthetic_columns(var.y, atm.n_0, compo, list(cfg.atom_list), [kz], [met], [co])
        var.y = y[0].copy()
        var.ymix = var.y / np.vstack(np.sum(var.y, axis=1))
        var = s.build_atm.InitialAbun().ele_sum(var)
        atm.Kzz = atm.Kzz * kz
        if cfg.use_photo:                                         # vulcan.py:170-176 on the column's own state
            solver.compute_tau(var, atm)
            solver.compute_flux(var, atm)
            solver.compute_J(var, atm)
            var = s.rate.remove_rate(var)
        t0 = time.time()
        s.integ(var, atm, para, s.make_atm)
        wall = time.time() - t0
    os.dup2(keep, 1)
    np.savez_compressed(out, index=index, params=PARAMS[index], y=var.y, ymix=var.ymix, t=var.t, dt=var.dt, count=para.count,
                        end_case=para.end_case, longdy=var.longdy, longdydt=var.longdydt, wall_s=wall, y_ini=y[0],
                        atom_ini=np.array([var.atom_ini[a] for a in cfg.atom_list]), atom_ini_device=atom_ini[0],
                        atom_loss=np.array([var.atom_loss[a] for a in cfg.atom_list]),
                        n_reject=para.delta_count + para.nega_count + para.loss_count)
    print("column %d (Kzz x %g, metallicity x %g, C/O %g): %d steps, t = %.4e s, end_case %d, longdy %.3e, %.0f s" % (
        index, kz, met, co, para.count, var.t, para.end_case, var.longdy, wall), flush=True)


def pack(src, src2=None):
    """src: one run per column; src2 (optional): second runs of some columns under another PYTHONHASHSEED (another summation order of tau /
    omega_0 inside the reference) - the reference's spread against ITSELF on those columns"""
    files = sorted(glob.glob(os.path.join(src, "col*.npz")))
    d = [dict(np.load(f)) for f in files]
    d.sort(key=lambda x: int(x["index"]))
    out = os.path.join(REPO, "tests", "golden", "HD189_ens8_reference.npz")
    extra = {}
    if src2:
        d2 = [dict(np.load(f)) for f in sorted(glob.glob(os.path.join(src2, "col*.npz")))]
        d2.sort(key=lambda x: int(x["index"]))
        pos = {int(x["index"]): q for q, x in enumerate(d)}
        extra = dict(seed2_column=np.array([pos[int(x["index"])] for x in d2]), seed2_ymix=np.array([x["ymix"] for x in d2]),
                     seed2_count=np.array([int(x["count"]) for x in d2]), seed2_t=np.array([float(x["t"]) for x in d2]))
    np.savez_compressed(out, **extra, params=np.array([x["params"] for x in d]), ymix=np.array([x["ymix"] for x in d]),
                        t=np.array([float(x["t"]) for x in d]), count=np.array([int(x["count"]) for x in d]),
                        end_case=np.array([int(x["end_case"]) for x in d]), longdy=np.array([float(x["longdy"]) for x in d]),
                        n_reject=np.array([int(x["n_reject"]) for x in d]), wall_s=np.array([float(x["wall_s"]) for x in d]),
                        atom_loss=np.array([x["atom_loss"] for x in d]), y_ini=np.array([x["y_ini"] for x in d]),
                        atom_ini=np.array([x["atom_ini"] for x in d]))
    print(out, os.path.getsize(out), "bytes;", len(d), "columns")


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--refdir", default=os.path.join(HERE, "_ref", "HD189"))
    ap.add_argument("--index", type=int)
    ap.add_argument("--out")
    ap.add_argument("--pack")
    ap.add_argument("--pack2", help="directory with second-seed runs of some columns")
    a = ap.parse_args()
    if a.pack:
        pack(a.pack, a.pack2)
    else:
        run_one(a.refdir, a.index, a.out)
