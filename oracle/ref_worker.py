#!/usr/bin/env python
"""TEST INFRASTRUCTURE - one host process timing the UNMODIFIED reference (oracle/_ref/<cfg>, staged by oracle/build_ref.py).

Protocol on stdin / stdout (driven by bench.py --impl reference): after the reference's own set-up sequence (vulcan.py:72-178 through
oracle/ref_session.py) the worker prints READY, then for every line `GO n` runs n calls of the reference's `op.Ros2.solver` - one
ATTEMPTED Ros2 step each (chemdf x2, neg_symjac, lhs_jac_tot, store_bandM, solve_banded x2; op.py:2860-3007), always from the same
state so that every call is the same work - prints `DONE <seconds>`, and exits on QUIT.  OMP_NUM_THREADS = 1 as vulcan.py:52 forces.
"""
import os
import sys
import time

os.environ["OMP_NUM_THREADS"] = "1"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)


def private_view(refdir):
    """a per-process view of the staged reference: symlinks to everything except fastchem_vulcan/ (copied: FastChem reads its input
    from and writes its output into that directory, build_atm.py:84-131, so concurrent workers must not share it) and output/"""
    import atexit
    import shutil
    import tempfile
    view = tempfile.mkdtemp(prefix="vk_ref_worker_")
    atexit.register(shutil.rmtree, view, True)
    for name in os.listdir(refdir):
        src = os.path.join(refdir, name)
        if name == "fastchem_vulcan":
            shutil.copytree(src, os.path.join(view, name), symlinks=True)
        elif name in ("output", "plot", "__pycache__"):
            os.makedirs(os.path.join(view, name), exist_ok=True)
        else:
            os.symlink(src, os.path.join(view, name))
    return view


def main():
    refdir = private_view(os.path.abspath(sys.argv[1]))
    warm = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    import contextlib
    import io
    import ref_session
    # the protocol goes over a private copy of stdout; fd 1 itself is pointed at /dev/null (the reference's FastChem subprocess and its
    # own prints write there)
    real_stdout = os.fdopen(os.dup(1), "w")
    devnull = os.open(os.devnull, os.O_WRONLY)
    os.dup2(devnull, 1)
    with contextlib.redirect_stdout(io.StringIO()):
        s = ref_session.setup(refdir)
        var, atm, para, solver = s.var, s.atm, s.para, s.solver
        for _ in range(warm):                                   # a few accepted steps: leaves the dt = 1e-10 start, pages everything in
            var = s.integ.backup(var)                           # op.py:937-941 (y_prev, atom_loss_prev) as Integration.__call__ does
            var, para = solver.one_step(var, atm, para)
            var = solver.step_size(var, para)
    y0, ymix0, dt0 = var.y.copy(), var.ymix.copy(), var.dt
    print("READY", flush=True, file=real_stdout)
    for line in sys.stdin:
        tok = line.split()
        if not tok:
            continue
        if tok[0] == "QUIT":
            break
        n = int(tok[1])
        with contextlib.redirect_stdout(io.StringIO()):
            t0 = time.time()
            for _ in range(n):
                var.y, var.ymix, var.dt = y0.copy(), ymix0.copy(), dt0
                solver.solver(var, atm, para)
            el = time.time() - t0
        print("DONE %.6f" % el, flush=True, file=real_stdout)


if __name__ == "__main__":
    main()
