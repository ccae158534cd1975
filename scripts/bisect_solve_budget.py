"""GPU diagnostic: element budget of the device block-tridiagonal solve on the reference's production-dt systems for
build variants of vk_solve.cu - FMA contraction off, exact 1/x in the 8 x 8 panel inverse - next to LAPACK and the C oracle.
    python scripts/bisect_solve_budget.py            (spawns one process per variant; builds happen in the build container beforehand:
    python scripts/bisect_solve_budget.py --build)"""
import os, subprocess, sys
REPO = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, "tests")); sys.path.insert(0, os.path.join(REPO, "oracle"))
VARIANTS = {"nofma": ["-fmad=false"], "exactrcp": ["-DVK_EXACT_RCP"], "nofma_exactrcp": ["-fmad=false", "-DVK_EXACT_RCP"]}

if "--build" in sys.argv:
    from vulcan_b200 import build
    for v, fl in VARIANTS.items():
        print(build.build(variant=v, solve_flags=fl))
    sys.exit(0)

if "--one" in sys.argv:
    import numpy as np
    from helpers import Case, gpu_columns
    from oracle import Oracle
    for tag, step in [("HD209S", 150), ("HD209S", 400), ("HD189", 300)]:
        c = Case(tag, step); o = Oracle(c.net)
        atm = o.make_atm(**c.atm_kwargs())
        D, up, dn = o.lhs(atm, c.y, c.k, c.dt)
        rhs = c.fx["chemdf"] + c.fx["diffdf"]
        xt = o.blocktri_truth(D, up, dn, rhs, 3)
        compo = c.st["compo"]; tot = (c.y[:, :, None] * compo[None]).sum(axis=(0, 1))
        bud = lambda v: (v[:, :, None] * compo[None]).sum(axis=(0, 1)) / tot
        col = gpu_columns(c)
        line = "%-16s %s-%d dt %.1e: LAPACK %.1e oracle %.1e | device" % (os.environ.get("VK_VARIANT", "product"), tag, step, c.dt,
                np.abs(bud(c.fx["k1"]) - bud(xt)).max(), np.abs(bud(o.blocktri_solve(o.blocktri_factor(D, up, dn), up, dn, rhs)) - bud(xt)).max())
        for rf in (0, 1, 2, 3):
            x, st = col.blocktri_solve(D, up, dn, rhs, refine=rf)
            e = np.abs(bud(x[0]) - bud(xt))
            line += " r%d %.1e" % (rf, e.max())
            if rf == 0:
                # where does the budget error sit?  element-weighted error of x per layer
                per = np.abs(((x[0] - xt)[:, :, None] * compo[None]).sum(axis=1) / tot[None]).max(axis=1)
                top = np.argsort(per)[-3:][::-1]
                line += " (layers %s: %s)" % (top.tolist(), ["%.1e" % per[q] for q in top])
        print(line, flush=True)
    sys.exit(0)

libdir = os.path.join(REPO, "vulcan_b200", "_lib")
for v in [None] + list(VARIANTS):
    env = dict(os.environ, VK_VARIANT=v or "product")
    if v:
        env["VK_LIB_PATH"] = os.path.join(libdir, "libvulcan_b200_%s.so" % v)
    subprocess.run([sys.executable, os.path.abspath(__file__), "--one"], env=env)
