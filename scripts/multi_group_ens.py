"""Experiment: the device-resident ensemble loop on G column groups (one vk_column handle + stream each, driven from G host
threads) - do kernels of different groups (FP64-bound factor, shared-memory-bound lhs, HBM-bound solves) overlap?
usage: python scripts/multi_group_ens.py [ncol] [steps] [G ...]"""
import os, sys, time, threading
import numpy as np
REPO = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, "tests"))
import torch
import bench
from vulcan_b200 import ensemble

ncol = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
Gs = [int(x) for x in sys.argv[3:]] or [1, 2, 4]
stagger = float(os.environ.get("STAGGER_MS", "0"))
case = bench.load_case()
cfg, st = case.cfg, case.st
y, atom_ini, kzz, kw = bench.build_columns(case, 0, ncol)
dt0 = float(cfg["dttry"])
for G in Gs:
    bounds = [ensemble.partition(ncol, G, g) for g in range(G)]
    runners = [ensemble.EnsembleRunner(case.net, case.nz, y[lo:hi], np.full(hi - lo, dt0), dict(kw), kzz[lo:hi], case.k, cfg,
                                       st["compo"], atom_ini[lo:hi], st["n_0"], device=0) for lo, hi in bounds]

    def go(n):
        def one(g):
            if stagger:
                time.sleep(1e-3 * stagger * g)
            runners[g].run(n)
        th = [threading.Thread(target=one, args=(g,)) for g in range(G)]
        for t in th: t.start()
        for t in th: t.join()
    go(3)
    for r, (lo, hi) in zip(runners, bounds):
        r.col.ens_set_state(y[lo:hi], np.full(hi - lo, dt0))
    go(1)
    torch.cuda.synchronize()
    t0 = time.time()
    go(steps)
    torch.cuda.synchronize()
    wall = time.time() - t0
    per = [r.col.last_kernel_ms()[0] / steps for r in runners]
    acc = sum(int(r.state(want_y=False)["n_accept"].sum()) for r in runners)
    print("G=%d: %.2f ms per ensemble step (wall), %.0f column-steps/s; per-group stream ms/step: %s; accepted %d" %
          (G, 1e3 * wall / steps, ncol * steps / wall, ["%.1f" % p for p in per], acc), flush=True)
    for r in runners:
        r.col.close()
    del runners
