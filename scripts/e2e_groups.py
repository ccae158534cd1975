"""e2e (host buffers in/out) throughput of ensemble.PipelinedHostSolver vs the number of column groups, for a given batch size:
python scripts/e2e_groups.py 512 1 2 4 8 16"""
import os, sys, time
import numpy as np
REPO = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, "tests"))
import torch
import bench
from vulcan_b200 import ensemble
ncol = int(sys.argv[1]); Gs = [int(x) for x in sys.argv[2:]] or [1, 2, 4, 8, 16]
case = bench.load_case(); cfg = case.cfg
y, atom_ini, kzz, kw = bench.build_columns(case, 0, ncol)
nv = ncol * case.nz * case.net.ni
pin = [torch.empty(nv, dtype=torch.float64).pin_memory() for _ in range(4)]
hy, hm, hs, ho = [p.numpy() for p in pin]
hy[:] = y.ravel(); hm[:] = (y / y.sum(axis=2, keepdims=True)).ravel()
hdt = np.full(ncol, float(cfg["dttry"])); hdelta = np.empty(ncol); hstat = np.zeros(ncol, dtype=np.int32)
for G in Gs:
    host = ensemble.PipelinedHostSolver(case.net, case.nz, dict(kw), kzz, case.k, cfg, n_groups=G, device=0, refine=0)
    for _ in range(2): host.solve_into(hy, hm, hdt, hs, ho, hdelta, hstat)
    torch.cuda.synchronize(); t0 = time.time()
    for _ in range(8): host.solve_into(hy, hm, hdt, hs, ho, hdelta, hstat)
    dt = (time.time() - t0) / 8
    print("ncol %d groups %2d: %.2f ms per step, %.0f column-steps/s" % (ncol, G, 1e3 * dt, ncol / dt), flush=True)
    host.close()
