import sys, os
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
from helpers import Case, GOLD
from test_gpu_parity import _columns
c = Case("HD189", 10); cfg = c.cfg
col = _columns(c, 1, refine=1)
col.ens_setup(cfg["rtol"], cfg["loss_eps"], cfg["dt_min"], cfg["dt_max"], cfg["dt_var_min"], cfg["dt_var_max"], cfg["pos_cut"], cfg["nega_cut"], c.st["compo"], c.st["atom_ini"], c.st["n_0"])
col.ens_set_state(c.y, c.dt)
col.ens_run(1)
s = col.ens_get_state()
print({k: v for k, v in s.items() if k != 'y'})
sol, ym, delta, status = col.ros2_solve(c.y, c.ymix, c.dt)
print('host-path delta', delta, status, 'ref', float(c.fx['delta']))
cl = col.clip_loss(sol, ym, c.st["compo"], cfg["pos_cut"], cfg["nega_cut"])
print('anyneg', cl['any_negative'], 'loss', (cl['atom_sum']-c.st['atom_ini'])/c.st['atom_ini'], 'rtol', cfg['rtol'], cfg['loss_eps'])
tr = np.load(GOLD + "/HD189_full.npz")["traj"]
col.ens_set_state(c.y, c.dt)
for n in range(10, 60):
    col.ens_run(1)
    s = col.ens_get_state(want_y=False)
    print(n, 'dt_next %.6e ref %.6e  rel %.1e  t %.6e ref %.6e  acc %d rej %d' % (s['dt'][0], tr[n+1,2], abs(s['dt'][0]-tr[n+1,2])/tr[n+1,2], s['t'][0], tr[n+1,1]-tr[10,1], s['n_accept'][0], s['n_reject'][0]))
