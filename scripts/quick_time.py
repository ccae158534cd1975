"""ad-hoc timing on the GPU box: single column and small ensembles (not the bench)."""
import sys, time, os
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
from helpers import Case
from vulcan_b200 import _abi

def make(case, ncol, refine):
    kw = case.atm_kwargs()
    net = _abi.DeviceNetwork(case.net)
    col = _abi.Columns(net, case.nz, ncol)
    col.set_atm(Kzz=kw["Kzz"], vz=kw["vz"], dzi=kw["dzi"], Dzz=kw["Dzz"], vs=kw["vs"], Tco=kw["Tco"], g=kw["g"], M=kw["M"],
                Ti=kw["Ti"], Hpi=kw["Hpi"], ms=kw["ms"], alpha=kw["alpha"], top_flux=kw["top_flux"], bot_flux=kw["bot_flux"],
                bot_vdep=kw["bot_vdep"], shared=True)
    col.set_k(case.k)
    col.set_step_opts(case.cfg["mtol"], case.cfg["atol"], refine=refine)
    return col

if __name__ == "__main__":
  c = Case("HD189", 100)
  for ncol in [1, 592, 888]:
      for refine in (0, 1):
          col = make(c, ncol, refine)
          y = np.repeat(c.y[None], ncol, 0); ym = np.repeat(c.ymix[None], ncol, 0); dt = np.full(ncol, c.dt)
          for _ in range(3): col.ros2_solve(y, ym, dt)
          t0 = time.time(); n = 5
          tot = fac = 0
          for _ in range(n):
              col.ros2_solve(y, ym, dt); a, b = col.last_kernel_ms(); tot += a; fac += b
          wall = (time.time() - t0) / n
          print("ncol %4d refine %d: device %.3f ms/step (factor %.3f ms)  wall(e2e host buffers) %.3f ms  -> %.1f col-steps/s device" % (ncol, refine, tot / n, fac / n, wall * 1e3, ncol / (tot / n) * 1e3), flush=True)
          col.close()
