import numpy as np, pytest
from helpers import Case
pytestmark = pytest.mark.gpu
from test_gpu_parity import _columns


def test_ens_matches_host_loop():
    """the device-resident controller equals the same logic driven from the host through vk_ros2_solve / vk_clip_loss."""
    c = Case("HD189", 100)
    cfg = c.cfg
    ncol = 3
    dev = _columns(c, ncol, refine=0)
    host = _columns(c, ncol, refine=0)
    scale = np.array([1.0, 0.5, 8.0])              # different dt per column -> column 2 is rejected at least once
    y0 = np.repeat(c.y[None], ncol, 0)
    dt0 = c.dt * scale
    dev.ens_setup(cfg["rtol"], cfg["loss_eps"], cfg["dt_min"], cfg["dt_max"], cfg["dt_var_min"], cfg["dt_var_max"],
                  cfg["pos_cut"], cfg["nega_cut"], c.st["compo"], c.st["atom_ini"], c.st["n_0"])
    dev.ens_set_state(y0, dt0)
    nsteps = 6
    dev.ens_run(nsteps)
    s = dev.ens_get_state()
    # host replica
    y, dt = y0.copy(), dt0.copy()
    ymix = y / y.sum(axis=2, keepdims=True)
    t = np.zeros(ncol); nacc = np.zeros(ncol, int); nrej = np.zeros(ncol, int)
    loss_prev = np.zeros((ncol, c.st["compo"].shape[1]))
    n_0 = c.st["n_0"]
    for it in range(nsteps):
        sol, ym, delta, status = host.ros2_solve(y, ymix, dt)
        cl = host.clip_loss(sol, ym, c.st["compo"], cfg["pos_cut"], cfg["nega_cut"])
        loss = (cl["atom_sum"] - c.st["atom_ini"]) / c.st["atom_ini"]
        for i in range(ncol):
            ok = (not cl["any_negative"][i]) and np.max(np.abs(loss[i] - loss_prev[i])) < cfg["loss_eps"] and delta[i] <= cfg["rtol"]
            ymix[i] = cl["ymix"][i]
            if ok:
                t[i] += dt[i]; nacc[i] += 1; loss_prev[i] = loss[i]
                y[i] = n_0[:, None] * cl["ymix"][i]
                d = delta[i] if delta[i] != 0 else 0.01 * cfg["rtol"]
                hf = min(max(0.9 * np.sqrt(cfg["rtol"] / d), cfg["dt_var_min"]), cfg["dt_var_max"])
                dt[i] = min(max(dt[i] * hf, cfg["dt_min"]), cfg["dt_max"])
            else:
                nrej[i] += 1; dt[i] *= cfg["dt_var_min"]
    assert np.array_equal(s["n_accept"], nacc) and np.array_equal(s["n_reject"], nrej)
    assert nrej.sum() >= 1
    assert np.allclose(s["dt"], dt, rtol=1e-12) and np.allclose(s["t"], t, rtol=1e-12)
    m = y > 1e-30
    assert np.max(np.abs(s["y"] - y)[m] / y[m]) < 1e-9
