"""Size-independent properties of the CUDA path at sizes the fixtures do not reach: every padded block size the factor kernel is
instantiated for (NIP = 48, 72, 96, 120), multi-wave batches, column-order invariance, linearity of the block-tridiagonal solve."""
import numpy as np
import pytest

from helpers import Case, gpu_columns

pytestmark = pytest.mark.gpu


def _synthetic_network(ni):
    """ni species X0..X{ni-1} coupled by two-body exchange reactions (enough structure for the network compiler; the rate tables are
    not used by vk_blocktri_solve)."""
    from vulcan_b200.network import Network
    lines = ["# synthetic", "# Two-body Reactions"]
    rid = 1
    for i in range(ni - 1):
        a, b, c, d = i, (i + 1) % ni, (i + 2) % ni, (i + 3) % ni
        lines.append("%d [ X%d + X%d -> X%d + X%d ]  1.0E-11 0.0 100.0" % (rid, a, b, c, d))
        rid += 2
    return Network.from_text("\n".join(lines) + "\n", name="synthetic%d" % ni)


def _dense_solve(D, up, dn, rhs):
    nz, n = up.shape
    A = np.zeros((nz * n, nz * n))
    idx = np.arange(n)
    for j in range(nz):
        A[j * n:(j + 1) * n, j * n:(j + 1) * n] = D[j]
        if j + 1 < nz:
            A[j * n + idx, (j + 1) * n + idx] = up[j]
        if j > 0:
            A[j * n + idx, (j - 1) * n + idx] = dn[j]
    return np.linalg.solve(A, rhs.ravel()).reshape(nz, n)


@pytest.mark.parametrize("ni", [7, 40, 48, 60, 72, 90, 100, 118])
def test_blocktri_solve_every_block_size(ni):
    """factor_kernel<48|72|96|120> + solve_kernel on random diagonally dominant systems vs a dense LAPACK solve."""
    from vulcan_b200 import _abi
    rng = np.random.default_rng(ni)
    nz, ncol = 11, 3
    net = _abi.DeviceNetwork(_synthetic_network(ni))
    col = _abi.Columns(net, nz, ncol)
    D = rng.standard_normal((ncol, nz, ni, ni)) * 0.3
    D[:, :, np.arange(ni), np.arange(ni)] += 3.0 + rng.random((ncol, nz, ni))
    up = rng.standard_normal((ncol, nz, ni)) * 0.2
    dn = rng.standard_normal((ncol, nz, ni)) * 0.2
    up[:, -1] = 0.0
    dn[:, 0] = 0.0
    rhs = rng.standard_normal((ncol, nz, ni))
    x, st = col.blocktri_solve(D, up, dn, rhs, refine=0)
    assert not st.any()
    for c in range(ncol):
        ref = _dense_solve(D[c], up[c], dn[c], rhs[c])
        assert np.max(np.abs(x[c] - ref)) <= 1e-12 * np.max(np.abs(ref))
    # linearity: solve(a r1 + r2) = a solve(r1) + solve(r2) to rounding
    r2 = rng.standard_normal((ncol, nz, ni))
    x2, _ = col.blocktri_solve(D, up, dn, r2, refine=0)
    x3, _ = col.blocktri_solve(D, up, dn, 2.5 * rhs + r2, refine=0)
    assert np.max(np.abs(x3 - (2.5 * x + x2))) <= 1e-12 * np.max(np.abs(x3))
    # a singular block is reported, not propagated as garbage
    Db = D.copy()
    Db[1, 0] = 0.0                                    # S_0 = D_0: an exactly singular first block
    _, stb = col.blocktri_solve(Db, up, dn, rhs, refine=0)
    assert stb[1] != 0 and stb[0] == 0 and stb[2] == 0


def test_multi_wave_batch_and_column_order():
    """700 columns (> 2 waves of resident factor blocks): identical inputs give bit-identical outputs in every slot, and permuting
    distinct columns permutes the results bit for bit."""
    c = Case("HD189", 10)
    ncol = 700
    col = gpu_columns(c, ncol)
    y = np.repeat(c.y[None], ncol, 0)
    ym = np.repeat(c.ymix[None], ncol, 0)
    scale = np.ones(ncol)
    scale[1::3] = 1.0 + 1e-3
    scale[2::3] = 1.0 - 2e-3
    y = y * scale[:, None, None]                     # three distinct column types, interleaved
    dt = np.full(ncol, c.dt)
    sol, ymo, delta, st = col.ros2_solve(y, ym, dt)
    assert not st.any()
    for k in range(3):
        assert np.all(sol[k::3] == sol[k]) and np.all(ymo[k::3] == ymo[k]) and np.all(delta[k::3] == delta[k])
    perm = np.random.default_rng(0).permutation(ncol)
    sol2, ymo2, delta2, _ = col.ros2_solve(y[perm], ym[perm], dt)
    assert np.array_equal(sol2, sol[perm]) and np.array_equal(ymo2, ymo[perm]) and np.array_equal(delta2, delta[perm])
    assert not np.array_equal(sol[0], sol[1])


def test_pipelined_host_solver_matches_single_handle(monkeypatch):
    """ensemble.PipelinedHostSolver (column groups on separate streams) returns exactly what one handle returns.  (Table-driven chemistry
    kernels on both sides: a handle of 37 columns would take the emitted kernels, groups of 7 - 8 columns would not, and the two sum in
    different orders - compared to rounding level in test_rhs_emitted_batch.)"""
    from vulcan_b200 import ensemble
    monkeypatch.setenv("VK_EMIT_JAC", "0")
    monkeypatch.setenv("VK_EMIT", "0")
    c = Case("HD189", 10)
    ncol = 37
    kw = c.atm_kwargs()
    kzz = np.repeat(np.asarray(kw["Kzz"])[None], ncol, 0) * np.linspace(0.5, 2.0, ncol)[:, None]
    y = np.repeat(c.y[None], ncol, 0) * np.linspace(0.999, 1.001, ncol)[:, None, None]
    ym = np.repeat(c.ymix[None], ncol, 0)
    dt = np.full(ncol, c.dt)
    one = ensemble.PipelinedHostSolver(c.net, c.nz, kw, kzz, c.k, c.cfg, n_groups=1)
    five = ensemble.PipelinedHostSolver(c.net, c.nz, kw, kzz, c.k, c.cfg, n_groups=5)
    outs = []
    for s in (one, five):
        sol, ymo = np.empty_like(y), np.empty_like(y)
        delta, st = np.empty(ncol), np.zeros(ncol, dtype=np.int32)
        s.solve_into(np.ascontiguousarray(y), np.ascontiguousarray(ym), dt, sol, ymo, delta, st)
        outs.append((sol, ymo, delta, st))
        s.close()
    for a, b in zip(outs[0], outs[1]):
        assert np.array_equal(a, b)
    assert not outs[0][3].any() and np.all(outs[0][2] > 0)


def test_grouped_ensemble_runner_matches_one_handle():
    """ensemble.GroupedEnsembleRunner (column groups on separate streams, advanced concurrently from host threads): every column ends exactly
    where it ends in one handle - columns are independent and both group sizes take the same kernels"""
    from vulcan_b200 import ensemble
    c = Case("HD189", 0)
    ncol = 80
    kz, met, co = [a[:ncol] for a in ensemble.sweep_grid()]
    y, atom_ini = ensemble.synthetic_columns(c.st["y_ini"], c.st["n_0"], c.st["compo"], c.cfg["atom_list"], kz, met, co)
    kw = c.atm_kwargs()
    kzz = kz[:, None] * np.asarray(kw["Kzz"])[None, :]
    dt = np.full(ncol, float(c.cfg["dttry"]))
    one = ensemble.EnsembleRunner(c.net, c.nz, y, dt, dict(kw), kzz, c.k, c.cfg, c.st["compo"], atom_ini, c.st["n_0"])
    two = ensemble.GroupedEnsembleRunner(c.net, c.nz, y, dt, dict(kw), kzz, c.k, c.cfg, c.st["compo"], atom_ini, c.st["n_0"], n_groups=2)
    assert two.n_groups == 2
    one.run(25)
    two.run(25)
    a, b = one.state(), two.state()
    for key in ("n_accept", "n_reject", "t", "dt", "y"):
        assert np.array_equal(a[key], b[key]), key
    assert a["n_accept"].min() > 10
    two.close()
    one.col.close()
