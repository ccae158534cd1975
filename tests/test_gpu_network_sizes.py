"""Random networks of every size class through the chemistry kernels + one Ros2 step on the device (round-1 staging script
scripts/gpu_network_sizes.py, now a parametrised -m gpu test).  The fixtures cover ni = 41, 65, 69, 71, 93, 99; the reference ships
networks up to ni = 99 and the kernels are instantiated up to the padded block size 120, so seeded random networks fill the gaps:
chemdf / diffdf / couplings bit-identical to the oracle, blocks to rounding, one small step to 1e-10."""
import numpy as np
import pytest

from helpers import Case
from oracle import Oracle
from vulcan_b200 import _abi
from vulcan_b200.network import Network

pytestmark = pytest.mark.gpu

R = 1.0 + 1.0 / 2 ** 0.5


def random_network(ni, seed):
    """ni species, ~13 reactions per species like the shipped NCHO / SNCHO files: two-body A + B -> C + D, three-body A + B + M -> C + M,
    dissociation A + M -> B + C + M, every species appears at least once."""
    rng = np.random.default_rng(seed)
    nm = lambda i: "X%d" % i
    lines, rid = ["# Two-body Reactions"], 1
    n2, n3 = 5 * ni, int(1.5 * ni)
    def pick(k):
        return [nm(int(i)) for i in rng.integers(0, ni, k)]
    for i in range(ni):                                   # chain so that every species exists, in index order
        lines.append("%d [ %s + %s -> %s + %s ] 1.0E-11 0.0 100.0" % (rid, nm(i), nm((i + 1) % ni), nm((i + 2) % ni), nm((i + 3) % ni)))
        rid += 2
    for _ in range(n2 - ni):
        a, b, c, d = pick(4)
        lines.append("%d [ %s + %s -> %s + %s ] 1.0E-11 0.0 100.0" % (rid, a, b, c, d))
        rid += 2
    lines.append("# 3-body and Disscoiation Reactions")
    for q in range(n3):
        a, b, c, d = pick(4)
        if q % 2:
            lines.append("%d [ %s + %s + M -> %s + M ] 1.0E-30 0.0 0.0 1.0E-11 0.0 0.0" % (rid, a, b, c))
        else:
            lines.append("%d [ %s + M -> %s + %s + M ] 1.0E-10 0.0 100.0 1.0E-9 0.0 0.0" % (rid, a, b, c))
        rid += 2
    return Network.from_text("\n".join(lines) + "\n", name="random%d" % ni)


def check(ni, seed=0, ncol=3):
    base = Case("HD189", 10)                              # atmosphere arrays of a real column; per-species arrays re-drawn for ni species
    nz = base.nz
    kw = base.atm_kwargs()
    rng = np.random.default_rng(100 + ni)
    net = random_network(ni, seed)
    assert net.ni == ni
    kw.update(Dzz=np.repeat(kw["Dzz"][:, :1], ni, 1) * rng.uniform(0.5, 2.0, (nz - 1, ni)), vs=np.zeros((nz - 1, ni)),
              ms=rng.uniform(1.0, 60.0, ni), alpha=np.full(ni, -0.25), top_flux=np.zeros(ni), bot_flux=np.zeros(ni), bot_vdep=np.zeros(ni),
              vm=np.zeros((nz, ni)), gas_indx=None, gas_indx_lhs=None, diff_esc_idx=[], use_vm_mol=False)
    n0 = base.y.sum(axis=1)
    mix = 10.0 ** rng.uniform(-12, 0, (nz, ni))
    y = n0[:, None] * mix / mix.sum(axis=1, keepdims=True)
    ymix = y / y.sum(axis=1, keepdims=True)
    k = np.zeros((nz, net.nr + 1))
    k[:, 1:] = 10.0 ** rng.uniform(-40, -32, (nz, net.nr))        # three-body terms k n M dt <= 1e-2 at n = M = 1e21, dt = 1e-12 s: a small step
    k[:, 2::2] *= 10.0 ** rng.uniform(-6, 0, (nz, net.nr // 2))
    dt = 1e-12                                                    # a genuinely small step for these random rates (1e-6 s gives delta ~ 1e5)
    o = Oracle(net)
    atm = o.make_atm(**kw)
    dev = _abi.DeviceNetwork(net, 0)
    col = _abi.Columns(dev, nz, ncol)
    col.set_atm(shared=True, **{a: kw[a] for a in ("Kzz", "vz", "dzi", "Dzz", "vs", "Tco", "g", "M", "Ti", "Hpi", "ms", "alpha", "top_flux",
                                                      "bot_flux", "bot_vdep", "use_moldiff", "use_settling", "use_topflux", "use_botflux",
                                                      "gas_indx", "gas_indx_lhs", "use_vm_mol", "vm", "diff_esc_idx")})
    col.set_k(k)
    col.set_step_opts(base.cfg["mtol"], base.cfg["atol"], refine=0, rhs_order=1)       # reference summation order: chemdf bit-identical
    Y = np.repeat(y[None], ncol, 0)
    chem, diff = col.eval_rhs(Y)
    ok_chem = np.array_equal(chem[0], o.chemdf(y, kw["M"], k)) and np.array_equal(chem[0], chem[-1])
    ok_diff = np.array_equal(diff[0], o.diffdf(atm, y))
    D, up, dn = col.eval_lhs(Y, np.full(ncol, dt))
    Do, upo, dno = o.lhs(atm, y, k, dt)
    scale = np.max(np.abs(Do), axis=2, keepdims=True)
    e_lhs = float(np.max(np.abs(D[0] - Do) / scale))
    ok_cpl = np.array_equal(up[0], upo) and np.array_equal(dn[0], dno)
    sol, ym, delta, status = col.ros2_solve(Y, np.repeat(ymix[None], ncol, 0), np.full(ncol, dt))
    ref = o.ros2_solver(atm, y, ymix, k, dt, base.cfg["mtol"], base.cfg["atol"], refine=0)
    m = ref["sol"] > 1e-30
    e_sol = float(np.max(np.abs(sol[0] - ref["sol"])[m] / ref["sol"][m]))
    print("ni %3d nr %4d: chemdf bit-identical %s, diffdf bit-identical %s, couplings bit-identical %s, blocks %.1e of the row scale, "
          "one Ros2 step %.1e (status %s, delta %.3e vs %.3e)" % (ni, net.nr, ok_chem, ok_diff, ok_cpl, e_lhs, e_sol, status.tolist(),
                                                                   delta[0], ref["delta"]), flush=True)
    return ok_chem and ok_diff and ok_cpl and e_lhs < 1e-13 and e_sol < 1e-10 and not status.any()


@pytest.mark.parametrize("ni", [20, 41, 56, 74, 93, 97, 99, 110, 120])
def test_random_network_of_size(ni):
    assert check(ni)
