"""The fused assembly + factorisation kernel (vulcan_b200/csrc/vk_factor_dev.cuh FUSED = true, producer in vk_chem.cu::lhs_produce) against
the two-kernel path it replaces: the blocks D, up, dn the producer warps form in shared memory are BIT-IDENTICAL to lhs_ml_kernel's (same
tables, same expression order) on every fixture config - all block sizes 48 / 72 / 96 / 120, settling / vm / no_mol / vz variants, fixed rows
- and a whole step through it agrees with the two-kernel path to the rounding of the factorisation."""
import os
import subprocess
import sys

import numpy as np
import pytest

from helpers import CASES, Case, REPO, case_id, gpu_columns

pytestmark = pytest.mark.gpu
ONE_PER_TAG = sorted({t: (t, s) for t, s in CASES}.values())       # the latest step of every config


@pytest.mark.parametrize("tag,step", ONE_PER_TAG, ids=[case_id(p) for p in ONE_PER_TAG])
def test_fused_blocks_bit_identical(tag, step):
    c = Case(tag, step)
    col = gpu_columns(c, 2)
    y = np.repeat(c.y[None], 2, 0) * np.array([1.0, 1.0 + 1e-3])[:, None, None]
    dt = np.array([c.dt, 2 * c.dt])
    D0, u0, l0 = col.eval_lhs(y, dt)
    from vulcan_b200 import _abi
    os.environ["VK_LHS_VIA_FUSED"] = "1"
    try:
        D1, u1, l1 = col.eval_lhs(y, dt)
    except _abi.VulcanB200Error as e:
        if e.code == _abi.VK_ERR_UNSUPPORTED:
            pytest.skip("block size %d has no idle warps for the producers: this config takes the two-kernel path" % (24 * ((c.ni + 23) // 24)))
        raise
    finally:
        del os.environ["VK_LHS_VIA_FUSED"]
    assert np.array_equal(u0, u1) and np.array_equal(l0, l1)
    assert np.array_equal(D0, D1)


SCRIPT = """
import sys, numpy as np
sys.path[:0] = [%r, %r]
from helpers import Case, gpu_columns
c = Case(%r, %d)
col = gpu_columns(c, 3, refine=%d)
y = np.repeat(c.y[None], 3, 0); ym = np.repeat(c.ymix[None], 3, 0)
sol, ymo, delta, st = col.ros2_solve(y, ym, np.full(3, c.dt))
np.save(sys.argv[1], np.concatenate([sol.ravel(), delta, st.astype(float)]))
"""


@pytest.mark.parametrize("tag,step,refine", [("HD189", 100, 0), ("HD189", 300, -1), ("HD209S", 30, 0), ("Earth", 30, 0), ("EarthS", 100, 0), ("HD189cho", 30, 0)])
def test_fused_step_matches_two_kernel_step(tag, step, refine, tmp_path):
    outs = []
    for fused in ("1", "0"):
        out = str(tmp_path / ("fused%s.npy" % fused))
        env = dict(os.environ, VK_FUSED=fused)
        r = subprocess.run([sys.executable, "-c", SCRIPT % (os.path.join(REPO, "tests"), REPO, tag, step, refine), out], env=env, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr[-2000:]
        outs.append(np.load(out))
    a, b = outs
    n = (len(a) - 6)
    sa, sb = a[:n], b[:n]
    assert not a[-3:].any() and not b[-3:].any()
    c = Case(tag, step)
    m = (np.abs(sb) > c.cfg["atol"])
    err = np.max(np.abs(sa - sb)[m] / np.abs(sb)[m])
    print("%s-%d dt %.2e refine %d: fused vs two-kernel step, max rel diff of sol above atol %.2e, delta %.6e vs %.6e" % (tag, step, c.dt, refine, err, a[n], b[n]))
    assert err < (1e-10 if c.dt <= 1e-2 else 1e-5)
    assert abs(a[n] - b[n]) <= 1e-6 * abs(b[n])
