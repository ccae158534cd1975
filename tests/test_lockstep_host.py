"""Host logic of the reference-facing layer (vulcan_b200/ros2.py + vulcan_b200/integration.py) on the CPU: the solver object is
wired to the oracle-backed stand-in of the C ABI (tests/oracle_columns.py) and must reproduce the reference's state after the
first N steps of every BASELINE single-column config.  The GPU twin of this test is tests/test_gpu_lockstep.py."""
import pytest

from helpers import have
from lockstep import LOCKSTEP, check_ion_photolysis, lockstep
from oracle_columns import oracle_backed_abi


@pytest.mark.parametrize("tag,nstep", LOCKSTEP, ids=["%s-%d" % p for p in LOCKSTEP])
def test_first_steps_reproduce_the_reference(tag, nstep):
    r = lockstep(tag, nstep, abi=oracle_backed_abi())
    print("%s: after %d steps  t %.1e  dt %.1e  y (masked) %.1e  y (>1e-30) %.1e  ymix %.1e  rejected %d  wall %.1f s" %
          (tag, nstep, r["t"], r["dt"], r["y"], r["y_all"], r["ymix"], r["rejected"], r["wall"]))
    assert r["t"] < 1e-9 and r["dt"] < 1e-6
    assert r["y"] < 1e-8 and r["ymix"] < 1e-8


@pytest.mark.skipif(not have("HD189ion", "photo0000.npz"), reason="fixture missing")
def test_compute_Jion_host_protocol():
    check_ion_photolysis(abi=oracle_backed_abi())


def test_stand_in_is_not_left_installed():
    from vulcan_b200 import _abi, ros2
    assert ros2._abi is _abi
