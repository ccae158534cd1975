"""GPU lock-step parity of the whole per-step protocol for every BASELINE single-column config (and the use_vm_mol variants):
the drop-in Ros2 object + Integration mirror, all arithmetic of the step on the B200 behind the C ABI, started from the
reference's initial state, must arrive at the state the UNMODIFIED reference had after the same number of steps
(tests/lockstep.py).  The CPU twin (tests/test_lockstep_host.py) runs the same host code on the oracle-backed stand-in."""
import pytest

from helpers import have
from lockstep import LOCKSTEP, check_ion_photolysis, lockstep

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("tag,nstep", LOCKSTEP, ids=["%s-%d" % p for p in LOCKSTEP])
def test_first_steps_reproduce_the_reference(tag, nstep):
    r = lockstep(tag, nstep)
    print("%s: after %d steps on the GPU  t %.1e  dt %.1e  y (masked) %.1e  y (>1e-30) %.1e  ymix %.1e  rejected %d  wall %.2f s" %
          (tag, nstep, r["t"], r["dt"], r["y"], r["y_all"], r["ymix"], r["rejected"], r["wall"]))
    assert r["t"] < 1e-9 and r["dt"] < 1e-6       # same accept / reject sequence and step sizes as the reference
    assert r["y"] < 1e-8 and r["ymix"] < 1e-8     # CPU twin measures 2e-15 ... 2e-9 (two different backward-stable solvers)


@pytest.mark.skipif(not have("HD189ion", "photo0000.npz"), reason="fixture missing")
def test_compute_Jion_on_the_gpu():
    """photo-ionisation rates (op.py:2789-2820): the ion branches ride in the device branch table of jrate_kernel"""
    check_ion_photolysis()
