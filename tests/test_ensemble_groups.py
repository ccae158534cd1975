"""Host logic of ensemble.GroupedEnsembleRunner (no GPU): the group policy and the partition of the columns."""
import numpy as np

from vulcan_b200 import ensemble


def test_auto_groups_policy():
    assert ensemble.auto_groups(1) == 1 and ensemble.auto_groups(255) == 1
    assert ensemble.auto_groups(256) == 2 and ensemble.auto_groups(512) == 2 and ensemble.auto_groups(4096) == 2


def test_group_bounds_cover_the_columns_once():
    for ncol in (256, 512, 1000, 4096):
        G = ensemble.auto_groups(ncol)
        b = [ensemble.partition(ncol, G, g) for g in range(G)]
        assert b[0][0] == 0 and b[-1][1] == ncol
        assert all(b[g][1] == b[g + 1][0] for g in range(G - 1))
        sizes = np.array([hi - lo for lo, hi in b])
        assert sizes.min() >= 32 and sizes.max() - sizes.min() <= 1      # every group takes the emitted kernels


def test_numa_binding_is_best_effort():
    """no GPU / no NVML / a single-node VM: the helper reports what it found and never raises (bench.py calls it on every rank)"""
    info = ensemble.bind_to_gpu_numa_node(0)
    assert isinstance(info, dict) and info["device"] == 0 and "node" in info
