"""CPU-only tests of the ensemble host logic: sweep grid, column partition, and the one collective (final gather) on a
world_size-2 gloo group."""
import os
import socket

import numpy as np
import pytest

from helpers import Case


def test_sweep_grid_ids():
    from vulcan_b200 import ensemble
    kz, met, co = ensemble.sweep_grid()
    assert kz.shape == (4096,)
    cid = (3 * 16 + 5) * 16 + 7
    assert np.isclose(kz[cid], np.logspace(-2, 2, 16)[3]) and np.isclose(met[cid], np.logspace(-1, 1, 16)[5])
    assert np.isclose(co[cid], np.linspace(0.1, 1.5, 16)[7])


@pytest.mark.parametrize("world", [1, 2, 3, 4, 8])
def test_partition_covers_everything(world):
    from vulcan_b200 import ensemble
    seen = []
    for r in range(world):
        lo, hi = ensemble.partition(4096, world, r)
        seen += list(range(lo, hi))
    assert seen == list(range(4096))
    sizes = [ensemble.partition(4096, world, r)[1] - ensemble.partition(4096, world, r)[0] for r in range(world)]
    assert max(sizes) - min(sizes) <= 1


def test_synthetic_columns_are_hydrostatic():
    from vulcan_b200 import ensemble
    c = Case("HD189", 100)
    kz, met, co = ensemble.sweep_grid(2, 2, 2)
    y, atom_ini = ensemble.synthetic_columns(c.y, c.st["n_0"], c.st["compo"], c.cfg["atom_list"], kz, met, co)
    assert y.shape == (8, c.nz, c.ni) and np.all(y >= 0)
    assert np.allclose(y.sum(axis=2), c.st["n_0"][None, :], rtol=1e-12)
    assert atom_ini.shape == (8, 4) and np.all(atom_ini > 0)
    assert not np.allclose(y[0], y[-1])


def _worker(rank, world, port, q):
    import torch.distributed as dist
    from vulcan_b200 import ensemble
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = ensemble.partition(11, world, rank)
    local = np.arange(lo, hi, dtype=np.float64)[:, None, None] * np.ones((1, 3, 2))
    out = ensemble.gather_final(local, world, rank, "cpu")
    if rank == 0:
        q.put(out[:, 0, 0].tolist())
    dist.barrier()
    dist.destroy_process_group()


def test_final_gather_gloo_world2():
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert got == [float(i) for i in range(11)]
