"""Parameter-sweep ensembles of independent columns (BASELINE config: 4096 HD189-like columns over a
Kzz x metallicity x C/O grid) and their partition across GPUs.

A single column does not shard (the block-tridiagonal system couples all layers every stage), so
multi-GPU work = a static partition of the column index across ranks: no collective inside the step, one final gather.
"""
import numpy as np

from . import _abi


def sweep_grid(n_k=16, n_z=16, n_co=16):
    """column id = ((iK*n_z + iZ)*n_co + iCO)  ->  (Kzz scale, metallicity scale, C/O)   """
    kzz = np.logspace(-2.0, 2.0, n_k)
    met = np.logspace(-1.0, 1.0, n_z)
    co = np.linspace(0.1, 1.5, n_co)
    ik, iz, ico = np.meshgrid(np.arange(n_k), np.arange(n_z), np.arange(n_co), indexing="ij")
    return kzz[ik.ravel()], met[iz.ravel()], co[ico.ravel()]


def partition(n_total, world_size, rank):
    """contiguous block of columns owned by `rank` (first ranks take the remainder)."""
    base, rem = divmod(n_total, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def synthetic_columns(y_base, n_0, compo, atom_names, kzz_scale, met_scale, c_to_o, solar_c_to_o=0.55):
    """HD189-like synthetic columns: the reference state `y_base` [nz, ni] re-weighted per column.

    Mixing ratios of species carrying C, N or O are multiplied by the metallicity scale, C-bearing species additionally
    by (C/O) / solar; every layer is renormalised to the hydrostatic density n_0 (op.py:909-914).  This keeps the
    sparsity pattern and the dynamic range of a real state (SURVEY.md §8d "synthetic single-column inputs")."""
    y_base = np.asarray(y_base, dtype=np.float64)
    compo = np.asarray(compo, dtype=np.float64)
    names = list(atom_names)
    heavy = np.zeros(compo.shape[0], dtype=bool)
    for a in ("C", "N", "O", "S"):
        if a in names:
            heavy |= compo[:, names.index(a)] > 0
    has_c = compo[:, names.index("C")] > 0 if "C" in names else np.zeros(compo.shape[0], dtype=bool)
    ymix = y_base / y_base.sum(axis=1, keepdims=True)
    ncol = len(kzz_scale)
    w = np.ones((ncol, compo.shape[0]))
    w[:, heavy] *= np.asarray(met_scale)[:, None]
    w[:, has_c] *= (np.asarray(c_to_o) / solar_c_to_o)[:, None]
    ym = ymix[None, :, :] * w[:, None, :]
    ym /= ym.sum(axis=2, keepdims=True)
    y = ym * np.asarray(n_0)[None, :, None]
    atom_ini = np.einsum("cji,ia->ca", y, compo)
    return y, atom_ini


class EnsembleRunner(object):
    """The columns [lo, hi) of an ensemble resident on one GPU, advanced by the device-resident controller."""

    def __init__(self, network, nz, y, dt, atm_common, kzz, k, cfg, compo, atom_ini, n_0, device=0, refine=0):
        self.ncol = y.shape[0]
        self.devnet = _abi.DeviceNetwork(network, device)
        self.col = _abi.Columns(self.devnet, nz, self.ncol)
        ni = network.ni
        rep = lambda a: np.ascontiguousarray(np.broadcast_to(np.asarray(a, dtype=np.float64), (self.ncol,) + np.shape(a)))
        a = atm_common
        self.col.set_atm(Kzz=np.ascontiguousarray(kzz), vz=rep(a["vz"]), dzi=rep(a["dzi"]), Dzz=rep(a["Dzz"]), vs=rep(a["vs"]),
                         Tco=rep(a["Tco"]), g=rep(a["g"]), M=rep(a["M"]), Ti=rep(a["Ti"]), Hpi=rep(a["Hpi"]), ms=rep(a["ms"]),
                         alpha=rep(a["alpha"]), top_flux=rep(a["top_flux"]), bot_flux=rep(a["bot_flux"]),
                         bot_vdep=rep(a["bot_vdep"]), use_moldiff=a["use_moldiff"], use_settling=a["use_settling"],
                         use_topflux=a["use_topflux"], use_botflux=a["use_botflux"], gas_indx=a.get("gas_indx"),
                         gas_indx_lhs=a.get("gas_indx_lhs"), shared=False)
        self.col.set_k(k)                                   # thermal + photolysis rates shared by the sweep (same T-P, same star)
        self.col.set_step_opts(cfg["mtol"], cfg["atol"], refine=refine)
        self.col.ens_setup(cfg["rtol"], cfg["loss_eps"], cfg["dt_min"], cfg["dt_max"], cfg["dt_var_min"], cfg["dt_var_max"],
                           cfg["pos_cut"], cfg["nega_cut"], compo, atom_ini, np.broadcast_to(n_0, (self.ncol, nz)))
        self.col.ens_set_state(y, dt)
        self.ni = ni

    def run(self, n_steps):
        self.col.ens_run(n_steps)
        return self.col.last_kernel_ms()[0]

    def state(self, want_y=True):
        return self.col.ens_get_state(want_y)


def gather_final(local_ymix, world_size, rank, device):
    """the ONE collective of an ensemble run: gather the final mixing ratios on rank 0 (torch.distributed / NCCL)."""
    import torch
    import torch.distributed as dist
    t = torch.from_numpy(np.ascontiguousarray(local_ymix)).to(device)
    if world_size == 1:
        return local_ymix
    sizes = [torch.zeros(1, dtype=torch.int64, device=device) for _ in range(world_size)]
    dist.all_gather(sizes, torch.tensor([t.shape[0]], dtype=torch.int64, device=device))
    sizes = [int(s.item()) for s in sizes]
    pad = torch.zeros((max(sizes),) + tuple(t.shape[1:]), dtype=t.dtype, device=device)
    pad[:t.shape[0]] = t
    out = [torch.empty_like(pad) for _ in range(world_size)]
    dist.all_gather(out, pad)
    if rank != 0:
        return None
    return torch.cat([o[:n] for o, n in zip(out, sizes)], 0).cpu().numpy()
