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


def _make_columns(devnet, nz, ncol, atm_common, kzz, k, cfg, refine, compo=None, k_static_rows_shared=False):
    """a vk_column handle for `ncol` columns of a sweep: per-column Kzz, everything else of the atmosphere replicated."""
    col = _abi.Columns(devnet, nz, ncol)
    rep = lambda a: np.ascontiguousarray(np.broadcast_to(np.asarray(a, dtype=np.float64), (ncol,) + np.shape(a)))
    a = atm_common
    col.set_atm(Kzz=np.ascontiguousarray(kzz), vz=rep(a["vz"]), dzi=rep(a["dzi"]), Dzz=rep(a["Dzz"]), vs=rep(a["vs"]),
                Tco=rep(a["Tco"]), g=rep(a["g"]), M=rep(a["M"]), Ti=rep(a["Ti"]), Hpi=rep(a["Hpi"]), ms=rep(a["ms"]),
                alpha=rep(a["alpha"]), top_flux=rep(a["top_flux"]), bot_flux=rep(a["bot_flux"]),
                bot_vdep=rep(a["bot_vdep"]), use_moldiff=a["use_moldiff"], use_settling=a["use_settling"],
                use_topflux=a["use_topflux"], use_botflux=a["use_botflux"], gas_indx=a.get("gas_indx"),
                gas_indx_lhs=a.get("gas_indx_lhs"), shared=False)
    col.set_k(k, static_rows_shared=k_static_rows_shared and np.ndim(k) == 3)   # thermal rates shared by the sweep (same T-P, same star)
    if refine < 0 and compo is None:
        raise ValueError("refine = -1 (auto) needs the element composition `compo` [ni, na]")
    col.set_step_opts(cfg["mtol"], cfg["atol"], refine=refine, compo=compo)
    return col


class EnsembleRunner(object):
    """The columns [lo, hi) of an ensemble resident on one GPU, advanced by the device-resident controller."""

    def __init__(self, network, nz, y, dt, atm_common, kzz, k, cfg, compo, atom_ini, n_0, device=0, refine=-1, k_static_rows_shared=False):
        self.ncol = y.shape[0]
        self.devnet = _abi.DeviceNetwork(network, device)
        self.col = _make_columns(self.devnet, nz, self.ncol, atm_common, kzz, k, cfg, refine, compo, k_static_rows_shared)
        self.col.ens_setup(cfg["rtol"], cfg["loss_eps"], cfg["dt_min"], cfg["dt_max"], cfg["dt_var_min"], cfg["dt_var_max"],
                           cfg["pos_cut"], cfg["nega_cut"], compo, atom_ini, np.broadcast_to(n_0, (self.ncol, nz)))
        self.col.ens_set_state(y, dt)
        self.ni = network.ni

    def run(self, n_steps):
        self.col.ens_run(n_steps)
        return self.col.last_kernel_ms()[0]

    def state(self, want_y=True):
        return self.col.ens_get_state(want_y)


def bind_to_gpu_numa_node(device):
    """One process per GPU on a multi-socket host: run this rank's host threads on the CPUs of the GPU's NUMA node and prefer that node for
    its page-locked buffers (first touch), so that the H2D / D2H copies of the host-buffer path do not cross the socket interconnect.
    Call before the pinned buffers and the worker threads exist.  Returns what was done (dict) - everything is best effort: a container
    may hide the topology or confine the process to CPUs of another node."""
    import ctypes
    import os
    info = {"device": int(device), "node": None, "cpus_bound": None, "mempolicy": None}
    try:
        import pynvml
        pynvml.nvmlInit()
        bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(int(device))).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        bus = bus.lower()
        if len(bus.split(":")[0]) == 8:                  # nvml prints an 8-digit PCI domain, sysfs a 4-digit one
            bus = bus[4:]
        with open("/sys/bus/pci/devices/%s/numa_node" % bus) as f:
            node = int(f.read().strip())
    except Exception as e:
        info["error"] = "no NUMA node for the device: %s" % e
        return info
    info["node"] = node
    if node < 0:
        return info
    try:
        with open("/sys/devices/system/node/node%d/cpulist" % node) as f:
            cpus = set()
            for part in f.read().strip().split(","):
                lo, _, hi = part.partition("-")
                cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = os.sched_getaffinity(0)
        both = cpus & allowed
        if both and both != allowed:
            os.sched_setaffinity(0, both)
            info["cpus_bound"] = len(both)
        else:
            info["cpus_bound"] = 0 if not both else len(both)
    except Exception as e:
        info["error"] = "affinity: %s" % e
    try:
        libc = ctypes.CDLL(None, use_errno=True)
        mask = ctypes.c_ulong(1 << node)
        MPOL_PREFERRED, SYS_set_mempolicy = 1, 238        # x86_64
        rc = libc.syscall(SYS_set_mempolicy, MPOL_PREFERRED, ctypes.byref(mask), ctypes.c_ulong(8 * ctypes.sizeof(mask)))
        info["mempolicy"] = "preferred node %d" % node if rc == 0 else "set_mempolicy failed (errno %d)" % ctypes.get_errno()
    except Exception as e:
        info["mempolicy"] = "unavailable: %s" % e
    return info


def auto_groups(ncol):
    """column groups for GroupedEnsembleRunner: two from 256 columns on.  Measured on a B200 (HD189 network, ms per step with 1 / 2 / 4 / 8
    groups): 512 columns 8.09 / 7.25 / 7.28 / -, 1024: 15.1 / 13.5 / 13.3 / -, 2048: 27.5 / 26.7 / 26.3 / 26.7, 4096: 54.0 / 52.2 / 52.8 /
    52.4 - two groups take most of the gain and keep every launch large (8 groups of 512 columns are 1.7 waves of the factor kernel each)"""
    return 2 if ncol >= 256 else 1


class GroupedEnsembleRunner(object):
    """EnsembleRunner over G column groups - one vk_column handle and stream per group, advanced concurrently from G host threads.  The
    kernels of a step are bound by different units (factorisation: FP64 pipe, solves: HBM, emitted chemistry: issue latency), and a batch
    leaves tails (512 columns are 1.7 waves of the factor kernel's 296 resident blocks): with several groups in flight the tail of one
    group's kernel runs next to another group's kernel of another kind.  Same arithmetic per column as one handle (columns are
    independent; the kernels a group takes depend on its size only through the >= 32 column threshold of the emitted kernels)."""

    def __init__(self, network, nz, y, dt, atm_common, kzz, k, cfg, compo, atom_ini, n_0, device=0, refine=-1, n_groups=None):
        self.ncol = y.shape[0]
        G = auto_groups(self.ncol) if n_groups is None else max(1, min(int(n_groups), self.ncol))
        self.bounds = [partition(self.ncol, G, g) for g in range(G)]
        dt = np.broadcast_to(np.asarray(dt, dtype=np.float64), (self.ncol,))
        self.runners = [EnsembleRunner(network, nz, y[lo:hi], np.ascontiguousarray(dt[lo:hi]), dict(atm_common), kzz[lo:hi], k, cfg, compo,
                                       atom_ini[lo:hi], n_0, device=device, refine=refine) for lo, hi in self.bounds]
        self.col = self.runners[0].col          # (profiling aids address one handle)
        self.ni = network.ni

    @property
    def n_groups(self):
        return len(self.runners)

    def run(self, n_steps):
        """n_steps attempted steps of every column; returns the device time (ms) of the slowest group's stream (the groups run concurrently)"""
        if len(self.runners) == 1:
            return self.runners[0].run(n_steps)
        import threading
        ms = [0.0] * len(self.runners)
        err = []

        def one(g):
            try:
                ms[g] = self.runners[g].run(n_steps)
            except Exception as e:          # surfaced in the calling thread
                err.append(e)
        th = [threading.Thread(target=one, args=(g,)) for g in range(len(self.runners))]
        for t in th:
            t.start()
        for t in th:
            t.join()
        if err:
            raise err[0]
        return max(ms)

    def set_state(self, y, dt):
        dt = np.broadcast_to(np.asarray(dt, dtype=np.float64), (self.ncol,))
        for r, (lo, hi) in zip(self.runners, self.bounds):
            r.col.ens_set_state(y[lo:hi], np.ascontiguousarray(dt[lo:hi]))

    def state(self, want_y=True):
        parts = [r.state(want_y) for r in self.runners]
        return {key: (None if parts[0][key] is None else np.concatenate([p[key] for p in parts], axis=0)) for key in parts[0]}

    def close(self):
        for r in self.runners:
            r.col.close()


class SteadyEnsemble(EnsembleRunner):
    """An ensemble advanced to per-column steady state entirely on the device (vk_ens_setup_steady / vk_ens_run_steady): every column
    stops on its own convergence test (Integration.stop / conv, op.py:1018-1087) and is frozen from then on; per column (steps, t, status)
    are gathered at the end.

    grid: dict(pico [nz+1], ms [ni], zco [nz+1], Hp [nz], dz [nz], pref_indx, gs) - what update_mu_dz reads (op.py:944-984);
    photo: None or the keyword arguments of _abi.Columns.photo_setup (the sweep shares the star and the cross sections; J rates are per
    column, so k is replicated per column)."""

    def __init__(self, network, nz, y, dt, atm_common, kzz, k, cfg, compo, atom_ini, n_0, grid, photo=None, device=0, refine=-1,
                 hist_cap=None, hist_stride=None, diff_esc_idx=(), conv_ignore_sp=None):
        ncol = y.shape[0]
        replicated = False
        if photo is not None and ncol > 1 and np.ndim(k) == 2:
            # one T-P profile: the thermal rows are the same in every column, the J rows are rewritten per column by the photolysis update
            k = np.ascontiguousarray(np.broadcast_to(np.asarray(k, dtype=np.float64), (ncol,) + np.shape(k)))
            replicated = True
        EnsembleRunner.__init__(self, network, nz, y, dt, atm_common, kzz, k, cfg, compo, atom_ini, n_0, device=device, refine=refine,
                                k_static_rows_shared=replicated)
        if photo is not None:
            self.col.photo_setup(**photo)
        conv_step = int(cfg["conv_step"])
        if hist_cap is None:
            # the exact look-back (every accepted state of the last conv_step steps) while it fits ~8 GB, else a thinned ring
            per_state = nz * network.ni * 8
            hist_cap = max(8, min(conv_step, int(8e9 // (per_state * ncol))))
        if hist_stride is None:
            hist_stride = max(1, -(-conv_step // hist_cap))
        self.hist_cap, self.hist_stride = int(hist_cap), int(hist_stride)
        self.col.ens_setup_steady(cfg, grid["pico"], grid["ms"], grid["zco"], grid["Hp"], grid["dz"], grid["pref_indx"], grid["gs"],
                                  conv_ignore_sp=conv_ignore_sp, diff_esc_idx=diff_esc_idx, hist_cap=self.hist_cap,
                                  hist_stride=self.hist_stride, use_photo=photo is not None)
        if photo is not None:
            self.col.ens_photo_update()                 # vulcan.py:170-176: one update at set-up, the loop updates again at count 0

    def run_to_steady_state(self, max_iterations=20000, chunk=64):
        import time
        t0 = time.time()
        left, it, dev_ms = self.ncol, 0, 0.0
        while left and it < max_iterations:
            left = self.col.ens_run_steady(chunk)
            dev_ms += self.col.last_kernel_ms()[0]
            it += chunk
        st = self.col.ens_get_state(want_y=True)
        sd = self.col.ens_get_steady()
        st.update(end_case=sd["end_case"], longdy=sd["longdy"], longdydt=sd["longdydt"], aflux_change=sd["aflux_change"], iterations=it,
                  wall_s=time.time() - t0, device_ms=dev_ms, columns_left=left)
        return st


class PipelinedHostSolver(object):
    """`vk_ros2_solve` on HOST buffers for a large batch of columns (the reference-facing call: y, ymix, dt in; sol, ymix, delta
    out).  The batch is split into `n_groups` vk_column handles, each with its own CUDA stream, driven from a thread pool (the
    ctypes calls release the GIL): the H2D copy of one group, the kernels of another and the D2H copy of a third overlap, so the
    host round trip costs max(transfer, compute) instead of their sum.  Pass page-locked arrays (e.g. numpy views of pinned torch
    tensors) to have them DMA'd directly."""

    def __init__(self, network, nz, atm_common, kzz, k, cfg, n_groups=None, device=0, refine=None, compo=None):
        from concurrent.futures import ThreadPoolExecutor
        if refine is None:
            refine = -1 if compo is not None else 0        # the product default (auto) needs the element composition
        self.ncol = kzz.shape[0]
        self.nz, self.ni = nz, network.ni
        self.devnet = _abi.DeviceNetwork(network, device)
        if n_groups is None:
            # measured on a B200 (scripts/e2e_groups.py): 512 columns best with 4-8 groups, 1024 with 4, 4096 with 16
            n_groups = max(4, min(16, self.ncol // 256))
        n_groups = max(1, min(n_groups, self.ncol))
        self.bounds = [partition(self.ncol, n_groups, g) for g in range(n_groups)]
        self.cols = [_make_columns(self.devnet, nz, hi - lo, atm_common, kzz[lo:hi], k, cfg, refine, compo) for lo, hi in self.bounds]
        self.pool = ThreadPoolExecutor(max_workers=n_groups)

    def solve_into(self, y, ymix, dt, sol, ymix_out, delta, status):
        """all arrays C-contiguous float64 ([ncol, nz, ni] / [ncol]); status int32 [ncol]"""
        per = self.nz * self.ni
        y, ymix, sol, ymix_out = (a.reshape(self.ncol * per) for a in (y, ymix, sol, ymix_out))

        def one(g):
            lo, hi = self.bounds[g]
            self.cols[g].ros2_solve_into(y[lo * per:hi * per], ymix[lo * per:hi * per], dt[lo:hi], sol[lo * per:hi * per],
                                         ymix_out[lo * per:hi * per], delta[lo:hi], status[lo:hi])
        list(self.pool.map(one, range(len(self.cols))))

    def close(self):
        self.pool.shutdown()
        for c in self.cols:
            c.close()


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
