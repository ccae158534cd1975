#!/usr/bin/env python
"""bench.py — throughput of the B200-native Ros2 hot path (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W]            our arm (N>1: launched by torch.distributed.run)
  python bench.py --impl reference [--gpus N] ...                CPU arm: the UNMODIFIED reference (oracle/_ref, numpy/scipy) on the host cores,
                                                                  one process per core; the C oracle port beside it / when oracle/_ref is absent

Workload (config.workload): BASELINE config "ensemble sweep: 4096 HD189-like columns over a Kzz x metallicity x C/O grid"
(NCHO_photo_network: ni=69, nr=878, nz=150), STRONG scaling: the 4096 columns are partitioned across the N ranks, no
collective inside the step, one final gather.  A "step" is one ATTEMPTED Ros2 step of every column through the whole
hot path (rhs -> lhs -> block-tridiagonal factor -> 2 solves (+refinement) -> epilogue -> clip -> accept/reject ->
rescale -> step size), inputs resident in HBM.  The per-step working set (D and W blocks: 2 x 6.2 MB per column) is far
larger than the 126 MB L2, so no explicit L2 flush is needed between timed steps.
`value`  = column-steps/s of the device-resident loop (vk_ens_run), CUDA-event timed, max over ranks.
`e2e`    = the same metric through the reference-facing call vk_ros2_solve with PINNED HOST buffers: H2D of y, ymix, dt and
           D2H of sol, ymix, delta inside the timed region every step.
Synthetic data: the reference's own HD189 state at step 100 (tests/golden fixture, produced by running the unmodified
reference) re-weighted per column (vulcan_b200/ensemble.py).  /root/reference is never read at run time.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

N_COLUMNS = 4096
CPU_STEPS_PER_THREAD = int(os.environ.get("VK_BENCH_CPU_STEPS_PER_THREAD", "48"))   # CPU legs: column-steps timed per host thread (~2-3 s wall, ~40 CPU-seconds on 16 threads)
BASE_STEP = 100          # fixture state the synthetic columns are derived from (dt = 4.84 s)
FLOP_FACTOR = lambda nz, ni: nz * (2.0 * ni ** 3 + ni ** 2)                 # getrf+getri count + scaled Schur update
FLOP_SOLVES = lambda nz, ni, nrhs: nrhs * nz * (2.0 * ni ** 2 + 2.0 * ni)   # forward/backward sweeps


def load_case():
    from vulcan_b200.fixtures import Case
    return Case("HD189", BASE_STEP)


def build_columns(case, lo, hi):
    from vulcan_b200 import ensemble
    kz, met, co = ensemble.sweep_grid()
    kz, met, co = kz[lo:hi], met[lo:hi], co[lo:hi]
    st, cfg = case.st, case.cfg
    y, atom_ini = ensemble.synthetic_columns(case.y, st["n_0"], st["compo"], cfg["atom_list"], kz, met, co)
    kw = case.atm_kwargs()
    kzz = kz[:, None] * np.asarray(kw["Kzz"])[None, :]
    return y, atom_ini, kzz, kw


class ClockSampler(object):
    """nvidia-smi sampling DURING the timed region (B200_PROFILING.md clocks line).  The sampler is started ahead of the region
    (nvidia-smi needs ~0.1 s to emit its first line) at a 20 ms period; every line is stamped when it is read and only the lines
    that fall inside [mark_begin, mark_end] are used - a region shorter than one period falls back to the nearest lines."""

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None
        self.t0 = self.t1 = None

    def start(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits",
                                          "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, bufsize=1)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def mark_begin(self):
        self.t0 = time.time()

    def mark_end(self):
        self.t1 = time.time()

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.1)
        self.proc.terminate()
        t0, t1 = self.t0 or 0.0, self.t1 or time.time()
        inside = [r for ts, r in self.rows if t0 <= ts <= t1 + 0.03]
        note = None
        if not inside and self.rows:       # region shorter than a sampling period: the lines closest to it
            mid = 0.5 * (t0 + t1)
            inside = [r for ts, r in sorted(self.rows, key=lambda x: abs(x[0] - mid))[:3]]
            note = "timed region shorter than the sampling period: nearest samples"
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in inside:
            f = [x.strip() for x in r.split(",")]
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
                for n, v in zip(names, f[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        out = {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
               "reasons": sorted(reasons), "samples": len(sm), "period_ms": 20}
        if note:
            out["note"] = note
        return out


def measured_fp64_peak(device):
    """MEASURED_PEAKS.json has no FP64 entry: time cuBLAS DGEMM 6144^3 in this run, best of 5."""
    import torch
    n = 6144
    a = torch.randn(n, n, dtype=torch.float64, device=device)
    b = torch.randn(n, n, dtype=torch.float64, device=device)
    torch.matmul(a, b)
    best = 1e9
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); torch.matmul(a, b); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return 2.0 * n ** 3 / (best * 1e-3) / 1e12


def cpu_port_rate(case, n_sample, threads, n_total=None):
    """column-steps/s of the CPU oracle port (oracle/vk_oracle.c, one attempted step + clip per column) on `threads`
    host threads; the ctypes calls release the GIL.  `n_total` column-steps are timed, cycling over `n_sample` distinct columns
    of the sweep (the per-column work does not depend on which column it is)."""
    n_total = n_sample if n_total is None else n_total
    from concurrent.futures import ThreadPoolExecutor
    if os.path.join(REPO, "oracle") not in sys.path:
        sys.path.insert(0, os.path.join(REPO, "oracle"))       # the CPU legs are the one place bench.py may execute oracle/
    from oracle import Oracle
    y, atom_ini, kzz, kw = build_columns(case, 0, n_sample)
    cfg = case.cfg
    o = Oracle(case.net)
    atms = []
    for i in range(n_sample):
        k2 = dict(kw); k2["Kzz"] = kzz[i]
        atms.append(o.make_atm(**k2))
    ymix = y / y.sum(axis=2, keepdims=True)

    def one(q):
        i = q % n_sample
        res = o.ros2_solver(atms[i], y[i], ymix[i], case.k, case.dt, cfg["mtol"], cfg["atol"], refine=0)
        o.clip_loss(res["sol"], res["ymix"], case.st["compo"], cfg["pos_cut"], cfg["nega_cut"], cfg["mtol"])
        return res["delta"]
    one(0)
    t0 = time.time()
    with ThreadPoolExecutor(max_workers=threads) as ex:
        list(ex.map(one, range(n_total)))
    return n_total / (time.time() - t0)


def host_threads():
    try:
        n = len(os.sched_getaffinity(0))
    except Exception:
        n = os.cpu_count() or 1
    return max(1, min(n, 64))


WORKLOAD = "ensemble sweep: 4096 HD189-like columns (NCHO_photo_network ni=69 nr=878 nz=150), Kzz x metallicity x C/O; one step = one " \
           "attempted Ros2 step (Ros2.solver, op.py:2860-3007) of every column"
REF_CALLS_PER_STEP = int(os.environ.get("VK_BENCH_REF_CALLS", "3"))    # reference arm: op.Ros2.solver calls per worker process and bench step


def reference_workers(n_proc, refdir):
    """n_proc host processes of the UNMODIFIED reference (oracle/ref_worker.py on oracle/_ref/HD189), each set up like vulcan.py does"""
    env = dict(os.environ, OMP_NUM_THREADS="1", PYTHONHASHSEED="0")
    procs = [subprocess.Popen([sys.executable, os.path.join(REPO, "oracle", "ref_worker.py"), refdir, "2"], stdin=subprocess.PIPE,
                              stdout=subprocess.PIPE, stderr=(None if os.environ.get("VK_BENCH_DEBUG") else subprocess.DEVNULL), text=True, bufsize=1, env=env) for _ in range(n_proc)]
    for p in procs:
        line = p.stdout.readline()
        if not line.startswith("READY"):
            for q in procs:
                q.kill()
            raise RuntimeError("reference worker failed to start: %r" % line)
    return procs


def reference_step(procs, n_calls):
    """every worker runs n_calls solver calls concurrently; returns the wall time of the slowest and the summed per-process time"""
    t0 = time.time()
    for p in procs:
        p.stdin.write("GO %d\n" % n_calls)
        p.stdin.flush()
    inner = [float(p.stdout.readline().split()[1]) for p in procs]
    return time.time() - t0, inner


def measure_cpu_arm(case, n_steps, n_warm, budget_s=150.0):
    """The reference's own CPU implementation of the path on the box's host cores (BASELINE north_star: "timed on the GPU box's own host
    cores in the same run, with the core count stated").  When oracle/_ref/HD189 is present (staged by oracle/build_ref.py in the build
    container; it travels with the snapshot) this times REAL op.Ros2.solver calls of the unmodified numpy/scipy reference: one process per
    host core (independent columns = independent processes, OMP_NUM_THREADS = 1 as vulcan.py:52 forces), `kind` "reference".  The C port of
    the GPU algorithm (oracle/vk_oracle.c) on the same cores is reported beside it (`cpu_port`), and alone when oracle/_ref is absent.
    Returns (value, ms_per_bench_step, bench steps done, column-steps per bench step, cpu_baseline dict, cpu_port dict or None)."""
    sys.path.insert(0, os.path.join(REPO, "oracle"))
    import build_ref
    threads = host_threads()
    refdir = build_ref.staged("HD189")
    port = None
    try:
        cpu_port_rate(case, max(2 * threads, 16), threads, 2 * threads)
        port = cpu_port_rate(case, max(2 * threads, 16), threads, max(CPU_STEPS_PER_THREAD // 4, 2) * threads)
    except Exception:
        port = None
    if refdir is None:
        n_sample = max(2 * threads, 16)
        n_total = CPU_STEPS_PER_THREAD * threads
        rates, walls = [], []
        for _ in range(max(1, n_warm > 0)):
            cpu_port_rate(case, n_sample, threads, 4 * threads)
        t_all = time.time()
        for _ in range(n_steps):
            t0 = time.time()
            rates.append(cpu_port_rate(case, n_sample, threads, n_total))
            walls.append(time.time() - t0)
            if time.time() - t_all > budget_s:
                break
        v, kind, cores = float(np.mean(rates)), "port", threads
        sample = "%d column-steps per bench step on %d host threads with the C oracle port (oracle/vk_oracle.c: same algorithm as the GPU " \
                 "path, refine=0); oracle/_ref absent, so the unmodified reference could not be timed here" % (n_total, threads)
        per_step = n_total
    else:
        try:
            avail = int([ln for ln in open("/proc/meminfo") if ln.startswith("MemAvailable")][0].split()[1]) // (1 << 20)     # GiB
        except Exception:
            avail = 64
        n_proc = max(1, min(threads, 32, avail // 3))          # ~2.5 GB per process (dense 857 MB Jacobian + band copy + LAPACK workspace)
        procs = reference_workers(n_proc, refdir)
        try:
            for _ in range(max(1, min(n_warm, 2))):
                reference_step(procs, 1)
            walls, inner = [], []
            t_all = time.time()
            for _ in range(n_steps):
                w, inn = reference_step(procs, REF_CALLS_PER_STEP)
                walls.append(w)
                inner += inn
                if time.time() - t_all > budget_s:
                    break
            one = reference_step(procs[:1], REF_CALLS_PER_STEP)[0] / REF_CALLS_PER_STEP if n_proc > 1 else None     # a single process alone
        finally:
            for p in procs:
                try:
                    p.stdin.write("QUIT\n"); p.stdin.flush()
                except Exception:
                    pass
            for p in procs:
                try:
                    p.wait(timeout=10)
                except Exception:
                    p.kill()
        per_step = n_proc * REF_CALLS_PER_STEP
        v, kind, cores = per_step / float(np.mean(walls)), "reference", n_proc
        sample = "%d op.Ros2.solver calls of the UNMODIFIED reference per bench step: %d host processes x %d calls (oracle/_ref/HD189, " \
                 "numpy/scipy, OMP_NUM_THREADS=1 each; %.2f s per call with all processes running%s; host has %d usable threads)" % (
                     per_step, n_proc, REF_CALLS_PER_STEP, float(np.mean(inner)) / REF_CALLS_PER_STEP,
                     "" if one is None else ", %.2f s per call for one process alone" % one, threads)
    cb = {"value": v, "unit": "column-steps/s", "cores": cores, "kind": kind, "sample": sample}
    cp = None if port is None else {"value": port, "unit": "column-steps/s", "cores": threads, "kind": "port",
                                    "sample": "C port of the GPU algorithm (oracle/vk_oracle.c) on %d host threads" % threads}
    return v, 1e3 * float(np.mean(walls)), len(walls), per_step, cb, cp


def run_reference_arm(args, rank, world):
    if rank != 0:
        return
    case = load_case()
    v, ms, n_done, per_step, cb, cp = measure_cpu_arm(case, args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": "ensemble column-steps/s", "value": v, "unit": "column-steps/s", "n_gpus": args.gpus,
        "steps": n_done, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "column_steps_per_bench_step": per_step,
                   "note": "ms_per_step is the MEASURED wall time of one bench step of this arm (a bounded sample of the workload: "
                           "%d column-steps), value = column-steps / s of that sample" % per_step},
        "cpu_baseline": cb, "cpu_port": cp,
        "e2e": {"value": v, "unit": "column-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--columns", type=int, default=N_COLUMNS)
    ap.add_argument("--groups", type=int, default=None, help="column groups (streams) of the device-resident runner; default: ensemble.auto_groups")
    ap.add_argument("--refine", type=int, default=-1, help="iterative refinement of the linear solves: -1 = the product default (auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-groups", type=int, default=None,
                    help="column groups (streams) of the pipelined host-buffer path (default: by batch size, ensemble.PipelinedHostSolver)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from vulcan_b200 import _abi, ensemble
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - the product has no CPU path (use --impl reference for the CPU arm)")
    args.warmup = max(args.warmup, 3)
    # one process per GPU: this rank's threads and page-locked buffers on the GPU's NUMA node (best effort, reported in config.numa)
    numa = ensemble.bind_to_gpu_numa_node(local_rank) if world > 1 else {"node": None}
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=device)

    case = load_case()
    cfg, st = case.cfg, case.st
    ncols_total = args.columns
    lo, hi = ensemble.partition(ncols_total, world, rank)
    y, atom_ini, kzz, kw = build_columns(case, lo, hi)
    atm_common = dict(kw)
    # every column starts like a fresh reference run: its own (re-weighted) state with dt = dttry (vulcan_cfg.dttry, store.py:32)
    dt0 = float(cfg["dttry"])
    # the public ensemble runner: column groups on separate streams (ensemble.auto_groups), advanced concurrently
    runner = ensemble.GroupedEnsembleRunner(case.net, case.nz, y, np.full(hi - lo, dt0), atm_common, kzz, case.k, cfg, st["compo"],
                                            atom_ini, st["n_0"], device=local_rank, refine=args.refine, n_groups=args.groups)
    ncol = hi - lo

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident loop ------------------------------------------------------------------------------------
    runner.run(args.warmup)
    runner.set_state(y, np.full(ncol, dt0))                 # every timed run starts from the same state
    runner.run(1)
    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.25)                                         # let nvidia-smi reach its sampling loop before the timed region
    barrier()
    sampler.mark_begin()
    t_wall0 = time.time()
    ms = runner.run(args.steps)                              # CUDA events on every group's stream bracket exactly K steps: the slowest group
    barrier()
    wall = time.time() - t_wall0
    sampler.mark_end()
    clocks = sampler.stop()
    tms = torch.tensor([ms], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    ms_max = float(tms.item())
    value = ncols_total * args.steps / (ms_max * 1e-3)
    s = runner.state(want_y=False)

    # ---- per-kernel time of the dominant kernel (factor) for the roofline: CUDA events inside the step ---------------
    # one group (stream): ev1..ev2 of the last step bracket the factor kernel inside the step.  Several groups: their kernels overlap inside
    # the step, so the kernel is timed on every group's resident state alone (vk_debug_time_kernel, CUDA events on the group's stream, the
    # step's own launch arguments) and the launches are summed - the burst peak applies to that number
    import ctypes
    fa, fb = ctypes.c_float(0), ctypes.c_float(0)
    if runner.n_groups == 1:
        for _ in range(3):
            runner.run(1)
        runner.col.lib.vk_last_kernel_ms(runner.col.handle, ctypes.byref(fa), ctypes.byref(fb))
        factor_ms, factor_how = float(fb.value), "CUDA events around the kernel inside the step"
    else:
        runner.run(1)
        factor_ms = float(sum(r.col.time_kernel(2, 3) for r in runner.runners))
        factor_how = "sum over the %d column groups, each group's launch timed alone on its resident state (inside the step the groups overlap)" % runner.n_groups

    # ---- the HBM-side kernels of the step, each timed alone on the resident state (library profiling aid vk_debug_time_kernel,
    # CUDA events on the column stream; scripts/kernel_times.py).  Supplementary: never allowed to break the bench line.
    kernel_ms = {}
    try:
        for which, name in ((0, "lhs: Jacobian + diagonal / couplings (emitted negjac + lhs_diag_kernel, or lhs_ml_kernel)"), (1, "rhs: chemdf + transport stencil (emitted chemdf + rhs_stencil_kernel, or rhs_warp_kernel)"), (6, "emitted chemdf kernel (part of the rhs line)"),
                            (7, "emitted Jacobian kernel (part of the lhs line)"),
                            (3, "lu_solve_kernel (first solve: backward sweep, forward fused into the factorisation)"),
                            (4, "lu_solve_kernel (forward + backward)")):
            try:
                msv = float(sum(r.col.time_kernel(which, 3) for r in runner.runners))      # every group's launch, timed alone
            except Exception:
                continue
            if msv > 0:
                kernel_ms[name] = msv
    except Exception:
        kernel_ms = {}

    runner_groups = runner.n_groups
    # ---- the one collective: final gather of the mixing ratios ----------------------------------------------------------
    fin = runner.state(want_y=True)
    ymix_local = fin["y"] / fin["y"].sum(axis=2, keepdims=True)
    gathered = ensemble.gather_final(ymix_local, world, rank, device)

    # ---- e2e through the reference-facing call (vk_ros2_solve on HOST buffers, pinned), H2D + D2H inside the timed region ------
    # the public API for large batches is ensemble.PipelinedHostSolver: column groups on separate streams so that copies overlap
    # the kernels of the other groups
    runner.close()
    nv = ncol * case.nz * case.net.ni
    pin = [torch.empty(nv, dtype=torch.float64).pin_memory() for _ in range(4)]
    hy, hm, hs, ho = [p.numpy() for p in pin]
    hy[:] = y.ravel()
    hm[:] = (y / y.sum(axis=2, keepdims=True)).ravel()
    hdt = np.full(ncol, dt0); hdelta = np.empty(ncol); hstat = np.zeros(ncol, dtype=np.int32)
    host = ensemble.PipelinedHostSolver(case.net, case.nz, atm_common, kzz, case.k, cfg, n_groups=args.e2e_groups,
                                        device=local_rank, refine=args.refine, compo=st["compo"])
    e2e_steps = max(2, min(args.steps, 5))
    for _ in range(2):
        host.solve_into(hy, hm, hdt, hs, ho, hdelta, hstat)
    barrier()
    t0 = time.time()
    for _ in range(e2e_steps):
        host.solve_into(hy, hm, hdt, hs, ho, hdelta, hstat)
    barrier()
    t_e2e = torch.tensor([time.time() - t0], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
    e2e_value = ncols_total * e2e_steps / float(t_e2e.item())
    host.close()

    if rank == 0:
        ni, nz = case.net.ni, case.nz
        refine = args.refine
        # kernels of one step: rhs x2, lhs, factor, solve x2, refinement (forced: resid + solve + axpy per pass and stage; auto: init + resid +
        # 4 x (solve, axpy, resid, select) per stage - blocks of columns below refine_dt_min exit at once), epilogue, clip, control, apply
        n_ref = 2 * 3 * refine if refine > 0 else (2 * (2 + 4 * 4) if refine == -1 else (2 * (2 + 4 * -refine) if refine < 0 else 0))
        # (rhs = emitted chemdf + layer scalars + stencil, lhs = emitted Jacobian + diagonal kernel for groups >= 32 columns of a registered network)
        emitted = (ncol // runner_groups) >= 32
        launches_per_step = (2 * 3 + 2 if emitted else 2 + 1) + 1 + 2 + n_ref + 1 + 1 + 1 + 1
        line = {
            "metric": "ensemble column-steps/s", "value": value, "unit": "column-steps/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD if ncols_total == N_COLUMNS else WORKLOAD.replace("4096", str(ncols_total)),
                       "partition": "columns partitioned across %d GPU(s), no collective in the step" % world, "refine": refine,
                       "columns_per_gpu": ncol, "column_groups": runner_groups, "numa_rank0": numa, "l2": "per-step working set (2 x %.1f MB per column) exceeds L2: no flush needed" % (nz * 72 * 72 * 8 / 1e6),
                       "accepted_fraction": float(np.sum(s["n_accept"])) / float(np.sum(s["n_accept"]) + np.sum(s["n_reject"]))},
            "e2e": {"value": e2e_value, "unit": "column-steps/s", "h2d_bytes_per_step": int(2 * nv * 8 * world + 8 * ncols_total),
                    "d2h_bytes_per_step": int(2 * nv * 8 * world + 12 * ncols_total)},
            "gpu_launches": launches_per_step * args.steps * runner_groups,
            "clocks": clocks, "wall_s_timed_region": wall,
            "final_gather_rows": None if gathered is None else int(gathered.shape[0]),
        }
        # roofline of the dominant kernel (block-tridiagonal factorisation: FP64 pipe bound)
        peak = measured_fp64_peak(device)
        fms = factor_ms
        flops = ncol * FLOP_FACTOR(nz, ni)
        # DRAM traffic of the kernel is not measurable without a profiler (a run under ncu is never a bench value): null here; the ncu
        # capture of the same kernel is committed under profiles/ and quoted in DESIGN.md section 6
        traffic = None
        line["roofline"] = {"bound": "tensor", "kernel": "factor_kernel (per-layer Gauss-Jordan inverse + Schur update)",
                            "achieved": flops / (fms * 1e-3) / 1e12, "peak": peak, "unit": "TFLOP/s",
                            "frac": flops / (fms * 1e-3) / 1e12 / peak, "traffic": traffic,
                            "peak_source": "cuBLAS DGEMM 6144^3 measured in this run (MEASURED_PEAKS.json has no FP64 entry)",
                            "flops_per_column_step": FLOP_FACTOR(nz, ni), "kernel_ms": fms, "kernel_ms_how": factor_how,
                            "note": "since round 2 the kernel also forms the forward elimination of the first solve for columns with dt < 1e3 s "
                                    "(z = W t, 2 nz ni^2 flop per column, NOT counted in the algorithmic flops; +1.3 ms of kernel time that saves a "
                                    "4.2 ms sweep of lu_solve_kernel)",
                            "share_of_step": fms / (ms_max / args.steps)}
        # HBM roofline of the streaming kernels (north_star: "achieved HBM GB/s for the rate/RHS/diffusion kernels"), algorithmic bytes
        # per column as DESIGN.md section 4 states them; peak = MEASURED_PEAKS.json (driver-written copy bandwidth) or the recipe's fallback
        try:
            hbm_peak, hbm_src = 6650.0, "fallback of B200_PROFILING.md (MEASURED_PEAKS.json absent)"
            pk = os.path.join(REPO, "MEASURED_PEAKS.json")
            if os.path.exists(pk):
                hbm_peak, hbm_src = float(json.load(open(pk))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs"
            nip = ((ni + 23) // 24) * 24 if ni > 48 else 48          # padded block size the kernels are instantiated for (48 / 72 / 96 / 120)
            alg = {"lhs: Jacobian + diagonal / couplings (emitted negjac + lhs_diag_kernel, or lhs_ml_kernel)": nz * nip * nip * 8.0 + nz * ni * 8.0,                        # writes D (+ up, dn), reads y; k is shared (L2)
                   "rhs: chemdf + transport stencil (emitted chemdf + rhs_stencil_kernel, or rhs_warp_kernel)": 2.0 * nz * ni * 8.0,                                       # reads y, writes f; k is shared (L2)
                   "emitted chemdf kernel (part of the rhs line)": 2.0 * nz * ni * 8.0,          # reads y, writes chemdf; k is shared (L2)
                   "emitted Jacobian kernel (part of the lhs line)": nz * ni * nip * 8.0 + nz * ni * 8.0,   # writes the ni dense rows of D, reads y
                   "lu_solve_kernel (first solve: backward sweep, forward fused into the factorisation)": 1.0 * nz * nip * (nip + 2) * 8.0,
                   "lu_solve_kernel (forward + backward)": 2.0 * nz * nip * (nip + 2) * 8.0}     # reads the factors F_j once per sweep
            line["hbm_kernels"] = {"peak_gbs": hbm_peak, "peak_source": hbm_src, "kernels": [
                {"kernel": name, "ms": msv, "algorithmic_bytes_per_column": alg[name], "achieved_gbs": alg[name] * ncol / (msv * 1e-3) / 1e9,
                 "frac": alg[name] * ncol / (msv * 1e-3) / 1e9 / hbm_peak} for name, msv in kernel_ms.items()]}
        except Exception:
            pass
        # photolysis update (compute_tau / compute_flux / compute_J, op.py:2580-2786) of a 64-column batch, CUDA events around flux_kernel +
        # jrate_kernel; algorithmic bytes per column: tau, sflux, dflux_u, dflux_d written + dflux_u read ((nz+1) x nbin x 8 B each), aflux
        # read (change) + written + read by the J contraction, J rows written; the cross-section tables are shared by the batch (L2)
        try:
            from vulcan_b200.fixtures import Case as _Case, steady_ensemble_from_fixture as _sef
            c0 = _Case("HD189", 0)
            npc = 64
            kz64, met64, co64 = [a[:npc] for a in ensemble.sweep_grid()]
            y64, ai64 = ensemble.synthetic_columns(c0.st["y_ini"], c0.st["n_0"], c0.st["compo"], c0.cfg["atom_list"], kz64, met64, co64)
            se64 = _sef(c0, y64, ai64, kz64, refine=refine, hist_cap=2, hist_stride=1)
            for _ in range(2):
                se64.col.ens_photo_update()
            ms_ph = []
            for _ in range(3):
                se64.col.ens_photo_update()
                ms_ph.append(se64.col.last_kernel_ms()[0])
            nbin = int(c0.st["nbin"])
            nbr = int(c0.st["cross_J"].shape[0])
            alg_ph = (5.0 * (nz + 1) * nbin + 3.0 * nz * nbin + nbr * nz) * 8.0
            msv = float(np.mean(ms_ph))
            line.setdefault("hbm_kernels", {"kernels": []})["kernels"].append(
                {"kernel": "photolysis update (flux_kernel + jrate_kernel), %d columns" % npc, "ms": msv, "algorithmic_bytes_per_column": alg_ph,
                 "achieved_gbs": alg_ph * npc / (msv * 1e-3) / 1e9, "frac": alg_ph * npc / (msv * 1e-3) / 1e9 / line["hbm_kernels"].get("peak_gbs", 6453.4),
                 "note": "one thread per (column, wavelength bin) marches the %d layers: latency-bound (exp, sqrt, divisions per layer), not a streaming kernel" % nz})
            se64.col.close()
        except Exception as e:
            line["photolysis_timing_error"] = repr(e)
        # single-column numbers (BASELINE metric part 1)
        one = ensemble.EnsembleRunner(case.net, case.nz, case.y[None], np.array([case.dt]), atm_common, np.asarray(kw["Kzz"])[None],
                                      case.k, cfg, st["compo"], st["atom_ini"][None], st["n_0"], device=local_rank, refine=refine)
        one.run(5)
        ms1 = one.run(30)
        col1 = one.col
        y1 = case.y[None].copy(); m1 = case.ymix[None].copy()
        for _ in range(3):
            col1.ros2_solve(y1, m1, case.dt)
        t0 = time.time()
        for _ in range(20):
            col1.ros2_solve(y1, m1, case.dt)
        e2e1 = 20 / (time.time() - t0)
        line["single_column"] = {"steps_per_s": 30 / (ms1 * 1e-3), "ms_per_step": ms1 / 30, "e2e_steps_per_s": e2e1,
                                 "note": "attempted Ros2 steps of ONE HD189 column (latency path: block cyclic reduction over the layers, vk_cr.inl)"}
        # time-to-steady-state: ONE HD189 column from the reference's own initial state to the reference's own stopping rule, the whole
        # loop device-resident (vk_ens_run_steady: steps, accept / reject, conv against the on-device history, photolysis cadence,
        # update_mu_dz); reference numbers from tests/golden/HD189_full.npz (unmodified reference, 1 core, build container)
        try:
            from vulcan_b200.fixtures import Case, GOLD, steady_ensemble_from_fixture
            c0 = Case("HD189", 0)
            t_ss = time.time()
            se = steady_ensemble_from_fixture(c0, c0.st["y_ini"][None], c0.st["atom_ini"][None], np.ones(1), refine=refine)
            out = se.run_to_steady_state(max_iterations=6000)
            wall_ss = time.time() - t_ss
            ref = np.load(os.path.join(GOLD, "HD189_full.npz"))
            ym = out["y"][0] / out["y"][0].sum(axis=1, keepdims=True)
            rel = np.abs(ym - ref["ymix"]) / np.maximum(ref["ymix"], 1e-300)
            line["single_column"]["time_to_steady_state"] = {
                "wall_s": wall_ss, "loop_wall_s": out["wall_s"], "device_ms": out["device_ms"], "accepted_steps": int(out["n_accept"][0]),
                "rejected_attempts": int(out["n_reject"][0]), "model_time_s": float(out["t"][0]), "converged": bool(out["end_case"][0] == 1),
                "attempts_per_s": float(out["n_accept"][0] + out["n_reject"][0]) / out["wall_s"],
                "vs_reference_final_state": {"max_rel_gt_1e-4": float(rel[ref["ymix"] > 1e-4].max()), "max_rel_gt_1e-12": float(rel[ref["ymix"] > 1e-12].max()),
                                             "median_gt_1e-20": float(np.median(rel[ref["ymix"] > 1e-20]))},
                "reference_cpu": {"wall_s": float(ref["wall_s"]), "accepted_steps": int(ref["count"]),
                                  "rejected_attempts": int(ref["delta_count"]) + int(ref["nega_count"]) + int(ref["loss_count"]),
                                  "model_time_s": float(ref["t"]), "where": "unmodified numpy/scipy reference, 1 core, build container"},
                "speedup_vs_reference_cpu": float(ref["wall_s"]) / wall_ss}
            se.col.close()
        except Exception as e:      # never lose the bench line over the auxiliary number
            line["single_column"]["time_to_steady_state"] = {"error": repr(e)}
        if not args.no_cpu_baseline:
            v, ms_cpu, n_done, per_step, cb, cp = measure_cpu_arm(case, 2, 1, budget_s=40.0)
            line["cpu_baseline"] = cb
            if cp is not None:
                line["cpu_port"] = cp
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
