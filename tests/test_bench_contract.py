"""bench.py contract checks that need no GPU: the reference arm (`--impl reference`: the CPU port of the path on the host cores) prints
ONE JSON line with the keys the driver reads, alone and under torch.distributed.run with 2 ranks (rank 0 prints, rank 1 exits 0 without
work); the product arm refuses to run without a CUDA device instead of falling back to a CPU path."""
import json
import os
import subprocess
import sys

from helpers import REPO

ENV = dict(os.environ, VK_BENCH_CPU_STEPS_PER_THREAD="2", OMP_NUM_THREADS="1")
KEYS = {"impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
        "dtype", "data", "config", "cpu_baseline", "e2e"}


def _check_line(out, n_gpus):
    lines = [ln for ln in out.splitlines() if ln.startswith("{")]
    assert len(lines) == 1, out
    d = json.loads(lines[0])
    assert KEYS <= set(d), KEYS - set(d)
    assert d["impl"] == "reference" and d["n_gpus"] == n_gpus and d["metric"] == "ensemble column-steps/s"
    assert d["value"] > 0 and d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["value"] == d["value"] and "workload" in d["config"]
    assert d["vs_baseline"] is None and d["dtype"] == "f64" and d["higher_is_better"] is True


def test_reference_arm_single_process():
    r = subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, env=ENV, cwd=REPO, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    _check_line(r.stdout, 1)


def test_reference_arm_under_torchrun_two_ranks():
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29631", os.path.join(REPO, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                        "--warmup", "1"], capture_output=True, text=True, env=ENV, cwd=REPO, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    _check_line(r.stdout, 2)


def test_product_arm_has_no_cpu_path():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("a CUDA device is present: the product arm would run")
    r = subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), "--steps", "1", "--warmup", "1"], capture_output=True, text=True,
                       env=ENV, cwd=REPO, timeout=600)
    assert r.returncode != 0 and "no CUDA device" in (r.stderr + r.stdout)
    assert not [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
