"""Builds libvulcan_b200.so (hand-written sm_100a CUDA behind the C ABI of include/vulcan_b200.h) in-tree with nvcc.

vk_chem.cu / vk_step.cu / vk_photo.cu are compiled with -fmad=false: they restate the reference's numpy expressions
operation by operation (bit-identical RHS); vk_solve.cu (the block-tridiagonal factor/solve) uses FMA.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "_lib")
LIB = os.path.join(LIBDIR, "libvulcan_b200.so")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"]
UNITS = {"vk_api.cu": [], "vk_chem.cu": ["-fmad=false"], "vk_step.cu": ["-fmad=false"], "vk_photo.cu": ["-fmad=false"],
         "vk_solve.cu": [], "vk_ens.cu": ["-fmad=false"], "vk_steady.cu": ["-fmad=false"], "vk_conden.cu": ["-fmad=false"], "vk_rates.cu": ["-fmad=false"],
         "vk_emit.cu": []}


def _newer(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, variant=None, solve_flags=()):
    """variant / solve_flags: diagnostic builds (scripts/bisect_solve_budget.py): extra nvcc flags for vk_solve.cu only, objects and
    library under a suffixed name (libvulcan_b200_<variant>.so); the product build takes neither."""
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    os.makedirs(LIBDIR, exist_ok=True)
    lib = LIB if not variant else os.path.join(LIBDIR, "libvulcan_b200_%s.so" % variant)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh")]
    headers.append(os.path.join(os.path.dirname(HERE), "include", "vulcan_b200.h"))
    objs, procs = [], []
    units = dict(UNITS)
    # emitted chemistry kernels: one generated unit per registered network (vulcan_b200/emit.py, networks/*.npz -> csrc/gen/*.cu)
    from . import emit
    for g in emit.generate_all(verbose=verbose):
        units[os.path.join("gen", os.path.basename(g))] = ["-fmad=false"]
    for src, extra in units.items():
        s = os.path.join(CSRC, src)
        o = os.path.join(LIBDIR, os.path.basename(src).replace(".cu", ".o"))
        if variant and src == "vk_solve.cu":
            o = os.path.join(LIBDIR, "vk_solve_%s.o" % variant)
            extra = list(extra) + list(solve_flags)
        if variant == "prodtrace" and src == "vk_chem.cu":
            o = os.path.join(LIBDIR, "vk_chem_prodtrace.o")
            extra = list(extra) + ["-DVK_PROD_TRACE"]
        objs.append(o)
        if force or _newer(o, [s] + headers):
            cmd = [nvcc] + ARCH + COMMON + extra + (["-Xptxas", "-v"] if verbose else []) + ["-c", s, "-o", o]
            procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            sys.stderr.write("---- %s\n%s\n" % (src, out))
        failed = failed or p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed")
    if force or procs or _newer(lib, objs):
        subprocess.run([nvcc] + ARCH + ["-shared", "-o", lib] + objs + ["-lcudart"], check=True)
    return lib


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
