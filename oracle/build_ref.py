#!/usr/bin/env python
"""TEST INFRASTRUCTURE - recipe for oracle/_ref/: the UNMODIFIED reference staged so that it can run on the GPU box's host cores.

/root/reference exists only in the build container, and the reference is pure Python whose generated chem_funs.py needs sympy minutes to
make.  This recipe (run by __graft_entry__.build() whenever /root/reference is present) stages a runnable copy of the reference for
the HD189 config - oracle/stage_reference.py's non-numerical shims, chem_funs.py pre-generated here - into oracle/_ref/HD189/.
oracle/_ref/ is git-ignored (it holds reference files, which never enter this repo's history) but NOT gpurun-ignored, so it travels to
the GPU box like the repo's own built .so files.  Consumers: bench.py --impl reference (oracle/ref_worker.py times real op.Ros2.solver
calls), oracle/dropin_in_reference.py (the drop-in class inside the reference's own op.Integration loop, CUDA library behind it).
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.path.join(HERE, "_ref")
PRUNE = ["demo", "plot_py", "output", "plot", "tools", "cfg_examples", "fastchem_vulcan/src", "fastchem_vulcan/obj", "README.md",
         "vulcan_cfg_README.txt", "GPL_license.txt", "__pycache__"]


def staged(config="HD189"):
    d = os.path.join(REF_ROOT, config)
    return d if os.path.exists(os.path.join(d, "chem_funs.py")) and os.path.exists(os.path.join(d, "op.py")) else None


def build(config="HD189", force=False):
    sys.path.insert(0, HERE)
    import stage_reference
    if not os.path.isdir(stage_reference.REF):
        return staged(config)                       # GPU box: use what travelled
    if staged(config) and not force:
        return staged(config)
    tmp = "/tmp/vulcan_ref_stage_%s_%d" % (config, os.getpid())
    stage_reference.stage(config, tmp)
    for rel in PRUNE:
        p = os.path.join(tmp, rel)
        if os.path.isdir(p):
            shutil.rmtree(p)
        elif os.path.exists(p):
            os.remove(p)
    for d in ("output", "plot", "plot/movie"):
        os.makedirs(os.path.join(tmp, d), exist_ok=True)
    dest = os.path.join(REF_ROOT, config)
    if os.path.exists(dest):
        shutil.rmtree(dest)
    os.makedirs(REF_ROOT, exist_ok=True)
    shutil.move(tmp, dest)
    return dest


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
