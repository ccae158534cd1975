"""CPU study (no GPU): a BASELINE config to steady state through the drop-in Ros2 class + Integration mirror on the oracle-backed
stand-in of the C ABI (tests/oracle_columns.py - test infrastructure, same algorithm as the GPU path), for a given refinement
setting: refine = 0 (shipped default), n > 0 (n fp64 passes), n < 0 (safeguarded study variant of oracle/vk_oracle.c).
    python scripts/study_refine_cpu.py HD209S -2
"""
import os, sys
REPO = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, "tests")); sys.path.insert(0, os.path.join(REPO, "oracle"))
from helpers import run_config            # noqa: E402
from oracle_columns import oracle_backed_abi  # noqa: E402

tag, refine = sys.argv[1], int(sys.argv[2])
count_max = int(sys.argv[3]) if len(sys.argv) > 3 else None
case, var, atm, para, integ, wall = run_config(tag, refine=refine, max_wall_s=3000, count_max=count_max, abi=oracle_backed_abi())
rej = getattr(para, "rejected", None)
print("%s refine %d: accepted %d, model time %.4e s, last dt %.4e, wall %.0f s, element loss %s, para: %s" % (
    tag, refine, para.count, var.t, var.dt, wall, {a: "%.2e" % v for a, v in var.atom_loss.items()},
    {k: v for k, v in vars(para).items() if isinstance(v, (int, float)) and ("rej" in k or "count" in k)}))
