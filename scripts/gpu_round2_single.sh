timeout 300 python scripts/prof_steady_single3.py 2>&1 | tail -6
timeout 900 python -m pytest tests/test_gpu_device_steady.py tests/test_gpu_steady_state.py tests/test_gpu_parity.py -q -p no:cacheprovider -k "hd189 or HD189 or device_loop or conserves or ros2_step or replay" 2>&1 | tail -4
python - <<'PY'
import sys, os, time, numpy as np
sys.path.insert(0, os.getcwd())
from vulcan_b200.fixtures import Case, steady_ensemble_from_fixture
c0 = Case("HD189", 0)
t0 = time.time()
se = steady_ensemble_from_fixture(c0, c0.st["y_ini"][None], c0.st["atom_ini"][None], np.ones(1))
out = se.run_to_steady_state(max_iterations=6000)
print("HD189 to steady state: %.3f s (loop %.3f s), %d accepted, %d rejected, end_case %d" % (time.time() - t0, out["wall_s"], out["n_accept"][0], out["n_reject"][0], out["end_case"][0]))
print("refine stats", se.col.refine_stats())
PY
