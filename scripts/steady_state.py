"""HD189 to steady state on the GPU (drop-in solver + Integration mirror); prints the comparison with the reference's full run
and writes gpurun_out/steady_state.json + the per-step trajectory."""
import json, os, sys, time
import numpy as np
REPO = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, "tests"))
from helpers import GOLD
from test_gpu_steady_state import run_hd189

refine = int(sys.argv[1]) if len(sys.argv) > 1 else 0
case, var, atm, para, integ, wall = run_hd189(refine=refine)
ref = np.load("%s/HD189_full.npz" % GOLD)
tr = ref["traj"]
t_ours = np.array(var.t_time)
n_rej = para.delta_count + para.nega_count + para.loss_count
ym, yr = var.ymix, ref["ymix"]
rel = np.abs(ym - yr) / np.maximum(yr, 1e-300)
out = dict(refine=refine, steps=para.count, rejected=n_rej, t=var.t, wall_s=wall, photo_updates=integ.n_photo_updates, photo_s=integ.t_photo,
           end_case=para.end_case, longdy=float(var.longdy), longdydt=float(var.longdydt),
           ref_steps=int(ref["count"]), ref_rejected=int(ref["delta_count"]) + int(ref["nega_count"]) + int(ref["loss_count"]),
           ref_t=float(ref["t"]), ref_wall_s=float(ref["wall_s"]), ref_longdy=float(ref["longdy"]), ref_longdydt=float(ref["longdydt"]))
for thr in (1e-20, 1e-12, 1e-8, 1e-4):
    m = yr > thr
    out["maxrel_gt_%g" % thr] = float(rel[m].max()); out["median_gt_%g" % thr] = float(np.median(rel[m]))
# lock-step length: how long does our accepted-step time grid follow the reference's?
n = min(len(t_ours), len(tr) - 1)
t_ref_after = tr[1:n + 1, 1]          # t_before of step c+1 = t after step c
dev = np.abs(t_ours[:n] - t_ref_after) / t_ref_after
first = int(np.argmax(dev > 1e-6)) if np.any(dev > 1e-6) else n
out["steps_in_lockstep_1e-6"] = first
print(json.dumps(out, indent=1))
j, i = np.unravel_index(np.argmax(np.where(yr > 1e-12, rel, 0)), rel.shape)
print("worst >1e-12: layer %d species %s ours %.4e ref %.4e" % (j, case.net.species[i], ym[j, i], yr[j, i]))
os.makedirs(os.path.join(REPO, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(REPO, "gpurun_out", "steady_state_refine%d.json" % refine), "w"), indent=1)
np.savez_compressed(os.path.join(REPO, "gpurun_out", "steady_state_traj_refine%d.npz" % refine), t=t_ours, ymix=ym)
