#!/usr/bin/env python
"""TEST INFRASTRUCTURE - packs the outputs of oracle/fixed_point_reference.py (two hash seeds of the unmodified reference run with the
tightened stopping rule) into one small committed fixture tests/golden/<cfg>_fixedpoint.npz: the two final states, three late snapshots of
each seed (to measure the drift WITHIN one reference run), the per-step (t, dt, longdy) traces and the spreads the reference has against itself.
usage: python oracle/pack_fixed_point.py HD189 /tmp/fp_out"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(os.path.dirname(HERE), "tests", "golden")


def spread(x, y):
    rel = np.abs(x - y) / np.maximum(y, 1e-300)
    return {"%g" % thr: [float(rel[y > thr].max()), float(np.median(rel[y > thr]))] for thr in (1e-20, 1e-12, 1e-8, 1e-4)}


def main():
    cfg, src = sys.argv[1], sys.argv[2]
    a = np.load(os.path.join(src, "%s_fp_seed0.npz" % cfg))
    b = np.load(os.path.join(src, "%s_fp_seed1.npz" % cfg))
    out = dict(tightened=str(a["tightened"]))
    info = {}
    for name, d in (("seed0", a), ("seed1", b)):
        out["ymix_" + name] = d["ymix"]
        sel = [q for q in range(len(d["snap_counts"])) if d["snap_counts"][q] % 500 == 0][-3:]
        out["snaps_" + name] = d["snaps"][sel]
        out["snap_counts_" + name] = d["snap_counts"][sel]
        out["traj_" + name] = d["traj"][:, [0, 1, 3, 6]].astype(np.float32)          # count, t_before, dt_used, longdy
        info[name] = dict(count=int(d["count"]), t=float(d["t"]), end_case=int(d["end_case"]), longdy=float(d["longdy"]), wall_s=float(d["wall_s"]),
                          rejected=int(d["delta_count"]) + int(d["nega_count"]) + int(d["loss_count"]), atom_loss=[float(v) for v in d["atom_loss"]],
                          longdy_p5_p50_p95_last1000=[float(v) for v in np.nanpercentile(d["traj"][-1000:, 6], [5, 50, 95])],
                          dt_p5_p50_p95_last1000=[float(v) for v in np.percentile(d["traj"][-1000:, 3], [5, 50, 95])],
                          drift_last_500_steps=spread(d["snaps"][sel[-1]], d["snaps"][sel[-2]]))
    info["seed0_vs_seed1_final"] = spread(a["ymix"], b["ymix"])
    out["info_json"] = json.dumps(info)
    np.savez_compressed(os.path.join(GOLD, "%s_fixedpoint.npz" % cfg), **out)
    print(json.dumps(info, indent=1))


if __name__ == "__main__":
    main()
