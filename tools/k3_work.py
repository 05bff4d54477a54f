#!/usr/bin/env python
"""Diagnostic (not a benchmark): per-query work of the K3 radius-1NN search on two scans of the bench scene.
Runs b2.find_correspondences with B2_K3_WORK set and summarises where the work concentrates (per query, per 128-query CTA)."""
import os, sys, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["B2_K3_WORK"] = "/tmp/k3work"
import dataset_pipeline_b200 as b2
from dataset_pipeline_b200 import synth

W = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
H = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
d = float(sys.argv[3]) if len(sys.argv) > 3 else 0.01
si, ti = (int(sys.argv[4]), int(sys.argv[5])) if len(sys.argv) > 5 else (0, 1)


def world(i):
    xyz, _, T = synth.room_scan(i, W, H)
    return (xyz.astype(np.float64) @ T[:3, :3].T + T[:3, 3]).astype(np.float32)


a, b = world(si), world(ti)
q, m, d2 = b2.find_correspondences(a, b, d)
for k in range(4):
    fn = "/tmp/k3work.%d.bin" % k
    if not os.path.exists(fn):
        continue
    raw = np.fromfile(fn, np.uint8)
    n = raw.size // 32
    pts = raw[: n * 16].view(np.float32).reshape(n, 4)
    wk = raw[n * 16:].view(np.uint32).reshape(n, 4).astype(np.int64)
    cost = wk[:, 0] + wk[:, 1] + wk[:, 2]            # points + box tests
    ncta = n // 128
    # a warp costs its slowest lane
    warp = cost[: ncta * 128].reshape(-1, 32).max(1).reshape(ncta, 4).sum(1)
    order = np.argsort(-warp)
    out = {"file": fn, "queries": int(n), "mean_points": float(wk[:, 0].mean()), "mean_box1": float(wk[:, 1].mean()), "mean_box2": float(wk[:, 2].mean()),
           "mean_cells": float(wk[:, 3].mean()), "p50_cost": float(np.percentile(cost, 50)), "p99_cost": float(np.percentile(cost, 99)),
           "max_cost": int(cost.max()), "cta_mean": float(warp.mean()), "cta_p99": float(np.percentile(warp, 99)), "cta_max": int(warp.max()),
           "top_ctas": []}
    for c in order[:12]:
        sl = slice(c * 128, c * 128 + 128)
        out["top_ctas"].append({"cta": int(c), "frac_of_grid": float(c) / ncta, "warp_cost": int(warp[c]), "centre": [round(float(v), 3) for v in pts[sl, :3].mean(0)],
                                "points_max": int(wk[sl, 0].max()), "box1_max": int(wk[sl, 1].max()), "box2_max": int(wk[sl, 2].max())})
    # histogram of CTA cost along the grid (16 equal slices)
    out["cta_cost_by_grid_slice"] = [float(x.mean()) for x in np.array_split(warp, 16)]
    print(json.dumps(out))
