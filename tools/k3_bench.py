#!/usr/bin/env python
"""K3 micro-benchmark: search phase of one outer iteration on the first N config-2 scans (all ordered pairs), repeated.
    B2_LIB_PATH=... python tools/k3_bench.py [--scans 4] [--reps 4]"""
import argparse, json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench as B

ap = argparse.ArgumentParser(); ap.add_argument("--scans", type=int, default=4); ap.add_argument("--reps", type=int, default=4)
a = ap.parse_args()
import dataset_pipeline_b200 as b2
B.ensure_scans(range(a.scans), 5000, 2000)
poses, _ = B.scene_poses(8)
g = b2.PointToPlaneICP(inner_max_iterations=1)
g.set_option("pack_overlap", int(os.environ.get("K3_BENCH_PACK_OVERLAP", "0")))      # 0: ms_search is K3 alone
for i in range(a.scans):
    xyz, nrm = B.load_scan(i, 5000, 2000)
    g.AddPointCloud(xyz, nrm, poses[i])
ms = []
for r in range(a.reps + 1):
    for i in range(a.scans):
        g.SetGlobalTCloud(i, poses[i])
    g.Run(0.01, r, 1, 1e-10, False)
    st = g.stats()
    if r:
        ms.append(st["ms_search"])
print(json.dumps({"lib": os.environ.get("B2_LIB_PATH", "default"), "env": {k: v for k, v in os.environ.items() if k.startswith("B2_K3")},
                  "ms_search": float(np.mean(ms)), "per_direction_ms": float(np.mean(ms)) / st["search_launches"], "corr": st["num_correspondences"],
                  "ms_index": st["ms_index"], "ms_pack": st["ms_pack"]}))
