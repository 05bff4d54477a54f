#!/usr/bin/env python
"""Secondary benchmark: PointCloudCleaner's filter chain and SplatCreator's per-point body on one room scan.
Metrics: points/s through b2_lsor_filter for ETH3D's own recipe `--filter 270,1.15 --filter 20,1.15` (README.md:372) and points/s through
b2_splat_create against the room's walls / floor / ceiling tessellated at 5 cm (the furniture is "not represented by the mesh").
Host buffers in and out (the copies are inside the timed region). One JSON line per tool."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def room_shell_mesh(step=0.05):
    """The six planes of the synthetic room [0,10] x [0,8] x [0,3] (dataset_pipeline_b200/synth/scene.c) as a triangle mesh."""
    V, F = [], []
    dims = (10.0, 8.0, 3.0)
    for a in range(3):
        b, c = (a + 1) % 3, (a + 2) % 3
        nb, nc = int(round(dims[b] / step)), int(round(dims[c] / step))
        ub, uc = np.meshgrid(np.linspace(0, dims[b], nb + 1), np.linspace(0, dims[c], nc + 1), indexing="ij")
        i0 = (np.arange(nb)[:, None] * (nc + 1) + np.arange(nc)[None, :]).ravel()
        for side in (0.0, dims[a]):
            P = np.zeros((ub.size, 3)); P[:, a] = side; P[:, b] = ub.ravel(); P[:, c] = uc.ravel()
            base = sum(len(v) for v in V)
            V.append(P)
            F.append(base + np.concatenate([np.stack([i0, i0 + 1, i0 + nc + 1], 1), np.stack([i0 + 1, i0 + nc + 2, i0 + nc + 1], 1)]))
    return np.concatenate(V).astype(np.float32), np.concatenate(F).astype(np.uint32)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scan-w", type=int, default=5000)
    ap.add_argument("--scan-h", type=int, default=2000)
    ap.add_argument("--filters", default="270,1.15;20,1.15")
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--cpu-points", type=int, default=200000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    a = ap.parse_args()
    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench_cleaner.py: no CUDA device — no CPU fallback")
    import dataset_pipeline_b200 as b2
    from dataset_pipeline_b200 import synth
    filters = [(int(f.split(",")[0]), float(f.split(",")[1])) for f in a.filters.split(";")]
    xyz, nrm, T = synth.room_scan(0, a.scan_w, a.scan_h, seed=20)
    xyz = (xyz.astype(np.float64) @ T[:3, :3].T + T[:3, 3]).astype(np.float32)          # global frame, like the mesh
    nrm = (nrm.astype(np.float64) @ T[:3, :3].T).astype(np.float32)
    n = xyz.shape[0]
    cores = os.cpu_count()

    def timed(fn):
        for _ in range(a.warmup):
            r = fn()
        torch.cuda.synchronize()
        t = time.perf_counter()
        for _ in range(a.steps):
            r = fn()
        return (time.perf_counter() - t) / a.steps, r

    dt, alive = timed(lambda: b2.clean_point_cloud(xyz, filters))
    res = {"metric": "PointCloudCleaner points/sec (filters %s, %d-pt scan)" % (a.filters, n), "value": n / dt, "unit": "points/s", "n_gpus": 1,
           "steps": a.steps, "warmup": a.warmup, "seconds_per_call": dt, "higher_is_better": True, "data": "synthetic", "dtype": "f32 distances, f64 sums",
           "config": {"workload": "room scan %dx%d rays, LocalStatisticalOutlierRemoval chain %s" % (a.scan_w, a.scan_h, a.filters), "points": n,
                      "kept": int(len(alive)), "h2d_bytes": n * 12}}
    if not a.no_cpu_baseline:
        from oracle import oracle as orc
        w = int((a.cpu_points * 2.5) ** 0.5); h = int(w / 2.5)
        sx, _, _ = synth.room_scan(0, w, h, seed=20)
        t = time.perf_counter()
        keep = np.arange(sx.shape[0])
        for k, f in filters:
            kk, _, _ = orc.lsor_filter(sx[keep], k, f); keep = keep[kk]
        tc = time.perf_counter() - t
        res["cpu_baseline"] = {"value": sx.shape[0] / tc, "unit": "points/s", "cores": cores, "kind": "port",
                               "sample": "oracle (kd-tree kNN with OpenMP over points; the reference's filter is serial) on a %d-point scan of the same scene" % sx.shape[0]}
    print(json.dumps(res))

    V, F = room_shell_mesh()
    dt, (corners, added, radius) = timed(lambda: b2.create_splats(xyz, nrm, V, F, 0.02, np.inf))
    res = {"metric": "SplatCreator points/sec (%d-pt scan, %d-triangle mesh)" % (n, len(F)), "value": n / dt, "unit": "points/s", "n_gpus": 1,
           "steps": a.steps, "warmup": a.warmup, "seconds_per_call": dt, "higher_is_better": True, "data": "synthetic", "dtype": "f32",
           "config": {"workload": "room scan %dx%d rays vs the room shell tessellated at 5 cm, distance_threshold 0.02" % (a.scan_w, a.scan_h),
                      "points": n, "triangles": int(len(F)), "splats": int(added.sum()), "h2d_bytes": n * 24 + V.nbytes + F.nbytes, "d2h_bytes": n * 53}}
    if not a.no_cpu_baseline:
        from oracle import oracle as orc
        m = min(n, 2000)
        sel = np.linspace(0, n - 1, m).astype(np.int64)
        t = time.perf_counter(); orc.mesh_squared_distance(xyz[sel], V, F); tc = time.perf_counter() - t
        res["cpu_baseline"] = {"value": m / (5 * tc), "unit": "points/s", "cores": cores, "kind": "port",
                               "sample": "oracle brute-force point-mesh distance (no tree; libigl's AABB tree is not restated) for %d points x <=5 queries each: a LOWER bound on the reference's rate, not a like-for-like baseline" % m}
    print(json.dumps(res))


if __name__ == "__main__":
    main()
