#!/usr/bin/env python
"""Secondary benchmark (BASELINE.json configs[2]): NormalEstimator, kNN = 16, one 50 M-point scan, 1 GPU.
Metric: points/s through b2_normals_estimate (host buffers in and out: H2D of 12 B/pt and D2H of 16 B/pt are inside the timed
region, as the tool would call it). Prints one JSON line for BASELINE.md."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scan-w", type=int, default=10000)
    ap.add_argument("--scan-h", type=int, default=5000)
    ap.add_argument("--k", type=int, default=16)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--cpu-points", type=int, default=2000000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--profile", action="store_true", help="cudaProfilerStart/Stop around one call (for `ncu --profile-from-start off`)")
    a = ap.parse_args()
    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench_normals.py: no CUDA device — no CPU fallback")
    import dataset_pipeline_b200 as b2
    from dataset_pipeline_b200 import synth
    t0 = time.perf_counter()
    xyz, nrm_true, _ = synth.room_scan(0, a.scan_w, a.scan_h, seed=20)
    t_gen = time.perf_counter() - t0
    n = xyz.shape[0]
    pinned = torch.empty(xyz.shape, dtype=torch.float32, pin_memory=True); pinned.numpy()[:] = xyz
    x = pinned.numpy()
    for _ in range(a.warmup):
        out = b2.estimate_normals(x, a.k, (0.0, 0.0, 0.0))
    torch.cuda.synchronize()
    if a.profile:
        torch.cuda.profiler.start()
        out = b2.estimate_normals(x, a.k, (0.0, 0.0, 0.0))
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        print(json.dumps({"profile": True, "points": int(n)}))
        return
    t = time.perf_counter()
    for _ in range(a.steps):
        out = b2.estimate_normals(x, a.k, (0.0, 0.0, 0.0))
    dt = (time.perf_counter() - t) / a.steps
    flat = out[:, 3] < 1e-3
    agree = float((np.abs((out[flat, :3] * nrm_true[flat]).sum(1)) > 0.99).mean())
    res = {"metric": "NormalEstimator points/sec (kNN=%d, %d-pt scan)" % (a.k, n), "value": n / dt, "unit": "points/s", "n_gpus": 1, "steps": a.steps,
           "warmup": a.warmup, "seconds_per_call": dt, "higher_is_better": True, "data": "synthetic", "dtype": "f32",
           "config": {"workload": "room scan %dx%d rays from one position, kNN=%d, viewpoint = scan origin" % (a.scan_w, a.scan_h, a.k),
                      "points": n, "generation_s": t_gen, "h2d_bytes": n * 12, "d2h_bytes": n * 16,
                      "normals_agree_with_analytic_on_flat_surfaces": agree}}
    # roofline of the device portion of the call (Morton sort + BVH + kn_knn_normals): SURVEY §8d counts 12 (k + 1) + 16 B per point.
    # The call's H2D / D2H legs are timed separately on the same pinned buffers and taken off the call time.
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    dev_in = torch.empty(xyz.shape, dtype=torch.float32, device="cuda"); host_out = torch.empty((n, 4), dtype=torch.float32, pin_memory=True)
    dev_out = torch.empty((n, 4), dtype=torch.float32, device="cuda")
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(3):
        dev_in.copy_(pinned, non_blocking=True); host_out.copy_(dev_out, non_blocking=True); torch.cuda.synchronize()
    t_copy = (time.perf_counter() - t) / 3
    t_dev = max(dt - t_copy, 1e-9)
    alg = (12.0 * (a.k + 1) + 16.0) * n
    res["roofline"] = {"bound": "hbm", "kernel": "device portion of b2_normals_estimate (Morton sort, implicit BVH, kn_knn_normals: exact kNN + two-pass covariance + eigenvector)",
                       "achieved": alg / t_dev / 1e9, "peak": peak, "unit": "GB/s", "frac": alg / t_dev / 1e9 / peak, "traffic": None,
                       "algorithmic_bytes_per_call": alg, "device_seconds": t_dev, "copy_seconds": t_copy,
                       "note": "gather / traversal-latency bound (8 of 32 lanes active in the per-thread BVH walk, profiles/r01d_k7_knn_normals_ncu.txt), not bandwidth bound"}
    res["e2e"] = {"value": n / dt, "unit": "points/s", "h2d_bytes_per_step": int(n * 12), "d2h_bytes_per_step": int(n * 16)}
    if not a.no_cpu_baseline:
        from oracle import oracle as orc
        w = int((a.cpu_points * 2) ** 0.5); h = w // 2
        sx, _, _ = synth.room_scan(0, w, h, seed=20)
        t = time.perf_counter(); orc.normals_knn(sx, a.k, (0.0, 0.0, 0.0)); tc = time.perf_counter() - t
        res["cpu_baseline"] = {"value": sx.shape[0] / tc, "unit": "points/s", "cores": os.cpu_count(), "kind": "port",
                               "sample": "oracle (kd-tree build + OpenMP over points, as the reference) on a %d-point scan of the same scene" % sx.shape[0]}
    print(json.dumps(res))


if __name__ == "__main__":
    main()
