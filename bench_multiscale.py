#!/usr/bin/env python
"""Secondary benchmark: multi-resolution point-cloud construction that produces ImageRegistrator's inputs —
CreateMultiScalePointCloud's scale loop (MergeClosePoints per scale) + DeterminePointNeighbors per scale, through the C ABI with host
buffers in and out. Workload: `--scans` synthetic room scans of `--scan-w x --scan-h` rays concatenated in scan order (the order the
reference processes them in), per-point minimum / maximum radii as ComputeMinMaxPointRadius would give for cameras ~2 m from the
surfaces (0.5 px footprint at f = 4400 px; maximum = minimum x 2^5). Metric: input points/s. Prints one JSON line for BASELINE.md."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def workload(nscans, w, h, seed=40):
    from dataset_pipeline_b200 import synth
    xs, ss = [], []
    for i in range(nscans):
        xyz, _, _ = synth.room_scan(i, w, h)
        T = np.eye(4); T[:3, :3] = synth.rot_xyz(0, 0, 0.35 * i); T[:3, 3] = synth.SCANNER_POSITIONS[i]
        xs.append((xyz.astype(np.float64) @ T[:3, :3].T + T[:3, 3]).astype(np.float32)); ss.append(np.full(len(xyz), i, np.uint8))
    x = np.concatenate(xs); s = np.concatenate(ss)
    rng = np.random.default_rng(seed)
    col = rng.uniform(0, 255, len(x)).astype(np.float32)
    cam = np.array([5.0, 4.0, 1.6], np.float32)
    depth = np.maximum(np.linalg.norm(x - cam, axis=1), 0.5)
    lo = (0.5 / 4400.0 * depth).astype(np.float32)                 # radius projecting to 0.5 px
    hi = (lo * 32.0).astype(np.float32)                             # / minimum_scaling_factor = 2^-(6-1)
    return x, col, s, lo, hi


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scans", type=int, default=3)
    ap.add_argument("--scan-w", type=int, default=5000)
    ap.add_argument("--scan-h", type=int, default=2000)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--cpu-w", type=int, default=1000)
    ap.add_argument("--cpu-h", type=int, default=400)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--profile", action="store_true", help="cudaProfilerStart/Stop around one pass (for `ncu --profile-from-start off`)")
    a = ap.parse_args()
    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench_multiscale.py: no CUDA device — no CPU fallback")
    import dataset_pipeline_b200 as b2
    t0 = time.perf_counter()
    x, col, s, lo, hi = workload(a.scans, a.scan_w, a.scan_h)
    t_gen = time.perf_counter() - t0
    n = len(x)

    def run():
        t = time.perf_counter()
        scales, st = b2.CreateMultiScalePointCloud(x, col, s, lo, hi, a.scans, return_stats=True)
        t1 = time.perf_counter()
        nb = []
        for (_, px, _, ps) in scales:
            counts = np.bincount(ps, minlength=a.scans)
            if counts.min() < 26:                                   # Problem::ComputeMultiResPointCloud drops such scales (problem.cc:205-241)
                continue
            nb.append(b2.DeterminePointNeighbors(a.scans, True, px, ps))
        t2 = time.perf_counter()
        return scales, st, nb, t1 - t, t2 - t1

    for _ in range(a.warmup):
        run()
    if a.profile:
        torch.cuda.profiler.start(); run(); torch.cuda.synchronize(); torch.cuda.profiler.stop()
        print(json.dumps({"profile": True, "points": int(n)}))
        return
    tm, tn = [], []
    for _ in range(a.steps):
        scales, st, nb, dm, dn = run()
        tm.append(dm); tn.append(dn)
    dt = float(np.mean(tm) + np.mean(tn))
    res = {"metric": "multi-resolution point cloud input points/sec (CreateMultiScalePointCloud + DeterminePointNeighbors)", "value": n / dt, "unit": "points/s",
           "n_gpus": 1, "steps": a.steps, "warmup": a.warmup, "higher_is_better": True, "data": "synthetic", "dtype": "f32 / u32",
           "config": {"workload": "%d room scans %dx%d concatenated (%d points), 0.5 px radii at f=4400, 6 image scales" % (a.scans, a.scan_w, a.scan_h, n),
                      "generation_s": t_gen, "scales": [{"radius": r, "points": int(len(p))} for (r, p, _, _) in scales], "neighbor_scales": len(nb)},
           "seconds": {"create_multi_scale": float(np.mean(tm)), "determine_neighbors": float(np.mean(tn)), "merge_device_ms": st["ms_device"]},
           "merge": {"neighbor_pairs": int(st["neighbor_pairs"]), "rounds": int(st["rounds"])}}
    if not a.no_cpu_baseline:
        from oracle import oracle as orc
        cx, cc, cs, clo, chi = workload(a.scans, a.cpu_w, a.cpu_h)
        # the same surface density per merge ball as the full workload: scale the radii with the sample spacing
        f = (a.scan_w * a.scan_h / float(a.cpu_w * a.cpu_h)) ** 0.5
        t = time.perf_counter(); sc = orc.ms_create(cx, cc, cs, (clo * f).astype(np.float32), (chi * f).astype(np.float32), a.scans); tc = time.perf_counter() - t
        t = time.perf_counter()
        for (_, px, _, ps) in sc:
            if np.bincount(ps, minlength=a.scans).min() >= 26:
                orc.ms_point_neighbors(px, ps, a.scans, True)
        tnb = time.perf_counter() - t
        res["cpu_baseline"] = {"value": len(cx) / (tc + tnb), "unit": "points/s", "cores": 1, "kind": "port",
                               "sample": "oracle (serial kd-tree sweeps, as the reference) on %d scans of %dx%d (%d points) with radii scaled to the same points per "
                                         "merge ball: create %.1f s, neighbours %.1f s" % (a.scans, a.cpu_w, a.cpu_h, len(cx), tc, tnb)}
    print(json.dumps(res))


if __name__ == "__main__":
    main()
