#!/usr/bin/env python
"""Secondary benchmark (Path B, BASELINE.json configs[3]): ImageRegistrator residual-evaluations/s.

A residual evaluation = one (image, point scale, observation) through AccumulateHAndBAndResidualsForObservations pass 1+2
(SURVEY.md §8d), i.e. one observation through kr_jacobians + kr_accumulate inside b2_reg_accumulate. The primary driver contract
is bench.py (ICP); this script prints one JSON line with the same style of keys for BASELINE.md.

Workload: N images (default 20 x 3008x2000, pinhole) of a textured plane rendered analytically on the GPU (harness only), a
multi-resolution point cloud of grids on that plane (~default 30 M points), no occlusion geometry (all visible).
"""
import argparse
import json
import math
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def build_scene(num_images, width, height, target_points, seed=31, camera_model=4, owns=lambda i: True):
    import torch
    from dataset_pipeline_b200.synth import reg_scene as rs
    dev = torch.device("cuda")
    fx = 0.8125 * width
    K = np.array([fx, fx, (width - 1) / 2.0, (height - 1) / 2.0], np.float32)
    dist_p = rs.DEFAULT_DISTORTION
    params = K if camera_model == rs.CAM_PINHOLE else np.concatenate([K, np.asarray(dist_p, np.float32)])
    rng = np.random.default_rng(seed)
    images, poses_gt, poses_init = [], [], []
    yy, xx = torch.meshgrid(torch.arange(height, device=dev, dtype=torch.float64), torch.arange(width, device=dev, dtype=torch.float64), indexing="ij")

    def tex(x, y):
        v = (torch.sin(7.0 * x) * torch.cos(5.0 * y) + 0.6 * torch.sin(19.0 * x + 1.3) * torch.sin(23.0 * y + 0.4) + 0.35 * torch.cos(41.0 * x - 29.0 * y)
             + 0.25 * torch.sin(83.0 * x + 61.0 * y) + 0.2 * torch.sin(211.0 * x - 173.0 * y) + 0.15 * torch.cos(431.0 * x + 389.0 * y))
        return 120.0 + 40.0 * v

    # pixel -> normalized ray coordinates (numerical inverse of the distortion for the non-pinhole models; harness only)
    ncx = (xx - float(K[2])) / float(K[0]); ncy = (yy - float(K[3])) / float(K[1])
    if camera_model != rs.CAM_PINHOLE:
        k1, k2, p1, p2, k3, k4, sx1, sy1 = [float(v) for v in dist_p]
        ux, uy = ncx.clone(), ncy.clone()
        for _ in range(60):
            x2, xy, y2 = ux * ux, ux * uy, uy * uy
            r2 = x2 + y2
            rad = 1 + r2 * (k1 + r2 * (k2 + r2 * (k3 + r2 * k4)))
            fx_ = ux * rad + 2 * p1 * xy + p2 * (r2 + 2 * x2) + sx1 * r2
            fy_ = uy * rad + 2 * p2 * xy + p1 * (r2 + 2 * y2) + sy1 * r2
            ux = ux + 0.8 * (ncx - fx_); uy = uy + 0.8 * (ncy - fy_)
        if camera_model == rs.CAM_BENCHMARK:
            r = torch.sqrt(ux * ux + uy * uy)
            f = torch.where(r > 1e-9, torch.tan(torch.clamp(r, max=1.5)) / torch.clamp(r, min=1e-9), torch.ones_like(r))
            ux, uy = ux * f, uy * f
        ncx, ncy = ux, uy
    for i in range(num_images):
        c = np.array([0.25 * math.cos(2.1 * i), 0.2 * math.sin(1.7 * i), 2.0 + 0.1 * math.sin(i)])
        R_wc = rs.rot(math.pi + 0.08 * math.sin(1.3 * i), 0.07 * math.cos(0.9 * i), 0.3 * i)
        R_cw = R_wc.T; t_cw = -R_cw @ c
        poses_gt.append(np.concatenate([rs.quat_from_R(R_cw), t_cw]).astype(np.float32))
        dR = rs.rot(*(rng.uniform(-0.0005, 0.0005, 3))); dt = rng.uniform(-0.001, 0.001, 3)
        poses_init.append(np.concatenate([rs.quat_from_R(dR @ R_cw), dR @ t_cw + dt]).astype(np.float32))
        if not owns(i):
            images.append(None)
            continue
        Rt = torch.tensor(R_wc, device=dev)
        dcx = ncx; dcy = ncy
        dwx = Rt[0, 0] * dcx + Rt[0, 1] * dcy + Rt[0, 2]; dwy = Rt[1, 0] * dcx + Rt[1, 1] * dcy + Rt[1, 2]; dwz = Rt[2, 0] * dcx + Rt[2, 1] * dcy + Rt[2, 2]
        s = -c[2] / dwz
        img = torch.clamp(torch.round(tex(c[0] + dwx * s, c[1] + dwy * s)), 0, 255).to(torch.uint8).cpu().numpy()
        images.append(img)
    # multi-resolution grids: scale k has radius r0 * 2^k; pixel footprint of scale 0 ~ 0.6 px at the finest image scale
    extent = (3.6, 2.6)
    nscales = 6
    # total points = sum_k (extent area / (2 r0 2^k)^2) = A/(4 r0^2) * 4/3
    r0 = math.sqrt(extent[0] * extent[1] / (3.0 * target_points))
    scales = []
    for k in range(nscales):
        radius = r0 * 2 ** k; step = 2 * radius
        nx = int(extent[0] / step); ny = int(extent[1] / step)
        if nx < 8 or ny < 8:
            break
        gx, gy = np.meshgrid(np.arange(nx, dtype=np.int64), np.arange(ny, dtype=np.int64), indexing="xy")
        x = ((gx.ravel() - (nx - 1) / 2.0) * step).astype(np.float32); y = ((gy.ravel() - (ny - 1) / 2.0) * step).astype(np.float32)
        xyz = np.stack([x, y, np.zeros_like(x)], 1)
        def nb(dx, dy):
            return (np.clip(gy + dy, 0, ny - 1) * nx + np.clip(gx + dx, 0, nx - 1)).ravel()
        idx = (gy * nx + gx).ravel()
        nbr = np.stack([nb(1, 0), nb(-1, 0), nb(0, 1), nb(0, -1), nb(1, 1)], 1)
        alt = np.stack([nb(-2, 0), nb(2, 0), nb(0, -2), nb(0, 2), nb(-1, -1)], 1)
        nbr = np.where(nbr == idx[:, None], alt, nbr).astype(np.uint64)
        colors = tex(torch.tensor(x, device=dev, dtype=torch.float64), torch.tensor(y, device=dev, dtype=torch.float64)).float().cpu().numpy()
        scales.append((xyz, np.float32(radius), nbr, colors))
    return {"intr": (width, height, params), "camera_model": camera_model, "images": images, "poses_gt": poses_gt, "poses_init": poses_init, "scales": scales}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--images", type=int, default=20)
    ap.add_argument("--width", type=int, default=3008)
    ap.add_argument("--height", type=int, default=2000)
    ap.add_argument("--points", type=float, default=30e6)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--camera", default="pinhole", choices=["pinhole", "thin_prism", "benchmark"])
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--profile", action="store_true", help="finest image scale only; cudaProfilerStart/Stop around one CreateObservations + ColorOptimizer + accumulate "
                    "(for `ncu --profile-from-start off`); prints no benchmark value")
    a = ap.parse_args()
    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench_reg.py: no CUDA device — no CPU fallback")
    import dataset_pipeline_b200 as b2
    from dataset_pipeline_b200 import registration as R
    from dataset_pipeline_b200.synth import reg_scene
    rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    comm = None
    if world > 1:      # one process per GPU (torchrun); images dealt round-robin, the library's own NCCL communicator for the sums
        import torch.distributed as dist
        from dataset_pipeline_b200.icp import Comm
        dist.init_process_group("gloo")
        ids = [Comm.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, 0)
        comm = Comm(rank, world, ids[0], device=local)
    model = {"pinhole": 4, "thin_prism": 14, "benchmark": 5}[a.camera]
    t0 = time.perf_counter()
    sc = build_scene(a.images, a.width, a.height, a.points, camera_model=model, owns=lambda i: i % world == rank)
    t_gen = time.perf_counter() - t0
    npts = sum(s[0].shape[0] for s in sc["scales"])
    g = b2.Registration(R.default_params(device=local))
    if comm is not None:
        g.set_comm(comm)
    nsc = reg_scene.load_into(g, sc, splats=False)
    if a.profile:
        g.set_image_scale(0)
        g.CreateObservationsForAllImages(1); g.ColorOptimizerApply(); g.accumulate()
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        g.CreateObservationsForAllImages(1); g.ColorOptimizerApply(); g.accumulate()
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        print(json.dumps({"profile": True, "stats": g.stats()}))
        return
    res = {}
    for scale in (nsc - 2, 0):
        g.set_image_scale(scale)
        g.CreateObservationsForAllImages(1)
        st_obs = g.stats()
        g.ColorOptimizerApply()
        for _ in range(a.warmup):
            g.accumulate()
        torch.cuda.synchronize()
        t = time.perf_counter()
        ms_j = ms_a = 0.0
        for _ in range(a.steps):
            g.accumulate()
            s = g.stats(); ms_j += s["ms_jacobian_kernel"]; ms_a += s["ms_accumulate_kernel"]
        dt = (time.perf_counter() - t) / a.steps
        if world > 1:   # the step time of the job is the slowest rank's
            tt = torch.tensor([dt, ms_j, ms_a], dtype=torch.float64); dist.all_reduce(tt, op=dist.ReduceOp.MAX); dt, ms_j, ms_a = [float(v) for v in tt]
        evals = g.stats()["residual_evaluations"]
        full = sum(int(g.observations(im, ps)[4].sum()) for im in range(min(2, a.images)) for ps in range(len(sc["scales"])))
        res[scale] = {"image_scale": scale, "observations": st_obs["observations"], "ms_create_observations": st_obs["ms_last_call"],
                      "residual_evaluations": evals, "s_per_accumulate": dt, "evals_per_s": evals / dt,
                      "ms_jacobian_kernels": ms_j / a.steps, "ms_accumulate_kernels": ms_a / a.steps, "fully_observed_first2_images": full}
    # one LM step + cost at the finest scale (end to end through the ABI)
    t = time.perf_counter(); ap_ = g.IntrinsicsAndPoseOptimizerApply(64.0); torch.cuda.synchronize(); t_apply = time.perf_counter() - t
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    fin = res[0]
    # K11 80 B/observation (pinhole K=4; 112 B with 12 intrinsics), K12 304 B/fully observed observation
    bpo = 80.0 if model == 4 else 112.0
    ach_j = bpo * (fin["residual_evaluations"] / world) / (fin["ms_jacobian_kernels"] * 1e-3) / 1e9 if fin["ms_jacobian_kernels"] else 0
    if rank != 0:
        if comm is not None:
            dist.barrier(); comm.close(); dist.destroy_process_group()
        return
    out = {"metric": "ImageRegistrator residual-evaluations/sec", "value": fin["evals_per_s"], "unit": "residual-evaluations/s", "n_gpus": world,
           "steps": a.steps, "warmup": a.warmup, "higher_is_better": True, "data": "synthetic", "dtype": "u8 images, f32 residuals/Jacobians, f64 accumulation",
           "scaling": "strong", "config": {"workload": "%d " % a.images + a.camera + " views %dx%d vs %.1fM-pt multi-resolution scan (%d scales), no occlusion geometry" % (a.width, a.height, npts / 1e6, len(sc["scales"])),
                      "image_scale_count": nsc, "scene_generation_s": t_gen},
           "per_scale": res, "lm_apply": {"applied": ap_[0], "tries": ap_[3], "seconds": t_apply},
           "roofline": {"bound": "hbm", "kernel": "kr_jacobians (K11, %d B/observation)" % int(bpo), "achieved": ach_j, "peak": peak, "unit": "GB/s", "frac": ach_j / peak,
                        "traffic": None}}
    if not a.no_cpu_baseline:
        from oracle import oracle as orc
        o = orc.Registration()
        sub = dict(sc); sub["images"] = sc["images"][:1]; sub["poses_init"] = sc["poses_init"][:1]; sub["poses_gt"] = sc["poses_gt"][:1]
        reg_scene.load_into(o, sub, splats=False)
        o.set_image_scale(0)
        t = time.perf_counter(); o.create_observations(1); t_obs = time.perf_counter() - t
        o.color_update()
        t = time.perf_counter(); o.accumulate(); t_acc = time.perf_counter() - t
        n_o = sum(len(o.observations(0, ps)[0]) for ps in range(len(sc["scales"])))
        out["cpu_baseline"] = {"value": n_o / t_acc, "unit": "residual-evaluations/s", "cores": 1, "kind": "port",
                               "sample": "oracle (serial, as the reference) on image 0 of the same workload at the finest image scale: %d observations, accumulate %.1f s, create_observations %.1f s" % (n_o, t_acc, t_obs)}
    print(json.dumps(out))
    if comm is not None:
        dist.barrier(); comm.close(); dist.destroy_process_group()


if __name__ == "__main__":
    main()
