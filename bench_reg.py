#!/usr/bin/env python
"""Path B benchmark (BASELINE.json configs[3] / [4]): ImageRegistrator residual-evaluations/s.

Workload: N pinhole views (default 20 x 6000x4000, fx = fy = 4400) of the config-2 room — the scene the ICP
benchmark scans — rendered by ray casting with a multi-octave albedo; the scan = three of the 10M-point room scans merged (30M points,
coloured with the same albedo), turned into the multi-resolution point cloud by ComputeMultiResPointCloud on the device; the occlusion
geometry = the room tessellated at 2 cm (1.8M triangles: depth pass K8 + boundary masking K9 per image). Initial state = ground truth
perturbed by U(+-2 mm), U(+-0.05 deg), fx, fy +-0.1 %.

A step = one iteration of Optimizer::RunOnCurrentScale at the finest image scale from that state (optimizer.cc:49-182):
CreateObservationsForAllImages (depth maps, visibility, scale selection) + ColorOptimizer::Apply + IntrinsicsAndPoseOptimizer::Apply
(accumulate H, b + LM tries). A residual evaluation = one (image, point scale, observation) through pass 1+2 of
AccumulateHAndBAndResidualsForObservations (K11 + K12). value = residual evaluations per second of whole steps.

    python bench_reg.py [--images 20 --width 6000 --height 4000 --steps 3 --warmup 1]      # one JSON line
bench.py calls secondary_line() and puts the result under "secondary" of its own line (image-sharded under torchrun).
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

METRIC = "ImageRegistrator residual-evaluations/sec"
UNIT = "residual-evaluations/s"
FX = 4400.0 / 6000.0          # focal length as a fraction of the image width (fx = fy = 4400 at 6000 x 4000)


def build_scene(num_images, width, height, scan_wh, device, need_image):
    """-> dict(intr, params_init, images [uint8 HxW or None], poses_gt, poses_init, mesh (v, f), scans [(xyz global f32, rgb u8)])."""
    import torch
    import bench as B
    from dataset_pipeline_b200.synth import room_views as rv
    fx = FX * width
    K = np.array([fx, fx, (width - 1) / 2.0, (height - 1) / 2.0], np.float32)
    Rs, cs, gt, init = rv.view_poses(num_images)
    images = [rv.render_view(width, height, fx, fx, float(K[2]), float(K[3]), Rs[i], cs[i], device, seed=31 + i).cpu().numpy() if need_image(i) else None
              for i in range(num_images)]
    rng = np.random.default_rng(32)
    K_init = K.copy(); K_init[:2] *= (1.0 + rng.uniform(-1e-3, 1e-3, 2)).astype(np.float32)
    W, H = scan_wh
    B.ensure_scans(range(3), W, H)
    _, gts = B.scene_poses(8)
    scans = []
    for i in range(3):
        xyz, _ = B.load_scan(i, W, H)
        T = gts[i].astype(np.float64)
        g = (xyz.astype(np.float64) @ T[:3, :3].T + T[:3, 3]).astype(np.float32)
        scans.append((g, rv.point_colors(g, device)))
    torch.cuda.synchronize()
    return {"intr": (width, height, K), "K_init": K_init, "images": images, "poses_gt": gt, "poses_init": init, "mesh": rv.room_mesh(0.02), "scans": scans}


def load(reg, sc, use_images, parts=None):
    w, h, _ = sc["intr"]
    t0 = time.perf_counter()
    reg.add_intrinsics(w, h, sc["K_init"])
    for img, T in zip(use_images, sc["poses_init"]):
        reg.add_image(0, img, None, T)
    t1 = time.perf_counter()
    count = reg.initialize()
    t2 = time.perf_counter()
    reg.set_mesh(*sc["mesh"])
    if parts is not None:
        parts.update({"add_images_ms": 1e3 * (t1 - t0), "initialize_ms": 1e3 * (t2 - t1), "set_mesh_ms": 1e3 * (time.perf_counter() - t2)})
    return count


def secondary_line(world=1, rank=0, comm=None, local=0, num_images=20, width=6000, height=4000, scan_wh=(5000, 2000), steps=3, warmup=1,
                   cpu_baseline=True, e2e=True):
    """The Path B numbers as a dict (rank 0; other ranks take part and return None)."""
    import torch
    import dataset_pipeline_b200 as b2
    from dataset_pipeline_b200 import multiscale as MS
    from dataset_pipeline_b200 import registration as R
    dev = torch.device("cuda", local)
    owns = lambda i: i % world == rank
    t0 = time.perf_counter()
    # the multi-resolution point cloud needs every image (ComputeMinMaxPointRadius): with several ranks each one builds it on a
    # handle of its own that holds all images, then keeps only its share in the sharded handle (set-up, untimed)
    sc = build_scene(num_images, width, height, scan_wh, dev, (lambda i: True))
    t_scene = time.perf_counter() - t0
    t0 = time.perf_counter()
    setup = b2.Registration(R.default_params(device=local))
    count = load(setup, sc, sc["images"])
    radii, pts, cols, sidx, nbrs = MS.ComputeMultiResPointCloud(setup, sc["scans"], count)
    if world > 1:
        setup.close(); del setup
    t_multires = time.perf_counter() - t0
    npts = int(sum(len(p) for p in pts))

    def make():
        if world == 1 and make.first is not None:
            g, make.first = make.first, None
        else:
            t0 = time.perf_counter()
            g = b2.Registration(R.default_params(device=local))
            if comm is not None:
                g.set_comm(comm)
            make.parts = {"create_ms": 1e3 * (time.perf_counter() - t0)}
            load(g, sc, [img if owns(i) else None for i, img in enumerate(sc["images"])], make.parts)
        t0 = time.perf_counter()
        for r, p, c, nb in zip(radii, pts, cols, nbrs):
            g.add_point_scale(p, float(r), nb, c)
        make.parts["add_point_scales_ms"] = 1e3 * (time.perf_counter() - t0)
        return g
    make.parts = {}
    make.first = setup if world == 1 else None

    g = make()
    state0 = g.get_state()
    g.set_image_scale(0)

    def step(g):
        g.set_state(*state0)
        g.CreateObservationsForAllImages(1)
        ms_obs = g.stats()["ms_last_call"]; nobs = g.stats()["observations"]
        g.ColorOptimizerApply()
        ms_col = g.stats()["ms_last_call"]
        ap = g.IntrinsicsAndPoseOptimizerApply(64.0)
        st = g.stats()
        return {"observations": int(nobs), "ms_create_observations": ms_obs, "ms_color": ms_col, "ms_apply": st["ms_last_call"], "lm_tries": int(ap[3]),
                "applied": bool(ap[0])}

    def accumulate_only(g):
        g.accumulate()
        st = g.stats()
        return int(st["residual_evaluations"]), float(st["ms_jacobian_kernel"]), float(st["ms_accumulate_kernel"]), float(st["ms_last_call"])

    for _ in range(warmup):
        step(g)
    torch.cuda.synchronize(dev)
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
    profile = os.environ.get("B2_BENCH_PROFILE") == "1"      # ncu --profile-from-start off: only the timed steps are captured
    if profile:
        torch.cuda.profiler.start()
    t0 = time.perf_counter()
    parts = [step(g) for _ in range(steps)]
    torch.cuda.synchronize(dev)
    dt = (time.perf_counter() - t0) / steps
    if profile:
        torch.cuda.profiler.stop()
    # the accumulate pass alone (K11 + K12), at the state the last step left the observations in
    g.set_state(*state0); g.CreateObservationsForAllImages(1); g.ColorOptimizerApply()
    acc = [accumulate_only(g) for _ in range(max(2, steps))][1:]
    evals = acc[-1][0]; ms_j = float(np.mean([a[1] for a in acc])); ms_a = float(np.mean([a[2] for a in acc])); ms_acc_call = float(np.mean([a[3] for a in acc]))
    if world > 1:
        tt = torch.tensor([dt, ms_j, ms_a, ms_acc_call], dtype=torch.float64, device=dev); dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt, ms_j, ms_a, ms_acc_call = [float(v) for v in tt]
    evals_local = evals / world          # (the library's counters are already summed over the ranks; images are dealt round-robin)
    g.close()

    e2e_out = None
    if e2e:
        times = []
        for s in range(2):
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize(dev); t0 = time.perf_counter()
            ge = make(); ge.set_image_scale(0)
            t1 = time.perf_counter()
            st_e = step(ge); _ = ge.get_state()
            torch.cuda.synchronize(dev); d = time.perf_counter() - t0
            e2e_parts = dict(make.parts); e2e_parts["iteration_ms"] = 1e3 * (t0 + d - t1)
            e2e_parts.update({"iteration_device_" + k: st_e[k] for k in ("ms_create_observations", "ms_color", "ms_apply")})
            ge.close()
            e2e_parts["destroy_ms"] = 1e3 * (time.perf_counter() - t0 - d)
            if world > 1:
                tt = torch.tensor([d], dtype=torch.float64, device=dev); dist.all_reduce(tt, op=dist.ReduceOp.MAX); d = float(tt.item())
            times.append(d)
        h2d = sum(img.size for i, img in enumerate(sc["images"]) if owns(i)) + sum(p.nbytes + nb.nbytes + c.nbytes for p, nb, c in zip(pts, nbrs, cols)) \
            + sc["mesh"][0].nbytes + sc["mesh"][1].nbytes
        e2e_out = {"value": evals / times[-1], "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(8 * (124 * 124 + 124 + 8) * 8 + 7 * 4 * num_images),
                   "seconds": times, "last_run_breakdown": {k: round(v, 2) for k, v in e2e_parts.items()}, "note": "b2_reg_create + intrinsics + images (host u8) + initialize (pyramids) + mesh + point scales (host arrays) + one optimizer "
                                             "iteration + b2_reg_get_state + destroy; second of two runs"}
    if rank != 0:
        return None
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    ach_j = 80.0 * evals_local / (ms_j * 1e-3) / 1e9 if ms_j > 0 else 0.0          # K11 80 B / observation (pinhole, 4 intrinsics)
    ach_a = 304.0 * evals_local / (ms_a * 1e-3) / 1e9 if ms_a > 0 else 0.0         # K12 304 B / fully observed observation (upper bound: all counted)

    def traffic(name, algorithmic):
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", name)))
            return tj["dram_over_algorithmic"] * algorithmic
        except Exception:
            return None
    out = {"metric": METRIC, "value": evals / dt, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warmup, "ms_per_step": 1e3 * dt,
           "higher_is_better": True, "scaling": "strong", "data": "synthetic", "dtype": "u8 images, f32 residuals/Jacobians, f64 accumulation",
           "config": {"workload": "ImageRegistrator: %d pinhole views %dx%d of the config-2 room vs %.1fM-pt multi-resolution scan (%d scales, from 3 x %.0fM-pt scans) + "
                                  "occlusion mesh (%d triangles); 1 optimizer iteration at image scale 0 per step" % (
                                      num_images, width, height, npts / 1e6, len(pts), scan_wh[0] * scan_wh[1] / 1e6, len(sc["mesh"][1])),
                      "image_scale_count": count, "points_per_scale": [int(len(p)) for p in pts], "parallelism": "images sharded over %d GPU(s)" % world,
                      "l2": "inputs larger than L2 (%.2f GB of image pyramids, %.2f GB of points / neighbours / descriptors)" % (
                          num_images * width * height * 4 / 3 / 1e9, sum(p.nbytes + nb.nbytes / 2 + 2 * c.nbytes * 5 for p, nb, c in zip(pts, nbrs, cols)) / 1e9),
                      "scene_generation_s": t_scene, "multires_build_s": t_multires},
           "per_step": parts, "residual_evaluations_per_step": evals,
           "accumulate_only": {"ms": ms_acc_call, "evals_per_s": evals / (ms_acc_call * 1e-3) if ms_acc_call else None, "ms_jacobian_kernels": ms_j, "ms_accumulate_kernels": ms_a},
           "roofline": {"bound": "hbm", "kernel": "kr_jacobians (K11, 80 B/observation: xyz, observation, 8 u8 taps, I + 10 Jacobian floats)", "achieved": ach_j, "peak": peak,
                        "unit": "GB/s", "frac": ach_j / peak, "traffic": traffic("reg_jacobians_traffic.json", 80.0 * evals_local),
                        "algorithmic_bytes_per_launch_set": 80.0 * evals_local, "kernel_ms_per_accumulate": ms_j},
           "roofline_second_kernel": {"bound": "hbm", "kernel": "kr_residual_weights + kr_accumulate_weighted (K12, 304 B/observation)", "achieved": ach_a, "peak": peak,
                                      "unit": "GB/s", "frac": ach_a / peak, "traffic": traffic("reg_accumulate_traffic.json", 304.0 * evals_local),
                                      "kernel_ms_per_accumulate": ms_a}}
    if e2e_out:
        out["e2e"] = e2e_out
    if cpu_baseline:
        out["cpu_baseline"] = cpu_sample(sc, radii, pts, cols, nbrs, evals)
    return out


def cpu_sample(sc, radii, pts, cols, nbrs, evals_full, views=(0, 1)):
    """The oracle (serial, as the reference) on a bounded sample: one outward- and one inward-looking view of the same workload at full
    resolution with the same occlusion mesh and point cloud — depth maps, observations, colour update, accumulate. Scaled by evaluations."""
    from oracle import oracle as orc
    orc.build()
    o = orc.Registration(orc.reg_default_params())
    w, h, _ = sc["intr"]
    o.add_intrinsics(w, h, sc["K_init"])
    for i in views:
        o.add_image(0, sc["images"][i], None, sc["poses_init"][i])
    o.initialize()
    o.set_mesh(*sc["mesh"])
    for r, p, c, nb in zip(radii, pts, cols, nbrs):
        o.add_point_scale(p, float(r), nb, c)
    o.set_image_scale(0)
    t0 = time.perf_counter(); o.create_observations(1); t_obs = time.perf_counter() - t0
    t0 = time.perf_counter(); o.color_update(); t_col = time.perf_counter() - t0
    t0 = time.perf_counter(); o.accumulate(); t_acc = time.perf_counter() - t0
    n_o = int(sum(len(o.observations(i, ps)[0]) for i in range(len(views)) for ps in range(len(pts))))
    t_all = t_obs + t_col + t_acc
    return {"value": n_o / t_all, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": "oracle (serial, as the reference's Path B) on views %s of the same workload at full resolution, same mesh and point cloud: %d residual evaluations; "
                      "depth maps + observations %.1f s, colour update %.1f s, accumulate %.1f s (one LM try's cost evaluation not included: a lower bound of the "
                      "reference's time per iteration)" % (list(views), n_o, t_obs, t_col, t_acc)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--images", type=int, default=20)
    ap.add_argument("--width", type=int, default=6000)
    ap.add_argument("--height", type=int, default=4000)
    ap.add_argument("--scan-w", type=int, default=5000)
    ap.add_argument("--scan-h", type=int, default=2000)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    a = ap.parse_args()
    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench_reg.py: no CUDA device — no CPU fallback")
    rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    comm = None
    if world > 1:
        import datetime
        import torch.distributed as dist
        import dataset_pipeline_b200 as b2
        dist.init_process_group("nccl", device_id=torch.device("cuda", local), timeout=datetime.timedelta(seconds=600))
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt.copy_(torch.frombuffer(bytearray(b2.Comm.unique_id()), dtype=torch.uint8))
        dist.broadcast(idt, 0)
        comm = b2.Comm(rank, world, bytes(idt.cpu().numpy().tobytes()), device=local)
    out = secondary_line(world, rank, comm, local, a.images, a.width, a.height, (a.scan_w, a.scan_h), a.steps, a.warmup, not a.no_cpu_baseline and world == 1,
                         not a.no_e2e)
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.barrier(); dist.destroy_process_group()


if __name__ == "__main__":
    main()
