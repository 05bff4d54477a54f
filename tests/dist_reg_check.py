"""Launched by torchrun (one process per GPU): image-sharded Path B (b2_reg_set_comm), sharded normals and pair-direction-sharded ICP
against the single-GPU results on rank 0.
Prints DIST_REG_OK from rank 0 on success. Used by tests/test_gpu_reg_dist.py; also runnable by hand:
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/dist_reg_check.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("gloo")           # only used to ship the NCCL id; the data path uses the library's own communicator
    import dataset_pipeline_b200 as b2
    from dataset_pipeline_b200 import registration as R
    from dataset_pipeline_b200.icp import Comm
    from dataset_pipeline_b200.synth import reg_scene
    ids = [Comm.unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, 0)
    comm = Comm(rank, world, ids[0], device=local)

    # ---- sharded normals (b2_normals_estimate_dist): every rank gets the full result, identical to the single-GPU call ----
    from dataset_pipeline_b200.normals import estimate_normals, estimate_normals_dist
    rng = np.random.default_rng(4)
    pts = np.concatenate([rng.uniform(-1, 1, (60000, 3)) * [1, 1, 0.02], [[50.0, 50.0, 50.0]]]).astype(np.float32)   # + one isolated point
    nd, dense_d = estimate_normals_dist(pts, 12, (0.0, 0.0, 5.0), comm, device=local)
    ns = estimate_normals(pts, 12, (0.0, 0.0, 5.0))
    assert np.array_equal(np.nan_to_num(nd, nan=7.0), np.nan_to_num(ns, nan=7.0)), "sharded normals differ from the single-GPU result"

    # ---- pair-direction-sharded ICP (b2_icp_config.rank / world_size / comm): poses, LM try sequence and counts as on one GPU ----
    from dataset_pipeline_b200 import synth
    clouds, poses, _ = synth.room_scans(3, 400, 160)
    def run_icp(**kw):
        g = b2.PointToPlaneICP(device=local, **kw)
        for (xyz, nrm), T in zip(clouds, poses):
            g.AddPointCloud(xyz, nrm, T)
        tries = []
        for it in range(3):
            g.Run(0.05, it, 1, 1e-10, False)
            tries.append((g.stats()["inner_iterations"], g.stats()["lm_tries_total"], g.stats()["num_correspondences"]))
        return [g.GetResultGlobalTCloud(i).astype(np.float64) for i in range(3)], tries
    pd, td = run_icp(rank=rank, world_size=world, comm=comm)
    if rank == 0:
        ps, ts = run_icp()
        assert td == ts, "sharded ICP: LM sequence / correspondence counts differ: %s vs %s" % (td, ts)
        for a, c in zip(pd, ps):
            assert np.linalg.norm(a - c) / np.linalg.norm(c) <= 1e-6, "sharded ICP poses differ"
        print("dist_icp_check", world, "ranks: (inner iterations, LM tries, correspondences) per outer iteration", td, flush=True)

    model = int(os.environ.get("B2_TEST_CAMERA", "5"))
    sc = reg_scene.make_rig_scene(num_sets=3, camera_model=model)          # 6 images: ranks own 3 each; rig sets span both ranks
    area = 320 * 240 // 4

    def build(comm_or_none):
        g = b2.Registration(R.default_params(max_initial_image_area_in_pixels=area, device=local))
        if comm_or_none is not None:
            g.set_comm(comm_or_none)
        w, h, K = sc["intr"]
        g.add_intrinsics(w, h, K, camera_model=model)
        for i, (img, T) in enumerate(zip(sc["images"], sc["poses_init"])):
            g.add_image(0, img if g.owns(i) else None, None, T)
        rig = g.add_rig(sc["rig_init"])
        for s in sc["rig_sets"]:
            g.add_rig_images(rig, s)
        g.initialize()
        for xyz, radius, nbr, colors in sc["scales"]:
            g.add_point_scale(xyz, float(radius), nbr, colors)
        g.set_splat_points(sc["scales"][0][0])
        g.set_image_scale(0)
        return g

    g = build(comm)
    g.CreateObservationsForAllImages(1)
    for i in range(len(sc["images"])):
        n = sum(len(g.observations(i, ps)[0]) for ps in range(3))
        assert (n > 0) == g.owns(i), (rank, i, n)                             # observation sets live on the owner only
    g.ColorOptimizerApply()
    cost, sums = g.ComputeCost()
    H, b, s2, c2 = g.accumulate()
    it, opt_cost, conv = g.RunOnCurrentScale(3, 1e-9, 100)
    ip, po = g.get_state(); rigs = g.get_rigs()
    # every rank holds the same replicated state
    gathered = [None] * world
    dist.all_gather_object(gathered, (ip.tobytes(), po.tobytes(), rigs.tobytes(), cost, opt_cost))
    assert all(x == gathered[0] for x in gathered), "ranks diverged"
    ok = True
    if rank == 0:
        r = build(None)
        r.CreateObservationsForAllImages(1); r.ColorOptimizerApply()
        rc, rs = r.ComputeCost()
        rH, rb, _, _ = r.accumulate()
        rit, ropt, rconv = r.RunOnCurrentScale(3, 1e-9, 100)
        rip, rpo = r.get_state(); rrigs = r.get_rigs()
        rel = lambda a, c: float(np.linalg.norm(np.asarray(a, np.float64) - np.asarray(c, np.float64)) / np.linalg.norm(np.asarray(c, np.float64)))
        # counts exact; sums differ only by the fp32 association of the descriptor means (sum over ranks of per-rank sums)
        checks = {"counts": float(abs(sums[1] - rs[1]) + abs(sums[3] - rs[3])), "cost": abs(cost - rc) / rc, "H": float(np.abs(H - rH).max() / np.abs(rH).max()),
                  "b": float(np.abs(b - rb).max() / np.abs(rb).max()), "iters": float(abs(it - rit)), "poses": rel(po, rpo), "intr": rel(ip, rip),
                  "rigs": rel(rigs, rrigs), "opt_cost": abs(opt_cost - ropt) / ropt}
        limits = {"counts": 0.0, "cost": 1e-6, "H": 1e-5, "b": 1e-5, "iters": 0.0, "poses": 1e-5, "intr": 1e-5, "rigs": 1e-5, "opt_cost": 1e-5}
        ok = all(checks[k] <= limits[k] for k in checks)
        print("dist_reg_check", world, "ranks:", {k: float("%.3g" % v) for k, v in checks.items()}, flush=True)
        if ok:
            print("DIST_REG_OK", flush=True)
    dist.barrier()
    comm.close()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
