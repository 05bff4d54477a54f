"""N>1 host-side logic on CPU (gloo, world_size 2): the pair-direction sharding of the C ABI planner covers every direction
exactly once, and summing the per-rank partial normal equations with an allreduce reproduces the single-rank system.
The per-rank partials are computed with the oracle's pieces (test infrastructure); the planner is the product's."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _partial_system(pairs, clouds_g, nv, owner_of, rank):
    """Normal equations from the pair-directions owned by `rank` (same algebra as k_finalize / impl.h:82-113 + :226)."""
    H = np.zeros((nv, nv)); b = np.zeros(nv); cost = 0.0
    for (s, t, q, m, _), owner in zip(pairs, owner_of):
        if owner != rank:
            continue
        ps, ns = clouds_g[s]; pt, nt = clouds_g[t]
        ps, ns, pt, nt = (a.astype(np.float64) for a in (ps[q], ns[q], pt[m], nt[m]))
        r1 = (ns * (pt - ps)).sum(1); r2 = (nt * (ps - pt)).sum(1)
        j1 = np.concatenate([ns, np.cross(pt, ns)], 1)          # d r1 / d target
        j2 = np.concatenate([nt, np.cross(ps, nt)], 1)          # d r2 / d source
        S = j1.T @ j1 + j2.T @ j2
        g = j1.T @ r1 - j2.T @ r2
        sv, tv = 6 * (s - 1), 6 * (t - 1)
        if sv >= 0:
            H[sv:sv + 6, sv:sv + 6] += S; b[sv:sv + 6] -= g
        if tv >= 0:
            H[tv:tv + 6, tv:tv + 6] += S; b[tv:tv + 6] += g
        if sv >= 0 and tv >= 0 and sv < tv:
            H[sv:sv + 6, tv:tv + 6] -= S; H[tv:tv + 6, sv:sv + 6] -= S
        cost += float((r1 * r1).sum() + (r2 * r2).sum())
    return H, b, cost


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from dataset_pipeline_b200 import synth
    from dataset_pipeline_b200.icp import plan_directions
    from oracle import oracle as orc
    clouds, poses, _ = synth.room_scans(3, 160, 64)
    o = orc.PointToPlaneICP(use_kdtree=True, inner_max_iterations=1)
    for (xyz, nrm), T in zip(clouds, poses):
        o.AddPointCloud(xyz, nrm, T)
    o.Run(0.08, 0, 1, 1e-10, False)
    pairs = o.pairs()
    plan = plan_directions(3, False, world)
    owner = {(s, t): w for s, t, w in plan}
    owner_of = [owner[(s, t)] for s, t, *_ in pairs]
    clouds_g = [orc.transform_cloud(xyz, nrm, T) for (xyz, nrm), T in zip(clouds, poses)]
    nv = 12
    H, b, cost = _partial_system(pairs, clouds_g, nv, owner_of, rank)
    buf = torch.from_numpy(np.concatenate([H.ravel(order="F"), b, [cost, sum(1 for w in owner_of if w == rank), 0.0]]))
    dist.all_reduce(buf, op=dist.ReduceOp.SUM)      # the ONE exchange of the data path: [H | b | cost | counters]
    if rank == 0:
        Ho, bo = o.normal_equations()
        tot = buf.numpy()
        ret["H_err"] = float(np.linalg.norm(tot[:nv * nv].reshape(nv, nv, order="F") - Ho) / np.linalg.norm(Ho))
        ret["b_err"] = float(np.linalg.norm(tot[nv * nv:nv * nv + nv] - bo) / np.linalg.norm(bo))
        ret["cost_err"] = abs(tot[nv * nv + nv] - o.stats()["first_cost"]) / o.stats()["first_cost"]
        ret["pairs"] = int(round(tot[nv * nv + nv + 1]))
        ret["expected_pairs"] = len(pairs)
    dist.destroy_process_group()


def test_direction_plan_partitions_all_directions():
    from dataset_pipeline_b200.icp import plan_directions
    for n, fixed in ((2, False), (8, False), (3, True), (1, True)):
        full = plan_directions(n, fixed, 1)
        assert len(full) == n * (n - 1) + (2 * n if fixed else 0)
        assert len(set((s, t) for s, t, _ in full)) == len(full)
        for world in (2, 4, 8):
            plan = plan_directions(n, fixed, world)
            assert [(s, t) for s, t, _ in plan] == [(s, t) for s, t, _ in full]      # same ik order on every rank
            owners = [w for *_, w in plan]
            assert all(0 <= w < world for w in owners)
            counts = np.bincount(owners, minlength=world)
            assert counts.max() - counts.min() <= 1                                  # balanced round-robin
    # 8 scans: the reference's 56 ordered pair-directions
    assert len(plan_directions(8, False, 8)) == 56


@pytest.mark.timeout(300)
def test_two_rank_allreduce_reproduces_single_rank_system(oracle):
    ctx = mp.get_context("spawn")
    mgr = ctx.Manager(); ret = mgr.dict()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, ret)) for r in range(2)]
    [p.start() for p in procs]
    [p.join(240) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    assert ret["pairs"] == ret["expected_pairs"]
    assert ret["H_err"] < 1e-5 and ret["b_err"] < 1e-5 and ret["cost_err"] < 1e-9


def _reg_rank(rank, world, port, ret):
    """One rank of the image-sharded Path B exchange on CPU: the oracle evaluates a problem holding only this rank's images (same
    points, same replicated state), the partial [H | b | sums] are summed with gloo, rank 0 compares with the full problem."""
    import torch
    import torch.distributed as dist
    from oracle import oracle as orc
    from dataset_pipeline_b200.synth import reg_scene
    from dataset_pipeline_b200._lib import lib
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    sc = reg_scene.make_scene(num_images=4, width=160, height=120, fx=130.0, num_scales=2, base_radius=0.008)
    area = 160 * 120 // 4
    owner = [lib().b2_reg_image_owner(i, world) for i in range(4)]

    def problem(images):
        r = orc.Registration(orc.reg_default_params(max_initial_image_area_in_pixels=area, variable_residuals_weight=0.0))
        w, h, K = sc["intr"]
        r.add_intrinsics(w, h, K)
        for i in images:
            r.add_image(0, sc["images"][i], None, sc["poses_init"][i])
        r.initialize()
        for xyz, radius, nbr, colors in sc["scales"]:
            r.add_point_scale(xyz, float(radius), nbr, colors)
        r.set_splat_points(sc["scales"][0][0])
        r.set_image_scale(0); r.create_observations(1)
        return r

    mine = [i for i in range(4) if owner[i] == rank]
    H, b, sums, _ = problem(mine).accumulate()
    # scatter the local system into the global variable layout [intrinsics(4) | 6 per image]
    nv = 4 + 6 * 4
    Hg = np.zeros((nv, nv)); bg = np.zeros(nv)
    idx = list(range(4)) + [4 + 6 * i + k for i in mine for k in range(6)]
    Hg[np.ix_(idx, idx)] = H; bg[idx] = b
    buf = torch.from_numpy(np.concatenate([Hg.ravel(), bg, sums[:4]]))
    dist.all_reduce(buf, op=dist.ReduceOp.SUM)        # the exchange of b2_reg_accumulate: [H | b | sums]
    if rank == 0:
        Hf, bf, sf, _ = problem(list(range(4))).accumulate()
        tot = buf.numpy()
        ret["H"] = float(np.abs(tot[:nv * nv].reshape(nv, nv) - Hf).max() / np.abs(Hf).max())
        ret["b"] = float(np.abs(tot[nv * nv:nv * nv + nv] - bf).max() / np.abs(bf).max())
        ret["sums"] = float(np.abs(tot[nv * nv + nv:] - sf[:4]).max() / np.abs(sf[:4]).max())
        ret["owners"] = owner
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_registration_exchange_reproduces_single_rank_system(oracle):
    """SURVEY §8e registration: images dealt round-robin, one sum-allreduce of [H | b | sums]. With fixed descriptors only (no
    colour update: its per-point means are a second allreduce, covered on the GPU by tests/dist_reg_check.py) the sharded sums
    equal the single-rank ones up to fp64 association."""
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    port = 29611
    ps = [ctx.Process(target=_reg_rank, args=(r, 2, port, ret)) for r in range(2)]
    for p in ps:
        p.start()
    for p in ps:
        p.join(240)
        assert p.exitcode == 0
    assert ret["owners"] == [0, 1, 0, 1]
    assert ret["H"] < 1e-12 and ret["b"] < 1e-12 and ret["sums"] < 1e-12, dict(ret)


# ---- sharded upload (cfg.shard_uploads): who reads its host buffers, who receives -------------------------------------------------
def _upload_rank(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from dataset_pipeline_b200.icp import upload_owner
    rng = np.random.default_rng(5)                      # every rank could read every cloud; only the owner does
    clouds = [rng.normal(size=(1000 + 17 * i, 3)).astype(np.float32) for i in range(5)]
    got, read = [], 0
    for i, c in enumerate(clouds):                      # the protocol of b2_icp_add_cloud: same order, same n on every rank
        owner = upload_owner(i, world)
        buf = torch.from_numpy(c.copy()) if owner == rank else torch.zeros(c.shape, dtype=torch.float32)
        read += c.nbytes if owner == rank else 0
        dist.broadcast(buf, src=owner)
        got.append(buf.numpy())
    ret[rank] = (all(np.array_equal(a, b) for a, b in zip(got, clouds)), read, sum(c.nbytes for c in clouds))
    dist.destroy_process_group()


def test_sharded_upload_protocol_delivers_every_cloud_once():
    from dataset_pipeline_b200.icp import upload_owner
    assert [upload_owner(i, 3) for i in range(7)] == [0, 1, 2, 0, 1, 2, 0] and upload_owner(2, 1) == 0 and upload_owner(-1, 2) == -1
    world = 2
    ret = mp.Manager().dict()
    mp.spawn(_upload_rank, args=(world, _free_port(), ret), nprocs=world, join=True)
    assert all(ret[r][0] for r in range(world))                                   # every rank ends up with every cloud
    assert sum(ret[r][1] for r in range(world)) == ret[0][2]                      # each byte crossed a host link exactly once
    assert max(ret[r][1] for r in range(world)) < 0.7 * ret[0][2]
