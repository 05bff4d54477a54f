"""Parity of the correspondence search where the benchmark actually runs ("What's weak" #1):
cells far above the inline threshold (chunk-box branches of scan_cell: > 64 and > 2048 points per cell), scans of 10^6
points with scanner-zenith clusters, and poses that are not the identity (the lookup runs in the target cloud's frame).
Bar: (query, match, d2) lists bit-identical to the oracle's kd-tree (FindCorrespondencesFast, icp_point_to_plane.cc:42-105).
Every test asserts through the kernel's work counters (B2_K3_WORK) that the branch it is about was really taken."""
import math
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def b2():
    import dataset_pipeline_b200 as b2
    return b2


def _rot(ax, ay, az):
    from dataset_pipeline_b200 import synth
    return synth.rot_xyz(ax, ay, az)


def _pose(R, t):
    T = np.eye(4, dtype=np.float64); T[:3, :3] = R; T[:3, 3] = t
    return T.astype(np.float32)


def _unit(rng, n):
    v = rng.normal(size=(n, 3)); v /= np.linalg.norm(v, axis=1, keepdims=True)
    return v.astype(np.float32)


def _search_both_ways(b2, oracle, clouds, poses, d, monkeypatch):
    """One outer iteration on the GPU with the work counters on; returns stats. Compares every pair-direction's list with the
    oracle's kd-tree search on the oracle-transformed (global-frame) clouds."""
    monkeypatch.setenv("B2_K3_WORK", "1")
    g = b2.PointToPlaneICP(keep_correspondences=True, inner_max_iterations=1)
    glob = []
    for (xyz, nrm), T in zip(clouds, poses):
        g.AddPointCloud(xyz, nrm, T)
        glob.append(oracle.transform_cloud(xyz, nrm, T)[0])
    g.Run(d, 0, 1, 1e-10, False)
    st = g.stats()
    seen = set()
    for s, t, q, m, d2 in g.pairs():
        qo, mo, do = oracle.find_correspondences(glob[s], glob[t], d, use_kdtree=True)
        assert np.array_equal(q, qo), "query indices differ for %d->%d" % (s, t)
        assert np.array_equal(m, mo), "match indices differ for %d->%d" % (s, t)
        assert np.array_equal(d2, do), "squared distances differ for %d->%d" % (s, t)
        seen.add((s, t))
    g.close()
    return st, seen


def _dense_target(rng):
    """Sparse background + a 70 000-point and a 3 000-point slab, each a few millimetres across (one or two grid cells)."""
    bg = rng.uniform(0, 1, (60000, 3))
    a = np.array([0.31, 0.52, 0.47]); b = np.array([0.72, 0.28, 0.61])
    ra = 0.006 * np.sqrt(rng.uniform(0, 1, 70000)); pa = rng.uniform(0, 2 * np.pi, 70000)
    slab_a = a + np.stack([ra * np.cos(pa), ra * np.sin(pa), rng.normal(0, 1e-4, 70000)], 1)
    rb = 0.006 * np.sqrt(rng.uniform(0, 1, 3000)); pb = rng.uniform(0, 2 * np.pi, 3000)
    slab_b = b + np.stack([rb * np.cos(pb), rng.normal(0, 1e-4, 3000), rb * np.sin(pb)], 1)
    tgt = np.concatenate([bg, slab_a, slab_b]).astype(np.float32)
    tgt = tgt[rng.permutation(len(tgt))]            # original indices carry no spatial order
    return tgt, a, b


def _dense_queries(rng, tgt, a, b):
    qs = [rng.uniform(0, 1, (40000, 3))]
    for c, ax in ((a, 2), (b, 1)):
        # range-noise outliers: 1-5 mm off the slab, spread over and beyond it
        q = c + rng.uniform(-0.012, 0.012, (6000, 3))
        q[:, ax] = c[ax] + rng.choice([-1, 1], 6000) * rng.uniform(0.001, 0.005, 6000)
        qs.append(q)
        qs.append(c + rng.normal(0, 2e-4, (3000, 3)))           # inside the slab
    src = np.concatenate(qs).astype(np.float32)
    src[:200] = tgt[rng.choice(len(tgt), 200, replace=False)]   # exact hits (d2 = 0)
    return src


def test_dense_cells_identity_pose(b2, oracle, monkeypatch):
    rng = np.random.default_rng(11)
    tgt, a, b = _dense_target(rng)
    tgt[-50:] = tgt[1000:1050]                                   # exact duplicates: lowest original index wins
    src = _dense_queries(rng, tgt, a, b)
    I = np.eye(4, dtype=np.float32)
    st, seen = _search_both_ways(b2, oracle, [(src, _unit(rng, len(src))), (tgt, _unit(rng, len(tgt)))], [I, I], 0.01, monkeypatch)
    assert (0, 1) in seen and (1, 0) in seen
    pts, box1, box2, cells, items = st["search_work"]
    assert box1 > 0, "no cell above the inline threshold was scanned"
    assert box2 > 0, "the > 2048-points-per-cell branch was never entered"
    assert items > 0


@pytest.mark.parametrize("case", ["rotated", "far", "tilted_far"])
def test_dense_cells_posed_clouds(b2, oracle, monkeypatch, case):
    """Same scene, clouds given in their own frames with non-trivial poses: the grid of each cloud is rotated against the other's
    and against the global axes; `far` puts the scene 900 m from the origin (rounding of the lookup >> rounding near the origin)."""
    rng = np.random.default_rng(12)
    tgt, a, b = _dense_target(rng)
    src = _dense_queries(rng, tgt, a, b)
    Rs, Rt = _rot(0.3, -0.2, 1.1), _rot(-0.7, 0.4, 0.35)
    ts, tt = np.array([0.5, -0.25, 0.125]), np.array([-1.0, 2.0, 0.5])
    if case == "far":
        Rs, Rt = np.eye(3), _rot(0, 0, 0.5)
    if case in ("far", "tilted_far"):
        ts = ts + np.array([900.0, -450.0, 30.0]); tt = tt + np.array([900.0, -450.0, 30.0])
    # cloud-frame coordinates such that pose * local ~ the common scene (float64, then rounded once)
    src_l = ((src.astype(np.float64) + (np.array([900.0, -450.0, 30.0]) if case != "rotated" else 0) - ts) @ Rs).astype(np.float32)
    tgt_l = ((tgt.astype(np.float64) + (np.array([900.0, -450.0, 30.0]) if case != "rotated" else 0) - tt) @ Rt).astype(np.float32)
    st, seen = _search_both_ways(b2, oracle, [(src_l, _unit(rng, len(src))), (tgt_l, _unit(rng, len(tgt)))], [_pose(Rs, ts), _pose(Rt, tt)],
                                 0.01, monkeypatch)
    assert (0, 1) in seen and (1, 0) in seen
    assert st["search_work"][1] > 0 and st["search_work"][2] > 0


def test_million_point_room_scans(b2, oracle, monkeypatch):
    """Two scans of the config-2 room at 1000 x 1000 rays (~10^6 points each, scanner-zenith / nadir clusters of thousands of points per
    2 cm cell), d = 0.01 as in BASELINE config 2, perturbed poses: every list bit-exact."""
    from dataset_pipeline_b200 import synth
    clouds, poses, _ = synth.room_scans(2, 1000, 1000)
    assert min(len(c[0]) for c in clouds) > 900000
    st, seen = _search_both_ways(b2, oracle, clouds, poses, 0.01, monkeypatch)
    assert seen == {(0, 1), (1, 0)}
    pts, box1, box2, cells, items = st["search_work"]
    assert box1 > 0, "no zenith cluster reached the chunk-box branches"      # (> 2048 per cell is pinned by the dense-cell tests above)
    assert st["num_correspondences"] > 500000


def test_three_scans_half_million_with_fixed_cloud(b2, oracle, monkeypatch):
    """A fixed cloud (concatenated in the global frame, identity pose) against two movable ones at 700 x 700 rays."""
    from dataset_pipeline_b200 import synth
    clouds, poses, _ = synth.room_scans(3, 700, 700)
    monkeypatch.setenv("B2_K3_WORK", "1")
    g = b2.PointToPlaneICP(keep_correspondences=True, inner_max_iterations=1)
    o = oracle.PointToPlaneICP(use_kdtree=True, inner_max_iterations=1)
    for k, ((xyz, nrm), T) in enumerate(zip(clouds, poses)):
        g.AddPointCloud(xyz, nrm, T, k == 0); o.AddPointCloud(xyz, nrm, T, k == 0)
    g.Run(0.01, 0, 1, 1e-10, False); o.Run(0.01, 0, 1, 1e-10, False)
    pg, po = g.pairs(), o.pairs()
    assert [(s, t) for s, t, *_ in pg] == [(s, t) for s, t, *_ in po]
    for (s, t, q1, m1, d1), (_, _, q2, m2, d2) in zip(pg, po):
        assert np.array_equal(q1, q2) and np.array_equal(m1, m2) and np.array_equal(d1, d2), "pair %d->%d" % (s, t)
    assert g.stats()["search_work"][1] > 0


def test_similarity_pose_is_searched_exactly(b2, oracle, monkeypatch):
    """A pose with scale 1.25 and a slight shear (Eigen::Affine3f admits it): cloud-frame distances are no longer the global ones;
    sigma = ||A^-1|| enlarges the cells instead of losing matches."""
    rng = np.random.default_rng(5)
    tgt = rng.uniform(-1, 1, (50000, 3)).astype(np.float32)
    src = rng.uniform(-1, 1, (30000, 3)).astype(np.float32)
    A = 1.25 * _rot(0.2, 0.1, -0.4); A[0, 1] += 0.02
    T = np.eye(4, dtype=np.float32); T[:3, :3] = A; T[:3, 3] = [0.1, 0.2, -0.3]
    st, seen = _search_both_ways(b2, oracle, [(src, _unit(rng, len(src))), (tgt, _unit(rng, len(tgt)))], [np.eye(4, dtype=np.float32), T], 0.05,
                                 monkeypatch)
    assert (0, 1) in seen and (1, 0) in seen


def test_index_survives_pose_updates_and_radius_change(b2, oracle, monkeypatch):
    """The static index is reused across outer iterations (poses change) and rebuilt when the radius changes; every iteration is
    compared on identical poses."""
    from dataset_pipeline_b200 import synth
    clouds, poses, _ = synth.room_scans(3, 300, 120)
    g = b2.PointToPlaneICP(keep_correspondences=True)
    o = oracle.PointToPlaneICP(use_kdtree=True)
    for (xyz, nrm), T in zip(clouds, poses):
        g.AddPointCloud(xyz, nrm, T); o.AddPointCloud(xyz, nrm, T)
    builds = []
    for it, d in enumerate([0.05, 0.05, 0.05, 0.02, 0.02, 0.08]):
        for i in range(3):
            g.SetGlobalTCloud(i, o.GetResultGlobalTCloud(i))
        g.Run(d, it, 1, 1e-10, False); o.Run(d, it, 1, 1e-10, False)
        builds.append(g.stats()["ms_index_build"] > 0)
        for (s, t, q1, m1, d1), (_, _, q2, m2, d2) in zip(g.pairs(), o.pairs()):
            assert np.array_equal(q1, q2) and np.array_equal(m1, m2) and np.array_equal(d1, d2), "iteration %d pair %d->%d" % (it, s, t)
    assert builds == [True, False, False, True, False, True]


def test_sparse_grid_uses_the_hash_layout(b2, oracle, monkeypatch):
    """A 30 m cube at d = 0.01 is 3.4 * 10^9 grid cells: above the rank-bitmap limit, so the occupied cells are hashed. Same bar."""
    rng = np.random.default_rng(21)
    tgt = rng.uniform(0, 30, (200000, 3)).astype(np.float32)
    tgt[:20000] = (np.array([11.0, 7.0, 23.0]) + rng.normal(0, 0.05, (20000, 3))).astype(np.float32)      # a dense blob too
    src = np.concatenate([tgt[rng.choice(len(tgt), 100000, replace=False)] + rng.normal(0, 0.003, (100000, 3)),
                          rng.uniform(0, 30, (50000, 3))]).astype(np.float32)
    Rs, Rt = _rot(0.1, 0.2, -0.3), _rot(-0.2, 0.05, 0.4)
    ts, tt = np.array([1.0, 2.0, 3.0]), np.array([-2.0, 0.5, 1.5])
    src_l = ((src.astype(np.float64) - ts) @ Rs).astype(np.float32)
    tgt_l = ((tgt.astype(np.float64) - tt) @ Rt).astype(np.float32)
    st, seen = _search_both_ways(b2, oracle, [(src_l, _unit(rng, len(src))), (tgt_l, _unit(rng, len(tgt)))], [_pose(Rs, ts), _pose(Rt, tt)],
                                 0.01, monkeypatch)
    assert (0, 1) in seen and (1, 0) in seen
    assert st["sparse_grids"] == 2
    assert st["num_correspondences"] > 50000


def _run_handle(b2, clouds, poses, d, iterations=2, move=None, **kw):
    g = b2.PointToPlaneICP(keep_correspondences=True, **kw)
    for (xyz, nrm), T in zip(clouds, poses):
        g.AddPointCloud(xyz, nrm, T)
    if move is not None:
        g.SetGlobalTCloud(*move)
    out = []
    for it in range(iterations):
        g.Run(d, it, 1, 1e-10, False)
        st = g.stats()
        out.append((st, g.pairs(), [g.GetResultGlobalTCloud(i) for i in range(len(clouds))], g.tries()))
    g.close()
    return out


def _same_iterations(a, b, exact=True):
    """exact: everything bit for bit (same index grids, hence the same record order and the same fp64 summation order).
    not exact: the lists bit for bit, the fp64 sums to rounding (the grids — and with them the order of the records — may differ)."""
    for it, ((sa, pa, Ta, ta), (sb, pb, Tb, tb)) in enumerate(zip(a, b)):
        assert [(s, t) for s, t, *_ in pa] == [(s, t) for s, t, *_ in pb]
        if exact or it == 0:
            for (s, t, q1, m1, d1), (_, _, q2, m2, d2) in zip(pa, pb):
                assert np.array_equal(q1, q2) and np.array_equal(m1, m2) and np.array_equal(d1, d2), "iteration %d pair %d->%d" % (it, s, t)
        for k in ("num_correspondences", "num_pairs", "inner_iterations", "lm_tries_total"):
            if exact or it == 0:
                assert sa[k] == sb[k], (it, k, sa[k], sb[k])
        for k in ("first_cost", "last_cost"):
            if exact:
                assert sa[k] == sb[k], (it, k, sa[k], sb[k])
            elif it == 0:
                assert abs(sa[k] - sb[k]) <= 1e-11 * abs(sa[k]), (it, k, sa[k], sb[k])
        if exact:
            assert list(ta) == list(tb)
            for A, B in zip(Ta, Tb):
                assert np.array_equal(A, B)


def test_searches_done_behind_the_uploads_are_the_same_searches(b2, oracle):
    """search_ahead (index_distance_hint given): the pair-directions among the clouds indexed while later clouds upload are searched at
    add time and adopted by the first Run. Lists, costs, LM decisions and poses are bit-identical to a handle that searches inside Run,
    and the adoption is dropped — not trusted — when a pose or the radius is not what it was searched with."""
    from dataset_pipeline_b200 import synth
    clouds, poses, _ = synth.room_scans(4, 400, 300)
    d = 0.02
    plain = _run_handle(b2, clouds, poses, d)
    assert plain[0][0]["searches_ahead"] == 0 and plain[0][0]["search_launches"] == 12
    # same handle configuration (hint: the indexes are built at AddPointCloud), searching inside Run: the bit-for-bit baseline
    off = _run_handle(b2, clouds, poses, d, index_distance_hint=d, search_ahead=False)
    assert off[0][0]["searches_ahead"] == 0 and off[0][0]["search_launches"] == 12
    ahead = _run_handle(b2, clouds, poses, d, index_distance_hint=d)
    # clouds 0..2 are indexed behind the uploads of clouds 1..3; cloud 3 is indexed by Run: 6 of the 12 directions were done ahead
    assert ahead[0][0]["searches_ahead"] == 6 and ahead[0][0]["search_launches"] == 6
    assert ahead[1][0]["searches_ahead"] == 0 and ahead[1][0]["search_launches"] == 12
    _same_iterations(off, ahead)
    # without a hint the index grids may be laid out differently (they are sized from all clouds' magnitudes instead of each cloud's
    # own), which reorders the records of a set: same lists, sums equal to rounding
    _same_iterations(plain, ahead, exact=False)
    # against the oracle too (first iteration's lists)
    o = oracle.PointToPlaneICP(use_kdtree=True)
    for (xyz, nrm), T in zip(clouds, poses):
        o.AddPointCloud(xyz, nrm, T)
    o.Run(d, 0, 1, 1e-10, False)
    for (s, t, q1, m1, d1), (_, _, q2, m2, d2) in zip(ahead[0][1], o.pairs()):
        assert np.array_equal(q1, q2) and np.array_equal(m1, m2) and np.array_equal(d1, d2), "pair %d->%d" % (s, t)

    # a pose changed between AddPointCloud and Run: the directions of that cloud are searched again, the others are adopted
    T1 = poses[1].copy(); T1[:3, 3] += np.array([0.004, -0.003, 0.002], np.float32)
    moved_off = _run_handle(b2, clouds, poses, d, move=(1, T1), index_distance_hint=d, search_ahead=False)
    moved_ahead = _run_handle(b2, clouds, poses, d, move=(1, T1), index_distance_hint=d)
    assert moved_ahead[0][0]["searches_ahead"] == 2 and moved_ahead[0][0]["search_launches"] == 10
    _same_iterations(moved_off, moved_ahead)

    # Run called with another radius than the hint: nothing is adopted (and the indexes are rebuilt for the new radius)
    other_off = _run_handle(b2, clouds, poses, 0.012, index_distance_hint=d, search_ahead=False)
    other_ahead = _run_handle(b2, clouds, poses, 0.012, index_distance_hint=d)
    assert other_ahead[0][0]["searches_ahead"] == 0 and other_ahead[0][0]["search_launches"] == 12
    _same_iterations(other_off, other_ahead)


def test_sets_packed_behind_the_searches_and_the_overflow_fallback(b2, oracle, monkeypatch):
    """K4 overlapped with K3 (optional, B2_PACK=overlap): sets are packed on their own stream at device-side running offsets while later sets are searched. Same
    records in the same places as the serial pack (costs, LM decisions, poses bit for bit, lists against the oracle), and when a
    set does not fit the record arrays of the previous iteration the whole iteration is packed the serial way."""
    from dataset_pipeline_b200 import synth
    clouds, poses, _ = synth.room_scans(3, 400, 300)
    radii = [0.01, 0.01, 0.05, 0.05]

    def run(mode):
        if mode is None:
            monkeypatch.delenv("B2_PACK", raising=False)
        else:
            monkeypatch.setenv("B2_PACK", mode)
        g = b2.PointToPlaneICP(keep_correspondences=True)
        for (xyz, nrm), T in zip(clouds, poses):
            g.AddPointCloud(xyz, nrm, T)
        out = []
        for it, d in enumerate(radii):
            for i in range(3):
                g.SetGlobalTCloud(i, poses[i])          # every iteration from the same poses: the same sets every time
            g.Run(d, it, 1, 1e-10, False)
            out.append((g.stats(), g.pairs(), [g.GetResultGlobalTCloud(i) for i in range(3)], g.tries()))
        g.close()
        return out

    serial, overlapped, nosize = run(None), run("overlap"), run("overlap_nosize")
    assert [o[0]["packs_overlapped"] for o in serial] == [0, 0, 0, 0]
    assert [o[0]["packs_overlapped"] for o in overlapped] == [6, 6, 6, 6]       # record arrays sized for the worst case up front
    # without the up-front sizing: first iteration serial (no arrays yet), second overlapped, third overflows (five times the
    # radius, several times the matches) and falls back, fourth overlapped again
    assert [o[0]["packs_overlapped"] for o in nosize] == [0, 6, 0, 6]
    assert nosize[2][0]["num_correspondences"] > 1.2 * nosize[1][0]["num_correspondences"]
    _same_iterations(serial, overlapped)
    _same_iterations(serial, nosize)
    o = oracle.PointToPlaneICP(use_kdtree=True)
    for (xyz, nrm), T in zip(clouds, poses):
        o.AddPointCloud(xyz, nrm, T)
    o.Run(radii[0], 0, 1, 1e-10, False)
    for (s, t, q1, m1, d1), (_, _, q2, m2, d2) in zip(overlapped[0][1], o.pairs()):
        assert np.array_equal(q1, q2) and np.array_equal(m1, m2) and np.array_equal(d1, d2), "pair %d->%d" % (s, t)
