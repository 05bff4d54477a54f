"""Parity tests proper: the CUDA path, called through the C ABI, against the oracle on the same seeded inputs.
Bar: correspondence lists bit-exact; normal equations and poses within 1e-5 relative Frobenius."""
import math
import os

import numpy as np
import pytest

from tests.golden import ref_test_inputs as rti

pytestmark = pytest.mark.gpu

TOL = 1e-5


def rel_fro(a, b):
    return float(np.linalg.norm(np.asarray(a, np.float64) - np.asarray(b, np.float64)) / max(np.linalg.norm(np.asarray(b, np.float64)), 1e-300))


@pytest.fixture(scope="module")
def b2():
    import dataset_pipeline_b200 as b2
    return b2


@pytest.fixture(scope="module")
def room3():
    from dataset_pipeline_b200 import synth
    return synth.room_scans(3, 300, 120)      # 3 x 36k points


def test_find_correspondences_bit_exact(b2, oracle):
    rng = np.random.default_rng(3)
    tgt = rng.uniform(-1, 1, (40000, 3)).astype(np.float32)
    src = rng.uniform(-1.05, 1.05, (30000, 3)).astype(np.float32)
    tgt[100:110] = tgt[50:60]            # exact duplicates: lowest-index tie-break
    src[:10] = tgt[50:60]
    for d in (0.01, 0.03, 0.1):
        qa, ma, da = b2.find_correspondences(src, tgt, d)
        qb, mb, db = oracle.find_correspondences(src, tgt, d, use_kdtree=True)
        assert np.array_equal(qa, qb)
        assert np.array_equal(ma, mb)
        assert np.array_equal(da, db)
    assert set(ma[:10].tolist()) <= set(range(50, 60))


def test_find_correspondences_edge_cases(b2, oracle):
    e = np.zeros((0, 3), np.float32)
    p = np.random.default_rng(0).uniform(0, 1, (100, 3)).astype(np.float32)
    assert len(b2.find_correspondences(e, p, 0.1)[0]) == 0
    assert len(b2.find_correspondences(p, e, 0.1)[0]) == 0
    assert len(b2.find_correspondences(p, p + 10.0, 0.1)[0]) == 0
    # boxes that do not intersect but lie within d of each other still match (no bbox gate in FindCorrespondencesFast)
    a = np.array([[0, 0, 0]], np.float32); b = np.array([[0.05, 0, 0]], np.float32)
    assert len(b2.find_correspondences(a, b, 0.1)[0]) == 1
    # strict radius
    t = np.array([[0.5, 0, 0]], np.float32)
    assert len(b2.find_correspondences(a, t, 0.5)[0]) == 0
    assert len(b2.find_correspondences(a, t, float(np.nextafter(np.float32(0.5), np.float32(1))))[0]) == 1
    # far-from-origin coordinates (large cell indices)
    big = (p * 50 + 4000).astype(np.float32)
    qa, ma, da = b2.find_correspondences(big, big[::-1].copy(), 0.5)
    qb, mb, db = oracle.find_correspondences(big, big[::-1].copy(), 0.5)
    assert np.array_equal(qa, qb) and np.array_equal(ma, mb) and np.array_equal(da, db)


def _setup_pair(b2, oracle, clouds, poses, fixed_first=False, inner=150):
    g = b2.PointToPlaneICP(keep_correspondences=True, inner_max_iterations=inner)
    o = oracle.PointToPlaneICP(use_kdtree=True, inner_max_iterations=inner)
    ids = []
    for k, ((xyz, nrm), T) in enumerate(zip(clouds, poses)):
        fx = fixed_first and k == 0
        ig = g.AddPointCloud(xyz, nrm, T, fx); io = o.AddPointCloud(xyz, nrm, T, fx)
        assert ig == io
        if not fx:
            ids.append(ig)
    return g, o, ids


def _check_iteration(g, o, ids):
    pg, po = g.pairs(), o.pairs()
    assert [(s, t) for s, t, *_ in pg] == [(s, t) for s, t, *_ in po]
    for (s, t, q1, m1, d1), (_, _, q2, m2, d2) in zip(pg, po):
        assert np.array_equal(q1, q2), "query indices differ for pair %d->%d" % (s, t)
        assert np.array_equal(m1, m2), "match indices differ for pair %d->%d" % (s, t)
        assert np.array_equal(d1, d2), "squared distances differ for pair %d->%d" % (s, t)
    Hg, bg, cg = g.normal_equations()
    Ho, bo = o.normal_equations()
    so = o.stats(); sg = g.stats()
    assert sg["num_correspondences"] == so["num_correspondences"]
    assert sg["num_pairs"] == so["num_pairs"]
    if Ho.size:
        assert rel_fro(Hg, Ho) <= TOL
        assert rel_fro(bg, bo) <= TOL
    assert abs(cg - so["first_cost"]) <= 1e-9 * max(1.0, abs(so["first_cost"]))
    for i in ids:
        assert rel_fro(g.GetResultGlobalTCloud(i), o.GetResultGlobalTCloud(i)) <= TOL
    return sg, so


def test_room_one_outer_iteration(b2, oracle, room3):
    clouds, poses, _ = room3
    g, o, ids = _setup_pair(b2, oracle, clouds, poses)
    g.Run(0.05, 0, 1, 1e-10, False); o.Run(0.05, 0, 1, 1e-10, False)
    sg, so = _check_iteration(g, o, ids)
    assert sg["num_correspondences"] > 20000
    # same LM accept/reject sequence on this scene
    assert np.array_equal(g.tries(), o.tries())
    assert abs(sg["last_cost"] - so["last_cost"]) <= 1e-6 * so["last_cost"]


def test_room_trajectory_with_pose_resync(b2, oracle, room3):
    """10 outer iterations; before each one both sides are given the oracle's poses, so every iteration is compared
    on identical inputs (correspondences bit-exact, pose update within 1e-5)."""
    clouds, poses, _ = room3
    g, o, ids = _setup_pair(b2, oracle, clouds, poses)
    for it in range(10):
        for i in ids:
            g.SetGlobalTCloud(i, o.GetResultGlobalTCloud(i))
        g.Run(0.05, it, 1, 1e-10, False); o.Run(0.05, it, 1, 1e-10, False)
        _check_iteration(g, o, ids)


def test_room_free_running_converges_like_oracle(b2, oracle, room3):
    clouds, poses, gts = room3
    g, o, ids = _setup_pair(b2, oracle, clouds, poses)
    g.Run(0.05, 0, 15, 1e-10, False); o.Run(0.05, 0, 15, 1e-10, False)
    for i in ids:
        assert rel_fro(g.GetResultGlobalTCloud(i), o.GetResultGlobalTCloud(i)) <= 1e-4


def test_fixed_cloud_both_directions(b2, oracle, room3):
    clouds, poses, _ = room3
    g, o, ids = _setup_pair(b2, oracle, clouds, poses, fixed_first=True)
    g.Run(0.05, 0, 1, 1e-10, False); o.Run(0.05, 0, 1, 1e-10, False)
    sg, _ = _check_iteration(g, o, ids)
    assert sg["num_variables"] == 12
    assert (1, 0) in [(s, t) for s, t, *_ in g.pairs()] and (0, 1) in [(s, t) for s, t, *_ in g.pairs()]


def test_ref_identical_cloud_alignment_gpu(b2, oracle):
    """The reference's own test (test_icp.cc:39-109) through the C ABI."""
    pts, nrm, poses = rti.identical_cloud_alignment_inputs()
    icp = b2.PointToPlaneICP()
    ids = [icp.AddPointCloud(pts, nrm, T, False) for T in poses]
    assert ids == list(range(20))
    icp.Run(float(np.float32(0.15) * np.float32(math.sqrt(3))), 0, 100, 1e-7, False)
    T0 = icp.GetResultGlobalTCloud(ids[0])
    for i in ids[1:]:
        assert np.abs(T0[:3, :] - icp.GetResultGlobalTCloud(i)[:3, :]).max() <= 1e-5


def test_ref_plane_with_single_point_gpu(b2):
    """The reference's own test (test_icp.cc:111-172) through the C ABI: rank-deficient planar case."""
    pts, nrm, poses = rti.plane_with_single_point_inputs()
    icp = b2.PointToPlaneICP()
    a = icp.AddPointCloud(pts, nrm, poses[0], False); b = icp.AddPointCloud(pts, nrm, poses[1], False)
    icp.Run(1.5, 0, 100, 1e-7, False)
    assert np.abs(icp.GetResultGlobalTCloud(a)[:3, :] - icp.GetResultGlobalTCloud(b)[:3, :]).max() <= 1e-5


def test_config1_relief_scans(b2, oracle):
    """BASELINE config 1: 2 x 50k scans, -d 0.01, 10 iterations, pose-resynced per iteration."""
    from dataset_pipeline_b200 import synth
    clouds, poses = synth.relief_scans()
    g, o, ids = _setup_pair(b2, oracle, clouds, poses)
    for it in range(10):
        for i in ids:
            g.SetGlobalTCloud(i, o.GetResultGlobalTCloud(i))
        g.Run(0.01, it, 1, 1e-10, False); o.Run(0.01, it, 1, 1e-10, False)
        _check_iteration(g, o, ids)


def test_pointnormal_strided_input(b2, oracle, room3):
    clouds, poses, _ = room3
    xyz, nrm = clouds[0]
    pn = np.zeros((len(xyz), 12), np.float32)
    pn[:, 0:3] = xyz; pn[:, 4:7] = nrm
    g1 = b2.PointToPlaneICP(); g2 = b2.PointToPlaneICP()
    for g in (g1, g2):
        if g is g1:
            g.AddPointNormalArray(pn, poses[0]); g.AddPointCloud(clouds[1][0], clouds[1][1], poses[1])
        else:
            g.AddPointCloud(xyz, nrm, poses[0]); g.AddPointCloud(clouds[1][0], clouds[1][1], poses[1])
        g.Run(0.05, 0, 1, 1e-10, False)
    assert np.array_equal(g1.GetResultGlobalTCloud(1), g2.GetResultGlobalTCloud(1))


def test_error_paths(b2):
    from dataset_pipeline_b200._lib import B2Error
    icp = b2.PointToPlaneICP()
    with pytest.raises(B2Error):
        icp.Run(0.1, 0, 1, 1e-6, False)          # reference: CHECK(!clouds_.empty())
    with pytest.raises(B2Error):
        icp.GetResultGlobalTCloud(3)
    # single movable cloud, no fixed: zero variables, nothing moves, converges immediately
    p = np.random.default_rng(0).uniform(0, 1, (100, 3)).astype(np.float32)
    icp.AddPointCloud(p, p, np.eye(4, dtype=np.float32))
    assert icp.Run(0.1, 0, 3, 1e-6, False) is True
    assert np.array_equal(icp.GetResultGlobalTCloud(0), np.eye(4, dtype=np.float32))


def test_deterministic_repeat(b2, room3):
    clouds, poses, _ = room3
    res = []
    for _ in range(2):
        g = b2.PointToPlaneICP()
        for (xyz, nrm), T in zip(clouds, poses):
            g.AddPointCloud(xyz, nrm, T)
        g.Run(0.05, 0, 3, 1e-10, False)
        res.append([g.GetResultGlobalTCloud(i) for i in range(3)] + [g.stats()["last_cost"]])
    for a, b in zip(res[0], res[1]):
        assert np.array_equal(a, b)


def _tries_equal_one_iteration(b2, oracle, clouds, poses, d, fixed_first=False):
    g, o, ids = _setup_pair(b2, oracle, clouds, poses, fixed_first=fixed_first)
    g.Run(d, 0, 1, 1e-10, False); o.Run(d, 0, 1, 1e-10, False)
    assert np.array_equal(g.tries(), o.tries()), (g.tries(), o.tries())
    sg, so = g.stats(), o.stats()
    assert sg["inner_iterations"] == so["inner_iterations"] and sg["lm_tries_total"] == so["lm_tries_total"]
    return sg


def test_lm_try_sequences_on_several_scenes(b2, oracle, room3):
    """K5 builds H from one symmetric S per correspondence set (the source-pose Jacobian is the negated target-pose one) where the
    reference evaluates both forms in fp32; the accept / reject sequence of the LM loop must nevertheless be the reference's. Checked on
    scenes of different conditioning: three room scans, a fixed cloud, the config-1 relief pair, the reference's rank-deficient plane
    (test_icp.cc:111-172) and its 20 identical clouds (:39-109, 114 variables)."""
    from dataset_pipeline_b200 import synth
    clouds, poses, _ = room3
    assert _tries_equal_one_iteration(b2, oracle, clouds, poses, 0.05)["inner_iterations"] > 3
    _tries_equal_one_iteration(b2, oracle, clouds, poses, 0.05, fixed_first=True)
    rc, rp = synth.relief_scans()
    _tries_equal_one_iteration(b2, oracle, rc, rp, 0.01)
    pts, nrm, pp = rti.plane_with_single_point_inputs()
    _tries_equal_one_iteration(b2, oracle, [(pts, nrm), (pts, nrm)], pp, 1.5)
    pts, nrm, pp = rti.identical_cloud_alignment_inputs()
    _tries_equal_one_iteration(b2, oracle, [(pts, nrm)] * 6, pp[:6], float(np.float32(0.15) * np.float32(math.sqrt(3))))


def test_packed_fp32_cost_path_returns_the_bits_of_the_scalar_one():
    """K5 evaluates the LM tries' costs with packed fp32 (FFMA2, two IEEE-rn lanes per instruction) on records staged by the bulk-copy
    engine; B2_K5=ldg runs the register-staged kernel with the scalar arithmetic and one try per pass. Costs, accept / reject sequence
    and poses must agree bit for bit over a whole alignment (each process reads the switch once, hence the subprocesses)."""
    import json
    import subprocess
    import sys
    code = r'''
import json, sys, numpy as np
sys.path.insert(0, %r)
import dataset_pipeline_b200 as b2
from dataset_pipeline_b200 import synth
clouds, poses, _ = synth.room_scans(4, 300, 200)
g = b2.PointToPlaneICP()
for (xyz, nrm), T in zip(clouds, poses):
    g.AddPointCloud(xyz, nrm, T)
out = []
for it in range(6):
    g.Run(0.03, it, 1, 1e-10, False)
    st = g.stats()
    out.append({"first": float(st["first_cost"]).hex(), "last": float(st["last_cost"]).hex(), "tries": [int(v) for v in g.tries()],
                "lambda": float(st["final_lambda"]).hex(), "n": int(st["num_correspondences"]),
                "poses": [g.GetResultGlobalTCloud(i).astype(np.float32).tobytes().hex() for i in range(4)]})
print("RESULT " + json.dumps(out))
''' % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    runs = []
    for mode in (None, "ldg"):
        env = dict(os.environ)
        env.pop("B2_K5", None)
        if mode:
            env["B2_K5"] = mode
        p = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=600)
        assert p.returncode == 0, p.stderr[-2000:]
        line = [l for l in p.stdout.splitlines() if l.startswith("RESULT ")][-1]
        runs.append(json.loads(line[len("RESULT "):]))
    assert len(runs[0]) == 6 and sum(len(r["tries"]) for r in runs[0]) > 6
    assert runs[0] == runs[1]
