"""The reference's own ICP acceptance tests (/root/reference/src/opt/test/test_icp.cc) run against the ORACLE,
plus self-checks of the oracle's building blocks. CPU only."""
import math

import numpy as np
import pytest

from tests.golden import ref_test_inputs as rti


def test_mt19937_matches_libstdcxx_known_values():
    # std::mt19937 default-constructed (seed 5489): 10000th output is 4123659995 (C++ standard, [rand.predef]).
    g = rti.MT19937(5489)
    v = 0
    for _ in range(10000):
        v = g()
    assert v == 4123659995


def test_ref_identical_cloud_alignment(oracle):
    """test_icp.cc:39-109 — 20 random-posed copies converge to the same pose within 1e-5 per entry."""
    pts, nrm, poses = rti.identical_cloud_alignment_inputs()
    icp = oracle.PointToPlaneICP(use_kdtree=True)
    ids = [icp.AddPointCloud(pts, nrm, T, False) for T in poses]
    assert ids == list(range(20))
    icp.Run(np.float32(0.15) * np.float32(math.sqrt(3)), 0, 100, 1e-7, False)
    T0 = icp.GetResultGlobalTCloud(ids[0])
    for i in ids[1:]:
        Ti = icp.GetResultGlobalTCloud(i)
        assert np.abs(T0[:3, :] - Ti[:3, :]).max() <= 1e-5


def test_ref_plane_with_single_point(oracle):
    """test_icp.cc:111-172 — rank-deficient planar case still aligns."""
    pts, nrm, poses = rti.plane_with_single_point_inputs()
    icp = oracle.PointToPlaneICP(use_kdtree=True)
    a = icp.AddPointCloud(pts, nrm, poses[0], False)
    b = icp.AddPointCloud(pts, nrm, poses[1], False)
    icp.Run(1.5, 0, 100, 1e-7, False)
    Ta, Tb = icp.GetResultGlobalTCloud(a), icp.GetResultGlobalTCloud(b)
    assert np.abs(Ta[:3, :] - Tb[:3, :]).max() <= 1e-5


def test_kdtree_equals_bruteforce(oracle):
    rng = np.random.default_rng(3)
    tgt = rng.uniform(-1, 1, (4000, 3)).astype(np.float32)
    src = rng.uniform(-1.1, 1.1, (3000, 3)).astype(np.float32)
    # exact duplicates in the target force the lowest-index tie-break
    tgt[100:110] = tgt[50:60]
    src[:10] = tgt[50:60]
    for d in (0.02, 0.1, 0.5):
        qa, ma, da = oracle.find_correspondences(src, tgt, d, use_kdtree=True)
        qb, mb, db = oracle.find_correspondences(src, tgt, d, use_kdtree=False)
        assert np.array_equal(qa, qb) and np.array_equal(ma, mb) and np.array_equal(da, db)
        r2 = np.float32(np.float64(np.float32(d)) ** 2)
        assert (da < r2).all()
    assert set(ma[:10]) <= set(range(50, 60))


def test_find_correspondences_against_numpy(oracle):
    rng = np.random.default_rng(5)
    tgt = rng.uniform(0, 1, (500, 3)).astype(np.float32)
    src = rng.uniform(0, 1, (300, 3)).astype(np.float32)
    d = np.float32(0.08)
    q, m, d2 = oracle.find_correspondences(src, tgt, d)
    diff = src[:, None, :] - tgt[None, :, :]
    D = (diff[..., 0] * diff[..., 0] + diff[..., 1] * diff[..., 1]) + diff[..., 2] * diff[..., 2]
    assert D.dtype == np.float32
    nn = D.argmin(1)
    ok = D[np.arange(300), nn] < np.float32(np.float64(d) * np.float64(d))
    assert np.array_equal(q, np.nonzero(ok)[0])
    assert np.array_equal(m, nn[ok])
    assert np.array_equal(d2, D[np.arange(300), nn][ok])


def test_empty_and_disjoint_inputs(oracle):
    a = np.zeros((0, 3), np.float32)
    b = np.random.default_rng(0).uniform(0, 1, (10, 3)).astype(np.float32)
    assert len(oracle.find_correspondences(a, b, 0.1)[0]) == 0
    assert len(oracle.find_correspondences(b, a, 0.1)[0]) == 0
    assert len(oracle.find_correspondences(b, b + 10.0, 0.1)[0]) == 0
    # radius is strict: a point at exactly distance d is NOT a correspondence
    p = np.array([[0, 0, 0]], np.float32); t = np.array([[0.5, 0, 0]], np.float32)
    assert len(oracle.find_correspondences(p, t, 0.5)[0]) == 0
    assert len(oracle.find_correspondences(p, t, np.nextafter(np.float32(0.5), np.float32(1)))[0]) == 1


def test_transform_cloud_op_order(oracle):
    rng = np.random.default_rng(7)
    xyz = rng.uniform(-5, 5, (1000, 3)).astype(np.float32)
    nrm = rng.normal(size=(1000, 3)).astype(np.float32)
    T = np.eye(4, dtype=np.float32)
    T[:3, :3] = rti.angle_axis_matrix(0.3, np.array([0.6, 0.0, 0.8], np.float32))
    T[:3, 3] = [1.5, -2.0, 0.25]
    ox, on = oracle.transform_cloud(xyz, nrm, T)
    f = np.float32
    for r in range(3):
        ex = xyz[:, 0] * T[r, 0] + (xyz[:, 1] * T[r, 1] + (xyz[:, 2] * T[r, 2] + T[r, 3]))
        en = nrm[:, 0] * T[r, 0] + (nrm[:, 1] * T[r, 1] + nrm[:, 2] * T[r, 2])
        assert ex.dtype == f and np.array_equal(ox[:, r], ex) and np.array_equal(on[:, r], en)


def test_se3_exp_and_ldlt(oracle):
    # small-angle and general branches against scipy-free closed forms
    q0 = np.array([0, 0, 0, 1], np.float32); t0 = np.zeros(3, np.float32)
    q, t = oracle.se3_exp_left_mul(np.array([0.1, -0.2, 0.3, 0, 0, 0]), q0, t0)
    assert np.allclose(q, [0, 0, 0, 1]) and np.allclose(t, [0.1, -0.2, 0.3])
    w = np.array([0.0, 0.0, math.pi / 2])
    q, t = oracle.se3_exp_left_mul(np.concatenate([[1.0, 0, 0], w]), q0, t0)
    assert np.allclose(q, [0, 0, math.sin(math.pi / 4), math.cos(math.pi / 4)], atol=1e-6)
    # V*u for rotation about z by 90deg: [sin/th, (1-cos)/th] * 1
    assert np.allclose(t, [1 / (math.pi / 2), 1 / (math.pi / 2), 0], atol=1e-6)
    rng = np.random.default_rng(11)
    A = rng.normal(size=(42, 60)); H = A @ A.T + 0.1 * np.eye(42); b = rng.normal(size=42)
    x = oracle.ldlt_solve_upper(np.triu(H), b)
    assert np.allclose(H @ x, b, atol=1e-9)
    # rank-deficient + damping
    A = rng.normal(size=(12, 5)); H = A @ A.T + 1e-3 * np.eye(12); b = rng.normal(size=12)
    x = oracle.ldlt_solve_upper(np.triu(H), b)
    assert np.allclose(H @ x, b, atol=1e-7)


def test_normal_equations_upper_triangle_quirk(oracle):
    """Row A8 (impl.h:82-113 + :226): for a pair i<k only the i->k set contributes cross terms."""
    rng = np.random.default_rng(2)
    base = rng.uniform(-1, 1, (400, 3)).astype(np.float32)
    nrm = rng.normal(size=(400, 3)); nrm = (nrm / np.linalg.norm(nrm, axis=1, keepdims=True)).astype(np.float32)
    icp = oracle.PointToPlaneICP(use_kdtree=False, inner_max_iterations=1)
    I = np.eye(4, dtype=np.float32)
    for k in range(3):
        T = I.copy(); T[:3, 3] = 0.001 * k
        icp.AddPointCloud(base, nrm, T, False)
    icp.Run(0.05, 0, 1, 0.0, False)
    H, b = icp.normal_equations()
    assert H.shape == (12, 12)
    assert np.allclose(H, H.T)
    pairs = icp.pairs()
    assert [(s, t) for s, t, *_ in pairs] == [(0, 1), (0, 2), (1, 0), (1, 2), (2, 0), (2, 1)]
    # recompute in numpy fp64 from the global-frame clouds with the quirk
    clouds = [oracle.transform_cloud(base, nrm, np.array(I + np.pad(np.zeros((3, 3)), ((0, 1), (0, 1))) ))[0] for _ in range(3)]
    He = np.zeros((12, 12)); be = np.zeros(12)
    for s, t, q, m, _ in pairs:
        T_s = I.copy(); T_s[:3, 3] = np.float32(0.001) * s
        T_t = I.copy(); T_t[:3, 3] = np.float32(0.001) * t
        ps, ns = oracle.transform_cloud(base, nrm, T_s); pt, nt = oracle.transform_cloud(base, nrm, T_t)
        ps, ns, pt, nt = (a.astype(np.float64) for a in (ps[q], ns[q], pt[m], nt[m]))
        r1 = (ns * (pt - ps)).sum(1); r2 = (nt * (ps - pt)).sum(1)
        j1t = np.concatenate([ns, np.cross(pt, ns)], 1); j1s = -j1t
        j2s = np.concatenate([nt, np.cross(ps, nt)], 1); j2t = -j2s
        sv, tv = 6 * (s - 1), 6 * (t - 1)
        for r, js, jt in ((r1, j1s, j1t), (r2, j2s, j2t)):
            if sv >= 0:
                He[sv:sv + 6, sv:sv + 6] += js.T @ js; be[sv:sv + 6] += js.T @ r
                if tv >= 0 and sv < tv:
                    He[sv:sv + 6, tv:tv + 6] += js.T @ jt; He[tv:tv + 6, sv:sv + 6] += jt.T @ js
            if tv >= 0:
                He[tv:tv + 6, tv:tv + 6] += jt.T @ jt; be[tv:tv + 6] += jt.T @ r
    assert np.linalg.norm(H - He) / np.linalg.norm(He) < 1e-5
    assert np.linalg.norm(b - be) / np.linalg.norm(be) < 1e-5
