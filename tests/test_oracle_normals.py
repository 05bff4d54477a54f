"""Oracle self-checks for the two-pass kNN normals (the reference has no test for them: parity UNPINNED, see
oracle/orc_normals.cc). CPU only."""
import numpy as np


def test_knn_indices_match_numpy(oracle):
    rng = np.random.default_rng(4)
    xyz = rng.uniform(0, 1, (600, 3)).astype(np.float32)
    xyz[10] = xyz[20]                       # duplicate point: (d2, index) tie-break
    out, idx = oracle.normals_knn(xyz, 8, return_indices=True)
    diff = xyz[:, None, :] - xyz[None, :, :]
    D = (diff[..., 0] * diff[..., 0] + diff[..., 1] * diff[..., 1]) + diff[..., 2] * diff[..., 2]
    order = np.lexsort((np.broadcast_to(np.arange(600), D.shape), D), axis=1)[:, :8]
    assert np.array_equal(idx, order)
    assert (idx[:, 0] == np.arange(600))[np.arange(600) != 20].all()   # self first, except the later duplicate


def test_planar_patch_normal_and_curvature(oracle):
    rng = np.random.default_rng(5)
    uv = rng.uniform(-1, 1, (2000, 2))
    n = np.array([0.3, -0.5, 0.81]); n /= np.linalg.norm(n)
    a = np.cross(n, [1, 0, 0]); a /= np.linalg.norm(a); b = np.cross(n, a)
    xyz = (uv[:, :1] * a + uv[:, 1:] * b + 3.0 * n).astype(np.float32)
    out = oracle.normals_knn(xyz, 16, viewpoint=(0, 0, 0))
    # plane at distance 3 along n from the origin: normals flipped towards the viewpoint = -n
    assert np.abs(out[:, :3] @ n + 1.0).max() < 2e-3
    assert out[:, 3].max() < 1e-3
    assert np.isfinite(out).all()


def test_too_few_points_gives_nan(oracle):
    xyz = np.array([[0, 0, 0], [1, 0, 0]], np.float32)
    out = oracle.normals_knn(xyz, 8)
    assert np.isnan(out).all()


def test_two_pass_covariance_matches_float_sequential(oracle):
    rng = np.random.default_rng(6)
    xyz = (rng.normal(size=(50, 3)) * [1, 1, 0.01] + [100, 50, 2]).astype(np.float32)   # far from the origin: two-pass matters
    out, idx = oracle.normals_knn(xyz, 50, return_indices=True)
    # all 50 points are everyone's neighbourhood -> same covariance up to summation order; normal ~ z
    assert (np.abs(np.abs(out[:, 2]) - 1) < 1e-3).all()
    assert (out[:, 2] * (0 - xyz[:, 2]) > 0).all()    # flipped towards the viewpoint (origin)
