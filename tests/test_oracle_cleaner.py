"""Oracle checks for the point-cloud tools: the C restatement in oracle/orc_cleaner.cc against independent
numpy restatements of local_statistical_outlier_removal.hpp:72-176, point_simplex_squared_distance.cpp:44-135 and splat_creator.cc:146-215.
The reference has no tests for these tools (parity unpinned); these pin the oracle to the written-down algorithm."""
import numpy as np
import pytest

from oracle import oracle as orc


def _knn_bruteforce(x, k):
    d2 = np.zeros((x.shape[0], x.shape[0]), np.float32)
    for a in range(3):                               # ((dx^2 + dy^2) + dz^2) in fp32
        diff = (x[:, None, a] - x[None, :, a]).astype(np.float32)
        d2 = (d2 + diff * diff).astype(np.float32) if a else (diff * diff).astype(np.float32)
    order = np.lexsort((np.broadcast_to(np.arange(x.shape[0]), d2.shape), d2), axis=1)[:, :k]
    return order, np.take_along_axis(d2, order, axis=1)


def _lsor_numpy(x, mean_k, factor, negative=False):
    fin = np.isfinite(x).all(axis=1)
    orig = np.nonzero(fin)[0]
    idx, d2 = _knn_bruteforce(x[fin], mean_k + 1)
    dist = np.zeros(x.shape[0], np.float32)
    for c, i in enumerate(orig):
        s = 0.0
        for j in range(1, mean_k + 1):
            s += float(np.sqrt(np.float32(d2[c, j])))
        dist[i] = np.float32(s / mean_k)
    keep, rem = [], []
    c = 0
    for i in range(x.shape[0]):
        if not fin[i]:
            rem.append(i); continue
        nb = orig[idx[c, 1:]]; c += 1
        dd = dist[nb].astype(np.float64)
        valid = dd > 0
        with np.errstate(invalid="ignore", divide="ignore"):
            mean = np.float64(dd[valid].sum() if valid.any() else 0.0) / np.float64(valid.sum())
        thr = mean * factor
        out = (float(dist[i]) <= thr) if negative else (float(dist[i]) > thr)
        (rem if out else keep).append(i)
    return np.array(keep, np.int32), np.array(rem, np.int32), dist


def _cloud(seed, n=500, outliers=25):
    rng = np.random.default_rng(seed)
    x = rng.uniform(-1, 1, (n, 3)).astype(np.float32)
    x[:, 2] = (0.05 * np.sin(3 * x[:, 0])).astype(np.float32)
    x[rng.choice(n, outliers, replace=False), 2] += rng.uniform(0.2, 0.6, outliers).astype(np.float32)
    return x


@pytest.mark.parametrize("mean_k,factor", [(8, 2.0), (20, 1.5), (1, 3.0)])
def test_lsor_matches_numpy_restatement(mean_k, factor):
    x = _cloud(mean_k)
    keep, rem, dist = orc.lsor_filter(x, mean_k, factor)
    k2, r2, d2 = _lsor_numpy(x, mean_k, factor)
    assert np.array_equal(dist, d2)
    assert np.array_equal(keep, k2) and np.array_equal(rem, r2)
    assert 0 < len(rem) < len(x) // 4               # the planted outliers go, the surface stays


def test_lsor_negative_and_nonfinite():
    x = _cloud(3)
    x[[5, 77, 400]] = [np.nan, 0, 0]
    x[9, 1] = np.inf
    keep, rem, dist = orc.lsor_filter(x, 10, 2.0)
    k2, r2, d2 = _lsor_numpy(x, 10, 2.0)
    assert np.array_equal(keep, k2) and np.array_equal(rem, r2) and np.array_equal(dist, d2)
    assert set([5, 9, 77, 400]) <= set(rem.tolist()) and (dist[[5, 9, 77, 400]] == 0).all()
    kn, rn, _ = orc.lsor_filter(x, 10, 2.0, negative=True)
    finite = np.isfinite(x).all(axis=1)
    assert set(kn.tolist()) == set(rem.tolist()) - set(np.nonzero(~finite)[0].tolist())   # the negative filter keeps exactly the outliers
    assert np.array_equal(np.sort(np.concatenate([kn, rn])), np.arange(len(x)))


def test_lsor_duplicates_have_zero_distance_and_do_not_count():
    x = _cloud(4, n=300, outliers=0)
    x[10:16] = x[10]                                  # six coincident points: mean distance to 3 neighbours is 0
    keep, rem, dist = orc.lsor_filter(x, 3, 2.0)
    assert (dist[10:16] == 0).all()
    k2, r2, d2 = _lsor_numpy(x, 3, 2.0)
    assert np.array_equal(keep, k2) and np.array_equal(dist, d2)


def _closest_f64(p, a, b, c):
    """Ericson's closest point in float64 (reference for the fp32 restatement)."""
    ab, ac, ap = b - a, c - a, p - a
    d1, d2 = ab @ ap, ac @ ap
    if d1 <= 0 and d2 <= 0: return a
    bp = p - b; d3, d4 = ab @ bp, ac @ bp
    if d3 >= 0 and d4 <= d3: return b
    vc = d1 * d4 - d3 * d2
    if (a != b).any() and vc <= 0 and d1 >= 0 and d3 <= 0: return a + d1 / (d1 - d3) * ab
    cp = p - c; d5, d6 = ab @ cp, ac @ cp
    if d6 >= 0 and d5 <= d6: return c
    vb = d5 * d2 - d1 * d6
    if vb <= 0 and d2 >= 0 and d6 <= 0: return a + d2 / (d2 - d6) * ac
    va = d3 * d6 - d5 * d4
    if va <= 0 and d4 - d3 >= 0 and d5 - d6 >= 0: return b + (d4 - d3) / ((d4 - d3) + (d5 - d6)) * (c - b)
    den = 1.0 / (va + vb + vc)
    return a + ab * (vb * den) + ac * (vc * den)


def _mesh(seed, nt=60):
    rng = np.random.default_rng(seed)
    g = 9
    u, v = np.meshgrid(np.linspace(-1, 1, g), np.linspace(-1, 1, g), indexing="ij")
    verts = np.stack([u.ravel(), v.ravel(), 0.1 * np.sin(2 * u.ravel())], axis=1).astype(np.float32)
    faces = []
    for i in range(g - 1):
        for j in range(g - 1):
            a = i * g + j
            faces += [[a, a + 1, a + g], [a + 1, a + g + 1, a + g]]
    extra = rng.uniform(-1, 1, (nt, 3, 3)).astype(np.float32) * 0.3 + rng.uniform(-1, 1, (nt, 1, 3)).astype(np.float32)
    base = len(verts)
    verts = np.concatenate([verts, extra.reshape(-1, 3)])
    faces += [[base + 3 * t, base + 3 * t + 1, base + 3 * t + 2] for t in range(nt)]
    faces.append([0, 0, 5])                           # degenerate (a == b): the AB edge region is skipped (point_simplex_squared_distance.cpp:71)
    return verts, np.array(faces, np.uint32)


def test_mesh_squared_distance_against_float64():
    verts, faces = _mesh(1)
    rng = np.random.default_rng(2)
    pts = rng.uniform(-1.3, 1.3, (300, 3)).astype(np.float32)
    pts[:20] = verts[rng.choice(len(verts), 20)]      # on vertices: distance exactly 0
    got = orc.mesh_squared_distance(pts, verts, faces)
    V = verts.astype(np.float64)
    ref = np.array([min(np.sum((p - _closest_f64(p, V[f[0]], V[f[1]], V[f[2]])) ** 2) for f in faces) for p in pts.astype(np.float64)])
    assert (got[:20] == 0).all()
    assert np.allclose(got, ref, rtol=2e-4, atol=1e-9)


def test_splat_geometry():
    verts, faces = _mesh(3)
    rng = np.random.default_rng(4)
    n = 400
    x = rng.uniform(-1, 1, (n, 3)).astype(np.float32); x[:, 2] = (0.1 * np.sin(2 * x[:, 0]) + rng.normal(0, 0.01, n)).astype(np.float32)
    x[:40, 2] += 0.5                                  # far above the surface: must become splats
    nr = rng.normal(0, 1, (n, 3)).astype(np.float32); nr /= np.linalg.norm(nr, axis=1, keepdims=True)
    nr[50] = [0, 0, 1]; nr[51] = [1e-7, -1e-7, 1]     # second branch of unitOrthogonal
    nr[52] = [np.nan, 0, 1]
    corners, added, radius = orc.create_splats(x, nr, verts, faces, 0.02, 0.15)
    idx, d2 = _knn_bruteforce(x, 5)
    assert np.array_equal(radius[np.arange(n) != 52], np.minimum(np.sqrt(d2[:, 4]), np.float32(0.15))[np.arange(n) != 52])
    assert not added[52] and (corners[52] == 0).all() and radius[52] == 0
    ok = np.arange(n) != 52
    right = (corners[ok, 0] - corners[ok, 3]) / (2 * radius[ok, None]); up = (corners[ok, 0] - corners[ok, 1]) / (2 * radius[ok, None])
    assert np.allclose(np.linalg.norm(right, axis=1), 1, atol=1e-3) and np.allclose(np.einsum("ij,ij->i", right, nr[ok]), 0, atol=1e-3)
    assert np.allclose(up, np.cross(nr[ok], right), atol=2e-3)
    assert np.allclose(corners[ok].mean(axis=1), x[ok], atol=1e-5)
    assert np.allclose(corners[50, 0] - corners[50, 3], [0, -2 * radius[50], 0], atol=1e-6)   # n = z: right = (0, -1, 0)
    # the mesh test: clear cases against float64
    V = verts.astype(np.float64)
    def far(p):
        return min(np.sum((p - _closest_f64(p, V[f[0]], V[f[1]], V[f[2]])) ** 2) for f in faces)
    for i in list(range(40)) + list(range(60, 90)):
        dmax = max(far(q.astype(np.float64)) for q in [x[i]] + list(corners[i]))
        if abs(dmax - 4e-4) > 1e-5:
            assert added[i] == (dmax > 4e-4), i
    assert added[:40].all()


def test_mesh_distance_properties():
    """Size-independent properties of the point-mesh distance: zero on the surface (vertices, edge midpoints, centroids), symmetric under
    a permutation of the faces and of the vertices inside a face's first edge... (the AB-edge branch is the only order-dependent one:
    checked to a tolerance), and never larger than the distance to any vertex."""
    verts, faces = _mesh(7)
    rng = np.random.default_rng(9)
    V = verts.astype(np.float64)
    tri = faces[:-1]                                                    # without the degenerate one
    mids = 0.5 * (V[tri[:, 0]] + V[tri[:, 1]]); cents = (V[tri[:, 0]] + V[tri[:, 1]] + V[tri[:, 2]]) / 3
    on = np.concatenate([V[:30], mids[:60], cents[:60]]).astype(np.float32)
    d_on = orc.mesh_squared_distance(on, verts, faces)
    assert d_on.max() < 1e-10
    pts = rng.uniform(-1.5, 1.5, (400, 3)).astype(np.float32)
    d = orc.mesh_squared_distance(pts, verts, faces)
    d_perm = orc.mesh_squared_distance(pts, verts, faces[rng.permutation(len(faces))])
    assert np.array_equal(d, d_perm)                                    # the minimum does not depend on the face order
    d_rot = orc.mesh_squared_distance(pts, verts, np.roll(faces, 1, axis=1))
    assert np.allclose(d, d_rot, rtol=1e-4, atol=1e-9)                  # rotating a face's vertex order changes the rounding only
    used = np.unique(faces)
    d_vert = ((pts[:, None, :].astype(np.float64) - V[None, used, :]) ** 2).sum(-1).min(1)
    assert (d <= d_vert * (1 + 1e-5) + 1e-9).all()


def test_lsor_scale_and_order_properties():
    """The filter is invariant under a uniform power-of-two scaling of the cloud (every distance scales exactly) and, for clouds without
    tied distances, under a permutation of the points (the kept SET is the same)."""
    x = _cloud(11, n=700, outliers=30)
    keep, rem, dist = orc.lsor_filter(x, 12, 1.6)
    k2, r2, d2 = orc.lsor_filter(x * np.float32(4.0), 12, 1.6)
    assert np.array_equal(keep, k2) and np.array_equal(d2, dist * np.float32(4.0))
    perm = np.random.default_rng(3).permutation(len(x))
    kp, _, dp = orc.lsor_filter(x[perm], 12, 1.6)
    assert set(perm[kp].tolist()) == set(keep.tolist())
    assert np.array_equal(dp, dist[perm])
