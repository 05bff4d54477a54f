"""GPU parity of the multi-resolution point-cloud construction (b2_ms_*: MergeClosePoints, CreateMultiScalePointCloud's scale loop,
DeterminePointNeighbors) against the oracle through the C ABI. Index / integer results and the fp32 averages are bit-exact: the centre
set is the unique greedy set of the reference's sweep and every average is summed in radiusSearch order (distance, then index)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _scan_cloud(num_scans=2, w=360, h=150):
    """Raster-ordered room scans (dense clusters at the scanner poles, long dependency chains along the scan lines)."""
    from dataset_pipeline_b200 import synth
    xs, ss = [], []
    for i in range(num_scans):
        xyz, _, _ = synth.room_scan(i, w, h)
        T = np.eye(4); T[:3, :3] = synth.rot_xyz(0, 0, 0.35 * i); T[:3, 3] = synth.SCANNER_POSITIONS[i]
        xs.append((xyz.astype(np.float64) @ T[:3, :3].T + T[:3, 3]).astype(np.float32)); ss.append(np.full(len(xyz), i, np.uint8))
    return np.concatenate(xs), np.concatenate(ss)


def _attrs(n, seed):
    rng = np.random.default_rng(seed)
    return rng.uniform(0, 255, n).astype(np.float32), rng.uniform(0.01, 0.3, n).astype(np.float32)


def _same(a, b):
    assert len(a[1]) == len(b[1]), (len(a[1]), len(b[1]))
    for u, v in zip(a, b):
        assert np.array_equal(u, v)


@pytest.mark.parametrize("md", [0.02, 0.08])
def test_merge_close_points_scan_order(oracle, md):
    import dataset_pipeline_b200 as b2
    x, s = _scan_cloud()
    col, maxr = _attrs(len(x), 1)
    got = b2.MergeClosePoints(md, 2, x, col, s, maxr, return_stats=True)
    ref = oracle.ms_merge_close_points(x, col, s, maxr, 2, md)
    _same(got[:4], ref)
    st = got[4]
    assert st["neighbor_pairs"] > len(x) and st["rounds"] >= 32 and 0 < len(ref[1]) < len(x)


def test_merge_close_points_random_duplicates_and_isolated(oracle):
    import dataset_pipeline_b200 as b2
    rng = np.random.default_rng(3)
    n = 20000
    x = rng.uniform(0, 1, (n, 3)).astype(np.float32); x[:, 2] *= np.float32(0.05)
    x[500:900] = x[500] + rng.normal(0, 2e-3, (400, 3)).astype(np.float32)        # dense cluster
    x[1000:1010] = x[3]                                                            # exact duplicates (d2 = 0 ties -> index order)
    x[-1] = (50.0, 50.0, 50.0)                                                      # isolated point: a centre of its own
    s = rng.integers(0, 5, n).astype(np.uint8)
    col, maxr = _attrs(n, 4)
    _same(b2.MergeClosePoints(0.03, 5, x, col, s, maxr), oracle.ms_merge_close_points(x, col, s, maxr, 5, 0.03))
    # a merge distance below the point spacing keeps every point (each its own centre, averages of one)
    got = b2.MergeClosePoints(1e-6, 5, x[2000:4000], col[2000:4000], s[2000:4000], maxr[2000:4000])
    assert len(got[1]) == 2000 and np.array_equal(got[0], x[2000:4000]) and np.array_equal(got[1], col[2000:4000])


def test_merge_is_deterministic_and_rejects_bad_arguments(oracle):
    import dataset_pipeline_b200 as b2
    from dataset_pipeline_b200._lib import B2Error
    x, s = _scan_cloud(1, 200, 80)
    col, maxr = _attrs(len(x), 9)
    a = b2.MergeClosePoints(0.05, 1, x, col, s, maxr); b = b2.MergeClosePoints(0.05, 1, x, col, s, maxr)
    _same(a, b)
    with pytest.raises(B2Error):
        b2.MergeClosePoints(0.0, 1, x, col, s, maxr)
    with pytest.raises(B2Error):
        b2.MergeClosePoints(0.05, 1, x, col, s + 1, maxr)                          # scan index outside [0, num_scans)
    e = b2.MergeClosePoints(0.05, 1, x[:0], col[:0], s[:0], maxr[:0])
    assert len(e[1]) == 0


def test_create_multi_scale_point_cloud(oracle):
    import dataset_pipeline_b200 as b2
    x, s = _scan_cloud(2, 300, 120)
    n = len(x)
    rng = np.random.default_rng(6)
    col = rng.uniform(0, 255, n).astype(np.float32)
    dist = np.linalg.norm(x - np.array([0.5, 0.3, 1.5], np.float32), axis=1)
    lo = (0.0015 * dist * (1 + 0.2 * rng.uniform(0, 1, n))).astype(np.float32); hi = (lo * rng.uniform(8, 64, n)).astype(np.float32)
    lo[::53] = np.inf; hi[::53] = -np.inf                                          # points no image observes never enter
    got, st = b2.CreateMultiScalePointCloud(x, col, s, lo, hi, 2, return_stats=True)
    ref = oracle.ms_create(x, col, s, lo, hi, 2)
    assert len(got) == len(ref) >= 4 and st["scales"] == len(ref)
    for (rg, xg, cg, sg), (rr, xr, cr, sr) in zip(got, ref):
        assert rg == rr and np.array_equal(xg, xr) and np.array_equal(cg, cr) and np.array_equal(sg, sr)


@pytest.mark.parametrize("limit", [False, True])
def test_determine_point_neighbors(oracle, limit):
    import dataset_pipeline_b200 as b2
    x, s = _scan_cloud(2, 240, 100)
    # DeterminePointNeighbors runs on merged clouds: no duplicates (the reference CHECKs the self-match)
    x, keep = np.unique(x, axis=0, return_index=True); s = s[keep]
    got = b2.DeterminePointNeighbors(2, limit, x, s)
    ref = oracle.ms_point_neighbors(x, s, 2, limit)
    assert got.shape == ref.shape == (len(x), 5) and np.array_equal(got, ref)
    if limit:
        assert np.array_equal(s[got.astype(np.int64)], np.repeat(s[:, None], 5, 1))


def test_reference_test_problem_through_the_abi(oracle):
    """/root/reference/src/opt/test/test_problem.cc:35-109."""
    import dataset_pipeline_b200 as b2
    pts = np.array([[i, 0, 0] for i in range(6)], np.float32)
    scan = np.array([0, 1, 0, 1, 0, 1], np.uint8)
    a = b2.DeterminePointNeighbors(2, True, pts, scan, 2, 2); a.sort(1)
    assert a.tolist() == [[2, 4], [3, 5], [0, 4], [1, 5], [0, 2], [1, 3]]
    b = b2.DeterminePointNeighbors(2, False, pts, scan, 2, 2); b.sort(1)
    assert b.tolist() == [[1, 2], [0, 2], [1, 3], [2, 4], [3, 5], [3, 4]]
    from dataset_pipeline_b200._lib import B2Error
    with pytest.raises(B2Error):
        b2.DeterminePointNeighbors(2, True, pts, scan, 25, 5)                      # fewer than 26 points per scan: reference CHECK_GE
