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


def _radius_scene(model, seed=8):
    """Images of the textured plane + two 'scans' of it (jittered grids with the texture as colour)."""
    from dataset_pipeline_b200.synth import reg_scene
    sc = reg_scene.make_scene(num_images=3, width=320, height=240, fx=260.0, camera_model=model, num_scales=1, base_radius=0.004)
    rng = np.random.default_rng(seed)
    scans = []
    for s, (step, x0, x1) in enumerate(((0.011, -1.3, 0.25), (0.012, -0.25, 1.3))):      # two scans with a strip of overlap
        gx, gy = np.meshgrid(np.arange(x0, x1, step), np.arange(-1.0, 1.0, step), indexing="xy")
        x = gx.ravel() + rng.uniform(-0.3, 0.3, gx.size) * step; y = gy.ravel() + rng.uniform(-0.3, 0.3, gx.size) * step
        xyz = np.stack([x, y, rng.normal(0, 2e-4, gx.size)], 1).astype(np.float32)
        g = np.clip(reg_scene.texture(x, y), 0, 255).astype(np.uint8)
        scans.append((xyz, np.stack([g, g, g], 1)))
    return sc, scans


def _load_images(reg, sc, splat_points):
    w, h, K = sc["intr"]
    reg.add_intrinsics(w, h, K, camera_model=sc["camera_model"])
    for img, T in zip(sc["images"], sc["poses_gt"]):
        reg.add_image(0, img, None, T)
    count = reg.initialize()
    reg.set_splat_points(splat_points)
    reg.set_image_scale(0)
    return count


@pytest.mark.parametrize("model", [4, 14, 5])
def test_compute_min_max_point_radius(oracle, model):
    import dataset_pipeline_b200 as b2
    from dataset_pipeline_b200 import registration as R
    sc, scans = _radius_scene(model)
    pts = np.concatenate([x for x, _ in scans])
    area = 320 * 240 // 64
    g = b2.Registration(R.default_params(max_initial_image_area_in_pixels=area)); o = oracle.Registration(oracle.reg_default_params(max_initial_image_area_in_pixels=area))
    cg = _load_images(g, sc, pts); co = _load_images(o, sc, pts)
    assert cg == co == 4
    msf = float(np.float32(2.0 ** (-(cg - 1))))
    lg, hg = g.ComputeMinMaxPointRadius(pts, msf); lo, ho = o.min_max_point_radius(pts, msf)
    seen = np.isfinite(lo)
    assert np.array_equal(np.isfinite(lg), seen) and 0.3 < seen.mean() <= 1.0
    if model == 5:        # the fisheye Undistort goes through tanf(): device and glibc differ by an ulp or two in the ray direction, and the
        # radius is a difference of two points ~2 m away (cancellation: 1e-7 * 2 m on a 4 mm radius) -> 1e-3 relative
        assert np.allclose(lg[seen], lo[seen], rtol=1e-3, atol=0) and np.allclose(hg[seen], ho[seen], rtol=1e-3, atol=0)
    else:
        assert np.array_equal(lg, lo) and np.array_equal(hg, ho)
    assert np.array_equal(hg[seen] >= lg[seen], np.ones(seen.sum(), bool))


def test_compute_multi_res_point_cloud_pipeline(oracle):
    """Problem::ComputeMultiResPointCloud (problem.cc:161-362) end to end: radii, scale loop, scale filter, neighbours, gradient filter."""
    import dataset_pipeline_b200 as b2
    from dataset_pipeline_b200 import multiscale as MS
    from dataset_pipeline_b200 import registration as R
    sc, scans = _radius_scene(4)
    pts = np.concatenate([x for x, _ in scans])
    area = 320 * 240 // 64
    g = b2.Registration(R.default_params(max_initial_image_area_in_pixels=area)); o = oracle.Registration(oracle.reg_default_params(max_initial_image_area_in_pixels=area))
    cg = _load_images(g, sc, pts); _load_images(o, sc, pts)
    got = MS.ComputeMultiResPointCloud(g, scans, cg)
    ref = oracle.ms_compute_multi_res_point_cloud(o, scans, cg)
    assert len(got[0]) == len(ref[0]) >= 2
    for k in range(len(ref[0])):
        assert got[0][k] == ref[0][k]
        for a in range(1, 5):
            assert np.array_equal(got[a][k], ref[a][k]), (k, a)
        assert len(ref[1][k]) >= 52


# ---- the reference's own tests (src/opt/test/test_multi_scale_point_cloud.cc) through the C ABI ---------------------------------
def test_reference_merge_close_points_through_the_abi():
    from dataset_pipeline_b200 import multiscale as MS
    from tests.test_oracle_multiscale import check_ref_merge, ref_merge_inputs
    xyz, colors, scans, max_radius = ref_merge_inputs()
    check_ref_merge(xyz, colors, scans, max_radius, *MS.MergeClosePoints(1.0, 2, xyz, colors, scans, max_radius))


def test_reference_preprocess_scans():
    """test_multi_scale_point_cloud.cc:109-151 (host glue of the tool, mirrored in dataset_pipeline_b200/multiscale.py)."""
    from dataset_pipeline_b200 import multiscale as MS
    scans = [(np.array([[1, 2, 3]], np.float32), np.array([[5, 5, 5]], np.uint8)), (np.array([[7, 8, 9]], np.float32), np.array([[11, 11, 11]], np.uint8))]
    pts, cols, idx = MS.PreprocessScans(scans)
    assert len(pts) == len(cols) == len(idx) == 2
    for i in range(2):
        k = int(idx[i])
        assert k in (0, 1) and np.allclose(pts[i], scans[k][0][0], rtol=5e-7) and np.isclose(cols[i], (5, 11)[k], rtol=5e-7)


def test_reference_create_multi_scale_point_cloud_through_the_abi():
    import dataset_pipeline_b200 as b2
    from dataset_pipeline_b200 import multiscale as MS
    from dataset_pipeline_b200 import registration as R
    from tests.test_oracle_multiscale import ref_create_inputs, run_ref_create
    run_ref_create(b2.Registration(R.default_params(image_scale_count_override=3)), MS.CreateMultiScalePointCloud, ref_create_inputs())
