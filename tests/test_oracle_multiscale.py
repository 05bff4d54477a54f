"""CPU checks of the multi-resolution point-cloud oracle (oracle/orc_multiscale.cc): the reference's own DeterminePointNeighbors test
(/root/reference/src/opt/test/test_problem.cc:35-109) and an independent numpy restatement of MergeClosePoints' greedy sweep."""
import numpy as np


def test_reference_determine_point_neighbors(oracle):
    """test_problem.cc:35-109: six points on a line, scans alternating; candidate count = neighbour count = 2."""
    pts = np.array([[i, 0, 0] for i in range(6)], np.float32)
    scan = np.array([0, 1, 0, 1, 0, 1], np.uint8)
    a = oracle.ms_point_neighbors(pts, scan, 2, True, 2, 2); a.sort(1)
    assert a.tolist() == [[2, 4], [3, 5], [0, 4], [1, 5], [0, 2], [1, 3]]                  # :67-78
    b = oracle.ms_point_neighbors(pts, scan, 2, False, 2, 2); b.sort(1)
    assert b.tolist() == [[1, 2], [0, 2], [1, 3], [2, 4], [3, 5], [3, 4]]                  # :97-108


def test_neighbors_are_a_subset_of_the_candidates_and_deterministic(oracle):
    rng = np.random.default_rng(5)
    x = rng.uniform(0, 1, (500, 3)).astype(np.float32)
    s = np.zeros(500, np.uint8)
    a = oracle.ms_point_neighbors(x, s, 1, False, 25, 5)
    b = oracle.ms_point_neighbors(x, s, 1, False, 25, 5)
    assert np.array_equal(a, b)
    d = ((x[:, None, :] - x[None, :, :]) ** 2).sum(-1)
    order = np.argsort(d, 1, kind="stable")[:, 1:26]
    for i in range(500):
        assert len(set(a[i].tolist())) == 5 and set(a[i].tolist()) <= set(order[i].tolist())
    # the shuffle really permutes: not simply the five nearest for every point
    assert any(set(a[i].tolist()) != set(order[i, :5].tolist()) for i in range(500))


def _merge_numpy(x, col, scan, maxr, num_scans, md):
    """Plain restatement of multi_scale_point_cloud.cc:44-124 with a brute-force radius search sorted by (d2, index)."""
    n = len(x)
    r2 = np.float32(np.float64(md) * np.float64(md))
    done = np.zeros(n, bool)
    out = []
    for i in range(n):
        if done[i]:
            continue
        dx = x[i, 0] - x[:, 0]; dy = x[i, 1] - x[:, 1]; dz = x[i, 2] - x[:, 2]
        d2 = ((dx * dx) + (dy * dy)) + (dz * dz)
        idx = np.nonzero(d2 < r2)[0]
        idx = idx[np.lexsort((idx, d2[idx]))]
        acc = np.zeros(3, np.float32); merged = np.zeros(num_scans, np.int64); csum = np.zeros(num_scans, np.float32)
        mx, best, best_scan = np.float32(-1), 0, -1
        for j in idx:
            acc = (acc + x[j]).astype(np.float32)
            csum[scan[j]] = np.float32(csum[scan[j]] + col[j])
            if maxr[j] > mx:
                mx = maxr[j]
            merged[scan[j]] += 1
            if merged[scan[j]] > best:
                best, best_scan = merged[scan[j]], scan[j]
            done[j] = True
        out.append((acc / np.float32(len(idx)), np.float32(csum[best_scan] / np.float32(merged[best_scan])), best_scan, mx))
    return out


def test_merge_close_points_against_numpy(oracle):
    rng = np.random.default_rng(11)
    n = 1500
    x = rng.uniform(0, 1, (n, 3)).astype(np.float32); x[:, 2] *= np.float32(0.02)
    x[100:140] = x[100] + rng.normal(0, 1e-3, (40, 3)).astype(np.float32)        # a dense cluster
    x[200] = x[7]                                                                 # an exact duplicate
    col = rng.uniform(0, 255, n).astype(np.float32); scan = rng.integers(0, 3, n).astype(np.uint8)
    maxr = rng.uniform(0.01, 0.2, n).astype(np.float32)
    ox, oc, os_, om = oracle.ms_merge_close_points(x, col, scan, maxr, 3, 0.06)
    ref = _merge_numpy(x, col, scan, maxr, 3, 0.06)
    assert len(ref) == len(oc) and 50 < len(oc) < n
    for k, (p, c, s, m) in enumerate(ref):
        assert np.array_equal(ox[k], p) and oc[k] == c and os_[k] == s and om[k] == m, k


def test_create_scale_loop_structure(oracle):
    rng = np.random.default_rng(2)
    n = 4000
    x = rng.uniform(-1, 1, (n, 3)).astype(np.float32); x[:, 2] = np.float32(2.0)
    col = rng.uniform(0, 255, n).astype(np.float32); scan = rng.integers(0, 2, n).astype(np.uint8)
    lo = (0.002 * (1 + rng.uniform(0, 1, n))).astype(np.float32); hi = (lo * 40).astype(np.float32)
    lo[::97] = np.inf; hi[::97] = -np.inf                                         # points no image observes
    scales = oracle.ms_create(x, col, scan, lo, hi, 2)
    assert len(scales) >= 4
    r0 = np.float32(np.float32(lo.min()) * np.float32(1.05))
    assert scales[0][0] == r0
    for a, b in zip(scales, scales[1:]):
        assert b[0] == np.float32(np.float64(a[0]) * 2) or abs(b[0] - 2 * a[0]) < 1e-9
        assert len(b[1]) <= len(a[1]) + n
    # merged points of one scale are at least the merge distance apart from every EARLIER centre's neighbourhood: centres of a scale
    # cannot be closer than... (they are averages, so only a weak sanity bound is asserted) and there are fewer of them at coarser scales
    assert len(scales[-1][1]) < len(scales[0][1])


# ---- the reference's own tests: src/opt/test/test_multi_scale_point_cloud.cc ----------------------------------------------------
def ref_merge_inputs():
    """test_multi_scale_point_cloud.cc:37-70: three points in a row that merge (the middle one from another scan) + one far point."""
    xyz = np.array([[0.1, 0, 0], [0.5, 0, 0], [0.9, 0, 0], [0.5, 0, 2]], np.float32)
    colors = np.array([0, 44, 2, 99], np.float32)
    scans = np.array([0, 1, 0, 1], np.uint8)
    max_radius = np.array([13, 12, 11, 47], np.float32)
    return xyz, colors, scans, max_radius


def check_ref_merge(xyz, colors, scans, max_radius, ox, oc, os_, om):
    """The assertions of test_multi_scale_point_cloud.cc:81-107 (EXPECT_FLOAT_EQ = 4 ulp)."""
    assert len(ox) == len(oc) == len(os_) == 2
    seen = set()
    for i in range(2):
        if os_[i] == 0:          # the merged first three input points; colour = mean over the winning scan's points only
            assert np.allclose(ox[i], [0.5, 0, 0], rtol=5e-7, atol=1e-7) and np.isclose(oc[i], 1, rtol=5e-7) and np.isclose(om[i], 13, rtol=5e-7)
        elif os_[i] == 1:        # the single far point of scan 1, untouched
            assert np.array_equal(ox[i], xyz[3]) and oc[i] == colors[3] and om[i] == max_radius[3]
        else:
            raise AssertionError("invalid scan index %d" % os_[i])
        seen.add(int(os_[i]))
    assert seen == {0, 1}


def test_reference_merge_close_points(oracle):
    xyz, colors, scans, max_radius = ref_merge_inputs()
    check_ref_merge(xyz, colors, scans, max_radius, *oracle.ms_merge_close_points(xyz, colors, scans, max_radius, 2, 1.0))


def ref_create_inputs():
    """test_multi_scale_point_cloud.cc:164-214: one 640x480 pinhole camera (fx = 640, fy = 480) in the default pose, a constant image,
    one point in front of the camera and one behind it; minimum scaling factor 2^-2, 3 image scales."""
    w, h = 640, 480
    return dict(w=w, h=h, K=np.array([w, h, w / 2 - 0.5, h / 2 - 0.5], np.float32), image=np.full((h, w), 100, np.uint8),
                pose=np.array([0, 0, 0, 1, 0, 0, 0], np.float32), points=np.array([[0, 0, 2], [0, 0, -2]], np.float32),
                colors=np.array([12, 33], np.float32), scans=np.zeros(2, np.uint8), min_scaling_factor=float(np.float32(2.0 ** -2)))


def run_ref_create(reg, create, ci):
    """CreateMultiScalePointCloud on the reference test's input, then the test's assertions (:241-289)."""
    reg.add_intrinsics(ci["w"], ci["h"], ci["K"])
    reg.add_image(0, ci["image"], None, ci["pose"])
    assert reg.initialize() == 3
    reg.set_splat_points(ci["points"])
    mmr = getattr(reg, "min_max_point_radius", None) or reg.ComputeMinMaxPointRadius
    lo, hi = mmr(ci["points"], ci["min_scaling_factor"])
    scales = create(ci["points"], ci["colors"], ci["scans"], lo, hi, 1)
    assert len(scales) == 2                                   # the point in front of the camera, at 2 scales
    for radius, xyz, col, si in scales:
        assert len(xyz) == len(col) == len(si) == 1
        assert col[0] == 12 and si[0] == 0 and np.array_equal(xyz[0], ci["points"][0])
        reg.add_point_scale(xyz, float(radius), np.zeros((1, 5), np.uint64), col)
    reg.set_image_scale(0)
    (getattr(reg, "create_observations", None) or reg.CreateObservationsForAllImages)(0)
    got = []
    for ps in range(2):
        idx, x, y, s, nb = reg.observations(0, ps)
        assert len(idx) == 1, "wrong number of observations on point scale %d" % ps
        got.append(float(s[0]))
    assert any(0 <= s < 1 for s in got) and any(1 <= s < 2 for s in got)


def test_reference_create_multi_scale_point_cloud(oracle):
    ci = ref_create_inputs()
    run_ref_create(oracle.Registration(oracle.reg_default_params(image_scale_count_override=3)), oracle.ms_create, ci)
