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
