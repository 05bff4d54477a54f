"""Host-side mirror of the multi-resolution point-cloud construction that feeds Path B:
opt::MergeClosePoints / opt::CreateMultiScalePointCloud (/root/reference/src/opt/multi_scale_point_cloud.cc:44-124, 214-368) and
opt::Problem::DeterminePointNeighbors (/root/reference/src/opt/problem.cc:706-786). Same argument meaning as the reference functions."""
import ctypes as C

import numpy as np

from . import _lib


class MsStats(C.Structure):
    _fields_ = [("neighbor_pairs", C.c_uint64), ("rounds", C.c_int32), ("scales", C.c_int32), ("ms_device", C.c_float)]


def _bind():
    L = _lib.lib()
    if getattr(L, "_ms_bound", False):
        return L
    fp, u8, u64 = C.POINTER(C.c_float), C.POINTER(C.c_uint8), C.POINTER(C.c_uint64)
    L.b2_ms_merge_close_points.argtypes = [fp, C.c_size_t, fp, u8, fp, C.c_int, C.c_float, fp, fp, u8, fp, C.POINTER(C.c_size_t), C.POINTER(MsStats)]
    L.b2_ms_create.argtypes = [fp, C.c_size_t, fp, u8, fp, fp, C.c_int, C.c_float, C.c_float, C.c_int, C.c_size_t, C.POINTER(C.c_int), fp, u64, fp, fp, u8,
                               C.POINTER(MsStats)]
    L.b2_ms_point_neighbors.argtypes = [fp, C.c_size_t, u8, C.c_int, C.c_int, C.c_int, C.c_int, u64]
    L._ms_bound = True
    return L


def _f(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _u8(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint8))


def _stats(s):
    return {k: getattr(s, k) for k, _ in MsStats._fields_}


def MergeClosePoints(merge_distance, num_scans, in_points, in_colors, in_scan_indices, in_max_radius, return_stats=False):
    """-> (out_points (m,3), out_colors, out_scan_indices, out_max_radius)   (multi_scale_point_cloud.cc:44-124)"""
    x = np.ascontiguousarray(in_points, np.float32); c = np.ascontiguousarray(in_colors, np.float32)
    s = np.ascontiguousarray(in_scan_indices, np.uint8); m = np.ascontiguousarray(in_max_radius, np.float32)
    n = x.shape[0]
    if x.shape != (n, 3) or c.shape != (n,) or s.shape != (n,) or m.shape != (n,):
        raise ValueError("expected (n,3) points and (n,) colours / scan indices / max radii")
    ox = np.zeros((max(n, 1), 3), np.float32); oc = np.zeros(max(n, 1), np.float32); os_ = np.zeros(max(n, 1), np.uint8); om = np.zeros(max(n, 1), np.float32)
    cnt = C.c_size_t(0); st = MsStats()
    _lib.check(_bind().b2_ms_merge_close_points(_f(x), n, _f(c), _u8(s), _f(m), int(num_scans), float(merge_distance), _f(ox), _f(oc), _u8(os_), _f(om),
                                                C.byref(cnt), C.byref(st)))
    k = cnt.value
    out = (ox[:k].copy(), oc[:k].copy(), os_[:k].copy(), om[:k].copy())
    return out + (_stats(st),) if return_stats else out


def CreateMultiScalePointCloud(points, colors, scan_indices, min_radius, max_radius, num_scans, min_radius_bias=1.05, merge_distance_factor=4.0,
                               max_scales=32, return_stats=False):
    """The scale loop of CreateMultiScalePointCloud (multi_scale_point_cloud.cc:263-368) on the per-point radii ComputeMinMaxPointRadius
    produced. -> list of (point_radius, points (m,3), colors, scan_indices), finest scale first."""
    x = np.ascontiguousarray(points, np.float32); c = np.ascontiguousarray(colors, np.float32); s = np.ascontiguousarray(scan_indices, np.uint8)
    lo = np.ascontiguousarray(min_radius, np.float32); hi = np.ascontiguousarray(max_radius, np.float32)
    n = x.shape[0]
    cap = max(n, 1) * 2 + 16          # every scale is at most the active set; the sets shrink ~4x per scale
    while True:
        rad = np.zeros(max_scales, np.float32); cnt = np.zeros(max_scales, np.uint64)
        ox = np.zeros((cap, 3), np.float32); oc = np.zeros(cap, np.float32); os_ = np.zeros(cap, np.uint8)
        k = C.c_int(0); st = MsStats()
        rc = _bind().b2_ms_create(_f(x), n, _f(c), _u8(s), _f(lo), _f(hi), int(num_scans), float(min_radius_bias), float(merge_distance_factor), int(max_scales),
                                  cap, C.byref(k), _f(rad), cnt.ctypes.data_as(C.POINTER(C.c_uint64)), _f(ox), _f(oc), _u8(os_), C.byref(st))
        if rc == 1 and b"capacity" in _lib.lib().b2_last_error() and cap < max(n, 1) * max_scales:
            cap *= 2
            continue
        _lib.check(rc)
        break
    out, off = [], 0
    for i in range(k.value):
        m = int(cnt[i])
        out.append((float(rad[i]), ox[off:off + m].copy(), oc[off:off + m].copy(), os_[off:off + m].copy()))
        off += m
    return (out, _stats(st)) if return_stats else out


def DeterminePointNeighbors(scan_count, limit_neighbors_to_same_scan_index, point_cloud, scan_indices, point_neighbor_candidate_count=25,
                            point_neighbor_count=5):
    """-> (n, point_neighbor_count) uint64 neighbour indices   (problem.cc:706-786)"""
    x = np.ascontiguousarray(point_cloud, np.float32); s = np.ascontiguousarray(scan_indices, np.uint8)
    n = x.shape[0]
    out = np.zeros((n, point_neighbor_count), np.uint64)
    _lib.check(_bind().b2_ms_point_neighbors(_f(x), n, _u8(s), int(scan_count), int(bool(limit_neighbors_to_same_scan_index)),
                                             int(point_neighbor_candidate_count), int(point_neighbor_count), out.ctypes.data_as(C.POINTER(C.c_uint64))))
    return out


def PreprocessScans(scans):
    """multi_scale_point_cloud.cc:186-212: scans = [(xyz (n,3) float32, rgb (n,3) uint8), ...] -> (points, colors, scan_indices);
    colour = 0.299 r + 0.587 g + 0.114 b evaluated in double, stored as float."""
    pts = np.concatenate([np.ascontiguousarray(x, np.float32) for x, _ in scans])
    cols = np.concatenate([(0.299 * c[:, 0].astype(np.float64) + 0.587 * c[:, 1].astype(np.float64) + 0.114 * c[:, 2].astype(np.float64)).astype(np.float32)
                           for _, c in scans])
    idx = np.concatenate([np.full(len(x), i, np.uint8) for i, (x, _) in enumerate(scans)])
    return pts, cols, idx


def _enough_points(scan_indices, num_scans, use_fixed_scan_colors, candidate_count):
    if use_fixed_scan_colors:
        return bool((np.bincount(scan_indices, minlength=num_scans)[:num_scans] >= candidate_count + 1).all())
    return len(scan_indices) >= candidate_count + 1


def ComputeMultiResPointCloud(reg, scans, image_scale_count, fixed_residuals_weight=1.0, point_neighbor_count=5, point_neighbor_candidate_count=25,
                              min_mean_intensity_difference_for_points=5, min_radius_bias=1.05, merge_distance_factor=4.0):
    """Problem::ComputeMultiResPointCloud (/root/reference/src/opt/problem.cc:161-362) on top of the C ABI: `reg` is an initialised
    Registration holding the images, intrinsics and occlusion geometry (ComputeMinMaxPointRadius needs them).
    -> (point_radii, points, colors, scan_indices, neighbor_indices), one entry per remaining point scale."""
    use_fixed = fixed_residuals_weight > 0
    num_scans = len(scans)
    pts, cols, sidx = PreprocessScans(scans)
    min_scaling = np.float32(2.0 ** (-1 * (image_scale_count - 1)))                    # float minimum_scaling_factor = pow(2, -(count - 1))  (:181-182)
    lo, hi = reg.ComputeMinMaxPointRadius(pts, float(min_scaling))
    scales = CreateMultiScalePointCloud(pts, cols, sidx, lo, hi, num_scans, min_radius_bias, merge_distance_factor)
    scales = [list(s) for s in scales if _enough_points(s[3], num_scans, use_fixed, point_neighbor_candidate_count)]          # :205-241
    K = point_neighbor_count
    for s in scales:                                                                     # :243-302
        _, p, c, si = s
        nb = DeterminePointNeighbors(num_scans, use_fixed, p, si, point_neighbor_candidate_count, K).astype(np.int64)
        diff = np.zeros(len(c), np.float32)
        for k in range(K):
            diff = (diff + np.abs(c[nb[:, k]] - c)).astype(np.float32)
        drop = (diff / np.float32(K)) < np.float32(min_mean_intensity_difference_for_points)
        keep = ~drop
        keep2 = keep.copy()
        keep2[nb[keep].ravel()] = True                                                    # neighbours of kept points stay too
        s[1], s[2], s[3] = p[keep2], c[keep2], si[keep2]
    scales = [s for s in scales if _enough_points(s[3], num_scans, use_fixed, point_neighbor_candidate_count)]                 # :304-350
    nbrs = [DeterminePointNeighbors(num_scans, use_fixed, s[1], s[3], point_neighbor_candidate_count, K) for s in scales]    # :352-361
    return [s[0] for s in scales], [s[1] for s in scales], [s[2] for s in scales], [s[3] for s in scales], nbrs
