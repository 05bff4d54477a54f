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
