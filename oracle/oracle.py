"""ORACLE — TEST INFRASTRUCTURE ONLY. Not part of the product path.

ctypes wrapper over oracle/_build/liboracle.so (the CPU restatement of the reference's
algorithms, see orc_icp.cc / orc_normals.cc headers for the file:line each function follows).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module; nothing under dataset_pipeline_b200/ does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None


def build(force=False):
    """Compile the restatement (gcc, reference flags). Building the checker is not using it."""
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".cc", ".h")) or f == "Makefile"]
    if (not force and os.path.exists(_LIB_PATH)
            and os.path.getmtime(_LIB_PATH) >= max(os.path.getmtime(s) for s in srcs)):
        return _LIB_PATH
    subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _LIB_PATH


class IcpStats(C.Structure):
    _fields_ = [("inner_iterations", C.c_int32), ("lm_tries_total", C.c_int32), ("num_pairs", C.c_int32),
                ("num_variables", C.c_int32), ("num_correspondences", C.c_uint64),
                ("first_cost", C.c_double), ("last_cost", C.c_double), ("final_lambda", C.c_double),
                ("t_transform", C.c_double), ("t_search", C.c_double), ("t_inner", C.c_double),
                ("t_acc", C.c_double), ("t_cost", C.c_double)]


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        build()
    L = C.CDLL(_LIB_PATH)
    fp, ip, dp = C.POINTER(C.c_float), C.POINTER(C.c_int), C.POINTER(C.c_double)
    L.orc_icp_create.restype = C.c_void_p
    L.orc_icp_destroy.argtypes = [C.c_void_p]
    L.orc_icp_set_options.argtypes = [C.c_void_p, C.c_int, C.c_int]
    L.orc_icp_set_query_stride.argtypes = [C.c_void_p, C.c_size_t]
    L.orc_icp_add_cloud.argtypes = [C.c_void_p, fp, fp, C.c_size_t, fp, C.c_int]
    L.orc_icp_run.argtypes = [C.c_void_p, C.c_float, C.c_int, C.c_int, C.c_float, ip]
    L.orc_icp_get_pose.argtypes = [C.c_void_p, C.c_int, fp]
    L.orc_icp_set_pose.argtypes = [C.c_void_p, C.c_int, fp]
    L.orc_icp_last_stats.argtypes = [C.c_void_p, C.POINTER(IcpStats)]
    L.orc_icp_last_tries.argtypes = [C.c_void_p, ip, C.c_int]
    L.orc_icp_last_pair_info.argtypes = [C.c_void_p, C.c_int, ip, ip, C.POINTER(C.c_uint64)]
    L.orc_icp_last_pair_corr.argtypes = [C.c_void_p, C.c_int, ip, ip, fp]
    L.orc_icp_last_normal_eq.argtypes = [C.c_void_p, dp, dp]
    L.orc_transform_cloud.argtypes = [fp, fp, C.c_size_t, fp, fp, fp]
    L.orc_find_correspondences.argtypes = [fp, C.c_size_t, fp, C.c_size_t, C.c_float, C.c_int, ip, ip, fp]
    L.orc_find_correspondences.restype = C.c_uint64
    L.orc_time_search.argtypes = [fp, C.c_size_t, fp, C.c_size_t, C.c_float, dp, dp, C.POINTER(C.c_uint64)]
    L.orc_se3_exp_left_mul.argtypes = [dp, fp, fp, fp, fp]
    L.orc_ldlt_solve_upper.argtypes = [dp, C.c_int, dp, dp]
    L.orc_normals_knn.argtypes = [fp, C.c_size_t, C.c_int, fp, fp, ip]
    _lib = L
    return L


def _f(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _i(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


def _d(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _c32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def mat_to_colmajor(T):
    """4x4 numpy (row, col) -> 16 floats column-major (Eigen::Affine3f storage)."""
    return np.ascontiguousarray(np.asarray(T, dtype=np.float32).T.reshape(16))


def colmajor_to_mat(v):
    return np.asarray(v, dtype=np.float32).reshape(4, 4).T.copy()


class PointToPlaneICP:
    """Mirror of icp::PointToPlaneICP (icp_point_to_plane.h:39-57) on the oracle."""

    def __init__(self, use_kdtree=True, inner_max_iterations=150):
        self._h = C.c_void_p(lib().orc_icp_create())
        lib().orc_icp_set_options(self._h, int(use_kdtree), int(inner_max_iterations))

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_icp_destroy(self._h)
            self._h = None

    def set_query_stride(self, stride):
        lib().orc_icp_set_query_stride(self._h, int(stride))

    def AddPointCloud(self, xyz, normals, global_T_cloud, fixed=False):
        xyz, normals = _c32(xyz), _c32(normals)
        T = mat_to_colmajor(global_T_cloud)
        return lib().orc_icp_add_cloud(self._h, _f(xyz), _f(normals), xyz.shape[0], _f(T), int(fixed))

    def Run(self, max_correspondence_distance, initial_iteration, max_num_iterations, convergence_threshold, print_progress=False):
        conv = C.c_int(0)
        rc = lib().orc_icp_run(self._h, max_correspondence_distance, initial_iteration, max_num_iterations,
                               convergence_threshold, C.byref(conv))
        if rc != 0:
            raise RuntimeError("oracle icp run failed (no clouds)")
        return bool(conv.value)

    def GetResultGlobalTCloud(self, cloud_index):
        T = np.zeros(16, np.float32)
        if lib().orc_icp_get_pose(self._h, cloud_index, _f(T)) != 0:
            raise IndexError(cloud_index)
        return colmajor_to_mat(T)

    def SetGlobalTCloud(self, cloud_index, T):
        v = mat_to_colmajor(T)
        if lib().orc_icp_set_pose(self._h, cloud_index, _f(v)) != 0:
            raise IndexError(cloud_index)

    def stats(self):
        s = IcpStats()
        lib().orc_icp_last_stats(self._h, C.byref(s))
        return {k: getattr(s, k) for k, _ in IcpStats._fields_}

    def tries(self):
        buf = np.zeros(256, np.int32)
        n = lib().orc_icp_last_tries(self._h, _i(buf), 256)
        return buf[:n].copy()

    def pairs(self):
        """[(src_impl_index, tgt_impl_index, q, m, d2)] of the last AlignMeshes, in ik order."""
        out = []
        n = self.stats()["num_pairs"]
        for k in range(n):
            s, t, c = C.c_int(), C.c_int(), C.c_uint64()
            lib().orc_icp_last_pair_info(self._h, k, C.byref(s), C.byref(t), C.byref(c))
            q = np.zeros(c.value, np.int32); m = np.zeros(c.value, np.int32); d2 = np.zeros(c.value, np.float32)
            lib().orc_icp_last_pair_corr(self._h, k, _i(q), _i(m), _f(d2))
            out.append((s.value, t.value, q, m, d2))
        return out

    def normal_equations(self):
        nv = self.stats()["num_variables"]
        H = np.zeros((nv, nv), np.float64, order="F"); b = np.zeros(nv, np.float64)
        lib().orc_icp_last_normal_eq(self._h, _d(H), _d(b))
        return np.asarray(H), b


def transform_cloud(xyz, normals, T):
    xyz, normals = _c32(xyz), _c32(normals)
    oxyz, onrm = np.empty_like(xyz), np.empty_like(normals)
    v = mat_to_colmajor(T)
    lib().orc_transform_cloud(_f(xyz), _f(normals), xyz.shape[0], _f(v), _f(oxyz), _f(onrm))
    return oxyz, onrm


def find_correspondences(src_xyz, tgt_xyz, max_dist, use_kdtree=True):
    src_xyz, tgt_xyz = _c32(src_xyz), _c32(tgt_xyz)
    n = src_xyz.shape[0]
    q = np.zeros(n, np.int32); m = np.zeros(n, np.int32); d2 = np.zeros(n, np.float32)
    c = lib().orc_find_correspondences(_f(src_xyz), n, _f(tgt_xyz), tgt_xyz.shape[0], max_dist, int(use_kdtree), _i(q), _i(m), _f(d2))
    return q[:c].copy(), m[:c].copy(), d2[:c].copy()


def time_search(src_xyz, tgt_xyz, max_dist):
    """(t_build, t_query, matched): single-thread kd-tree build on tgt + nearest-within-radius for every src point."""
    src_xyz, tgt_xyz = _c32(src_xyz), _c32(tgt_xyz)
    tb, tq, m = C.c_double(), C.c_double(), C.c_uint64()
    lib().orc_time_search(_f(src_xyz), src_xyz.shape[0], _f(tgt_xyz), tgt_xyz.shape[0], max_dist, C.byref(tb), C.byref(tq), C.byref(m))
    return tb.value, tq.value, m.value


def se3_exp_left_mul(x, q, t):
    x = np.ascontiguousarray(x, np.float64); q = _c32(q); t = _c32(t)
    qo = np.zeros(4, np.float32); to = np.zeros(3, np.float32)
    lib().orc_se3_exp_left_mul(_d(x), _f(q), _f(t), _f(qo), _f(to))
    return qo, to


def ldlt_solve_upper(A, b):
    A = np.asfortranarray(A, np.float64); b = np.ascontiguousarray(b, np.float64)
    x = np.zeros_like(b)
    lib().orc_ldlt_solve_upper(_d(A), A.shape[0], _d(b), _d(x))
    return x


def normals_knn(xyz, k, viewpoint=(0.0, 0.0, 0.0), return_indices=False):
    xyz = _c32(xyz)
    n = xyz.shape[0]
    out = np.zeros((n, 4), np.float32)
    vp = np.asarray(viewpoint, np.float32)
    idx = np.zeros((n, k), np.int32) if return_indices else None
    lib().orc_normals_knn(_f(xyz), n, k, _f(vp), _f(out), _i(idx) if return_indices else None)
    return (out, idx) if return_indices else out


def normals_radius(xyz, radius, viewpoint=(0.0, 0.0, 0.0), return_counts=False):
    """setRadiusSearch mode: all neighbours within `radius` (strict), sorted by (distance, index)."""
    x = _c32(xyz); vp = _c32(viewpoint)
    n = x.shape[0]
    out = np.zeros((n, 4), np.float32); cnt = np.zeros(n, np.int32)
    L = lib()
    L.orc_normals_radius.argtypes = [C.POINTER(C.c_float), C.c_size_t, C.c_float, C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_int)]
    L.orc_normals_radius(_f(x), n, float(radius), _f(vp), _f(out), cnt.ctypes.data_as(C.POINTER(C.c_int)))
    return (out, cnt) if return_counts else out


# ---------------------------------------------------------------------------------------------------------------------
# Point-cloud tools (orc_cleaner.cc): LocalStatisticalOutlierRemoval, SplatCreator
# ---------------------------------------------------------------------------------------------------------------------
def lsor_filter(xyz, mean_k, factor, negative=False):
    """Returns (kept indices, removed indices, pass-1 mean distances)."""
    x = _c32(xyz); n = x.shape[0]
    keep = np.zeros(n, np.int32); rem = np.zeros(n, np.int32); dist = np.zeros(n, np.float32)
    nk, nr = C.c_uint64(0), C.c_uint64(0)
    L = lib()
    L.orc_lsor_filter.argtypes = [C.POINTER(C.c_float), C.c_size_t, C.c_int, C.c_double, C.c_int, C.POINTER(C.c_int32), C.POINTER(C.c_uint64),
                                  C.POINTER(C.c_int32), C.POINTER(C.c_uint64), C.POINTER(C.c_float)]
    L.orc_lsor_filter.restype = C.c_int
    rc = L.orc_lsor_filter(_f(x), n, int(mean_k), float(factor), int(bool(negative)), keep.ctypes.data_as(C.POINTER(C.c_int32)), C.byref(nk),
                           rem.ctypes.data_as(C.POINTER(C.c_int32)), C.byref(nr), _f(dist))
    if rc != 0:
        raise ValueError("fewer than mean_k + 1 finite points")
    return keep[:nk.value].copy(), rem[:nr.value].copy(), dist


def mesh_squared_distance(points, vertices, faces):
    p = _c32(points); v = _c32(vertices); f = np.ascontiguousarray(faces, np.uint32)
    out = np.zeros(p.shape[0], np.float32)
    L = lib()
    L.orc_mesh_squared_distance.argtypes = [C.POINTER(C.c_float), C.c_size_t, C.POINTER(C.c_float), C.c_void_p, C.c_size_t, C.POINTER(C.c_float)]
    L.orc_mesh_squared_distance.restype = None
    L.orc_mesh_squared_distance(_f(p), p.shape[0], _f(v), f.ctypes.data, f.shape[0], _f(out))
    return out


def create_splats(xyz, normals, vertices, faces, distance_threshold=0.02, max_splat_size=np.inf):
    x = _c32(xyz); nr = _c32(normals); v = _c32(vertices); f = np.ascontiguousarray(faces, np.uint32)
    n = x.shape[0]
    corners = np.zeros((n, 4, 3), np.float32); added = np.zeros(n, np.uint8); radius = np.zeros(n, np.float32)
    thr2 = np.float32(distance_threshold) * np.float32(distance_threshold)
    L = lib()
    L.orc_splat_create.argtypes = [C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_size_t, C.POINTER(C.c_float), C.c_void_p, C.c_size_t, C.c_float,
                                   C.c_float, C.POINTER(C.c_float), C.c_void_p, C.POINTER(C.c_float)]
    L.orc_splat_create.restype = C.c_uint64
    cnt = L.orc_splat_create(_f(x), _f(nr), n, _f(v), f.ctypes.data, f.shape[0], float(max_splat_size), float(thr2), _f(corners), added.ctypes.data,
                             _f(radius))
    assert cnt == int(added.sum())
    return corners, added.astype(bool), radius


# ---------------------------------------------------------------------------------------------------------------------
# Multi-resolution point cloud (orc_multiscale.cc)
# ---------------------------------------------------------------------------------------------------------------------
def _u8(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint8))


def ms_merge_close_points(xyz, colors, scan_indices, max_radius, num_scans, merge_distance):
    """MergeClosePoints (multi_scale_point_cloud.cc:44-124) -> (xyz, colors, scan_indices, max_radius) of the merged cloud."""
    x = _c32(xyz); c = _c32(colors); m = _c32(max_radius); s = np.ascontiguousarray(scan_indices, np.uint8)
    n = x.shape[0]
    ox = np.zeros((n, 3), np.float32); oc = np.zeros(n, np.float32); om = np.zeros(n, np.float32); os_ = np.zeros(n, np.uint8)
    L = lib()
    L.orc_ms_merge_close_points.restype = C.c_uint64
    L.orc_ms_merge_close_points.argtypes = [C.POINTER(C.c_float), C.c_size_t, C.POINTER(C.c_float), C.POINTER(C.c_uint8), C.POINTER(C.c_float), C.c_int,
                                            C.c_float, C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_uint8), C.POINTER(C.c_float)]
    k = int(L.orc_ms_merge_close_points(_f(x), n, _f(c), _u8(s), _f(m), int(num_scans), float(merge_distance), _f(ox), _f(oc), _u8(os_), _f(om)))
    return ox[:k].copy(), oc[:k].copy(), os_[:k].copy(), om[:k].copy()


def ms_create(xyz, colors, scan_indices, min_radius, max_radius, num_scans, min_radius_bias=1.05, merge_distance_factor=4.0, max_scales=32):
    """CreateMultiScalePointCloud's scale loop (multi_scale_point_cloud.cc:263-368) -> list of (radius, xyz, colors, scan_indices)."""
    x = _c32(xyz); c = _c32(colors); lo = _c32(min_radius); hi = _c32(max_radius); s = np.ascontiguousarray(scan_indices, np.uint8)
    n = x.shape[0]
    cap = max(1, n) * max_scales
    rad = np.zeros(max_scales, np.float32); cnt = np.zeros(max_scales, np.uint64)
    ox = np.zeros((cap, 3), np.float32); oc = np.zeros(cap, np.float32); os_ = np.zeros(cap, np.uint8)
    L = lib()
    L.orc_ms_create.argtypes = [C.POINTER(C.c_float), C.c_size_t, C.POINTER(C.c_float), C.POINTER(C.c_uint8), C.POINTER(C.c_float), C.POINTER(C.c_float),
                                C.c_int, C.c_float, C.c_float, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_uint64), C.POINTER(C.c_float),
                                C.POINTER(C.c_float), C.POINTER(C.c_uint8)]
    k = L.orc_ms_create(_f(x), n, _f(c), _u8(s), _f(lo), _f(hi), int(num_scans), float(min_radius_bias), float(merge_distance_factor), int(max_scales),
                        _f(rad), cnt.ctypes.data_as(C.POINTER(C.c_uint64)), _f(ox), _f(oc), _u8(os_))
    if k < 0:
        raise RuntimeError("orc_ms_create: more than %d scales" % max_scales)
    out, off = [], 0
    for i in range(k):
        m = int(cnt[i])
        out.append((float(rad[i]), ox[off:off + m].copy(), oc[off:off + m].copy(), os_[off:off + m].copy()))
        off += m
    return out


def ms_point_neighbors(xyz, scan_indices, scan_count, limit_to_same_scan, candidate_count=25, neighbor_count=5):
    """Problem::DeterminePointNeighbors (problem.cc:706-786) -> (n, neighbor_count) uint64."""
    x = _c32(xyz); s = np.ascontiguousarray(scan_indices, np.uint8)
    n = x.shape[0]
    out = np.zeros((n, neighbor_count), np.uint64)
    L = lib()
    L.orc_ms_point_neighbors.argtypes = [C.POINTER(C.c_float), C.c_size_t, C.POINTER(C.c_uint8), C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_uint64)]
    rc = L.orc_ms_point_neighbors(_f(x), n, _u8(s), int(scan_count), int(bool(limit_to_same_scan)), int(candidate_count), int(neighbor_count),
                                  out.ctypes.data_as(C.POINTER(C.c_uint64)))
    if rc != 0:
        raise RuntimeError("orc_ms_point_neighbors failed (%d): too few points per scan or no self-match" % rc)
    return out


def ms_compute_multi_res_point_cloud(reg, scans, image_scale_count, fixed_residuals_weight=1.0, K=5, candidates=25, min_diff=5, bias=1.05, factor=4.0):
    """Problem::ComputeMultiResPointCloud (problem.cc:161-362) restated on the oracle's pieces; `reg` = oracle Registration (initialised).
    scans: [(xyz float32 (n,3), rgb uint8 (n,3))]. Loops are written point by point where the reference's order matters."""
    use_fixed = fixed_residuals_weight > 0
    ns = len(scans)
    pts = np.concatenate([_c32(x) for x, _ in scans])
    cols = np.concatenate([np.array([np.float32(0.299 * float(r) + 0.587 * float(g) + 0.114 * float(b)) for r, g, b in c], np.float32) for _, c in scans])
    sidx = np.concatenate([np.full(len(x), i, np.uint8) for i, (x, _) in enumerate(scans)])
    lo, hi = reg.min_max_point_radius(pts, float(np.float32(2.0 ** (-(image_scale_count - 1)))))
    scales = [list(t) for t in ms_create(pts, cols, sidx, lo, hi, ns, bias, factor)]

    def enough(si):
        if use_fixed:
            return all(int((si == k).sum()) >= candidates + 1 for k in range(ns))
        return len(si) >= candidates + 1

    scales = [s for s in scales if enough(s[3])]
    for s in scales:
        _, p, c, si = s
        nb = ms_point_neighbors(p, si, ns, use_fixed, candidates, K)
        n = len(c)
        delete1 = np.zeros(n, bool)
        for i in range(n):
            acc = np.float32(0)
            for k in range(K):
                acc = np.float32(acc + np.float32(abs(np.float32(c[int(nb[i, k])] - c[i]))))
            delete1[i] = np.float32(acc / np.float32(K)) < min_diff
        delete2 = np.ones(n, bool)
        for i in range(n):
            if not delete1[i]:
                delete2[i] = False
                for k in range(K):
                    delete2[int(nb[i, k])] = False
        s[1], s[2], s[3] = p[~delete2], c[~delete2], si[~delete2]
    scales = [s for s in scales if enough(s[3])]
    nbrs = [ms_point_neighbors(s[1], s[3], ns, use_fixed, candidates, K) for s in scales]
    return [s[0] for s in scales], [s[1] for s in scales], [s[2] for s in scales], [s[3] for s in scales], nbrs


# ---------------------------------------------------------------------------------------------------------------------
# Path B (orc_reg.cc)
# ---------------------------------------------------------------------------------------------------------------------
class RegParams(C.Structure):
    _fields_ = [("point_neighbor_count", C.c_int32), ("fixed_residuals_weight", C.c_float), ("variable_residuals_weight", C.c_float),
                ("robust_weighting_type", C.c_int32), ("robust_weighting_parameter", C.c_float),
                ("maximum_valid_intensity", C.c_float), ("occlusion_depth_threshold", C.c_float),
                ("min_occlusion_check_image_scale", C.c_int32), ("max_initial_image_area_in_pixels", C.c_int32),
                ("splat_radius", C.c_float), ("image_scale_count_override", C.c_int32),
                ("min_occlusion_depth", C.c_float), ("max_occlusion_depth", C.c_float), ("mask_occlusion_boundaries", C.c_int32)]


_reg_bound = False


def _bind_reg():
    global _reg_bound
    L = lib()
    if _reg_bound:
        return L
    fp, ip, dp, vp = C.POINTER(C.c_float), C.POINTER(C.c_int), C.POINTER(C.c_double), C.c_void_p
    u8, u64p = C.POINTER(C.c_uint8), C.POINTER(C.c_uint64)
    L.orc_reg_default_params.argtypes = [C.POINTER(RegParams)]
    L.orc_reg_create.argtypes = [C.POINTER(RegParams)]; L.orc_reg_create.restype = vp
    L.orc_reg_destroy.argtypes = [vp]
    L.orc_reg_add_intrinsics.argtypes = [vp, C.c_int, C.c_int, fp]
    L.orc_reg_add_intrinsics_model.argtypes = [vp, C.c_int, C.c_int, C.c_int, fp]
    L.orc_cam_param_count.argtypes = [C.c_int]
    L.orc_cam_cutoff.argtypes = [C.c_int, C.c_int, C.c_int, fp, fp]
    L.orc_cam_eval.argtypes = [C.c_int, C.c_int, C.c_int, fp, C.c_int, fp, C.c_size_t, fp]
    L.orc_reg_add_image.argtypes = [vp, C.c_int, u8, u8, fp]
    L.orc_reg_initialize.argtypes = [vp]
    L.orc_reg_set_camera_mask.argtypes = [vp, C.c_int, u8]
    L.orc_reg_add_rig.argtypes = [vp, C.c_int, fp]
    L.orc_reg_add_rig_images.argtypes = [vp, C.c_int, ip]
    L.orc_reg_get_rigs.argtypes = [vp, fp]
    L.orc_reg_set_rigs.argtypes = [vp, fp]
    L.orc_reg_variable_index.argtypes = [vp, C.c_int, C.c_int]
    L.orc_reg_point_jacobians_rig.argtypes = [vp, C.c_int, C.c_int, C.c_uint64, fp, fp, fp, fp]
    L.orc_reg_add_point_scale.argtypes = [vp, fp, C.c_size_t, C.c_float, u64p, fp]
    L.orc_reg_set_splat_points.argtypes = [vp, fp, C.c_size_t]
    L.orc_reg_set_mesh.argtypes = [vp, fp, C.c_size_t, C.POINTER(C.c_uint32), C.c_size_t]
    L.orc_reg_mesh_edges.argtypes = [vp] + [C.POINTER(C.c_uint32)] * 4 + [u8]
    L.orc_reg_mesh_edges.restype = C.c_uint64
    L.orc_reg_set_depth_map.argtypes = [vp, C.c_int, C.c_int, C.c_int, fp]
    L.orc_reg_set_image_scale.argtypes = [vp, C.c_int]
    L.orc_reg_image_scale_count.argtypes = [vp]
    L.orc_reg_num_variables.argtypes = [vp]
    L.orc_reg_render_depth.argtypes = [vp, C.c_int, ip, ip, fp]
    L.orc_reg_create_observations.argtypes = [vp, C.c_int]
    L.orc_reg_num_observations.argtypes = [vp, C.c_int, C.c_int]; L.orc_reg_num_observations.restype = C.c_uint64
    L.orc_reg_get_observations.argtypes = [vp, C.c_int, C.c_int, u64p, fp, fp, fp, u8]
    L.orc_reg_color_update.argtypes = [vp]
    L.orc_reg_get_descriptors.argtypes = [vp, C.c_int, fp, fp, ip]
    L.orc_reg_cost.argtypes = [vp, dp]; L.orc_reg_cost.restype = C.c_double
    L.orc_reg_accumulate.argtypes = [vp, dp, dp, dp]; L.orc_reg_accumulate.restype = C.c_double
    L.orc_reg_get_state.argtypes = [vp, fp, fp]
    L.orc_reg_set_state.argtypes = [vp, fp, fp]
    L.orc_reg_cost_for_delta.argtypes = [vp, dp]; L.orc_reg_cost_for_delta.restype = C.c_double
    L.orc_reg_apply.argtypes = [vp, fp, fp, ip]
    L.orc_reg_run_on_current_scale.argtypes = [vp, C.c_int, C.c_float, C.c_int, dp, ip]
    L.orc_reg_point_jacobians.argtypes = [vp, C.c_int, C.c_int, C.c_uint64, fp, fp, fp]
    L.orc_interp_bilinear.argtypes = [u8, C.c_int, C.c_int, C.c_float, C.c_float, fp, fp, fp]
    L.orc_interp_trilinear.argtypes = [u8, C.c_int, C.c_int, u8, C.c_float, C.c_float, C.c_float, fp, fp, fp, fp]
    L.orc_robust.argtypes = [C.c_int, C.c_float, C.c_float, C.c_int]; L.orc_robust.restype = C.c_float
    L.orc_image_pyramid_level.argtypes = [u8, C.c_int, C.c_int, u8]
    _reg_bound = True
    return L


def _u8(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint8))


CAM_FOV, CAM_POLYNOMIAL, CAM_POLYNOMIAL_TANGENTIAL, CAM_FISHEYE_POLYNOMIAL_TANGENTIAL, CAM_PINHOLE, CAM_BENCHMARK, CAM_FISHEYE_POLYNOMIAL_4 = 0, 1, 2, 3, 4, 5, 6
CAM_SIMPLE_PINHOLE, CAM_RADIAL, CAM_SIMPLE_RADIAL, CAM_FULL_OPENCV, CAM_POLYNOMIAL_4, CAM_RADIAL_FISHEYE, CAM_SIMPLE_RADIAL_FISHEYE, CAM_THIN_PRISM = 7, 8, 9, 10, 11, 12, 13, 14
# (camera::CameraBase::Type values, camera_base.h:67-84)


def cam_param_count(model):
    n = _bind_reg().orc_cam_param_count(model)
    if n < 0:
        raise ValueError("unsupported camera model %d" % model)
    return n


def cam_cutoff(model, w, h, params):
    """(radius_cutoff_squared of the camera, of a fisheye camera's inner model) after construction."""
    p = _c32(params); out = np.zeros(2, np.float32)
    if _bind_reg().orc_cam_cutoff(model, w, h, _f(p), _f(out)) != 0:
        raise ValueError("unsupported camera model")
    return float(out[0]), float(out[1])


_CAM_OPS = {"distort": (0, 2, 2), "project": (1, 2, 2), "d_by_world": (2, 3, 6), "d_by_intrinsics": (3, 3, None), "undistort": (4, 2, 2),
            "distort_deriv": (5, 2, 4)}


def cam_eval(model, w, h, params, op, pts):
    code, nin, nout = _CAM_OPS[op]
    p = _c32(params); x = _c32(pts).reshape(-1, nin)
    if nout is None:
        nout = 2 * cam_param_count(model)
    out = np.zeros((x.shape[0], nout), np.float32)
    if _bind_reg().orc_cam_eval(model, w, h, _f(p), code, _f(x), x.shape[0], _f(out)) != 0:
        raise ValueError("orc_cam_eval failed")
    return out


def reg_default_params(**kw):
    p = RegParams()
    _bind_reg().orc_reg_default_params(C.byref(p))
    for k, v in kw.items():
        setattr(p, k, v)
    return p


class Registration:
    """Oracle of the Path B problem + optimizer pieces (opt::Problem / VisibilityEstimator / IntrinsicsAndPoseOptimizer /
    CostCalculator / ColorOptimizer / Optimizer), pinhole cameras."""

    def __init__(self, params=None):
        L = _bind_reg()
        self.params = params or reg_default_params()
        self._h = C.c_void_p(L.orc_reg_create(C.byref(self.params)))
        self.K = self.params.point_neighbor_count
        self.n_intr = 0; self.n_img = 0; self.scale_sizes = []; self.intr_np = []; self.rig_cams = []

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_reg_destroy(self._h); self._h = None

    def add_intrinsics(self, w, h, params, camera_model=CAM_PINHOLE):
        model = camera_model
        p = _c32(params)
        if p.size != cam_param_count(model):
            raise ValueError("camera model %d takes %d parameters" % (model, cam_param_count(model)))
        self.n_intr += 1; self.intr_np.append(int(p.size))
        return lib().orc_reg_add_intrinsics_model(self._h, model, w, h, _f(p))

    def set_camera_mask(self, intr_id, mask):
        m = np.ascontiguousarray(mask, np.uint8)
        if lib().orc_reg_set_camera_mask(self._h, intr_id, _u8(m)) != 0:
            raise ValueError("bad intrinsics id")

    def add_rig(self, image_T_rig):
        """image_T_rig: (num_cameras, 7) qx qy qz qw tx ty tz, camera 0 = reference (identity)."""
        T = _c32(image_T_rig).reshape(-1, 7)
        rid = lib().orc_reg_add_rig(self._h, T.shape[0], _f(T))
        if rid < 0:
            raise ValueError("a rig needs at least two cameras")
        self.rig_cams.append(T.shape[0])
        return rid

    def add_rig_images(self, rig_id, image_ids):
        ids = np.ascontiguousarray(image_ids, np.int32)
        r = lib().orc_reg_add_rig_images(self._h, rig_id, ids.ctypes.data_as(C.POINTER(C.c_int)))
        if r < 0:
            raise ValueError("bad rig image set")
        return r

    def get_rigs(self):
        out = np.zeros((sum(self.rig_cams), 7), np.float32)
        lib().orc_reg_get_rigs(self._h, _f(out))
        return out

    def set_rigs(self, image_T_rig_all):
        T = _c32(image_T_rig_all).reshape(-1, 7)
        assert T.shape[0] == sum(self.rig_cams)
        lib().orc_reg_set_rigs(self._h, _f(T))

    def variable_index(self, kind, idx):
        """kind: "intrinsics" | "rig" | "image" -> first variable of that block."""
        return lib().orc_reg_variable_index(self._h, {"intrinsics": 0, "rig": 1, "image": 2}[kind], idx)

    def _intr_shape(self, flat):
        return flat.reshape(self.n_intr, -1) if len(set(self.intr_np)) == 1 else flat

    def add_image(self, intr_id, gray, mask, image_T_global):
        g = np.ascontiguousarray(gray, np.uint8); T = _c32(image_T_global)
        m = np.ascontiguousarray(mask, np.uint8) if mask is not None else None
        self.n_img += 1
        return lib().orc_reg_add_image(self._h, intr_id, _u8(g), _u8(m) if m is not None else None, _f(T))

    def initialize(self):
        r = lib().orc_reg_initialize(self._h)
        if r < 0:
            raise ValueError("odd pyramid parent size")
        return r

    def add_point_scale(self, xyz, radius, neighbors, colors):
        xyz = _c32(xyz); nb = np.ascontiguousarray(neighbors, np.uint64); col = _c32(colors)
        self.scale_sizes.append(xyz.shape[0])
        return lib().orc_reg_add_point_scale(self._h, _f(xyz), xyz.shape[0], radius, nb.ctypes.data_as(C.POINTER(C.c_uint64)), _f(col))

    def set_splat_points(self, xyz):
        xyz = _c32(xyz); lib().orc_reg_set_splat_points(self._h, _f(xyz), xyz.shape[0])

    def set_mesh(self, vertices, faces):
        v = _c32(vertices); f = np.ascontiguousarray(faces, np.uint32)
        lib().orc_reg_set_mesh(self._h, _f(v), v.shape[0], f.ctypes.data_as(C.POINTER(C.c_uint32)), f.shape[0])

    def mesh_edges(self):
        n = lib().orc_reg_mesh_edges(self._h, None, None, None, None, None)
        a = [np.zeros(n, np.uint32) for _ in range(4)]; fl = np.zeros(n, np.uint8)
        lib().orc_reg_mesh_edges(self._h, *[x.ctypes.data_as(C.POINTER(C.c_uint32)) for x in a], _u8(fl))
        return a[0], a[1], a[2], a[3], fl

    def set_depth_map(self, image, depth):
        d = _c32(depth); lib().orc_reg_set_depth_map(self._h, image, d.shape[1], d.shape[0], _f(d))

    def set_image_scale(self, s):
        lib().orc_reg_set_image_scale(self._h, s)

    def image_scale_count(self):
        return lib().orc_reg_image_scale_count(self._h)

    def num_variables(self):
        return lib().orc_reg_num_variables(self._h)

    def min_max_point_radius(self, points, min_scaling_factor, min_radius=None, max_radius=None):
        x = _c32(points)
        n = x.shape[0]
        lo = np.full(n, np.inf, np.float32) if min_radius is None else _c32(min_radius).copy()
        hi = np.full(n, -np.inf, np.float32) if max_radius is None else _c32(max_radius).copy()
        L = lib()
        L.orc_reg_min_max_point_radius.argtypes = [C.c_void_p, C.POINTER(C.c_float), C.c_size_t, C.c_double, C.POINTER(C.c_float), C.POINTER(C.c_float)]
        L.orc_reg_min_max_point_radius.restype = None
        L.orc_reg_min_max_point_radius(self._h, _f(x), n, float(min_scaling_factor), _f(lo), _f(hi))
        return lo, hi

    ComputeMinMaxPointRadius = min_max_point_radius

    # ---- GroundTruthCreator (ground_truth_creator.cc:44-215) ----
    def gt_accumulate_observations(self, image, points, counts):
        x = _c32(points); c = np.ascontiguousarray(counts, np.int32).copy()
        L = lib()
        L.orc_reg_gt_accumulate_observations.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_float), C.c_size_t, C.POINTER(C.c_int32)]
        L.orc_reg_gt_accumulate_observations.restype = None
        L.orc_reg_gt_accumulate_observations(self._h, int(image), _f(x), x.shape[0], c.ctypes.data_as(C.POINTER(C.c_int32)))
        return c

    def gt_create(self, image, points, rgb, counts, radius, size_wh, rendering_bgr=None):
        x = _c32(points); c = np.ascontiguousarray(counts, np.int32); col = np.ascontiguousarray(rgb, np.uint8)
        w, h = size_wh
        occ = np.zeros((h, w), np.float32); gt = np.zeros((h, w), np.float32)
        ren = np.ascontiguousarray(rendering_bgr, np.uint8).copy() if rendering_bgr is not None else None
        u8 = C.POINTER(C.c_uint8)
        L = lib()
        L.orc_reg_gt_create.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_float), u8, C.c_size_t, C.POINTER(C.c_int32), C.c_int, C.POINTER(C.c_float),
                                        C.POINTER(C.c_float), u8]
        L.orc_reg_gt_create.restype = None
        L.orc_reg_gt_create(self._h, int(image), _f(x), col.ctypes.data_as(u8), x.shape[0], c.ctypes.data_as(C.POINTER(C.c_int32)), int(radius), _f(occ), _f(gt),
                            ren.ctypes.data_as(u8) if ren is not None else None)
        return occ, gt, ren

    def render_depth(self, image):
        w, h = C.c_int(), C.c_int()
        lib().orc_reg_render_depth(self._h, image, C.byref(w), C.byref(h), None)
        out = np.zeros((h.value, w.value), np.float32)
        s = lib().orc_reg_render_depth(self._h, image, C.byref(w), C.byref(h), _f(out))
        return out, s

    def create_observations(self, border=1):
        lib().orc_reg_create_observations(self._h, border)

    def observations(self, image, ps):
        n = lib().orc_reg_num_observations(self._h, image, ps)
        idx = np.zeros(n, np.uint64); x = np.zeros(n, np.float32); y = np.zeros(n, np.float32); s = np.zeros(n, np.float32); nb = np.zeros(n, np.uint8)
        lib().orc_reg_get_observations(self._h, image, ps, idx.ctypes.data_as(C.POINTER(C.c_uint64)), _f(x), _f(y), _f(s), _u8(nb))
        return idx, x, y, s, nb

    def color_update(self):
        lib().orc_reg_color_update(self._h)

    def descriptors(self, ps):
        n = self.scale_sizes[ps]
        f = np.zeros(n * self.K, np.float32); v = np.zeros(n * self.K, np.float32); c = np.zeros(n, np.int32)
        lib().orc_reg_get_descriptors(self._h, ps, _f(f), _f(v), _i(c))
        return f, v, c

    def cost(self):
        s = np.zeros(6, np.float64)
        return lib().orc_reg_cost(self._h, _d(s)), s

    def accumulate(self):
        nv = self.num_variables()
        H = np.zeros((nv, nv), np.float64, order="F"); b = np.zeros(nv); s = np.zeros(6)
        c = lib().orc_reg_accumulate(self._h, _d(H), _d(b), _d(s))
        return np.asarray(H), b, s, c

    def get_state(self):
        ip = np.zeros(sum(self.intr_np), np.float32); po = np.zeros((self.n_img, 7), np.float32)
        lib().orc_reg_get_state(self._h, _f(ip), _f(po))
        return self._intr_shape(ip), po

    def set_state(self, intr_params, poses):
        ip = _c32(intr_params).reshape(-1); po = _c32(poses)
        assert ip.size == sum(self.intr_np)
        lib().orc_reg_set_state(self._h, _f(ip), _f(po))

    def cost_for_delta(self, delta):
        d = np.ascontiguousarray(delta, np.float64)
        return lib().orc_reg_cost_for_delta(self._h, _d(d))

    def apply(self, lam):
        l = C.c_float(lam); mc = C.c_float(0); ap = C.c_int(0)
        tries = lib().orc_reg_apply(self._h, C.byref(l), C.byref(mc), C.byref(ap))
        return bool(ap.value), l.value, mc.value, tries

    def run_on_current_scale(self, max_it, max_change_thr, no_opt_thr):
        oc = C.c_double(0); cv = C.c_int(0)
        it = lib().orc_reg_run_on_current_scale(self._h, max_it, max_change_thr, no_opt_thr, C.byref(oc), C.byref(cv))
        return it, oc.value, bool(cv.value)

    def point_jacobians_rig(self, image, ps, obs_index, np_intr=4):
        I = C.c_float(0); jk = np.zeros(np_intr, np.float32); jp = np.zeros(6, np.float32); jr = np.zeros(6, np.float32)
        lib().orc_reg_point_jacobians_rig(self._h, image, ps, obs_index, C.byref(I), _f(jk), _f(jp), _f(jr))
        return I.value, jk, jp, jr

    def point_jacobians(self, image, ps, obs_index, np_intr=4):
        I = C.c_float(0); jk = np.zeros(np_intr, np.float32); jp = np.zeros(6, np.float32)
        lib().orc_reg_point_jacobians(self._h, image, ps, obs_index, C.byref(I), _f(jk), _f(jp))
        return I.value, jk, jp


def interp_bilinear(img, x, y):
    img = np.ascontiguousarray(img, np.uint8)
    v, dx, dy = C.c_float(), C.c_float(), C.c_float()
    ok = _bind_reg().orc_interp_bilinear(_u8(img), img.shape[1], img.shape[0], x, y, C.byref(v), C.byref(dx), C.byref(dy))
    return ok, v.value, dx.value, dy.value


def interp_trilinear(img0, img1, x, y, z):
    img0 = np.ascontiguousarray(img0, np.uint8); img1 = np.ascontiguousarray(img1, np.uint8)
    v, dx, dy, dz = C.c_float(), C.c_float(), C.c_float(), C.c_float()
    _bind_reg().orc_interp_trilinear(_u8(img0), img0.shape[1], img0.shape[0], _u8(img1), x, y, z, C.byref(v), C.byref(dx), C.byref(dy), C.byref(dz))
    return v.value, dx.value, dy.value, dz.value


def robust(type_, p, r, weight=False):
    return _bind_reg().orc_robust(type_, p, r, int(weight))


def image_pyramid_level(img):
    img = np.ascontiguousarray(img, np.uint8)
    out = np.zeros((img.shape[0] // 2, img.shape[1] // 2), np.uint8)
    _bind_reg().orc_image_pyramid_level(_u8(img), img.shape[1], img.shape[0], _u8(out))
    return out
