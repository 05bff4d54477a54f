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
