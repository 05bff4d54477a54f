"""Host-side mirror of pcl::NormalEstimationTwoPassOMP (/root/reference/src/geometry/two_pass_normal_3d_omp.h:53-99) as the
tools drive it (icp_scan_aligner.cc:323-330, normal_estimator.cc:177-194): setInputCloud / setKSearch / setViewPoint / compute."""
import ctypes as C

import numpy as np

from . import _lib


class NormalEstimationTwoPassOMP:
    def __init__(self):
        self._xyz = None
        self._k = 0
        self._radius = 0.0
        self._vp = np.zeros(3, np.float32)
        self.is_dense = True

    def setInputCloud(self, xyz):
        self._xyz = np.ascontiguousarray(xyz, np.float32)
        if self._xyz.ndim != 2 or self._xyz.shape[1] != 3:
            raise ValueError("expected (n,3) float32")

    def setSearchMethod(self, tree=None):   # the spatial index is internal (implicit BVH on the GPU)
        pass

    def setKSearch(self, k):
        self._k = int(k); self._radius = 0.0

    def setRadiusSearch(self, radius):
        """All neighbours within `radius` (normal_estimator.cc:181-182) instead of the k nearest."""
        self._radius = float(radius); self._k = 0

    def setViewPoint(self, x, y, z):
        self._vp = np.array([x, y, z], np.float32)

    def compute(self, return_indices=False):
        """Returns (n,4) float32: normal_x, normal_y, normal_z, curvature (NaN rows where < 3 neighbours)."""
        if self._xyz is None or (self._k <= 0 and self._radius <= 0):
            raise _lib.B2Error(2, "setInputCloud and setKSearch / setRadiusSearch must be called first")
        n = self._xyz.shape[0]
        out = np.zeros((n, 4), np.float32)
        if self._radius > 0:
            cnt = np.zeros(n, np.int32); dense = C.c_int32(1)
            fp, ip = C.POINTER(C.c_float), C.POINTER(C.c_int32)
            L = _lib.lib()
            L.b2_normals_estimate_radius.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_float, fp, C.c_void_p, C.c_int, fp, ip, ip]
            _lib.check(L.b2_normals_estimate_radius(self._xyz.ctypes.data, n, 12, self._radius, self._vp.ctypes.data_as(fp), None, -1,
                                                    out.ctypes.data_as(fp), cnt.ctypes.data_as(ip), C.byref(dense)))
            self.is_dense = bool(dense.value)
            return (out, cnt) if return_indices else out
        idx = np.zeros((n, self._k), np.int32) if return_indices else None
        dense = C.c_int32(1)
        fp, ip = C.POINTER(C.c_float), C.POINTER(C.c_int32)
        _lib.check(_lib.lib().b2_normals_estimate(self._xyz.ctypes.data, n, 12, self._k, self._vp.ctypes.data_as(fp),
                                                  out.ctypes.data_as(fp), idx.ctypes.data_as(ip) if return_indices else None,
                                                  C.byref(dense)))
        self.is_dense = bool(dense.value)
        return (out, idx) if return_indices else out


def estimate_normals_dist(xyz, k, viewpoint, comm, device=-1):
    """Multi-GPU variant (one process per GPU): every rank passes the whole cloud and gets the whole (n,4) result; the ranks share
    the queries and merge with one allreduce over `comm` (dataset_pipeline_b200.icp.Comm)."""
    x = np.ascontiguousarray(xyz, np.float32); vp = np.asarray(viewpoint, np.float32)
    out = np.zeros((x.shape[0], 4), np.float32); dense = C.c_int32(1)
    fp = C.POINTER(C.c_float)
    L = _lib.lib()
    L.b2_normals_estimate_dist.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_int, fp, C.c_void_p, C.c_int, fp, C.POINTER(C.c_int32)]
    _lib.check(L.b2_normals_estimate_dist(x.ctypes.data, x.shape[0], 12, int(k), vp.ctypes.data_as(fp), comm._c, device, out.ctypes.data_as(fp),
                                          C.byref(dense)))
    return out, bool(dense.value)


def estimate_normals_radius(xyz, radius, viewpoint=(0.0, 0.0, 0.0), return_counts=False):
    ne = NormalEstimationTwoPassOMP()
    ne.setInputCloud(xyz); ne.setRadiusSearch(radius); ne.setViewPoint(*viewpoint)
    return ne.compute(return_counts)


def estimate_normals(xyz, k, viewpoint=(0.0, 0.0, 0.0), return_indices=False):
    ne = NormalEstimationTwoPassOMP()
    ne.setInputCloud(xyz); ne.setKSearch(k); ne.setViewPoint(*viewpoint)
    return ne.compute(return_indices)
