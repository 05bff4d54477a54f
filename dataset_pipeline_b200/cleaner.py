"""Host-side mirrors of the point-cloud tools next to the hot paths:

* pcl::LocalStatisticalOutlierRemoval (/root/reference/src/geometry/local_statistical_outlier_removal.h, .hpp:72-176) as PointCloudCleaner
  drives it (/root/reference/src/exe/point_cloud_cleaner.cc:80-96): setInputCloud / setMeanK / setDistanceFactorThresh / setNegative / filter;
* the per-point body of SplatCreator (/root/reference/src/exe/splat_creator.cc:118-215) and igl::AABB::squared_distance.
Everything runs in libeth3d_b200.so; there is no CPU fallback."""
import ctypes as C

import numpy as np

from . import _lib

_fp, _ip = C.POINTER(C.c_float), C.POINTER(C.c_int32)


def _c32(a, cols=3):
    a = np.ascontiguousarray(a, np.float32)
    if a.ndim != 2 or a.shape[1] != cols:
        raise ValueError("expected (n,%d) float32" % cols)
    return a


class LocalStatisticalOutlierRemoval:
    def __init__(self, extract_removed_indices=True):
        self._xyz = None
        self._mean_k = 1
        self._factor = 3.0            # local_statistical_outlier_removal.h: distance_factor_threshold_ default
        self._negative = False
        self._removed = np.zeros(0, np.int32)
        self.mean_distances = None

    def setInputCloud(self, xyz):
        self._xyz = _c32(xyz)

    def setMeanK(self, k):
        self._mean_k = int(k)

    def setDistanceFactorThresh(self, factor):
        self._factor = float(factor)

    def setNegative(self, negative):
        self._negative = bool(negative)

    def getRemovedIndices(self):
        return self._removed

    def filter(self):
        """Returns the ascending indices of the kept points (applyFilterIndices); the filtered cloud is xyz[indices]."""
        if self._xyz is None:
            raise _lib.B2Error(2, "setInputCloud must be called first")
        n = self._xyz.shape[0]
        keep = np.zeros(n, np.int32); removed = np.zeros(n, np.int32); dist = np.zeros(n, np.float32)
        nk, nr = C.c_size_t(0), C.c_size_t(0)
        L = _lib.lib()
        L.b2_lsor_filter.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_int, C.c_double, C.c_int, _ip, C.POINTER(C.c_size_t), _ip,
                                     C.POINTER(C.c_size_t), _fp]
        _lib.check(L.b2_lsor_filter(self._xyz.ctypes.data, n, 12, self._mean_k, self._factor, int(self._negative), keep.ctypes.data_as(_ip),
                                    C.byref(nk), removed.ctypes.data_as(_ip), C.byref(nr), dist.ctypes.data_as(_fp)))
        self._removed = removed[:nr.value].copy()
        self.mean_distances = dist
        return keep[:nk.value].copy()


def clean_point_cloud(xyz, filters):
    """PointCloudCleaner's loop (point_cloud_cleaner.cc:80-96): apply (knn, factor) filters one after the other. Returns the indices, into
    the input, of the points that survive all of them."""
    xyz = _c32(xyz)
    alive = np.arange(xyz.shape[0], dtype=np.int64)
    for knn, factor in filters:
        sor = LocalStatisticalOutlierRemoval()
        sor.setInputCloud(xyz[alive]); sor.setMeanK(knn); sor.setDistanceFactorThresh(factor)
        alive = alive[sor.filter()]
    return alive


def mesh_squared_distance(points, vertices, faces):
    p = _c32(points); v = _c32(vertices); f = np.ascontiguousarray(faces, np.uint32)
    out = np.zeros(p.shape[0], np.float32)
    L = _lib.lib()
    L.b2_mesh_squared_distance.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, _fp]
    _lib.check(L.b2_mesh_squared_distance(p.ctypes.data, p.shape[0], v.ctypes.data, v.shape[0], f.ctypes.data, f.shape[0], out.ctypes.data_as(_fp)))
    return out


def create_splats(xyz, normals, vertices, faces, distance_threshold=0.02, max_splat_size=np.inf):
    """Returns (corners (n,4,3), added (n,) bool, radius (n,)). The tool's output mesh is corners[added] with faces (2,1,0), (0,3,2) per splat;
    the threshold is squared in fp32 as in splat_creator.cc:91."""
    x = _c32(xyz); nr = _c32(normals); v = _c32(vertices); f = np.ascontiguousarray(faces, np.uint32)
    n = x.shape[0]
    corners = np.zeros((n, 4, 3), np.float32); added = np.zeros(n, np.uint8); radius = np.zeros(n, np.float32)
    thr2 = np.float32(distance_threshold) * np.float32(distance_threshold)
    cnt = C.c_size_t(0)
    L = _lib.lib()
    L.b2_splat_create.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_float,
                                  C.c_float, _fp, C.c_void_p, _fp, C.POINTER(C.c_size_t)]
    _lib.check(L.b2_splat_create(x.ctypes.data, nr.ctypes.data, n, 12, v.ctypes.data, v.shape[0], f.ctypes.data, f.shape[0], float(max_splat_size),
                                 float(thr2), corners.ctypes.data_as(_fp), added.ctypes.data, radius.ctypes.data_as(_fp), C.byref(cnt)))
    assert cnt.value == int(added.sum())
    return corners, added.astype(bool), radius
