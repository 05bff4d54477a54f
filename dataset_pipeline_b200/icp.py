"""Host-side mirror of icp::PointToPlaneICP (/root/reference/src/icp/icp_point_to_plane.h:39-57) over the C ABI.

Same method names, argument meaning and return values as the reference class:
    AddPointCloud(point_cloud, global_T_cloud, fixed) -> id (-1 for fixed clouds)
    Run(max_correspondence_distance, initial_iteration, max_num_iterations, convergence_threshold_max_movement, print_progress) -> converged
    GetResultGlobalTCloud(cloud_index) -> 4x4
The reference aborts (glog CHECK) on misuse; here that is a B2Error.
"""
import ctypes as C

import numpy as np

from . import _lib


def _f(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _i(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32))


def _colmajor(T):
    return np.ascontiguousarray(np.asarray(T, dtype=np.float32).T.reshape(16))


class Comm:
    """Library-owned NCCL communicator (b2_comm_*): one process per GPU; rank 0 creates the id and ships it to the others."""

    @staticmethod
    def unique_id():
        buf = C.create_string_buffer(128)
        _lib.check(_lib.lib().b2_comm_unique_id(buf))
        return buf.raw

    def __init__(self, rank, world_size, unique_id, device=-1):
        self._c = C.c_void_p()
        self.rank, self.world_size = rank, world_size
        _lib.check(_lib.lib().b2_comm_create(rank, world_size, unique_id, device, C.byref(self._c)))

    def allreduce_f64(self, ptr, count, stream=0):
        _lib.check(_lib.lib().b2_comm_allreduce_f64(self._c, C.c_void_p(ptr), count, C.c_void_p(stream)))

    def close(self):
        if getattr(self, "_c", None) is not None and self._c:
            _lib.lib().b2_comm_destroy(self._c); self._c = None


class PointToPlaneICP:
    def __init__(self, device=-1, inner_max_iterations=150, keep_correspondences=False, rank=0, world_size=1,
                 allreduce=None, stream=None, comm=None, index_distance_hint=0.0, shard_uploads=False, search_ahead=True):
        L = _lib.lib()
        cfg = _lib.IcpConfig()
        L.b2_icp_default_config(C.byref(cfg))
        cfg.device = device
        cfg.inner_max_iterations = inner_max_iterations
        cfg.keep_correspondences = int(keep_correspondences)
        cfg.rank, cfg.world_size = rank, world_size
        cfg.index_distance_hint = float(index_distance_hint)
        cfg.shard_uploads = int(bool(shard_uploads))
        cfg.search_ahead = int(bool(search_ahead))
        self._cb = None
        if allreduce is not None:
            # allreduce(ptr:int, count:int, stream:int) -> None ; wrapped into the C hook
            def _hook(user, buf, count, strm):
                try:
                    allreduce(int(buf), int(count), int(strm) if strm else 0)
                    return 0
                except Exception as e:  # pragma: no cover - surfaced as B2_ERR_COMM
                    print("allreduce hook failed:", e)
                    return 1
            self._cb = _lib.ALLREDUCE_FN(_hook)
            cfg.allreduce = self._cb
        if stream is not None:
            cfg.stream = C.c_void_p(int(stream))
        self._comm = comm
        if comm is not None:
            cfg.comm = comm._c
            cfg.rank, cfg.world_size = comm.rank, comm.world_size
        self._h = C.c_void_p()
        _lib.check(L.b2_icp_create(C.byref(cfg), C.byref(self._h)))

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            _lib.lib().b2_icp_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- reference API -------------------------------------------------------------------------------------
    def AddPointCloud(self, xyz, normals, global_T_cloud, fixed=False):
        """xyz, normals: (n,3) float32 host arrays (or a (n,12) float32 pcl::PointNormal-layout array via AddPointNormalArray)."""
        xyz = np.ascontiguousarray(xyz, np.float32)
        normals = np.ascontiguousarray(normals, np.float32)
        if xyz.shape != normals.shape or xyz.ndim != 2 or xyz.shape[1] != 3:
            raise ValueError("xyz and normals must both be (n,3)")
        T = _colmajor(global_T_cloud)
        out = C.c_int32(0)
        _lib.check(_lib.lib().b2_icp_add_cloud(self._h, xyz.ctypes.data, normals.ctypes.data, xyz.shape[0], 12, _f(T), int(fixed), C.byref(out)))
        return out.value

    def AddPointNormalArray(self, point_normals, global_T_cloud, fixed=False):
        """point_normals: (n,12) float32 in pcl::PointNormal layout (x y z _ nx ny nz _ curvature _ _ _), 48 B stride."""
        pn = np.ascontiguousarray(point_normals, np.float32)
        if pn.ndim != 2 or pn.shape[1] != 12:
            raise ValueError("expected (n,12) float32")
        T = _colmajor(global_T_cloud)
        out = C.c_int32(0)
        _lib.check(_lib.lib().b2_icp_add_cloud(self._h, pn.ctypes.data, pn.ctypes.data + 16, pn.shape[0], 48, _f(T), int(fixed), C.byref(out)))
        return out.value

    def AddPointCloudDevice(self, xyz_ptr, normals_ptr, n, global_T_cloud, fixed=False):
        """Packed float3 arrays already resident in HBM (raw device pointers, e.g. torch tensor .data_ptr())."""
        T = _colmajor(global_T_cloud)
        out = C.c_int32(0)
        _lib.check(_lib.lib().b2_icp_add_cloud_dev(self._h, C.c_void_p(xyz_ptr), C.c_void_p(normals_ptr), n, _f(T), int(fixed), C.byref(out)))
        return out.value

    def Run(self, max_correspondence_distance, initial_iteration, max_num_iterations, convergence_threshold_max_movement, print_progress=False):
        conv = C.c_int32(0)
        _lib.check(_lib.lib().b2_icp_run(self._h, max_correspondence_distance, initial_iteration, max_num_iterations,
                                         convergence_threshold_max_movement, int(print_progress), C.byref(conv)))
        return bool(conv.value)

    def GetResultGlobalTCloud(self, cloud_index):
        T = np.zeros(16, np.float32)
        _lib.check(_lib.lib().b2_icp_get_pose(self._h, cloud_index, _f(T)))
        return T.reshape(4, 4).T.copy()

    # ---- introspection (parity dumps; not in the reference API) ---------------------------------------------
    def SetGlobalTCloud(self, cloud_index, T):
        _lib.check(_lib.lib().b2_icp_set_pose(self._h, cloud_index, _f(_colmajor(T))))

    def set_option(self, name, value):
        """Scheduling switches for A/B measurements ("pack_overlap", "lpt_order"); results never depend on them."""
        _lib.check(_lib.lib().b2_icp_set_option(self._h, name.encode(), int(value)))

    def stats(self):
        s = _lib.IcpStats()
        _lib.check(_lib.lib().b2_icp_last_stats(self._h, C.byref(s)))
        out = {k: getattr(s, k) for k, _ in _lib.IcpStats._fields_}
        out["search_work"] = [int(v) for v in s.search_work]
        return out

    def tries(self):
        buf = np.zeros(256, np.int32)
        n = C.c_int32(0)
        _lib.check(_lib.lib().b2_icp_get_lm_tries(self._h, _i(buf), 256, C.byref(n)))
        return buf[:n.value].copy()

    def pairs(self, with_lists=True):
        out = []
        k = 0
        while True:
            s, t, c = C.c_int32(), C.c_int32(), C.c_uint64()
            if _lib.lib().b2_icp_get_pair_info(self._h, k, C.byref(s), C.byref(t), C.byref(c)) != 0:
                break
            if with_lists:
                q = np.zeros(c.value, np.int32); m = np.zeros(c.value, np.int32); d2 = np.zeros(c.value, np.float32)
                _lib.check(_lib.lib().b2_icp_get_pair_correspondences(self._h, k, _i(q), _i(m), _f(d2)))
                out.append((s.value, t.value, q, m, d2))
            else:
                out.append((s.value, t.value, c.value))
            k += 1
        return out

    def pair_correspondences(self, k):
        """(src_impl_index, tgt_impl_index, q, m, d2) of the k-th non-empty correspondence set of the last outer iteration."""
        s, t, c = C.c_int32(), C.c_int32(), C.c_uint64()
        _lib.check(_lib.lib().b2_icp_get_pair_info(self._h, k, C.byref(s), C.byref(t), C.byref(c)))
        q = np.zeros(c.value, np.int32); m = np.zeros(c.value, np.int32); d2 = np.zeros(c.value, np.float32)
        _lib.check(_lib.lib().b2_icp_get_pair_correspondences(self._h, k, _i(q), _i(m), _f(d2)))
        return s.value, t.value, q, m, d2

    def normal_equations(self):
        nv = C.c_int32(0)
        _lib.check(_lib.lib().b2_icp_get_normal_equations(self._h, None, None, None, C.byref(nv)))
        n = nv.value
        H = np.zeros((n, n), np.float64, order="F"); b = np.zeros(n, np.float64); cost = C.c_double(0)
        _lib.check(_lib.lib().b2_icp_get_normal_equations(self._h, H.ctypes.data_as(C.POINTER(C.c_double)),
                                                          b.ctypes.data_as(C.POINTER(C.c_double)), C.byref(cost), C.byref(nv)))
        return np.asarray(H), b, cost.value


def find_correspondences(src_xyz, tgt_xyz, max_correspondence_distance):
    """FindCorrespondencesFast (/root/reference/src/icp/icp_point_to_plane.cc:42-105) on the GPU."""
    src = np.ascontiguousarray(src_xyz, np.float32); tgt = np.ascontiguousarray(tgt_xyz, np.float32)
    n = src.shape[0]
    q = np.zeros(max(n, 1), np.int32); m = np.zeros(max(n, 1), np.int32); d2 = np.zeros(max(n, 1), np.float32)
    c = C.c_uint64(0)
    _lib.check(_lib.lib().b2_find_correspondences(_f(src), n, _f(tgt), tgt.shape[0], max_correspondence_distance, _i(q), _i(m), _f(d2), C.byref(c)))
    return q[:c.value].copy(), m[:c.value].copy(), d2[:c.value].copy()


def plan_directions(n_movable, has_fixed=False, world_size=1):
    """[(src_impl_index, tgt_impl_index, owner_rank)] in the reference's ik order (host-only call, no GPU needed)."""
    cap = n_movable * n_movable + 2 * n_movable + 1
    s = np.zeros(cap, np.int32); t = np.zeros(cap, np.int32); o = np.zeros(cap, np.int32)
    c = C.c_int32(0)
    _lib.check(_lib.lib().b2_icp_plan_directions(n_movable, int(has_fixed), world_size, _i(s), _i(t), _i(o), cap, C.byref(c)))
    return [(int(s[k]), int(t[k]), int(o[k])) for k in range(c.value)]


def upload_owner(cloud_id, world_size):
    """Rank that uploads movable cloud `cloud_id` under shard_uploads (host-only call, no GPU needed)."""
    return int(_lib.lib().b2_icp_upload_owner(int(cloud_id), int(world_size)))
