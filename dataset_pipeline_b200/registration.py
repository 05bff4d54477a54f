"""Host-side mirror of the Path B seams (opt::Problem + opt::Optimizer and the pieces it drives,
/root/reference/src/opt/optimizer.h:36-57, visibility_estimator.h:46-48, intrinsics_and_pose_optimizer.h:48-51,
cost_calculator.cc:44-47, color_optimizer.cc:40-43) over the C ABI (b2_reg_*). Method names follow the reference."""
import ctypes as C

import numpy as np

from . import _lib


class RegParams(C.Structure):
    _fields_ = [("point_neighbor_count", C.c_int32), ("fixed_residuals_weight", C.c_float), ("variable_residuals_weight", C.c_float),
                ("robust_weighting_type", C.c_int32), ("robust_weighting_parameter", C.c_float),
                ("maximum_valid_intensity", C.c_float), ("occlusion_depth_threshold", C.c_float),
                ("min_occlusion_check_image_scale", C.c_int32), ("max_initial_image_area_in_pixels", C.c_int32),
                ("splat_radius", C.c_float), ("image_scale_count_override", C.c_int32), ("device", C.c_int32),
                ("min_occlusion_depth", C.c_float), ("max_occlusion_depth", C.c_float), ("mask_occlusion_boundaries", C.c_int32)]


class RegStats(C.Structure):
    _fields_ = [("observations", C.c_uint64), ("residual_evaluations", C.c_uint64), ("kernel_launches", C.c_int32),
                ("ms_last_call", C.c_float), ("ms_jacobian_kernel", C.c_float), ("ms_accumulate_kernel", C.c_float)]


_bound = False


def _L():
    global _bound
    L = _lib.lib()
    if _bound:
        return L
    fp, ip, dp, vp = C.POINTER(C.c_float), C.POINTER(C.c_int32), C.POINTER(C.c_double), C.c_void_p
    u8, u64p = C.POINTER(C.c_uint8), C.POINTER(C.c_uint64)
    L.b2_reg_default_params.argtypes = [C.POINTER(RegParams)]; L.b2_reg_default_params.restype = None
    L.b2_reg_create.argtypes = [C.POINTER(RegParams), C.POINTER(vp)]
    L.b2_reg_destroy.argtypes = [vp]
    L.b2_reg_add_intrinsics.argtypes = [vp, C.c_int, C.c_int, C.c_int, fp, C.c_int, ip]
    L.b2_reg_add_image.argtypes = [vp, C.c_int, u8, u8, fp, ip]
    L.b2_reg_initialize.argtypes = [vp, ip]
    L.b2_reg_add_point_scale.argtypes = [vp, fp, C.c_size_t, C.c_float, u64p, fp, ip]
    L.b2_reg_set_splat_points.argtypes = [vp, fp, C.c_size_t]
    L.b2_reg_set_mesh.argtypes = [vp, fp, C.c_size_t, C.POINTER(C.c_uint32), C.c_size_t]
    L.b2_reg_set_depth_map.argtypes = [vp, C.c_int, C.c_int, C.c_int, fp]
    L.b2_reg_set_image_scale.argtypes = [vp, C.c_int]
    L.b2_reg_num_variables.argtypes = [vp, ip]
    L.b2_reg_render_depth.argtypes = [vp, C.c_int, ip, ip, ip, fp]
    L.b2_reg_min_max_point_radius.argtypes = [vp, fp, C.c_size_t, C.c_double, fp, fp]
    L.b2_reg_gt_accumulate_observations.argtypes = [vp, C.c_int, fp, C.c_size_t, ip]
    L.b2_reg_gt_create.argtypes = [vp, C.c_int, fp, C.POINTER(C.c_uint8), C.c_size_t, ip, C.c_int, fp, fp, C.POINTER(C.c_uint8)]
    L.b2_reg_create_observations.argtypes = [vp, C.c_int]
    L.b2_reg_num_observations.argtypes = [vp, C.c_int, C.c_int, u64p]
    L.b2_reg_get_observations.argtypes = [vp, C.c_int, C.c_int, u64p, fp, fp, fp, u8]
    L.b2_reg_get_point_jacobians.argtypes = [vp, C.c_int, C.c_int, fp, fp, fp]
    L.b2_reg_color_update.argtypes = [vp]
    L.b2_reg_get_descriptors.argtypes = [vp, C.c_int, fp, fp, ip]
    L.b2_reg_cost.argtypes = [vp, dp, dp]
    L.b2_reg_accumulate.argtypes = [vp, dp, dp, dp, dp]
    L.b2_reg_get_state.argtypes = [vp, fp, fp]
    L.b2_reg_set_state.argtypes = [vp, fp, fp]
    L.b2_reg_cost_for_delta.argtypes = [vp, dp, dp]
    L.b2_reg_apply.argtypes = [vp, fp, fp, ip, ip]
    L.b2_reg_run_on_current_scale.argtypes = [vp, C.c_int, C.c_float, C.c_int, C.c_int, dp, ip, ip]
    L.b2_reg_last_stats.argtypes = [vp, C.POINTER(RegStats)]
    L.b2_reg_add_rig.argtypes = [vp, C.c_int, fp, ip]
    L.b2_reg_add_rig_images.argtypes = [vp, C.c_int, ip, ip]
    L.b2_reg_get_rigs.argtypes = [vp, fp]
    L.b2_reg_set_rigs.argtypes = [vp, fp]
    L.b2_reg_variable_index.argtypes = [vp, C.c_int, C.c_int, ip]
    L.b2_reg_get_point_jacobians_rig.argtypes = [vp, C.c_int, C.c_int, fp, fp, fp, fp]
    L.b2_reg_set_camera_mask.argtypes = [vp, C.c_int, u8]
    L.b2_reg_set_comm.argtypes = [vp, vp]
    L.b2_reg_image_owner.argtypes = [C.c_int, C.c_int]
    L.b2_camera_eval.argtypes = [C.c_int, C.c_int, C.c_int, fp, C.c_int, C.c_int, fp, C.c_size_t, fp, fp]
    _bound = True
    return L


def _f(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _d(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _u8(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint8))


CAM_FOV, CAM_POLYNOMIAL, CAM_POLYNOMIAL_TANGENTIAL, CAM_FISHEYE_POLYNOMIAL_TANGENTIAL, CAM_PINHOLE, CAM_BENCHMARK, CAM_FISHEYE_POLYNOMIAL_4 = 0, 1, 2, 3, 4, 5, 6
CAM_SIMPLE_PINHOLE, CAM_RADIAL, CAM_SIMPLE_RADIAL, CAM_FULL_OPENCV, CAM_POLYNOMIAL_4, CAM_RADIAL_FISHEYE, CAM_SIMPLE_RADIAL_FISHEYE, CAM_THIN_PRISM = 7, 8, 9, 10, 11, 12, 13, 14
# (camera::CameraBase::Type values, camera_base.h:67-84)


_CAM_OPS = {"cutoff": (0, 2, 0), "project": (1, 2, 2), "d_by_world": (2, 3, 6), "d_by_intrinsics": (3, 3, None)}


def camera_eval(camera_model, width, height, params, op="cutoff", pts=None):
    """camera::CameraBase evaluation on the device (b2_camera_eval): returns (out, (cutoff2, inner_cutoff2)).
    op: "cutoff" | "project" (n x 2 normalized -> pixels) | "d_by_world" (n x 3 -> n x 6) | "d_by_intrinsics" (n x 3 -> n x 2np)."""
    code, nin, nout = _CAM_OPS[op]
    p = np.ascontiguousarray(params, np.float32)
    x = np.ascontiguousarray(pts if pts is not None else np.zeros((0, nin)), np.float32).reshape(-1, nin)
    if nout is None:
        nout = 2 * p.size
    out = np.zeros((x.shape[0], nout), np.float32); cut = np.zeros(2, np.float32)
    _lib.check(_L().b2_camera_eval(camera_model, width, height, _f(p), p.size, code, _f(x), x.shape[0], _f(out), _f(cut)))
    return out, (float(cut[0]), float(cut[1]))


def default_params(**kw):
    p = RegParams()
    _L().b2_reg_default_params(C.byref(p))
    for k, v in kw.items():
        setattr(p, k, v)
    return p


class Registration:
    """opt::Problem state in HBM + the optimizer pieces. PINHOLE / THIN_PRISM / BENCHMARK (thin-prism fisheye) cameras, no rigs."""

    def __init__(self, params=None):
        L = _L()
        self.params = params or default_params()
        self._h = C.c_void_p()
        _lib.check(L.b2_reg_create(C.byref(self.params), C.byref(self._h)))
        self.K = self.params.point_neighbor_count
        self.n_intr = 0; self.n_img = 0; self.scale_sizes = []; self.intr_np = []; self.image_intr = []; self.rig_cams = []

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            _lib.lib().b2_reg_destroy(self._h); self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- problem set-up (opt::Problem) ----
    def add_intrinsics(self, width, height, params, camera_model=CAM_PINHOLE):
        """camera_model: CAM_PINHOLE (4 params), CAM_THIN_PRISM or CAM_BENCHMARK (12 params) — camera::CameraBase::Type values."""
        p = np.ascontiguousarray(params, np.float32); out = C.c_int32(0)
        _lib.check(_L().b2_reg_add_intrinsics(self._h, camera_model, width, height, _f(p), p.size, C.byref(out)))
        self.n_intr += 1; self.intr_np.append(int(p.size))
        self.__dict__.setdefault("_intr_sizes", []).append((int(width), int(height)))
        return out.value

    def set_camera_mask(self, intrinsics_id, mask):
        m = np.ascontiguousarray(mask, np.uint8)
        _lib.check(_L().b2_reg_set_camera_mask(self._h, intrinsics_id, _u8(m)))

    # ---- multi-GPU ----
    def set_comm(self, comm):
        """comm: dataset_pipeline_b200.icp.Comm (library-owned NCCL communicator) or None. Before add_image. Images are dealt
        round-robin (image_owner); every rank makes the same calls, and may pass gray=None for images it does not own."""
        _lib.check(_L().b2_reg_set_comm(self._h, comm._c if comm is not None else None))
        self._comm = comm
        self._rank, self._world = (comm.rank, comm.world_size) if comm is not None else (0, 1)

    def owns(self, image_id):
        return image_id % getattr(self, "_world", 1) == getattr(self, "_rank", 0)

    # ---- rigs (opt::Rig / opt::RigImages) ----
    def add_rig(self, image_T_rig):
        """image_T_rig: (num_cameras, 7) qx qy qz qw tx ty tz, camera 0 = reference (identity)."""
        T = np.ascontiguousarray(image_T_rig, np.float32).reshape(-1, 7); out = C.c_int32(0)
        _lib.check(_L().b2_reg_add_rig(self._h, T.shape[0], _f(T), C.byref(out)))
        self.rig_cams.append(T.shape[0])
        return out.value

    def add_rig_images(self, rig_id, image_ids):
        ids = np.ascontiguousarray(image_ids, np.int32); out = C.c_int32(0)
        if ids.size != self.rig_cams[rig_id]:
            raise ValueError("one image id per rig camera")
        _lib.check(_L().b2_reg_add_rig_images(self._h, rig_id, ids.ctypes.data_as(C.POINTER(C.c_int32)), C.byref(out)))
        return out.value

    def get_rigs(self):
        out = np.zeros((sum(self.rig_cams), 7), np.float32)
        if out.size:
            _lib.check(_L().b2_reg_get_rigs(self._h, _f(out)))
        return out

    def set_rigs(self, image_T_rig_all):
        T = np.ascontiguousarray(image_T_rig_all, np.float32).reshape(-1, 7)
        if T.shape[0] != sum(self.rig_cams):
            raise ValueError("expected %d rig camera poses" % sum(self.rig_cams))
        _lib.check(_L().b2_reg_set_rigs(self._h, _f(T)))

    def variable_index(self, kind, idx):
        """kind: "intrinsics" | "rig" | "image" -> first variable of that block in the optimizer's vector."""
        out = C.c_int32(0)
        _lib.check(_L().b2_reg_variable_index(self._h, {"intrinsics": 0, "rig": 1, "image": 2}[kind], idx, C.byref(out)))
        return out.value

    def _intr_shape(self, flat):
        return flat.reshape(self.n_intr, -1) if len(set(self.intr_np)) == 1 else flat

    def add_image(self, intrinsics_id, gray, mask, image_T_global):
        T = np.ascontiguousarray(image_T_global, np.float32)
        g = np.ascontiguousarray(gray, np.uint8) if gray is not None else None
        m = np.ascontiguousarray(mask, np.uint8) if mask is not None else None
        out = C.c_int32(0)
        _lib.check(_L().b2_reg_add_image(self._h, intrinsics_id, _u8(g) if g is not None else None, _u8(m) if m is not None else None, _f(T),
                                         C.byref(out)))
        self.n_img += 1; self.image_intr.append(int(intrinsics_id))
        return out.value

    def initialize(self):
        c = C.c_int32(0)
        _lib.check(_L().b2_reg_initialize(self._h, C.byref(c)))
        return c.value

    def add_point_scale(self, xyz, radius, neighbors, colors):
        xyz = np.ascontiguousarray(xyz, np.float32); nb = np.ascontiguousarray(neighbors, np.uint64); col = np.ascontiguousarray(colors, np.float32)
        out = C.c_int32(0)
        _lib.check(_L().b2_reg_add_point_scale(self._h, _f(xyz), xyz.shape[0], radius, nb.ctypes.data_as(C.POINTER(C.c_uint64)), _f(col), C.byref(out)))
        self.scale_sizes.append(xyz.shape[0])
        return out.value

    def set_splat_points(self, xyz):
        xyz = np.ascontiguousarray(xyz, np.float32)
        _lib.check(_L().b2_reg_set_splat_points(self._h, _f(xyz), xyz.shape[0]))

    def set_mesh(self, vertices, faces):
        """OcclusionGeometry::AddMesh: triangle mesh in the global frame ((nv,3) float32, (nf,3) uint32)."""
        v = np.ascontiguousarray(vertices, np.float32); f = np.ascontiguousarray(faces, np.uint32)
        _lib.check(_L().b2_reg_set_mesh(self._h, _f(v), v.shape[0], f.ctypes.data_as(C.POINTER(C.c_uint32)), f.shape[0]))

    def set_depth_map(self, image, depth):
        d = np.ascontiguousarray(depth, np.float32)
        _lib.check(_L().b2_reg_set_depth_map(self._h, image, d.shape[1], d.shape[0], _f(d)))

    def set_image_scale(self, s):
        _lib.check(_L().b2_reg_set_image_scale(self._h, s))

    def num_variables(self):
        n = C.c_int32(0); _lib.check(_L().b2_reg_num_variables(self._h, C.byref(n))); return n.value

    # ---- OcclusionGeometry::RenderDepthMap ----
    def ComputeMinMaxPointRadius(self, points, min_scaling_factor, min_radius=None, max_radius=None):
        """ComputeMinMaxPointRadius over all images (multi_scale_point_cloud.cc:126-184, 232-255) -> (min_radius, max_radius); points no
        image observes keep +inf / -inf."""
        x = np.ascontiguousarray(points, np.float32)
        n = x.shape[0]
        lo = np.full(n, np.inf, np.float32) if min_radius is None else np.ascontiguousarray(min_radius, np.float32).copy()
        hi = np.full(n, -np.inf, np.float32) if max_radius is None else np.ascontiguousarray(max_radius, np.float32).copy()
        _lib.check(_L().b2_reg_min_max_point_radius(self._h, _f(x), n, float(min_scaling_factor), _f(lo), _f(hi)))
        return lo, hi

    # ---- GroundTruthCreator (src/exe/ground_truth_creator.cc) ----
    def AccumulateScanObservationsForImage(self, image, points, observation_counts):
        """observation_counts (int32, n) += 1 for the scan points visible in `image` (:44-82); returns the updated array."""
        x = np.ascontiguousarray(points, np.float32)
        c = np.ascontiguousarray(observation_counts, np.int32).copy()
        _lib.check(_L().b2_reg_gt_accumulate_observations(self._h, int(image), _f(x), x.shape[0], c.ctypes.data_as(C.POINTER(C.c_int32))))
        return c

    def CreateGroundTruthForImage(self, image, points, colors_rgb, observation_counts, scan_point_radius=2, scan_rendering_bgr=None,
                                  write_depth_maps=True, write_occlusion_depth=True):
        """-> (occlusion_depth or None, ground_truth_depth or None, scan_rendering or None)   (:84-215, without the file I/O)"""
        x = np.ascontiguousarray(points, np.float32)
        c = np.ascontiguousarray(observation_counts, np.int32)
        w, h = self._gt_size(image)
        occ = np.zeros((h, w), np.float32) if write_occlusion_depth else None
        gt = np.zeros((h, w), np.float32) if write_depth_maps else None
        rgb = np.ascontiguousarray(colors_rgb, np.uint8) if colors_rgb is not None else None
        ren = np.ascontiguousarray(scan_rendering_bgr, np.uint8).copy() if scan_rendering_bgr is not None else None
        u8 = C.POINTER(C.c_uint8)
        _lib.check(_L().b2_reg_gt_create(self._h, int(image), _f(x), rgb.ctypes.data_as(u8) if rgb is not None else None, x.shape[0],
                                        c.ctypes.data_as(C.POINTER(C.c_int32)), int(scan_point_radius), _f(occ) if occ is not None else None,
                                        _f(gt) if gt is not None else None, ren.ctypes.data_as(u8) if ren is not None else None))
        return occ, gt, ren

    def _gt_size(self, image):
        return self._intr_sizes[self.image_intr[int(image)]]

    def render_depth(self, image):
        w, h, s = C.c_int32(), C.c_int32(), C.c_int32()
        _lib.check(_L().b2_reg_render_depth(self._h, image, C.byref(w), C.byref(h), C.byref(s), None))
        out = np.zeros((h.value, w.value), np.float32)
        _lib.check(_L().b2_reg_render_depth(self._h, image, C.byref(w), C.byref(h), C.byref(s), _f(out)))
        return out, s.value

    # ---- VisibilityEstimator ----
    def CreateObservationsForAllImages(self, border_size=1):
        _lib.check(_L().b2_reg_create_observations(self._h, border_size))

    def observations(self, image, ps):
        n = C.c_uint64(0)
        _lib.check(_L().b2_reg_num_observations(self._h, image, ps, C.byref(n)))
        n = n.value
        idx = np.zeros(n, np.uint64); x = np.zeros(n, np.float32); y = np.zeros(n, np.float32); s = np.zeros(n, np.float32); nb = np.zeros(n, np.uint8)
        if n:
            _lib.check(_L().b2_reg_get_observations(self._h, image, ps, idx.ctypes.data_as(C.POINTER(C.c_uint64)), _f(x), _f(y), _f(s), _u8(nb)))
        return idx, x, y, s, nb

    def point_jacobians(self, image, ps):
        n = len(self.observations(image, ps)[0])
        I = np.zeros(n, np.float32); jK = np.zeros((n, self.intr_np[self.image_intr[image]]), np.float32); jP = np.zeros((n, 6), np.float32)
        if n:
            _lib.check(_L().b2_reg_get_point_jacobians(self._h, image, ps, _f(I), _f(jK), _f(jP)))
        return I, jK, jP

    def point_jacobians_rig(self, image, ps):
        n = len(self.observations(image, ps)[0])
        I = np.zeros(n, np.float32); jK = np.zeros((n, self.intr_np[self.image_intr[image]]), np.float32)
        jP = np.zeros((n, 6), np.float32); jR = np.zeros((n, 6), np.float32)
        if n:
            _lib.check(_L().b2_reg_get_point_jacobians_rig(self._h, image, ps, _f(I), _f(jK), _f(jP), _f(jR)))
        return I, jK, jP, jR

    # ---- ColorOptimizer / CostCalculator / IntrinsicsAndPoseOptimizer ----
    def ColorOptimizerApply(self):
        _lib.check(_L().b2_reg_color_update(self._h))

    def descriptors(self, ps):
        n = self.scale_sizes[ps]
        f = np.zeros(n * self.K, np.float32); v = np.zeros(n * self.K, np.float32); c = np.zeros(n, np.int32)
        _lib.check(_L().b2_reg_get_descriptors(self._h, ps, _f(f), _f(v), c.ctypes.data_as(C.POINTER(C.c_int32))))
        return f, v, c

    def ComputeCost(self):
        c = C.c_double(0); s = np.zeros(6)
        _lib.check(_L().b2_reg_cost(self._h, C.byref(c), _d(s)))
        return c.value, s

    def accumulate(self):
        nv = self.num_variables()
        H = np.zeros((nv, nv), np.float64, order="F"); b = np.zeros(nv); s = np.zeros(6); c = C.c_double(0)
        _lib.check(_L().b2_reg_accumulate(self._h, _d(H), _d(b), _d(s), C.byref(c)))
        return np.asarray(H), b, s, c.value

    def get_state(self):
        ip = np.zeros(sum(self.intr_np), np.float32); po = np.zeros((self.n_img, 7), np.float32)
        _lib.check(_L().b2_reg_get_state(self._h, _f(ip), _f(po)))
        return self._intr_shape(ip), po

    def set_state(self, intr_params, poses):
        ip = np.ascontiguousarray(intr_params, np.float32).reshape(-1); po = np.ascontiguousarray(poses, np.float32)
        if ip.size != sum(self.intr_np):
            raise ValueError("intrinsics parameter vector has %d entries, expected %d" % (ip.size, sum(self.intr_np)))
        _lib.check(_L().b2_reg_set_state(self._h, _f(ip), _f(po)))

    def cost_for_delta(self, delta):
        d = np.ascontiguousarray(delta, np.float64); c = C.c_double(0)
        _lib.check(_L().b2_reg_cost_for_delta(self._h, _d(d), C.byref(c)))
        return c.value

    def IntrinsicsAndPoseOptimizerApply(self, lam):
        l = C.c_float(lam); mc = C.c_float(0); ap = C.c_int32(0); tr = C.c_int32(0)
        _lib.check(_L().b2_reg_apply(self._h, C.byref(l), C.byref(mc), C.byref(ap), C.byref(tr)))
        return bool(ap.value), l.value, mc.value, tr.value

    # ---- Optimizer ----
    def RunOnCurrentScale(self, max_num_iterations, max_change_convergence_threshold, iterations_without_new_optimum_threshold, print_progress=False):
        oc = C.c_double(0); cv = C.c_int32(0); it = C.c_int32(0)
        _lib.check(_L().b2_reg_run_on_current_scale(self._h, max_num_iterations, max_change_convergence_threshold,
                                                    iterations_without_new_optimum_threshold, int(print_progress), C.byref(oc), C.byref(cv), C.byref(it)))
        return it.value, oc.value, bool(cv.value)

    def stats(self):
        s = RegStats()
        _lib.check(_L().b2_reg_last_stats(self._h, C.byref(s)))
        return {k: getattr(s, k) for k, _ in RegStats._fields_}
