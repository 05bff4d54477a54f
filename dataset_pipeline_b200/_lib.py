"""ctypes binding of libeth3d_b200.so (include/eth3d_b200.h). The CUDA library is the product path: if it is missing or
cannot be loaded this module raises — there is no CPU / eager fallback anywhere in this package."""
import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("B2_LIB_PATH") or os.path.join(_HERE, "_build", "libeth3d_b200.so")   # (override: A/B builds of the library)
_lib = None

B2_OK = 0
ERRORS = {1: "B2_ERR_ARG", 2: "B2_ERR_STATE", 3: "B2_ERR_CUDA", 4: "B2_ERR_NO_DEVICE", 5: "B2_ERR_ALLOC", 6: "B2_ERR_COMM"}

ALLREDUCE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p)


class IcpConfig(C.Structure):
    _fields_ = [("device", C.c_int32), ("inner_max_iterations", C.c_int32), ("keep_correspondences", C.c_int32),
                ("rank", C.c_int32), ("world_size", C.c_int32), ("allreduce", ALLREDUCE_FN), ("allreduce_user", C.c_void_p),
                ("stream", C.c_void_p), ("comm", C.c_void_p), ("index_distance_hint", C.c_float), ("shard_uploads", C.c_int32),
                ("search_ahead", C.c_int32)]


class IcpStats(C.Structure):
    _fields_ = [("inner_iterations", C.c_int32), ("lm_tries_total", C.c_int32), ("num_pairs", C.c_int32),
                ("num_variables", C.c_int32), ("num_correspondences", C.c_uint64), ("local_correspondences", C.c_uint64),
                ("first_cost", C.c_double), ("last_cost", C.c_double), ("final_lambda", C.c_double),
                ("passes", C.c_int32), ("kernel_launches", C.c_int32),
                ("ms_index", C.c_float), ("ms_search", C.c_float), ("ms_pack", C.c_float), ("ms_inner", C.c_float),
                ("ms_total", C.c_float), ("ms_accum_kernel_avg", C.c_float), ("ms_search_kernel_avg", C.c_float),
                ("search_launches", C.c_int32), ("search_algorithmic_bytes", C.c_uint64),
                ("ms_index_build", C.c_float), ("sparse_grids", C.c_int32), ("search_work", C.c_uint64 * 5),
                ("searches_ahead", C.c_int32), ("packs_overlapped", C.c_int32)]


class B2Error(RuntimeError):
    def __init__(self, code, text):
        super().__init__("%s: %s" % (ERRORS.get(code, code), text))
        self.code = code


def build(force=False):
    """nvcc -gencode arch=compute_100a,code=sm_100a (cross-compiles without a GPU). In-tree output so it travels to the GPU box."""
    csrc = os.path.join(_HERE, "csrc")
    subprocess.check_call(["make", "-C", csrc, "-s"] + (["-B"] if force else []))
    return LIB_PATH


# every symbol include/eth3d_b200.h declares
EXPORTS = [
    "b2_abi_version",
    "b2_camera_eval",
    "b2_comm_allreduce",
    "b2_comm_allreduce_f64",
    "b2_comm_broadcast",
    "b2_comm_create",
    "b2_comm_destroy",
    "b2_comm_info",
    "b2_comm_unique_id",
    "b2_device_info",
    "b2_find_correspondences",
    "b2_icp_add_cloud",
    "b2_icp_add_cloud_dev",
    "b2_icp_create",
    "b2_icp_default_config",
    "b2_icp_destroy",
    "b2_icp_get_lm_tries",
    "b2_icp_get_normal_equations",
    "b2_icp_get_pair_correspondences",
    "b2_icp_get_pair_info",
    "b2_icp_get_pose",
    "b2_icp_last_stats",
    "b2_icp_plan_directions",
    "b2_icp_run",
    "b2_icp_set_option",
    "b2_icp_set_pose",
    "b2_icp_upload_owner",
    "b2_last_error",
    "b2_lsor_filter",
    "b2_mesh_squared_distance",
    "b2_ms_create",
    "b2_ms_merge_close_points",
    "b2_ms_point_neighbors",
    "b2_normals_estimate",
    "b2_normals_estimate_dist",
    "b2_normals_estimate_radius",
    "b2_reg_accumulate",
    "b2_reg_add_image",
    "b2_reg_add_intrinsics",
    "b2_reg_add_point_scale",
    "b2_reg_add_rig",
    "b2_reg_add_rig_images",
    "b2_reg_apply",
    "b2_reg_color_update",
    "b2_reg_cost",
    "b2_reg_cost_for_delta",
    "b2_reg_create",
    "b2_reg_create_observations",
    "b2_reg_default_params",
    "b2_reg_destroy",
    "b2_reg_get_descriptors",
    "b2_reg_get_observations",
    "b2_reg_get_point_jacobians",
    "b2_reg_get_point_jacobians_rig",
    "b2_reg_get_rigs",
    "b2_reg_get_state",
    "b2_reg_gt_accumulate_observations",
    "b2_reg_gt_create",
    "b2_reg_image_owner",
    "b2_reg_initialize",
    "b2_reg_last_stats",
    "b2_reg_min_max_point_radius",
    "b2_reg_num_observations",
    "b2_reg_num_variables",
    "b2_reg_render_depth",
    "b2_reg_run_on_current_scale",
    "b2_reg_set_camera_mask",
    "b2_reg_set_comm",
    "b2_reg_set_depth_map",
    "b2_reg_set_image_scale",
    "b2_reg_set_mesh",
    "b2_reg_set_rigs",
    "b2_reg_set_splat_points",
    "b2_reg_set_state",
    "b2_reg_variable_index",
    "b2_splat_create",
    "b2_trim",
]


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError("libeth3d_b200.so is not built (%s). Run `python -c 'import __graft_entry__ as g; g.build()'`; "
                          "there is no CPU fallback." % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    fp, ip, dp, vp = C.POINTER(C.c_float), C.POINTER(C.c_int32), C.POINTER(C.c_double), C.c_void_p
    L.b2_last_error.restype = C.c_char_p
    L.b2_abi_version.restype = C.c_int
    L.b2_device_info.argtypes = [ip, C.c_char_p, C.c_size_t, ip, ip, ip]
    L.b2_comm_unique_id.argtypes = [C.c_char_p]
    L.b2_comm_create.argtypes = [C.c_int, C.c_int, C.c_char_p, C.c_int, C.POINTER(vp)]
    L.b2_comm_destroy.argtypes = [vp]
    L.b2_comm_allreduce_f64.argtypes = [vp, vp, C.c_size_t, vp]
    L.b2_comm_info.argtypes = [vp, ip, ip]
    L.b2_icp_default_config.argtypes = [C.POINTER(IcpConfig)]
    L.b2_icp_default_config.restype = None
    L.b2_icp_create.argtypes = [C.POINTER(IcpConfig), C.POINTER(vp)]
    L.b2_icp_destroy.argtypes = [vp]
    L.b2_icp_add_cloud.argtypes = [vp, vp, vp, C.c_size_t, C.c_size_t, fp, C.c_int, ip]
    L.b2_icp_add_cloud_dev.argtypes = [vp, vp, vp, C.c_size_t, fp, C.c_int, ip]
    L.b2_icp_run.argtypes = [vp, C.c_float, C.c_int, C.c_int, C.c_float, C.c_int, ip]
    L.b2_icp_get_pose.argtypes = [vp, C.c_int, fp]
    L.b2_icp_set_pose.argtypes = [vp, C.c_int, fp]
    L.b2_icp_last_stats.argtypes = [vp, C.POINTER(IcpStats)]
    L.b2_icp_set_option.argtypes = [vp, C.c_char_p, C.c_int]
    L.b2_icp_get_lm_tries.argtypes = [vp, ip, C.c_int, ip]
    L.b2_icp_get_pair_info.argtypes = [vp, C.c_int, ip, ip, C.POINTER(C.c_uint64)]
    L.b2_icp_get_pair_correspondences.argtypes = [vp, C.c_int, ip, ip, fp]
    L.b2_icp_get_normal_equations.argtypes = [vp, dp, dp, dp, ip]
    L.b2_icp_plan_directions.argtypes = [C.c_int, C.c_int, C.c_int, ip, ip, ip, C.c_int, ip]
    L.b2_find_correspondences.argtypes = [fp, C.c_size_t, fp, C.c_size_t, C.c_float, ip, ip, fp, C.POINTER(C.c_uint64)]
    L.b2_normals_estimate.argtypes = [vp, C.c_size_t, C.c_size_t, C.c_int, fp, fp, ip, ip]
    _lib = L
    return L


def check(rc):
    if rc != B2_OK:
        raise B2Error(rc, lib().b2_last_error().decode("utf-8", "replace"))
