"""Port of the reference's end-to-end registration test: src/opt/test/test_alignment.cc:50-84 (TestPairAlignment, the body of
TEST(Alignment, SimpleTwoFrame)) with test_alignment_util.cc:134-330 (ProcessOnePair, DetermineErrorMetrics). Shared by the oracle test
(CPU) and the C-ABI test (GPU): `make_reg` builds a Registration-like object, `multires` is the matching ComputeMultiResPointCloud."""
import math
import os

import numpy as np

FIXTURE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_alignment_pair.npz")
TRANSLATION_THRESHOLD = 1e-2       # relative to the average scene depth (test_alignment.cc:52)
ROTATION_THRESHOLD_DEG = 1.0       # (:53)


def load_pair(key):
    z = np.load(FIXTURE)
    g = lambda k: z[key + "_" + k]
    w, h, fx, fy, cx, cy, depth_factor = [float(v) for v in g("calibration")]
    return dict(w=int(w), h=int(h), fx=np.float32(fx), fy=np.float32(fy), cx=np.float32(cx), cy=np.float32(cy), depth_factor=np.float32(depth_factor),
                a_gray=g("a_gray"), b_gray=g("b_gray"), a_bgr=g("a_bgr"), a_depth=g("a_depth"), a_t_b=g("a_t_b"), average_scene_depth=float(g("average_scene_depth")))


def point_cloud_of(info):
    """test_alignment_util.cc:174-192: every pixel with depth != 0 -> (depth nx, depth ny, depth), pinhole ImageToNormalized in fp32
    (camera_pinhole.h: fx_inv x + cx_inv), the products with the double depth rounded to float; colour = the model image's pixel."""
    depth = (np.float64(info["depth_factor"]) * info["a_depth"].astype(np.float64))
    ys, xs = np.nonzero(depth != 0)                       # row-major order = the reference's y, x loops
    d = depth[ys, xs]
    fx_inv = np.float32(1.0) / info["fx"]; fy_inv = np.float32(1.0) / info["fy"]
    cx_inv = -info["cx"] / info["fx"]; cy_inv = -info["cy"] / info["fy"]
    nx = (fx_inv * xs.astype(np.float32) + cx_inv).astype(np.float32); ny = (fy_inv * ys.astype(np.float32) + cy_inv).astype(np.float32)
    xyz = np.stack([(d * nx.astype(np.float64)).astype(np.float32), (d * ny.astype(np.float64)).astype(np.float32), d.astype(np.float32)], 1)
    rgb = info["a_bgr"][ys, xs][:, ::-1].copy()
    return xyz, rgb


def quat_to_R(q):
    x, y, z, w = [float(v) for v in q]
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def process_one_pair(info, make_reg, multires, run, get_poses):
    """ProcessOnePair (test_alignment_util.cc:134-300) with max_initial_image_area_in_pixels = 80 * 60. Returns the estimated a_T_b (3x4)."""
    xyz, rgb = point_cloud_of(info)
    reg = make_reg(max_initial_image_area_in_pixels=80 * 60)
    reg.add_intrinsics(info["w"], info["h"], np.array([info["fx"], info["fy"], info["cx"], info["cy"]], np.float32))
    ident = np.array([0, 0, 0, 1, 0, 0, 0], np.float32)
    reg.add_image(0, info["a_gray"], None, ident)
    reg.add_image(0, info["b_gray"], None, ident)
    count = reg.initialize()
    reg.set_splat_points(xyz)                                        # occlusion_geometry->SetSplatPoints(point_cloud)
    radii, pts, cols, sidx, nbrs = multires(reg, [(xyz, rgb)], count)    # Problem::SetScanGeometryAndInitialize -> ComputeMultiResPointCloud
    assert len(radii) >= 1
    for r, p, c, nb in zip(radii, pts, cols, nbrs):
        reg.add_point_scale(p, float(r), nb, c)
    scale = (count - 1) - 1                                          # Optimizer(problem.max_image_scale() - 1, ...)
    scale = max(scale, 0)
    log = []
    while True:
        reg.set_image_scale(scale)
        log.append((scale,) + tuple(run(reg, 300, 0.0, 10)))         # kMaxIterations, kMaxChangeConvergenceThreshold, kIterationsWithoutNewOptimumThreshold
        if scale == 0:                                               # Optimizer::NextScale
            break
        scale -= 1
    poses = get_poses(reg)                                           # (n_images, 7): qx qy qz qw tx ty tz = image_T_global
    Ra, ta = quat_to_R(poses[0][:4]), poses[0][4:].astype(np.float64)
    Rb, tb = quat_to_R(poses[1][:4]), poses[1][4:].astype(np.float64)
    R = Ra @ Rb.T                                                    # model.image_T_global * query.global_T_image
    t = ta - R @ tb
    return np.concatenate([R, t[:, None]], 1), log


def error_metrics(info, est):
    """DetermineErrorMetrics (test_alignment_util.cc:302-322)."""
    terr = float(np.linalg.norm(est[:, 3] - info["a_t_b"][:, 3]) / info["average_scene_depth"])
    D = est[:, :3].T @ info["a_t_b"][:, :3]
    ang = math.degrees(math.acos(max(-1.0, min(1.0, (np.trace(D) - 1) / 2))))
    return terr, ang
