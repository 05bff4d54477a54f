"""Port of the reference's four-frame registration tests: src/opt/test/test_alignment.cc:86-634 (Test4FrameAlignment, the body of
TEST(Alignment, FourFrame_FixedColorsOnly / _FixedAndVariableColors / _FixedColorsOnly_Rig / _FixedAndVariableColors_Rig)).

What is the reference's: the scene (a 61 x 61 random height map pulled back at its borders, random vertex colours), the four pinhole
views (two image sets 0.5 apart, two cameras 0.1 apart in each), the point cloud (30 % of the rendered pixels, unprojected), the
perturbation that survives the test's three assignments (the last one wins: +0.002 / +0.006 in x and y for camera 0 / camera 1), the
parameters, the scale loop, and the pass criteria (every component of log(result * truth^-1) <= 0.0016, mean optical flow <= 0.07 px).
What cannot be: the std::mt19937 / libstdc++ distribution streams are replaced by numpy's (same distributions, seed 0), and the OpenGL
colour + depth render by the rasteriser below (perspective-correct barycentric interpolation, pixel centres at integers as
renderer.cc's projection has them). The depth-residual variant needs B8's depth branch, which is not built.
Shared by the oracle test (CPU) and, with the library's objects, a C-ABI test."""
import math

import numpy as np

W = H = 256
FX = FY = np.float32(0.5 * W)
CX = np.float32(0.5 * W - 0.5)
CY = np.float32(0.5 * H - 0.5)
NV = 61
TEST_THRESHOLD = 0.0016            # test_alignment.cc:541
FLOW_THRESHOLD = 0.07              # :592


def build_scene(seed=0):
    rng = np.random.default_rng(seed)
    ys, xs = np.meshgrid(np.arange(NV), np.arange(NV), indexing="ij")
    u = (xs / np.float32(NV - 1.0)).astype(np.float32); v = (ys / np.float32(NV - 1.0)).astype(np.float32)
    x = ((u - np.float32(0.5)) * np.float32(5.0)).astype(np.float32)
    y = ((v - np.float32(0.5)) * np.float32(5.0)).astype(np.float32)
    z = (np.float32(1.0) + rng.uniform(-0.05, 0.05, (NV, NV))).astype(np.float32)
    z = (z - 6 * np.sqrt((u.astype(np.float64) - 0.5) ** 2 + (v.astype(np.float64) - 0.5) ** 2)).astype(np.float32)
    verts = np.stack([x, y, z], -1).reshape(-1, 3)
    colors = rng.integers(0, 256, (NV * NV, 3)).astype(np.float32)             # r, g, b per vertex
    faces = []
    for yy in range(NV - 1):
        for xx in range(NV - 1):
            faces.append((xx + (yy + 1) * NV, (xx + 1) + yy * NV, xx + yy * NV))
            faces.append((xx + (yy + 1) * NV, (xx + 1) + (yy + 1) * NV, (xx + 1) + yy * NV))
    return verts, colors, np.array(faces, np.int64), rng


def render(verts, colors, faces, t_image_global, near=0.1):
    """Colour (h, w, 3 uint8, r g b) and linear depth (h, w float32, 0 = background) of the mesh from a camera with identity rotation
    at image_T_global = (I, t)."""
    cam = verts.astype(np.float64) + np.asarray(t_image_global, np.float64)
    depth = np.zeros((H, W), np.float64); color = np.zeros((H, W, 3), np.float64)
    zbuf = np.full((H, W), np.inf)
    px = float(FX) * cam[:, 0] / cam[:, 2] + float(CX); py = float(FY) * cam[:, 1] / cam[:, 2] + float(CY)
    for f in faces:
        zc = cam[f, 2]
        if (zc <= near).any():
            continue                      # such triangles project outside the image in this scene (see the module docstring)
        X, Y = px[f], py[f]
        x0, x1 = max(0, int(math.ceil(X.min()))), min(W - 1, int(math.floor(X.max())))
        y0, y1 = max(0, int(math.ceil(Y.min()))), min(H - 1, int(math.floor(Y.max())))
        if x0 > x1 or y0 > y1:
            continue
        gx, gy = np.meshgrid(np.arange(x0, x1 + 1), np.arange(y0, y1 + 1))
        area = (X[1] - X[0]) * (Y[2] - Y[0]) - (X[2] - X[0]) * (Y[1] - Y[0])
        if area == 0:
            continue
        w0 = ((X[1] - gx) * (Y[2] - gy) - (X[2] - gx) * (Y[1] - gy)) / area
        w1 = ((X[2] - gx) * (Y[0] - gy) - (X[0] - gx) * (Y[2] - gy)) / area
        w2 = 1.0 - w0 - w1
        inside = (w0 >= 0) & (w1 >= 0) & (w2 >= 0)
        if not inside.any():
            continue
        iz = w0 / zc[0] + w1 / zc[1] + w2 / zc[2]
        zpix = 1.0 / iz
        sub = zbuf[y0:y1 + 1, x0:x1 + 1]
        take = inside & (zpix < sub)
        if not take.any():
            continue
        c = (w0[..., None] * colors[f[0]] / zc[0] + w1[..., None] * colors[f[1]] / zc[1] + w2[..., None] * colors[f[2]] / zc[2]) / iz[..., None]
        sub[take] = zpix[take]
        depth[y0:y1 + 1, x0:x1 + 1][take] = zpix[take]
        color[y0:y1 + 1, x0:x1 + 1][take] = c[take]
    return np.clip(np.rint(color), 0, 255).astype(np.uint8), depth.astype(np.float32)


def gray_of(rgb):
    """cv::imread(IMREAD_GRAYSCALE) of the saved colour PNG (image.cc:48): OpenCV's fixed-point BGR -> gray."""
    import cv2
    return cv2.cvtColor(np.ascontiguousarray(rgb[:, :, ::-1]), cv2.COLOR_BGR2GRAY)


def se3_log(R, t):
    """Sophus::SE3::log -> (upsilon, omega), the six components the test thresholds."""
    cos_th = max(-1.0, min(1.0, (np.trace(R) - 1) / 2)); th = math.acos(cos_th)
    if th < 1e-10:
        w = np.array([R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1]]) / 2
        Vinv = np.eye(3)
    else:
        w = th / (2 * math.sin(th)) * np.array([R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1]])
        K = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]])
        Vinv = np.eye(3) - 0.5 * K + (1 - th * math.cos(th / 2) / (2 * math.sin(th / 2))) / (th * th) * K @ K
    return np.concatenate([Vinv @ t, w])


def quat_to_R(q):
    x, y, z, w = [float(v) for v in q]
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def run_four_frame(make_reg, multires, run, get_state, use_fixed_colors, use_variable_colors, use_rig, max_iterations=500, no_optimum=25):
    """Returns (worst |log component|, mean optical flow in px, per-scale log)."""
    verts, colors, faces, rng = build_scene(0)
    # image_T_global = inverse of global_T_image; all rotations are the identity (test_alignment.cc:176-219)
    t_global_image = {(0, 0): np.array([0, -0.25, 0]), (0, 1): np.array([0.1, -0.25, 0]), (1, 0): np.array([0, 0.25, 0]), (1, 1): np.array([0.1, 0.25, 0])}
    order = [(0, 0), (0, 1), (1, 0), (1, 1)]
    t_image_global = {k: -v for k, v in t_global_image.items()}
    images, depths = {}, {}
    for k in order:
        images[k], depths[k] = render(verts, colors, faces, t_image_global[k])
    fx_inv = np.float32(1.0) / FX; cx_inv = -CX / FX; fy_inv = np.float32(1.0) / FY; cy_inv = -CY / FY
    pts, rgb = [], []
    for k in order:                                                        # :262-298, row-major pixels, probability 0.3
        d = depths[k]
        sel = (d > 0) & (rng.uniform(0, 1, d.shape) < 0.3)
        ys, xs = np.nonzero(sel)
        nx = (fx_inv * xs.astype(np.float32) + cx_inv).astype(np.float32); ny = (fy_inv * ys.astype(np.float32) + cy_inv).astype(np.float32)
        dd = d[ys, xs]
        cam = np.stack([dd * nx, dd * ny, dd], 1).astype(np.float32)
        pts.append((cam + t_global_image[k].astype(np.float32)).astype(np.float32))
        rgb.append(images[k][ys, xs])
    xyz = np.concatenate(pts); rgb = np.concatenate(rgb)
    # the perturbation that is in effect when the problem is built (:383-397): translation only, +0.002 (camera 0) / +0.006 (camera 1)
    pert = {k: np.array([(1.0 if k[1] == 0 else 3.0) * 0.002] * 2 + [0.0]) for k in order}
    start_t = {k: pert[k] + t_image_global[k] for k in order}

    reg = make_reg(point_neighbor_count=5, robust_weighting_type=2, robust_weighting_parameter=5.0, max_initial_image_area_in_pixels=64 * 64,
                   occlusion_depth_threshold=0.05, fixed_residuals_weight=1.0 if use_fixed_colors else 0.0,
                   variable_residuals_weight=1.0 if use_variable_colors else 0.0)
    reg.add_intrinsics(W, H, np.array([FX, FY, CX, CY], np.float32))
    ids = {}
    for k in order:
        ids[k] = reg.add_image(0, gray_of(images[k]), None, np.array([0, 0, 0, 1, *start_t[k]], np.float32))
    if use_rig:
        # rig->image_T_rig[1] = perturbed[0][1] * perturbed[0][0]^-1 (:428-430): a pure translation here
        rel = start_t[(0, 1)] - start_t[(0, 0)]
        rig = reg.add_rig(np.array([[0, 0, 0, 1, 0, 0, 0], [0, 0, 0, 1, *rel]], np.float32))
        reg.add_rig_images(rig, [ids[(0, 0)], ids[(0, 1)]])
        reg.add_rig_images(rig, [ids[(1, 0)], ids[(1, 1)]])
    count = reg.initialize()
    reg.set_splat_points(xyz)
    radii, P, C, sidx, nbrs = multires(reg, [(xyz, rgb)], count, 1.0 if use_fixed_colors else 0.0)
    assert len(radii) >= 1
    for r, p, c, nb in zip(radii, P, C, nbrs):
        reg.add_point_scale(p, float(r), nb, c)
    scale = max((count - 1) - 1, 0)                                        # Optimizer(problem.max_image_scale() - 1, ...)
    log = []
    while True:
        reg.set_image_scale(scale)
        log.append((scale,) + tuple(run(reg, max_iterations, 1e-20, no_optimum)))
        if scale == 0:
            break
        scale -= 1
    intr, poses = get_state(reg)
    worst = 0.0
    flow_sum, flow_n = 0.0, 0
    fx, fy, cx, cy = [float(v) for v in np.asarray(intr).reshape(-1)[:4]]
    for k in order:
        q = poses[ids[k]]
        R = quat_to_R(q[:4]); t = q[4:].astype(np.float64)
        # delta = result_image_T_global * image_T_global(truth)^-1 = (R, t) * (I, -t_truth)
        worst = max(worst, float(np.abs(se3_log(R, t - R @ t_image_global[k])).max()))
        d = depths[k]
        ys, xs = np.nonzero(d > 0)
        nx = (fx_inv * xs.astype(np.float32) + cx_inv).astype(np.float64); ny = (fy_inv * ys.astype(np.float32) + cy_inv).astype(np.float64)
        dd = d[ys, xs].astype(np.float64)
        glob = np.stack([dd * nx, dd * ny, dd], 1) + t_global_image[k]
        res = glob @ R.T + t
        ok = res[:, 2] > 0
        u = fx * res[ok, 0] / res[ok, 2] + cx; v = fy * res[ok, 1] / res[ok, 2] + cy
        flow_sum += float(np.sqrt((u - xs[ok]) ** 2 + (v - ys[ok]) ** 2).sum()); flow_n += int(ok.sum())
    return worst, flow_sum / max(flow_n, 1), log
