"""Synthetic Path B scene (data generator, not on the product path): a textured plane z = 0 seen by cameras looking down from
~2 m (pinhole, thin-prism or thin-prism-fisheye "benchmark" model); images are rendered analytically (every pixel is
unprojected — numerically for the distorted models — and its ray intersected with the plane; texture = sum of sinusoids); the
multi-resolution point cloud is a set of regular grids on the plane with the texture as grey colour and 5 grid neighbours."""
import math

import numpy as np


def texture(x, y):
    v = (np.sin(7.0 * x) * np.cos(5.0 * y) + 0.6 * np.sin(19.0 * x + 1.3) * np.sin(23.0 * y + 0.4) + 0.35 * np.cos(41.0 * x - 29.0 * y)
         + 0.25 * np.sin(83.0 * x + 61.0 * y))
    return 120.0 + 45.0 * v


def quat_from_R(R):
    w = math.sqrt(max(0.0, 1 + R[0, 0] + R[1, 1] + R[2, 2])) / 2
    x = (R[2, 1] - R[1, 2]) / (4 * w); y = (R[0, 2] - R[2, 0]) / (4 * w); z = (R[1, 0] - R[0, 1]) / (4 * w)
    return np.array([x, y, z, w])


def rot(rx, ry, rz):
    cx, sx, cy, sy, cz, sz = math.cos(rx), math.sin(rx), math.cos(ry), math.sin(ry), math.cos(rz), math.sin(rz)
    Rx = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]]); Ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
    Rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]])
    return Rz @ Ry @ Rx


CAM_PINHOLE, CAM_BENCHMARK, CAM_THIN_PRISM = 4, 5, 14
# the distortion of the reference's benchmark-camera test (camera/test/test_camera.cc:508-515)
DEFAULT_DISTORTION = (0.221184, 0.128597, 0.000531602, -0.000388873, 0.0623079, 0.20419, -0.000805024, 4.07704e-05)


def _thin_prism(d, x, y):
    k1, k2, p1, p2, k3, k4, sx1, sy1 = d
    x2, xy, y2 = x * x, x * y, y * y
    r2 = x2 + y2
    radial = 1 + r2 * (k1 + r2 * (k2 + r2 * (k3 + r2 * k4)))
    return (x * radial + 2 * p1 * xy + p2 * (r2 + 2 * x2) + sx1 * r2, y * radial + 2 * p2 * xy + p1 * (r2 + 2 * y2) + sy1 * r2)


# camera::CameraBase::Type -> (distortion function, fisheye wrapper, single focal length); mirrors b2_camera.cuh:cam_model
_MODELS = {0: ("fov", False, False), 1: ("radial", False, False), 2: ("tangential", False, False), 3: ("tangential", True, False),
           4: ("none", False, False), 5: ("thin_prism", True, False), 6: ("radial", True, False), 7: ("none", False, True),
           8: ("radial", False, True), 9: ("radial", False, True), 10: ("opencv", False, False), 11: ("radial", False, False),
           12: ("radial", True, True), 13: ("radial", True, True), 14: ("thin_prism", False, False)}


def distort(camera_model, dist, x, y):
    """Normalized -> distorted coordinates in float64 (scene generation only; the product path is the CUDA library)."""
    kind, fisheye, _ = _MODELS[camera_model]
    if fisheye:
        r = np.sqrt(x * x + y * y)
        f = np.where(r > 1e-9, np.arctan(r) / np.maximum(r, 1e-9), 1.0)
        x, y = x * f, y * f
    if kind == "none":
        return x, y
    if kind == "thin_prism":
        return _thin_prism(dist, x, y)
    r2 = x * x + y * y
    if kind == "radial":
        fac = 0.0
        for k in reversed(list(dist)):
            fac = r2 * (k + fac)
        return x * (1 + fac), y * (1 + fac)
    if kind == "tangential":
        k1, k2, p1, p2 = dist
        radial = 1 + r2 * (k1 + r2 * k2)
        return x * radial + 2 * p1 * x * y + p2 * (r2 + 2 * x * x), y * radial + 2 * p2 * x * y + p1 * (r2 + 2 * y * y)
    if kind == "opencv":
        k1, k2, p1, p2, k3, k4, k5, k6 = dist
        radial = (1 + r2 * (k1 + r2 * (k2 + r2 * k3))) / (1 + r2 * (k4 + r2 * (k5 + r2 * k6)))
        return x * radial + 2 * p1 * x * y + p2 * (r2 + 2 * x * x), y * radial + 2 * p2 * x * y + p1 * (r2 + 2 * y * y)
    omega = dist[0]                                                     # FOV
    r = np.sqrt(r2)
    f = np.where(r > 1e-9, np.arctan(r * 2 * math.tan(0.5 * omega)) / (np.maximum(r, 1e-9) * omega), 1.0)
    return x * f, y * f


def unproject(camera_model, dist, dx, dy):
    """Distorted (f-normalised) coordinates -> normalized ray coordinates (float64, damped fixed-point inversion)."""
    if _MODELS[camera_model][0] == "none" and not _MODELS[camera_model][1]:
        return dx, dy
    if camera_model in (CAM_BENCHMARK, CAM_THIN_PRISM):                 # (kept as it was: the committed scenes depend on it)
        ux, uy = dx.copy(), dy.copy()
        for _ in range(60):
            fx_, fy_ = _thin_prism(dist, ux, uy)
            ux = ux + 0.8 * (dx - fx_); uy = uy + 0.8 * (dy - fy_)
        if camera_model == CAM_BENCHMARK:
            r = np.sqrt(ux * ux + uy * uy)
            f = np.where(r > 1e-9, np.tan(np.minimum(r, 1.5)) / np.maximum(r, 1e-9), 1.0)
            ux, uy = ux * f, uy * f
        return ux, uy
    ux, uy = dx.copy(), dy.copy()
    for _ in range(80):
        fx_, fy_ = distort(camera_model, dist, ux, uy)
        ux = ux + 0.8 * (dx - fx_); uy = uy + 0.8 * (dy - fy_)
    return ux, uy


def split_params(camera_model, params):
    """GetParameters order -> ((fx, fy, cx, cy), distortion parameters)."""
    p = [float(v) for v in params]
    return ((p[0], p[0], p[1], p[2]), p[3:]) if _MODELS[camera_model][2] else ((p[0], p[1], p[2], p[3]), p[4:])


def render_image(width, height, K, camera_model, distortion, R_wc, c):
    """Image of the textured plane z = 0 from the camera with centre c and camera-to-world rotation R_wc."""
    yy, xx = np.mgrid[0:height, 0:width].astype(np.float64)
    ux, uy = unproject(camera_model, distortion, (xx - K[2]) / K[0], (yy - K[3]) / K[1])
    d_w = np.stack([ux, uy, np.ones_like(xx)], -1) @ R_wc.T
    s = -c[2] / d_w[..., 2]
    pw = c + d_w * s[..., None]
    return np.clip(np.rint(texture(pw[..., 0], pw[..., 1])), 0, 255).astype(np.uint8)


def pose7(R_cw, t_cw):
    return np.concatenate([quat_from_R(R_cw), t_cw]).astype(np.float32)


def make_rig_scene(num_sets=2, width=320, height=240, fx=260.0, camera_model=CAM_PINHOLE, distortion=DEFAULT_DISTORTION, num_scales=3,
                   base_radius=0.004, seed=5, perturb=(0.002, 0.002)):
    """Two-camera rig recorded `num_sets` times. Returns the make_scene dict (images ordered set0-cam0, set0-cam1, set1-cam0, ...)
    plus rig_gt / rig_init ((2,7) image_T_rig, camera 0 identity) and rig_sets (image ids per set)."""
    rng = np.random.default_rng(seed)
    base = make_scene(num_images=0, width=width, height=height, fx=fx, num_scales=num_scales, base_radius=base_radius, camera_model=camera_model,
                      distortion=distortion)
    K = base["intr"][2][:4]
    E_R = rot(0.02, -0.05, 0.03); E_t = np.array([-0.18, 0.01, 0.005])          # cam1_T_cam0 (image_T_rig of camera 1)
    images, poses_gt, poses_init, sets = [], [], [], []
    dE = rot(*(rng.uniform(-perturb[1], perturb[1], 3))); dEt = rng.uniform(-perturb[0], perturb[0], 3)
    Ei_R, Ei_t = dE @ E_R, dE @ E_t + dEt
    for i in range(num_sets):
        c = np.array([0.2 * math.cos(1.9 * i), 0.15 * math.sin(1.3 * i), 2.0 + 0.08 * i])
        R_wc = rot(math.pi + 0.06 * math.sin(1.1 * i + 0.3), 0.05 * math.cos(0.7 * i), 0.4 * i)
        R0 = R_wc.T; t0 = -R0 @ c
        dR = rot(*(rng.uniform(-perturb[1], perturb[1], 3))); dt = rng.uniform(-perturb[0], perturb[0], 3)
        R0i, t0i = dR @ R0, dR @ t0 + dt
        ids = []
        for cam, (R, t, Ri, ti) in enumerate([(R0, t0, R0i, t0i), (E_R @ R0, E_R @ t0 + E_t, Ei_R @ R0i, Ei_R @ t0i + Ei_t)]):
            images.append(render_image(width, height, K, camera_model, distortion, R.T, -R.T @ t))
            poses_gt.append(pose7(R, t)); poses_init.append(pose7(Ri, ti))
            ids.append(len(images) - 1)
        sets.append(ids)
    ident = np.array([0, 0, 0, 1, 0, 0, 0], np.float32)
    base.update(images=images, poses_gt=poses_gt, poses_init=poses_init, rig_sets=sets,
                rig_gt=np.stack([ident, pose7(E_R, E_t)]), rig_init=np.stack([ident, pose7(Ei_R, Ei_t)]))
    return base


def load_rig_into(reg, scene, use_init=True):
    """load_into for a make_rig_scene scene: images, then the rig and its image sets (which define the dependent images' poses)."""
    w, h, K = scene["intr"]
    reg.add_intrinsics(w, h, K, camera_model=scene.get("camera_model", CAM_PINHOLE))
    for img, T in zip(scene["images"], scene["poses_init"] if use_init else scene["poses_gt"]):
        reg.add_image(0, img, None, T)
    rig = reg.add_rig(scene["rig_init"] if use_init else scene["rig_gt"])
    for ids in scene["rig_sets"]:
        reg.add_rig_images(rig, ids)
    count = reg.initialize()
    for xyz, radius, nbr, colors in scene["scales"]:
        reg.add_point_scale(xyz, float(radius), nbr, colors)
    reg.set_splat_points(scene["scales"][0][0])
    return count


def make_scene(num_images=3, width=640, height=480, fx=520.0, extent=(2.4, 1.8), base_radius=0.0025, num_scales=3, seed=31,
               perturb=(0.002, 0.002), camera_model=CAM_PINHOLE, distortion=DEFAULT_DISTORTION, camera_params=None):
    """Returns dict(intr=(w,h,params), camera_model, images=[uint8 HxW], poses_gt, poses_init (qx qy qz qw tx ty tz),
    scales=[(xyz, radius, nbr, colors)]). params = fx fy cx cy (+ k1 k2 p1 p2 k3 k4 sx1 sy1 for the thin-prism models), or camera_params
    (the model's full GetParameters vector) for any of the 15 camera models."""
    rng = np.random.default_rng(seed)
    K = np.array([fx, fx, (width - 1) / 2.0, (height - 1) / 2.0], np.float32)
    params = K if camera_model == CAM_PINHOLE else np.concatenate([K, np.asarray(distortion, np.float32)])
    if camera_params is not None:
        params = np.asarray(camera_params, np.float32)
        K4, distortion = split_params(camera_model, params)
        K = np.array(K4, np.float32)
    images, poses_gt, poses_init = [], [], []
    for i in range(num_images):
        # camera centre above the plane, looking down (camera z axis = -world z), small tilts
        c = np.array([0.25 * math.cos(2.1 * i), 0.2 * math.sin(1.7 * i), 2.0 + 0.1 * math.sin(i)])
        R_wc = rot(math.pi + 0.08 * math.sin(1.3 * i), 0.07 * math.cos(0.9 * i), 0.3 * i)      # camera-to-world
        R_cw = R_wc.T; t_cw = -R_cw @ c
        images.append(render_image(width, height, K, camera_model, distortion, R_wc, c))
        q = quat_from_R(R_cw)
        poses_gt.append(np.concatenate([q, t_cw]).astype(np.float32))
        dR = rot(*(rng.uniform(-perturb[1], perturb[1], 3))); dt = rng.uniform(-perturb[0], perturb[0], 3)
        Rp = dR @ R_cw; tp = dR @ t_cw + dt
        poses_init.append(np.concatenate([quat_from_R(Rp), tp]).astype(np.float32))
    scales = []
    for sidx in range(num_scales):
        radius = base_radius * (2 ** sidx)
        step = 2 * radius
        nx = int(extent[0] / step); ny = int(extent[1] / step)
        gx, gy = np.meshgrid(np.arange(nx), np.arange(ny), indexing="xy")
        x = (gx.ravel() - (nx - 1) / 2.0) * step; y = (gy.ravel() - (ny - 1) / 2.0) * step
        xyz = np.stack([x, y, np.zeros_like(x)], 1).astype(np.float32)
        idx = (gy * nx + gx).ravel()
        def nb(dx, dy):
            return (np.clip(gy + dy, 0, ny - 1) * nx + np.clip(gx + dx, 0, nx - 1)).ravel()
        nbr = np.stack([nb(1, 0), nb(-1, 0), nb(0, 1), nb(0, -1), nb(1, 1)], 1).astype(np.uint64)
        same = nbr == idx[:, None].astype(np.uint64)           # clipped border neighbours: point elsewhere so they differ from the centre
        alt = np.stack([nb(-2, 0), nb(2, 0), nb(0, -2), nb(0, 2), nb(-1, -1)], 1).astype(np.uint64)
        nbr = np.where(same, alt, nbr)
        colors = texture(x, y).astype(np.float32)
        scales.append((xyz, np.float32(radius), nbr, colors))
    return {"intr": (width, height, params), "camera_model": camera_model, "images": images, "poses_gt": poses_gt, "poses_init": poses_init, "scales": scales}


def load_into(reg, scene, use_init=True, splats=True):
    """Feeds a scene into a Registration-like object (product mirror or oracle: same method names)."""
    w, h, K = scene["intr"]
    reg.add_intrinsics(w, h, K, camera_model=scene.get("camera_model", CAM_PINHOLE))
    for img, T in zip(scene["images"], scene["poses_init"] if use_init else scene["poses_gt"]):
        reg.add_image(0, img, None, T)
    count = reg.initialize()
    for xyz, radius, nbr, colors in scene["scales"]:
        reg.add_point_scale(xyz, float(radius), nbr, colors)
    if splats:
        reg.set_splat_points(scene["scales"][0][0])
    return count
