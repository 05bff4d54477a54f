"""Pins the oracle's camera models (oracle/orc_camera.h) on the reference's own camera tests,
/root/reference/src/camera/test/test_camera.cc:40-420 (RunCameraModelTests) with the cameras of :424-427 (Pinhole) and
:508-515 (Benchmark); the thin-prism model (the benchmark camera's inner model) runs through the same checks."""
import numpy as np
import pytest

W, H = 640, 480
PINHOLE = [250.0, 200.0, 319.5, 239.5]
BENCH = [340.926, 341.124, 302.4, 201.6, 0.221184, 0.128597, 0.000531602, -0.000388873, 0.0623079, 0.20419, -0.000805024, 4.07704e-05]


K1, K2, K3 = 0.13, -0.66, 0.64
PT = [340.926, 341.124, 302.4, 201.6, -0.101082, 0.0703954, 0.000438661, -0.000680887]


def cameras(orc):
    """The cameras of the reference's test_camera.cc:422-505 (one TEST per model) + the thin-prism model (the benchmark camera's inner
    model) + Polynomial4 (FisheyePolynomial4's inner model)."""
    return [(orc.CAM_PINHOLE, PINHOLE),
            (orc.CAM_SIMPLE_PINHOLE, [250.0, 319.5, 239.5]),
            (orc.CAM_RADIAL, [250.0, 319.5, 239.5, K1, -1e-2]),
            (orc.CAM_RADIAL_FISHEYE, [250.0, 319.5, 239.5, -K1, -K2]),
            (orc.CAM_SIMPLE_RADIAL, [450.0, 319.5, 239.5, K1]),
            (orc.CAM_SIMPLE_RADIAL_FISHEYE, [450.0, 319.5, 239.5, K1]),
            (orc.CAM_POLYNOMIAL, PINHOLE + [K1, K2, K3]),
            (orc.CAM_FOV, PINHOLE + [1.0]),
            (orc.CAM_POLYNOMIAL_TANGENTIAL, PT),
            (orc.CAM_FULL_OPENCV, PT[:4] + [-0.101082, 0.0703954, 0.0438661, -0.0680887, -0.00101082, .1, .001, -.001]),
            (orc.CAM_FISHEYE_POLYNOMIAL_4, PT[:4] + [0.221184, 0.128597, 0.0623079, 0.20419]),
            (orc.CAM_POLYNOMIAL_4, PT[:4] + [0.221184, 0.128597, 0.0623079, 0.20419]),
            (orc.CAM_FISHEYE_POLYNOMIAL_TANGENTIAL, PT),
            (orc.CAM_THIN_PRISM, BENCH), (orc.CAM_BENCHMARK, BENCH)]


def base_intrinsics(p, model, orc):
    unique = model in (orc.CAM_SIMPLE_PINHOLE, orc.CAM_RADIAL, orc.CAM_RADIAL_FISHEYE, orc.CAM_SIMPLE_RADIAL, orc.CAM_SIMPLE_RADIAL_FISHEYE)
    return (p[0], p[0], p[1], p[2]) if unique else tuple(p[:4])


def test_parameter_counts(oracle):
    orc = oracle
    assert orc.cam_param_count(orc.CAM_PINHOLE) == 4           # camera_pinhole.h ParameterCount
    assert orc.cam_param_count(orc.CAM_THIN_PRISM) == 12       # camera_thin_prism.h:52-54
    assert orc.cam_param_count(orc.CAM_BENCHMARK) == 12        # FisheyeBase::ParameterCount -> inner model's
    for model, p in cameras(oracle):
        assert orc.cam_param_count(model) == len(p)            # test_camera.cc:361-376 TestParameterStorage
    with pytest.raises(ValueError):
        orc.cam_param_count(15)


def test_undistort_distort_image_corners(oracle):
    orc = oracle
    # test_camera.cc:40-70: Distort(Undistort(n)) == n at the four image corners, 1e-5
    for model, p in cameras(oracle):
        if model == orc.CAM_FOV:
            continue                                           # test_camera.cc:384-388: the corners show nothing with these parameters
        fx, fy, cx, cy = base_intrinsics(p, model, orc)
        fxi, fyi = np.float32(1.0 / fx), np.float32(1.0 / fy)
        cxi, cyi = np.float32(-1.0 * cx / fx), np.float32(-1.0 * cy / fy)
        corners = np.array([[0, 0], [W - 1, 0], [0, H - 1], [W - 1, H - 1]], np.float32)
        n = np.stack([fxi * corners[:, 0] + cxi, fyi * corners[:, 1] + cyi], 1)
        back = orc.cam_eval(model, W, H, p, "distort", orc.cam_eval(model, W, H, p, "undistort", n))
        assert np.abs(back - n).max() < 1e-5, (model, back, n)


def test_distort_undistort(oracle):
    orc = oracle
    # test_camera.cc:123-150: Undistort(Distort(n)) == n for nine image positions, 1e-5
    rel = np.array([[0, 0], [1, 1], [0, 1], [1, 0], [.5, .5], [.1, .2], [.8, .9], [.5, .6], [.1, .9]], np.float32)
    for model, p in cameras(oracle):
        fx, fy, cx, cy = base_intrinsics(p, model, orc)
        fxi, fyi = np.float32(1.0 / fx), np.float32(1.0 / fy)
        cxi, cyi = np.float32(-1.0 * cx / fx), np.float32(-1.0 * cy / fy)
        n = np.stack([fxi * (rel[:, 0] * W) + cxi, fyi * (rel[:, 1] * H) + cyi], 1)
        back = orc.cam_eval(model, W, H, p, "undistort", orc.cam_eval(model, W, H, p, "distort", n))
        assert np.abs(back - n).max() < 1e-5, (model, back, n)


def test_image_derivative_by_world(oracle):
    orc = oracle
    # test_camera.cc:206-268: analytic ImageDerivativeByWorld vs central differences (step 1e-3), tolerance 250e-3
    pts = np.array([[0.0, 0.0, 3.0], [1.0, 3.0, 8.0], [-0.1, 0.7, -0.8]], np.float32)
    k = np.float32(0.001)
    for model, p in cameras(oracle):
        ana = orc.cam_eval(model, W, H, p, "d_by_world", pts).reshape(-1, 2, 3)
        for i, at in enumerate(pts):
            if model == orc.CAM_FOV and at[0] == 0 and at[1] == 0:
                continue                                       # test_camera.cc:262-265
            num = np.zeros((2, 3), np.float32)
            for a in range(3):
                plus, minus = at.copy(), at.copy()
                plus[a] += k; minus[a] -= k
                pp = orc.cam_eval(model, W, H, p, "project", [[plus[0] / plus[2], plus[1] / plus[2]]])[0]
                pm = orc.cam_eval(model, W, H, p, "project", [[minus[0] / minus[2], minus[1] / minus[2]]])[0]
                num[:, a] = (pp - pm) / (2 * k)
            if not np.all(np.isfinite(num)):
                continue   # beyond the cut-off radius for this camera (the reference comments such a point out, :259)
            assert np.abs(ana[i] - num).max() < 0.25, (model, at, ana[i], num)


def test_image_derivative_by_intrinsics(oracle):
    orc = oracle
    # test_camera.cc:282-346: analytic ImageDerivativeByIntrinsics vs central differences over re-constructed cameras, 2.5e-3
    pts = np.array([[0.0, 0.0, 3.0], [1.0, 2.5, 4.0], [1.0, 3.0, 8.0], [-0.1, 0.4, 0.8]], np.float32)
    for model, p in cameras(oracle):
        npar = len(p)
        n = pts[:, :2] / pts[:, 2:3]
        pix = orc.cam_eval(model, W, H, p, "project", n)
        inside = (pix[:, 0] >= 0) & (pix[:, 1] >= 0) & (pix[:, 0] < W) & (pix[:, 1] < H)
        ana = orc.cam_eval(model, W, H, p, "d_by_intrinsics", pts).reshape(-1, 2, npar)
        step = np.float32(0.01)
        for c in range(npar):
            pp, pm = np.array(p, np.float32), np.array(p, np.float32)
            pp[c] += step; pm[c] -= step
            num = (orc.cam_eval(model, W, H, pp, "project", n) - orc.cam_eval(model, W, H, pm, "project", n)) / (2 * step)
            for i in range(len(pts)):
                if inside[i]:
                    assert abs(ana[i, 0, c] - num[i, 0]) < 2.5e-3 and abs(ana[i, 1, c] - num[i, 1]) < 2.5e-3, (model, i, c, ana[i, :, c], num[i])
        assert inside.sum() >= 2


def test_cutoff_placement(oracle):
    orc = oracle
    # InitCutoff (camera_base_impl.h:410-462): every border pixel unprojects below the cut-off, and the cut-off is tight (1.01x the
    # largest border radius unless a second solution caps it). Only the INNER model of the benchmark camera carries one.
    own, inner = orc.cam_cutoff(orc.CAM_BENCHMARK, W, H, BENCH)
    assert own == float("inf") and np.isfinite(inner)
    tp_own, tp_inner = orc.cam_cutoff(orc.CAM_THIN_PRISM, W, H, BENCH)
    assert tp_own == inner and tp_inner == float("inf")
    assert orc.cam_cutoff(orc.CAM_PINHOLE, W, H, PINHOLE) == (float("inf"), float("inf"))
    fx, fy, cx, cy = BENCH[:4]
    border = np.array([[x, y] for x in range(0, W, 7) for y in (0, H - 1)] + [[x, y] for y in range(0, H, 7) for x in (0, W - 1)], np.float32)
    n = np.stack([np.float32(1.0 / fx) * border[:, 0] + np.float32(-cx / fx), np.float32(1.0 / fy) * border[:, 1] + np.float32(-cy / fy)], 1)
    und = orc.cam_eval(orc.CAM_THIN_PRISM, W, H, BENCH, "undistort", n)
    r2 = (und ** 2).sum(1)
    assert r2.max() <= tp_own <= 1.0101 * r2.max() * 1.02
    # points beyond the cut-off project to infinity, just inside they stay finite
    r_in, r_out = np.sqrt(tp_own) * 0.999, np.sqrt(tp_own) * 1.001
    assert np.all(np.isfinite(orc.cam_eval(orc.CAM_THIN_PRISM, W, H, BENCH, "project", [[r_in, 0]])))
    assert np.all(np.isinf(orc.cam_eval(orc.CAM_THIN_PRISM, W, H, BENCH, "project", [[r_out, 0.0001]])))
    # fisheye: the cut-off applies to theta = atan(r)
    t_in, t_out = np.tan(np.sqrt(inner) * 0.999), np.tan(np.sqrt(inner) * 1.001)
    assert np.all(np.isfinite(orc.cam_eval(orc.CAM_BENCHMARK, W, H, BENCH, "project", [[t_in, 0]])))
    assert np.all(np.isinf(orc.cam_eval(orc.CAM_BENCHMARK, W, H, BENCH, "project", [[t_out, 0.0001]])))


def test_pyramid_scaling_matches_pinhole_rule(oracle):
    orc = oracle
    # ScaledBy(0.5) (camera_base_impl.h:70-89) keeps the distortion and halves f; (c+0.5)/2-0.5. Checked through projection:
    # a normalized point must land at (x+0.5)/2-0.5 of its full-resolution pixel.
    reg = orc.Registration(orc.reg_default_params(max_initial_image_area_in_pixels=W * H // 16))
    i = reg.add_intrinsics(W, H, BENCH, camera_model=orc.CAM_BENCHMARK)
    img = np.zeros((H, W), np.uint8)
    reg.add_image(i, img, None, [0, 0, 0, 1, 0, 0, 0])
    assert reg.initialize() == 3
    n = np.array([[0.2, -0.1], [0.0, 0.0], [-0.4, 0.3]], np.float32)
    full = orc.cam_eval(orc.CAM_BENCHMARK, W, H, BENCH, "project", n)
    half_p = np.array(BENCH, np.float32); half_p[0] *= 0.5; half_p[1] *= 0.5
    half_p[2] = np.float32(0.5) * (half_p[2] + np.float32(0.5)) - np.float32(0.5); half_p[3] = np.float32(0.5) * (half_p[3] + np.float32(0.5)) - np.float32(0.5)
    half = orc.cam_eval(orc.CAM_BENCHMARK, W // 2, H // 2, half_p, "project", n)
    assert np.abs(half - ((full + 0.5) / 2 - 0.5)).max() < 1e-3


def test_radial_and_closed_form_cutoffs(oracle):
    """RadialBase::InitCutoff (camera_base_impl_radial.h:143-171) and SimpleRadialCamera::InitCutoff (camera_simple_radial.cc:53-57)."""
    orc = oracle
    # closed form: k < 0 -> -1 / (3 k) (where d(distorted r)/dr = 1 + 3 k r^2 changes sign); k >= 0 -> no cut-off
    assert orc.cam_cutoff(orc.CAM_SIMPLE_RADIAL, W, H, [450.0, 319.5, 239.5, -0.2]) == (np.float32(-1.0) / (np.float32(3) * np.float32(-0.2)), float("inf"))
    assert orc.cam_cutoff(orc.CAM_SIMPLE_RADIAL, W, H, [450.0, 319.5, 239.5, K1]) == (float("inf"), float("inf"))
    own, inner = orc.cam_cutoff(orc.CAM_SIMPLE_RADIAL_FISHEYE, W, H, [450.0, 319.5, 239.5, -0.2])
    assert own == float("inf") and inner == np.float32(-1.0) / (np.float32(3) * np.float32(-0.2))      # the fisheye camera's INNER model carries it
    # radial search: the cut-off lies just above the undistorted radius of the farthest corner (x 1.01 on the square), or at the
    # second solution if that is nearer; projections are finite inside and infinite outside
    for model, p in [(orc.CAM_RADIAL, [250.0, 319.5, 239.5, K1, -1e-2]), (orc.CAM_POLYNOMIAL, PINHOLE + [K1, K2, K3]),
                     (orc.CAM_POLYNOMIAL_4, PT[:4] + [0.221184, 0.128597, 0.0623079, 0.20419])]:
        own, inner = orc.cam_cutoff(model, W, H, p)
        assert np.isfinite(own) and inner == float("inf"), (model, own, inner)
        fx, fy, cx, cy = base_intrinsics(p, model, orc)
        corners = np.array([[0, 0], [0, H], [W, 0], [W, H]], np.float32)
        n = np.stack([np.float32(1.0 / fx) * corners[:, 0] + np.float32(-cx / fx), np.float32(1.0 / fy) * corners[:, 1] + np.float32(-cy / fy)], 1)
        far = n[np.argmax((n ** 2).sum(1))][None]
        und = orc.cam_eval(model, W, H, p, "undistort", far)
        r2 = float((und ** 2).sum())
        if np.isfinite(r2) and np.abs(orc.cam_eval(model, W, H, p, "distort", und) - far).max() < 1e-4:
            assert own <= 1.0101 * r2 * 1.001, (model, own, r2)
        r_in, r_out = np.sqrt(own) * 0.999, np.sqrt(own) * 1.001
        assert np.all(np.isfinite(orc.cam_eval(model, W, H, p, "project", [[r_in, 0]])))
        assert np.all(np.isinf(orc.cam_eval(model, W, H, p, "project", [[r_out, 1e-4]])))
    # the fisheye wrappers apply the inner cut-off to theta = atan(r)
    own, inner = orc.cam_cutoff(orc.CAM_RADIAL_FISHEYE, W, H, [250.0, 319.5, 239.5, -K1, -K2])
    assert own == float("inf") and np.isfinite(inner)
    t_in, t_out = np.tan(min(np.sqrt(inner) * 0.999, 1.5)), np.tan(min(np.sqrt(inner) * 1.001, 1.55))
    assert np.all(np.isfinite(orc.cam_eval(orc.CAM_RADIAL_FISHEYE, W, H, [250.0, 319.5, 239.5, -K1, -K2], "project", [[t_in, 0]])))
    if np.sqrt(inner) * 1.001 < 1.55:
        assert np.all(np.isinf(orc.cam_eval(orc.CAM_RADIAL_FISHEYE, W, H, [250.0, 319.5, 239.5, -K1, -K2], "project", [[t_out, 1e-4]])))
