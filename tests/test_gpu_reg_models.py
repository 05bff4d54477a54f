"""GPU parity for the camera models beyond pinhole / thin prism through the C ABI: every model of the reference's
camera tests (/root/reference/src/camera/test/test_camera.cc:422-505) — cut-off searches, projection and its derivatives bit-exact
against the oracle (oracle/orc_camera.h, pinned on those tests in tests/test_oracle_camera.py) — and Path B (observations, Jacobians,
normal equations, LM steps) with each distortion family, including the single-focal-length parameter layouts."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

W, H = 640, 480
PINHOLE = [250.0, 200.0, 319.5, 239.5]
K1, K2, K3 = 0.13, -0.66, 0.64
PT = [340.926, 341.124, 302.4, 201.6, -0.101082, 0.0703954, 0.000438661, -0.000680887]
BENCH = [340.926, 341.124, 302.4, 201.6, 0.221184, 0.128597, 0.000531602, -0.000388873, 0.0623079, 0.20419, -0.000805024, 4.07704e-05]
# (model, GetParameters vector): the cameras of test_camera.cc + the inner models of the fisheye cameras
CAMERAS = [(4, PINHOLE), (7, [250.0, 319.5, 239.5]), (8, [250.0, 319.5, 239.5, K1, -1e-2]), (12, [250.0, 319.5, 239.5, -K1, -K2]),
           (9, [450.0, 319.5, 239.5, K1]), (9, [450.0, 319.5, 239.5, -0.2]), (13, [450.0, 319.5, 239.5, K1]), (1, PINHOLE + [K1, K2, K3]),
           (0, PINHOLE + [1.0]), (2, PT), (10, PT[:4] + [-0.101082, 0.0703954, 0.0438661, -0.0680887, -0.00101082, .1, .001, -.001]),
           (6, PT[:4] + [0.221184, 0.128597, 0.0623079, 0.20419]), (11, PT[:4] + [0.221184, 0.128597, 0.0623079, 0.20419]), (3, PT),
           (14, BENCH), (5, BENCH)]


def _b2():
    import dataset_pipeline_b200 as b2
    from dataset_pipeline_b200 import registration
    return b2, registration


def test_all_models_cutoff_projection_derivatives(oracle):
    _, R = _b2()
    rng = np.random.default_rng(5)
    nrm = rng.uniform(-1.3, 1.3, (3000, 2)).astype(np.float32)
    nrm[:5] = [[0, 0], [1e-7, 0], [0, -1e-7], [3.0, 3.0], [-2.5, 0.1]]
    z = rng.uniform(0.5, 6.0, (3000, 1)).astype(np.float32)
    pts = np.concatenate([nrm * z, z], 1).astype(np.float32)
    for model, p in CAMERAS:
        for (w, h, q) in [(W, H, p), (161, 97, None)]:
            if q is None:                                        # a quarter-size camera with the same distortion: different border, different cut-off
                q = list(p)
                nb = 3 if model in (7, 8, 9, 12, 13) else 4
                for i in range(nb):
                    q[i] = q[i] * 0.25
            _, cut = R.camera_eval(model, w, h, q, "cutoff")
            assert cut == oracle.cam_cutoff(model, w, h, q), (model, w, h, cut, oracle.cam_cutoff(model, w, h, q))
        for op, x in [("project", nrm), ("d_by_world", pts), ("d_by_intrinsics", pts)]:
            got, _ = R.camera_eval(model, W, H, p, op, x)
            ref = oracle.cam_eval(model, W, H, p, op, x)
            assert got.shape == ref.shape, (model, op)
            fin = np.isfinite(ref)
            assert np.array_equal(np.isfinite(got), fin), (model, op)
            assert np.array_equal(np.isnan(got), np.isnan(ref)), (model, op)
            if model == 0:
                # FOV: atan of the device's double library vs glibc's, rounded to fp32: equal except for rare double-rounding cases
                assert np.abs(got[fin] - ref[fin]).max() <= 2e-6 * max(1.0, np.abs(ref[fin]).max()), (model, op)
                assert (got[fin] != ref[fin]).mean() < 1e-3, (model, op)
            else:
                assert np.array_equal(got[fin], ref[fin]), (model, op, np.abs(got[fin] - ref[fin]).max())


# one camera per distortion family / parameter layout, scaled to the 320 x 240 test scene (focal length ~260, mild distortion)
SCENE_CAMERAS = {
    7: [260.0, 159.5, 119.5],                                                    # SIMPLE_PINHOLE: 3 parameters, single focal length
    9: [260.0, 159.5, 119.5, -0.08],                                             # SIMPLE_RADIAL: closed-form cut-off
    8: [260.0, 159.5, 119.5, 0.05, -0.01],                                       # RADIAL: RadialBase cut-off search
    13: [260.0, 159.5, 119.5, 0.03],                                             # SIMPLE_RADIAL_FISHEYE
    1: [260.0, 258.0, 159.5, 119.5, 0.04, -0.02, 0.005],                         # POLYNOMIAL
    0: [260.0, 258.0, 159.5, 119.5, 0.9],                                        # FOV
    2: [260.0, 258.0, 159.5, 119.5, -0.05, 0.02, 4e-4, -6e-4],                   # POLYNOMIAL_TANGENTIAL: generic cut-off search
    10: [260.0, 258.0, 159.5, 119.5, -0.05, 0.02, 4e-3, -6e-3, -1e-3, 0.05, 1e-3, -1e-3],   # FULL_OPENCV
    6: [260.0, 258.0, 159.5, 119.5, 0.02, 0.01, 0.005, 0.002],                   # FISHEYE_POLYNOMIAL_4
}


def _pair(oracle, model, **kw):
    b2, R = _b2()
    from dataset_pipeline_b200.synth import reg_scene
    sc = reg_scene.make_scene(num_images=2, width=320, height=240, camera_model=model, camera_params=SCENE_CAMERAS[model], num_scales=3, base_radius=0.004)
    area = 320 * 240 // 4
    g = b2.Registration(R.default_params(max_initial_image_area_in_pixels=area, **kw))
    o = oracle.Registration(oracle.reg_default_params(max_initial_image_area_in_pixels=area, **kw))
    assert reg_scene.load_into(g, sc) == reg_scene.load_into(o, sc)
    return g, o


@pytest.mark.parametrize("model", sorted(SCENE_CAMERAS))
def test_path_b_with_every_distortion_family(oracle, model):
    g, o = _pair(oracle, model)
    npar = len(SCENE_CAMERAS[model])
    exact = model != 0
    g.set_image_scale(0); o.set_image_scale(0)
    g.CreateObservationsForAllImages(1); o.create_observations(1)
    total = 0
    for im in range(2):
        for ps in range(3):
            gi, gx, gy, gs, gn = g.observations(im, ps)
            oi, ox, oy, os_, on = o.observations(im, ps)
            if exact:
                assert np.array_equal(gi, oi) and np.array_equal(gx, ox) and np.array_equal(gy, oy) and np.array_equal(gs, os_) and np.array_equal(gn, on), (im, ps)
            else:
                common, ga, oa = np.intersect1d(gi, oi, return_indices=True)
                assert len(gi) - len(common) <= 3 and len(oi) - len(common) <= 3, (im, ps, len(gi), len(oi), len(common))
            total += len(oi)
            if len(oi) and exact:
                I, jK, jP = g.point_jacobians(im, ps)
                assert jK.shape == (len(oi), npar)
                for k in np.linspace(0, len(oi) - 1, 25).astype(int):
                    rI, rK, rP = o.point_jacobians(im, ps, int(k), np_intr=npar)
                    assert I[k] == rI and np.array_equal(jK[k], rK) and np.array_equal(jP[k], rP), (im, ps, k, jK[k], rK)
    assert total > 20000
    g.ColorOptimizerApply(); o.color_update()
    Hg, bg, sg, cg = g.accumulate()
    Ho, bo, so, co = o.accumulate()
    assert Hg.shape == Ho.shape == (npar + 12, npar + 12)
    tol = 1e-6 if exact else 1e-3
    assert np.abs(Hg - Ho).max() <= tol * np.abs(Ho).max()
    assert np.abs(bg - bo).max() <= tol * np.abs(bo).max()
    assert abs(cg - co) <= tol * abs(co)
    # LM steps + the outer loop: same number of iterations, states within the north-star tolerance
    ng, c2g, okg = g.RunOnCurrentScale(3, 0.0, 100)
    no, c2o, oko = o.run_on_current_scale(3, 0.0, 100)
    assert ng == no and oko == okg
    gi, gp = g.get_state(); oi, op = o.get_state()
    rel = lambda a, b: np.linalg.norm(np.asarray(a, np.float64) - np.asarray(b, np.float64)) / np.linalg.norm(np.asarray(b, np.float64))
    assert np.asarray(gi).size == npar
    assert rel(gp, op) < 1e-5 and rel(gi, oi) < 1e-5, (rel(gp, op), rel(gi, oi))
    assert abs(c2g - c2o) <= 1e-5 * abs(c2o)


@pytest.mark.parametrize("model", [2, 13, 1])
def test_mesh_occlusion_with_further_models(oracle, model):
    """K8 / K9 with the vertex-stage distortion of the other renderer programs (opengl/renderer.cc:154-560: z * Distort(x/z, y/z), pushed
    out by 99 beyond the cut-off): depth maps and the observation sets behind them identical to the oracle's software rasteriser."""
    b2, R = _b2()
    from dataset_pipeline_b200.synth import reg_scene
    import os, sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from mesh_util import _grid_mesh, _box_mesh
    sc = reg_scene.make_scene(num_images=2, width=320, height=240, camera_model=model, camera_params=SCENE_CAMERAS[model], num_scales=3, base_radius=0.004)
    Vp, Fp = _grid_mesh(-1.3, 1.3, -1.0, 1.0, 0.0, 30)
    Vb, Fb = _box_mesh((0.2, 0.1, 0.5), 0.15)
    V = np.concatenate([Vp, Vb]); F = np.concatenate([Fp, Fb + len(Vp)])
    area = 320 * 240 // 4
    kw = dict(max_initial_image_area_in_pixels=area, mask_occlusion_boundaries=1)
    g = b2.Registration(R.default_params(**kw)); o = oracle.Registration(oracle.reg_default_params(**kw))
    for r in (g, o):
        reg_scene.load_into(r, sc, splats=False)
        r.set_mesh(V, F); r.set_image_scale(0)
    for im in range(2):
        dg, sg = g.render_depth(im); do, so = o.render_depth(im)
        assert sg == so and np.array_equal(dg, do), (model, im, int((dg != do).sum()))
        assert (do > 0).mean() > 0.5 and (do == -1).sum() > 200
    g.CreateObservationsForAllImages(1); o.create_observations(1)
    n = 0
    for im in range(2):
        for ps in range(3):
            go, oo = g.observations(im, ps), o.observations(im, ps)
            assert all(np.array_equal(a, b) for a, b in zip(go, oo)), (model, im, ps)
            n += len(oo[0])
    assert n > 10000


@pytest.mark.parametrize("model", [7, 9, 10, 0])
def test_min_max_point_radius_with_further_models(oracle, model):
    """ComputeMinMaxPointRadius (multi_scale_point_cloud.cc:126-184): ImageToNormalized through the undistortion lookup (generic
    IterativeUndistort per pixel), directly for SimplePinhole, in closed form for the FOV camera."""
    import dataset_pipeline_b200 as b2
    from dataset_pipeline_b200 import registration as R
    from dataset_pipeline_b200.synth import reg_scene
    sc = reg_scene.make_scene(num_images=3, width=320, height=240, camera_model=model, camera_params=SCENE_CAMERAS[model], num_scales=1, base_radius=0.004)
    rng = np.random.default_rng(8)
    gx, gy = np.meshgrid(np.arange(-1.3, 1.3, 0.011), np.arange(-1.0, 1.0, 0.011), indexing="xy")
    pts = np.stack([gx.ravel() + rng.uniform(-0.003, 0.003, gx.size), gy.ravel() + rng.uniform(-0.003, 0.003, gx.size), rng.normal(0, 2e-4, gx.size)], 1).astype(np.float32)
    area = 320 * 240 // 64
    g = b2.Registration(R.default_params(max_initial_image_area_in_pixels=area)); o = oracle.Registration(oracle.reg_default_params(max_initial_image_area_in_pixels=area))
    counts = []
    for reg in (g, o):
        w, h, K = sc["intr"]
        reg.add_intrinsics(w, h, K, camera_model=model)
        for img, T in zip(sc["images"], sc["poses_gt"]):
            reg.add_image(0, img, None, T)
        counts.append(reg.initialize())
        reg.set_splat_points(pts); reg.set_image_scale(0)
    assert counts[0] == counts[1] == 4
    msf = float(np.float32(2.0 ** (-(counts[0] - 1))))
    lg, hg = g.ComputeMinMaxPointRadius(pts, msf); lo, ho = o.min_max_point_radius(pts, msf)
    seen = np.isfinite(lo)
    assert np.array_equal(np.isfinite(lg), seen) and 0.3 < seen.mean() <= 1.0
    if model == 0:      # FOV: tan / atan of the device's double library vs glibc's, and a difference of two nearby rays (cancellation)
        assert np.allclose(lg[seen], lo[seen], rtol=1e-3, atol=0) and np.allclose(hg[seen], ho[seen], rtol=1e-3, atol=0)
    else:
        assert np.array_equal(lg, lo) and np.array_equal(hg, ho)
