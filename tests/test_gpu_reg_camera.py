"""GPU parity of Path B with the distorted camera models (THIN_PRISM, BENCHMARK = thin-prism fisheye) against the oracle.
The radius cut-off search, projection, derivatives, observation sets and Jacobians must be bit-identical (fp32, same evaluation
order; the fisheye atan is the correctly rounded value on both sides, oracle/orc_camera.h); normal equations agree to the fp64
summation order; states after LM steps within the north_star tolerance (1e-5 relative)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

W, H = 640, 480
BENCH = [340.926, 341.124, 302.4, 201.6, 0.221184, 0.128597, 0.000531602, -0.000388873, 0.0623079, 0.20419, -0.000805024, 4.07704e-05]
PINHOLE = [250.0, 200.0, 319.5, 239.5]


def _b2():
    import dataset_pipeline_b200 as b2
    from dataset_pipeline_b200 import registration
    return b2, registration


def _grid_points(n=4000, seed=3):
    rng = np.random.default_rng(seed)
    nrm = rng.uniform(-1.3, 1.3, (n, 2)).astype(np.float32)
    z = rng.uniform(0.5, 6.0, (n, 1)).astype(np.float32)
    return nrm, np.concatenate([nrm * z, z], 1).astype(np.float32)


def test_cutoff_search_bit_exact(oracle):
    _, R = _b2()
    for (w, h, p) in [(W, H, BENCH), (320, 240, [170.4, 170.6, 151.0, 100.5] + BENCH[4:]), (161, 97, [90.0, 88.0, 80.2, 48.1, -0.12, 0.03, 1e-3, -2e-3, 0.01, 0.0, 1e-3, 2e-3])]:
        for model in (oracle.CAM_THIN_PRISM, oracle.CAM_BENCHMARK):
            _, cut = R.camera_eval(model, w, h, p, "cutoff")
            assert cut == oracle.cam_cutoff(model, w, h, p), (model, w, h)
    _, cut = R.camera_eval(oracle.CAM_PINHOLE, W, H, PINHOLE, "cutoff")
    assert cut == (float("inf"), float("inf"))


def test_projection_and_derivatives(oracle):
    _, R = _b2()
    nrm, pts = _grid_points()
    for model, p, exact in [(oracle.CAM_PINHOLE, PINHOLE, True), (oracle.CAM_THIN_PRISM, BENCH, True), (oracle.CAM_BENCHMARK, BENCH, True)]:
        for op, x in [("project", nrm), ("d_by_world", pts), ("d_by_intrinsics", pts)]:
            got, _ = R.camera_eval(model, W, H, p, op, x)
            ref = oracle.cam_eval(model, W, H, p, op, x)
            fin = np.isfinite(ref)
            assert np.array_equal(np.isfinite(got), fin), (model, op)          # same points beyond the cut-off
            assert fin.mean() > 0.3
            if exact:
                assert np.array_equal(got[fin], ref[fin]), (model, op)
            else:
                # 1-ulp atan differences (5 % of the calls) propagate to a few ulp of the largest magnitudes involved:
                # <= 5e-4 px on pixel coordinates of up to ~10^3, and 2e-6 of the largest derivative entry
                err = np.abs(got[fin] - ref[fin]).max()
                assert err <= (5e-4 if op == "project" else 2e-6 * np.abs(ref[fin]).max()), (model, op, err)


def _pair(oracle, model, **kw):
    b2, R = _b2()
    from dataset_pipeline_b200.synth import reg_scene
    sc = reg_scene.make_scene(num_images=2, width=320, height=240, fx=260.0, camera_model=model, num_scales=3, base_radius=0.004)
    area = 320 * 240 // 4
    g = b2.Registration(R.default_params(max_initial_image_area_in_pixels=area, **kw))
    o = oracle.Registration(oracle.reg_default_params(max_initial_image_area_in_pixels=area, **kw))
    assert reg_scene.load_into(g, sc) == reg_scene.load_into(o, sc)
    return g, o


# K12 arithmetic (b2_reg_kernels.cuh, kr_accumulate): "f64" = default, products formed in fp64 from the fp32 Jacobian differences with
# the fixed / variable weights merged; "f32" (B2_K12=f32) = the reference's own order, every product rounded to fp32 first. The
# second must match the oracle to the fp64 summation order (1e-9); the first differs by the rounding of the individual fp32 products
# (<= 2 ulp_fp32 each, averaging out over the sum) and is held to 1e-6 — the north_star tolerance on the resulting updates is 1e-5.
K12_TOL = {"f64": 1e-6, "f32": 1e-9}


@pytest.mark.parametrize("k12", ["f64", "f32"])
@pytest.mark.parametrize("model", [14, 5])
def test_observations_jacobians_normal_equations(oracle, model, k12, monkeypatch):
    monkeypatch.setenv("B2_K12", k12)
    g, o = _pair(oracle, model)
    exact = True
    g.set_image_scale(0); o.set_image_scale(0)
    g.CreateObservationsForAllImages(1); o.create_observations(1)
    total = 0
    for im in range(2):
        for ps in range(3):
            gi, gx, gy, gs, gn = g.observations(im, ps)
            oi, ox, oy, os_, on = o.observations(im, ps)
            if exact:
                assert np.array_equal(gi, oi) and np.array_equal(gx, ox) and np.array_equal(gy, oy) and np.array_equal(gs, os_) and np.array_equal(gn, on)
            else:
                # a 1-ulp shift of a projection can move a point across a pixel / scale boundary: allow a handful
                common, ga, oa = np.intersect1d(gi, oi, return_indices=True)
                assert len(gi) - len(common) <= 3 and len(oi) - len(common) <= 3, (im, ps, len(gi), len(oi), len(common))
                assert np.abs(gx[ga] - ox[oa]).max(initial=0) < 1e-3 and np.abs(gs[ga] - os_[oa]).max(initial=0) < 1e-4
            total += len(oi)
            if len(oi) and exact:
                I, jK, jP = g.point_jacobians(im, ps)
                for k in np.linspace(0, len(oi) - 1, 25).astype(int):
                    rI, rK, rP = o.point_jacobians(im, ps, int(k), np_intr=12)
                    assert I[k] == rI and np.array_equal(jK[k], rK) and np.array_equal(jP[k], rP), (im, ps, k)
    assert total > 20000
    g.ColorOptimizerApply(); o.color_update()
    Hg, bg, sg, cg = g.accumulate()
    Ho, bo, so, co = o.accumulate()
    assert Hg.shape == (24 + 12, 24 + 12) or Hg.shape == (12 + 12, 12 + 12)
    # fisheye: the entries are sums of products of Jacobian DIFFERENCES (neighbour - centre, ~1e-3 of the Jacobians themselves), so the
    # few-ulp projection differences of the atan() implementations show up ~1e3 times larger here; the LM test below holds the states
    # to 1e-5 all the same. Thin prism (no transcendental) must agree to fp64 summation order.
    tol = K12_TOL[k12] if exact else 1e-3
    assert np.abs(Hg - Ho).max() <= tol * np.abs(Ho).max()
    assert np.abs(bg - bo).max() <= tol * np.abs(bo).max()
    assert abs(cg - co) <= tol * abs(co)
    if exact:
        assert sg[1] == so[1] and sg[3] == so[3]


@pytest.mark.parametrize("model", [14, 5])
def test_lm_step_and_outer_loop(oracle, model):
    g, o = _pair(oracle, model)
    g.set_image_scale(0); o.set_image_scale(0)
    ng, cg, okg = g.RunOnCurrentScale(4, 0.0, 100)
    no, co, oko = o.run_on_current_scale(4, 0.0, 100)
    assert ng == no and oko == okg
    gi, gp = g.get_state(); oi, op = o.get_state()
    rel = lambda a, b: np.linalg.norm(np.asarray(a, np.float64) - np.asarray(b, np.float64)) / np.linalg.norm(np.asarray(b, np.float64))
    # pose update within 1e-5 relative Frobenius (north_star tolerance); intrinsics likewise
    assert rel(gp, op) < 1e-5, rel(gp, op)
    assert rel(gi, oi) < 1e-5, rel(gi, oi)
    assert abs(cg - co) <= 1e-5 * abs(co)
    # the optimisation did something: intrinsics moved away from the initial values
    assert np.abs(np.asarray(oi).reshape(-1)[:4] - np.array([260.0, 260.0, 159.5, 119.5])).max() > 1e-4


def test_mixed_models_variable_layout(oracle):
    """Two intrinsics with different parameter counts: the variable vector is [4 | 12 | 6 | 6] (ids in order)."""
    b2, R = _b2()
    from dataset_pipeline_b200.synth import reg_scene
    sp = reg_scene.make_scene(num_images=1, width=320, height=240, fx=260.0, camera_model=4, num_scales=3, base_radius=0.004)
    sb = reg_scene.make_scene(num_images=2, width=320, height=240, fx=260.0, camera_model=5, num_scales=3, base_radius=0.004)
    area = 320 * 240 // 4
    g = b2.Registration(R.default_params(max_initial_image_area_in_pixels=area))
    o = oracle.Registration(oracle.reg_default_params(max_initial_image_area_in_pixels=area))
    for reg in (g, o):
        a = reg.add_intrinsics(320, 240, sp["intr"][2], camera_model=4)
        b = reg.add_intrinsics(320, 240, sb["intr"][2], camera_model=5)
        reg.add_image(a, sp["images"][0], None, sp["poses_init"][0])
        reg.add_image(b, sb["images"][1], None, sb["poses_init"][1])
        reg.initialize()
        for xyz, radius, nbr, colors in sp["scales"]:
            reg.add_point_scale(xyz, float(radius), nbr, colors)
        reg.set_image_scale(0)
    g.CreateObservationsForAllImages(1); o.create_observations(1)
    g.ColorOptimizerApply(); o.color_update()
    Hg, bg, _, cg = g.accumulate(); Ho, bo, _, co = o.accumulate()
    assert Hg.shape == Ho.shape == (28, 28)
    assert np.abs(Hg - Ho).max() <= K12_TOL["f64"] * np.abs(Ho).max() and np.abs(bg - bo).max() <= K12_TOL["f64"] * np.abs(bo).max()
    # the pinhole image's rows couple only to its own intrinsics block [0,4) and pose block [16,22)
    assert np.all(Hg[0:4, 4:16] == 0) and np.all(Hg[0:4, 22:28] == 0) and np.abs(Hg[0:4, 16:22]).max() > 0
    ip, _ = g.get_state()
    assert ip.shape == (16,)


def test_rejects_unknown_model_and_wrong_count():
    b2, R = _b2()
    g = b2.Registration()
    with pytest.raises(Exception):
        g.add_intrinsics(64, 48, [50, 50, 32, 24, 0.1], camera_model=15)         # not a camera::CameraBase::Type
    with pytest.raises(Exception):
        g.add_intrinsics(64, 48, [50, 50, 32, 24], camera_model=8)                # RADIAL is f cx cy k1 k2
    with pytest.raises(Exception):
        g.add_intrinsics(64, 48, [50, 50, 32, 24], camera_model=5)                # BENCHMARK needs 12


@pytest.mark.parametrize("model", [14, 5])
def test_mesh_occlusion_with_distorted_cameras(oracle, model):
    """K8/K9 with the renderer's vertex-stage distortion (opengl/renderer.cc:630-653): depth maps (incl. boundary masking) and the
    observation sets behind them stay bit-identical to the oracle's software rasteriser. This is the configuration of the real ETH3D
    pipeline (THIN_PRISM_FISHEYE images + mesh occlusion geometry)."""
    b2, R = _b2()
    from dataset_pipeline_b200.synth import reg_scene
    import os, sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from mesh_util import _grid_mesh, _box_mesh
    sc = reg_scene.make_scene(num_images=2, width=320, height=240, fx=260.0, camera_model=model, num_scales=3, base_radius=0.004)
    Vp, Fp = _grid_mesh(-1.3, 1.3, -1.0, 1.0, 0.0, 30)                 # the textured plane itself
    Vb, Fb = _box_mesh((0.2, 0.1, 0.5), 0.15)                          # a box above it: silhouettes + hidden points
    Vg = np.array([[-4, -3, 1.9], [4, -3, 1.9], [4, 3, 2.3], [-4, 3, 2.3]], np.float32); Fg = np.array([[0, 1, 2], [0, 2, 3]], np.uint32)   # passes behind the cameras' near plane
    V = np.concatenate([Vp, Vb, Vg]); F = np.concatenate([Fp, Fb + len(Vp), Fg + len(Vp) + len(Vb)])
    area = 320 * 240 // 4
    for mask_flag in (1, 0):
        kw = dict(max_initial_image_area_in_pixels=area, mask_occlusion_boundaries=mask_flag)
        g = b2.Registration(R.default_params(**kw)); o = oracle.Registration(oracle.reg_default_params(**kw))
        for r in (g, o):
            reg_scene.load_into(r, sc, splats=False)
            r.set_mesh(V, F); r.set_image_scale(0)
        for im in range(2):
            dg, sg = g.render_depth(im); do, so = o.render_depth(im)
            assert sg == so and np.array_equal(dg, do), (model, mask_flag, im, int((dg != do).sum()))
            assert (do > 0).mean() > 0.5
            if mask_flag:
                assert (do == -1).sum() > 200
        g.CreateObservationsForAllImages(1); o.create_observations(1)
        n = 0
        for im in range(2):
            for ps in range(3):
                go, oo = g.observations(im, ps), o.observations(im, ps)
                assert all(np.array_equal(a, b) for a, b in zip(go, oo)), (model, mask_flag, im, ps)
                n += len(oo[0])
        assert n > 10000
    # the distortion matters: a pinhole camera with the same f, c sees a different depth map
    gp = b2.Registration(R.default_params(max_initial_image_area_in_pixels=area, mask_occlusion_boundaries=0))
    gp.add_intrinsics(320, 240, sc["intr"][2][:4]); gp.add_image(0, sc["images"][0], None, sc["poses_init"][0]); gp.initialize()
    gp.set_mesh(V, F); gp.set_image_scale(0)
    dp, _ = gp.render_depth(0)
    g = b2.Registration(R.default_params(max_initial_image_area_in_pixels=area, mask_occlusion_boundaries=0))
    reg_scene.load_into(g, sc, splats=False); g.set_mesh(V, F); g.set_image_scale(0)
    dd, _ = g.render_depth(0)
    assert dd.shape == dp.shape and (np.abs(dd - dp) > 0.005).mean() > 0.01
