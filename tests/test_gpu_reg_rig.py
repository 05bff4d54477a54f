"""GPU parity of Path B with camera rigs (opt::Rig / opt::RigImages: dependent images share the reference image's pose block and
add a 6-variable extrinsics block per rig camera) against the oracle, whose rig Jacobians are pinned on the reference's
ComputePointIntensityAndJacobiansForRig test (tests/test_oracle_reg.py)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _pair(oracle, model):
    import dataset_pipeline_b200 as b2
    from dataset_pipeline_b200 import registration as R
    from dataset_pipeline_b200.synth import reg_scene
    sc = reg_scene.make_rig_scene(camera_model=model)
    area = 320 * 240 // 4
    g = b2.Registration(R.default_params(max_initial_image_area_in_pixels=area))
    o = oracle.Registration(oracle.reg_default_params(max_initial_image_area_in_pixels=area))
    assert reg_scene.load_rig_into(g, sc) == reg_scene.load_rig_into(o, sc)
    g.set_image_scale(0); o.set_image_scale(0)
    return g, o, sc


def test_variable_layout_and_dependent_poses(oracle):
    g, o, sc = _pair(oracle, 4)
    assert g.num_variables() == o.num_variables() == 4 + 6 + 2 * 6
    for kind, n in (("intrinsics", 1), ("rig", 1), ("image", 4)):
        for i in range(n + 1):
            assert g.variable_index(kind, i) == o.variable_index(kind, i), (kind, i)
    assert [g.variable_index("image", i) for i in range(4)] == [10, 10, 16, 16]      # dependent images use their reference's block
    gi, gp = g.get_state(); oi, op = o.get_state()
    assert np.array_equal(gp, op) and np.array_equal(gi, oi)                            # dependent poses = image_T_rig * reference pose
    assert np.array_equal(g.get_rigs(), o.get_rigs()) and np.array_equal(g.get_rigs(), sc["rig_init"])


@pytest.mark.parametrize("k12", ["f64", "f32"])      # K12 arithmetic: see tests/test_gpu_reg_camera.py::K12_TOL
@pytest.mark.parametrize("model", [4, 14, 5])
def test_rig_jacobians_and_normal_equations(oracle, model, k12, monkeypatch):
    monkeypatch.setenv("B2_K12", k12)
    g, o, _ = _pair(oracle, model)
    exact = True
    npar = 4 if model == 4 else 12
    g.CreateObservationsForAllImages(1); o.create_observations(1)
    if exact:
        for im in range(4):
            for ps in range(3):
                go, oo = g.observations(im, ps), o.observations(im, ps)
                assert all(np.array_equal(a, b) for a, b in zip(go, oo))
                if len(oo[0]) == 0:
                    continue
                I, jK, jP, jR = g.point_jacobians_rig(im, ps)
                for k in np.linspace(0, len(oo[0]) - 1, 12).astype(int):
                    rI, rK, rP, rR = o.point_jacobians_rig(im, ps, int(k), np_intr=npar)
                    assert I[k] == rI and np.array_equal(jK[k], rK) and np.array_equal(jP[k], rP) and np.array_equal(jR[k], rR), (im, ps, k)
                if im % 2 == 0:
                    assert not jR.any()                 # reference images have no extrinsics Jacobian
                else:
                    assert np.abs(jR).max() > 0
    g.ColorOptimizerApply(); o.color_update()
    Hg, bg, sg, cg = g.accumulate(); Ho, bo, so, co = o.accumulate()
    nv = npar + 6 + 12
    assert Hg.shape == Ho.shape == (nv, nv)
    tol = {"f64": 1e-6, "f32": 1e-9}[k12]    # f32: fp64 summation order only; f64: + rounding of the reference's fp32 products
    assert np.abs(Hg - Ho).max() <= tol * np.abs(Ho).max() and np.abs(bg - bo).max() <= tol * np.abs(bo).max()
    assert abs(cg - co) <= tol * abs(co)
    # every block of the rig structure is populated: K-R, R-R, R-P(reference)
    r0, p0 = npar, npar + 6
    assert np.abs(Hg[:npar, r0:r0 + 6]).max() > 0 and np.abs(Hg[r0:r0 + 6, r0:r0 + 6]).max() > 0 and np.abs(Hg[r0:r0 + 6, p0:p0 + 12]).max() > 0
    assert np.array_equal(Hg, Hg.T)


@pytest.mark.parametrize("model", [4, 5])
def test_rig_lm_steps(oracle, model):
    g, o, sc = _pair(oracle, model)
    ng, cg, okg = g.RunOnCurrentScale(4, 1e-9, 100)
    no, co, oko = o.run_on_current_scale(4, 1e-9, 100)
    assert ng == no and okg == oko
    rel = lambda a, b: np.linalg.norm(np.asarray(a, np.float64) - np.asarray(b, np.float64)) / np.linalg.norm(np.asarray(b, np.float64))
    gi, gp = g.get_state(); oi, op = o.get_state()
    assert rel(gp, op) < 1e-5 and rel(gi, oi) < 1e-5 and rel(g.get_rigs(), o.get_rigs()) < 1e-5     # north_star tolerance on the updates
    assert abs(cg - co) <= 1e-5 * abs(co)
    # dependent images stay tied to the rig: pose(cam1) == image_T_rig[1] * pose(cam0) after the optimisation (checked through the oracle's algebra)
    rig = g.get_rigs()[1]
    for a, b in sc["rig_sets"]:
        q, t = oracle.se3_exp_left_mul(np.zeros(6), gp[a][:4], gp[a][4:])      # identity * pose: normalises like the product does
        assert np.allclose(gp[b][4:], _rot(rig[:4]) @ gp[a][4:] + rig[4:], atol=1e-5)
    # and the optimisation moved the extrinsics
    assert np.abs(g.get_rigs()[1] - sc["rig_init"][1]).max() > 1e-6


def _rot(q):
    x, y, z, w = [float(v) for v in q]
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def test_rig_argument_errors():
    import dataset_pipeline_b200 as b2
    g = b2.Registration()
    i = g.add_intrinsics(64, 48, [50, 50, 32, 24])
    ident = [0, 0, 0, 1, 0, 0, 0]
    a = g.add_image(i, np.zeros((48, 64), np.uint8), None, ident)
    b = g.add_image(i, np.zeros((48, 64), np.uint8), None, ident)
    with pytest.raises(Exception):
        g.add_rig([ident])                                   # single cameras get no rig (rig.cc:31-34)
    r = g.add_rig([ident, [0, 0, 0, 1, 0.1, 0, 0]])
    with pytest.raises(Exception):
        g.add_rig_images(r, [a, a])                          # one image per camera
    g.add_rig_images(r, [a, b])
    with pytest.raises(Exception):
        g.add_rig_images(r, [a, b])                          # already assigned
