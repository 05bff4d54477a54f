"""GPU parity of the GroundTruthCreator kernels (b2_reg_gt_*: visibility counts per scan point, occlusion depth, ground-truth depth map, scan
rendering; /root/reference/src/exe/ground_truth_creator.cc:44-215) against the oracle. Counts, depth maps and renderings are bit-exact:
the depth map is a per-pixel minimum and the rendering keeps, per pixel, the last point of the reference's scan-order painting."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _scene(model, with_mask):
    from dataset_pipeline_b200.synth import reg_scene
    sc = reg_scene.make_scene(num_images=3, width=320, height=240, fx=260.0, camera_model=model, num_scales=1, base_radius=0.004)
    rng = np.random.default_rng(12)
    step = 0.009
    gx, gy = np.meshgrid(np.arange(-1.3, 1.3, step), np.arange(-1.0, 1.0, step), indexing="xy")
    x = gx.ravel() + rng.uniform(-0.3, 0.3, gx.size) * step; y = gy.ravel() + rng.uniform(-0.3, 0.3, gx.size) * step
    z = rng.normal(0, 2e-4, gx.size)
    z[::7] -= 0.3                       # a layer of points 30 cm behind the surface: occluded by the splats of the front layer
    xyz = np.stack([x, y, z], 1).astype(np.float32)
    rgb = rng.integers(0, 256, (len(xyz), 3)).astype(np.uint8)
    masks = None
    if with_mask:
        masks = []
        for i in range(3):
            m = np.zeros((240, 320), np.uint8); m[40:90, 60 + 20 * i:160] = 2; m[150:170, 200:260] = 1      # kEvalObs hides points, kObs does not
            masks.append(m)
    front = xyz[np.arange(len(xyz)) % 7 != 0]
    return sc, xyz, rgb, masks, front


def _load(reg, sc, masks, splats):
    w, h, K = sc["intr"]
    reg.add_intrinsics(w, h, K, camera_model=sc["camera_model"])
    for i, (img, T) in enumerate(zip(sc["images"], sc["poses_gt"])):
        reg.add_image(0, img, masks[i] if masks else None, T)
    reg.initialize()
    reg.set_splat_points(splats)


@pytest.mark.parametrize("model,with_mask", [(4, False), (4, True), (5, True)])
def test_ground_truth_creator(oracle, model, with_mask):
    import dataset_pipeline_b200 as b2
    from dataset_pipeline_b200 import registration as R
    sc, xyz, rgb, masks, front = _scene(model, with_mask)
    g = b2.Registration(R.default_params()); o = oracle.Registration(oracle.reg_default_params())
    _load(g, sc, masks, front); _load(o, sc, masks, front)
    cg = np.zeros(len(xyz), np.int32); co = np.zeros(len(xyz), np.int32)
    for im in range(3):
        cg = g.AccumulateScanObservationsForImage(im, xyz, cg); co = o.gt_accumulate_observations(im, xyz, co)
    assert np.array_equal(cg, co)
    # the layer behind the surface is occluded in the two images that look down on it (image 0 of the synthetic scene looks at the plane
    # from the other side), so its points stay below the two observations CreateGroundTruthForImage requires
    assert (co >= 2).mean() > 0.3 and (co[::7] <= 1).mean() > 0.9
    rng = np.random.default_rng(1)
    for im in range(3):
        base = rng.integers(0, 256, (240, 320, 3)).astype(np.uint8)
        occ_g, gt_g, ren_g = g.CreateGroundTruthForImage(im, xyz, rgb, cg, 2, base)
        occ_o, gt_o, ren_o = o.gt_create(im, xyz, rgb, co, 2, (320, 240), base)
        assert np.array_equal(occ_g, occ_o)
        assert np.array_equal(gt_g, gt_o) and np.isfinite(gt_o).mean() > 0.2
        assert np.array_equal(ren_g, ren_o) and (ren_o != base).any() and (ren_o == base).any()
    # depth map only, no rendering, radius 0
    occ_g, gt_g, ren_g = g.CreateGroundTruthForImage(0, xyz, None, cg, 0, None)
    assert ren_g is None and np.array_equal(gt_g, o.gt_create(0, xyz, rgb, co, 0, (320, 240))[1])
