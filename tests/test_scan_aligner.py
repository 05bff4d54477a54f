"""The ICPScanAligner tool body (host mirror dataset_pipeline_b200/scan_aligner.py of icp_scan_aligner.cc:274-375): the coarse-to-fine
schedule, checked on the CPU with the oracle as ICP / normals backend against a literal restatement of the tool's loop, and — `gpu` —
end to end through the library (normals K7 -> ICP K1..K6 at every scale) against the oracle-backed run."""
import math

import numpy as np
import pytest

from dataset_pipeline_b200 import scan_aligner as SA
from dataset_pipeline_b200 import synth


def _oracle_backends(oracle):
    return (lambda d: oracle.PointToPlaneICP(use_kdtree=True)), (lambda xyz, k: oracle.normals_knn(xyz, k, (0.0, 0.0, 0.0)))


def test_scale_schedule_follows_the_tools_arithmetic():
    # README.md:671-673 recipe: -d 0.01 --number_of_scales 4 (downscale_step 4, factor 2 by default)
    s = SA.scale_schedule(4, 0.01)
    assert [st for _, st in s] == [64, 16, 4, 1]
    assert [float(d) for d, _ in s] == [float(np.float32(8.0 * float(np.float32(0.01)))), float(np.float32(4.0 * float(np.float32(0.01)))),
                                        float(np.float32(2.0 * float(np.float32(0.01)))), float(np.float32(0.01))]
    assert SA.scale_schedule(1, 0.1) == [(np.float32(0.1), 1)]
    # int step = std::pow(3, 2) = 9; float distance = (float)(pow(1.5f, 2) * 0.02f)
    assert SA.scale_schedule(3, 0.02, 3, 1.5)[0] == (np.float32(math.pow(1.5, 2) * float(np.float32(0.02))), 9)


def test_rotation_of_is_the_orthogonal_polar_factor():
    rng = np.random.default_rng(2)
    Q, _ = np.linalg.qr(rng.normal(size=(3, 3)))
    if np.linalg.det(Q) < 0:
        Q[:, 0] = -Q[:, 0]
    T = np.eye(4); T[:3, :3] = Q @ (np.eye(3) + 1e-4 * rng.normal(size=(3, 3)))
    R = SA.rotation_of(T)
    assert np.allclose(R @ R.T, np.eye(3), atol=1e-6) and np.linalg.det(R) > 0.999
    assert np.allclose(R, Q, atol=5e-4)


def test_scale_loop_matches_a_literal_restatement_of_the_tool(oracle):
    clouds, start, gt = synth.room_scans(3, 160, 80)
    scans = [c[0] for c in clouds]
    icp_factory, normals = _oracle_backends(oracle)
    kw = dict(max_correspondence_distance=0.02, max_num_iterations=12, convergence_threshold_max_movement=1e-6,
              normal_estimation_neighbor_count=12, number_of_scales=3, downscale_step=2, search_distance_increase_factor_per_scale=2.0)
    poses, log = SA.align_scans(scans, start, fixed=[True, False, False], icp_factory=icp_factory, estimate_normals=normals, **kw)

    # the tool's loop, written out (icp_scan_aligner.cc:274-375)
    R = [np.asarray(T, np.float64)[:3, :3].copy() for T in start]
    t = [np.asarray(T, np.float64)[:3, 3].copy() for T in start]
    want_log = []
    for scale_index in range(3):
        dist = np.float32(math.pow(2.0, 3 - 1 - scale_index) * float(np.float32(0.02)))
        icp = oracle.PointToPlaneICP(use_kdtree=True)
        ids = []
        pts = []
        for i in range(3):
            xyz = scans[i]
            if scale_index < 2:
                xyz = xyz[::int(math.pow(2, 3 - 1 - scale_index))]
            nrm = oracle.normals_knn(np.ascontiguousarray(xyz), 12, (0.0, 0.0, 0.0))[:, :3]
            T = np.eye(4); T[:3, :3] = R[i]; T[:3, 3] = t[i]
            ids.append(icp.AddPointCloud(np.ascontiguousarray(xyz), np.ascontiguousarray(nrm), T.astype(np.float32), i == 0))
            pts.append(len(xyz))
        its, conv = 0, False
        for iteration in range(12):
            conv = icp.Run(float(dist), iteration, 1, 1e-6, False)
            its += 1
            for i in (1, 2):
                G = icp.GetResultGlobalTCloud(ids[i])
                R[i] = SA.rotation_of(G); t[i] = G[:3, 3].astype(np.float64)
            if conv:
                break
        want_log.append({"max_correspondence_distance": float(dist), "stride": 4 >> scale_index, "points": pts, "iterations": its, "converged": conv})
    assert log == want_log
    assert [e["stride"] for e in log] == [4, 2, 1] and log[0]["points"][0] == len(scans[0][::4])
    for i in range(3):
        assert np.array_equal(poses[i][:3, :3], R[i]) and np.array_equal(poses[i][:3, 3], t[i])
    assert np.array_equal(poses[0], np.asarray(start[0], np.float64))            # the fixed cloud keeps its pose
    # and the alignment does what the tool is for: the movable scans end closer to the scanner poses than they started
    for i in (1, 2):
        before = np.linalg.norm(np.asarray(start[i])[:3, 3] - np.asarray(start[0])[:3, 3] - (gt[i][:3, 3] - gt[0][:3, 3]))
        after = np.linalg.norm(poses[i][:3, 3] - poses[0][:3, 3] - (gt[i][:3, 3] - gt[0][:3, 3]))
        assert after < before


@pytest.mark.gpu
def test_scale_loop_through_the_library_lands_where_the_oracle_does(oracle):
    """Normals (K7) -> ICP (K1..K6) on the device at every scale, default backends of align_scans, against the same schedule on the oracle.
    The two runs are free-running (no pose resynchronisation), so they are compared where both converge, not bit for bit."""
    clouds, start, gt = synth.room_scans(3, 240, 120)
    scans = [c[0] for c in clouds]
    kw = dict(max_correspondence_distance=0.02, max_num_iterations=30, convergence_threshold_max_movement=1e-6,
              normal_estimation_neighbor_count=16, number_of_scales=3, downscale_step=2)
    got, glog = SA.align_scans(scans, start, fixed=[True, False, False], **kw)
    icp_factory, normals = _oracle_backends(oracle)
    want, wlog = SA.align_scans(scans, start, fixed=[True, False, False], icp_factory=icp_factory, estimate_normals=normals, **kw)
    assert [(e["max_correspondence_distance"], e["stride"], e["points"]) for e in glog] == \
           [(e["max_correspondence_distance"], e["stride"], e["points"]) for e in wlog]
    for i in range(3):
        # metres (10 m room, d = 2 cm). Normals perturbed by 1e-4 move the converged poses by ~1e-6 / 2e-7 on this scene: 50x margin
        assert np.abs(got[i][:3, 3] - want[i][:3, 3]).max() < 5e-5, (i, got[i][:3, 3], want[i][:3, 3])
        assert np.abs(got[i][:3, :3] - want[i][:3, :3]).max() < 2e-5
    assert np.array_equal(got[0], np.asarray(start[0], np.float64))
