"""Host-side mirror of the body of the ICPScanAligner tool (/root/reference/src/exe/icp_scan_aligner.cc:274-375): the coarse-to-fine
scale loop around icp::PointToPlaneICP — per scale the scans are subsampled by a stride, their normals re-estimated, a fresh ICP object is
filled and iterated until it reports convergence, and the poses carry over to the next scale. File handling (MeshLab project, PLY) is
the tool's and stays with it.

The ICP and the normal estimation are passed in as callables so that the same schedule runs on the library (the defaults below) and, in
the tests, on the CPU restatement; nothing here imports the latter."""
import math

import numpy as np


def scale_schedule(number_of_scales, max_correspondence_distance, downscale_step=4, search_distance_increase_factor_per_scale=2.0):
    """[(scaled max correspondence distance (float32), subsampling stride)] from the coarsest scale to the finest
    (icp_scan_aligner.cc:285-286: std::pow(float, int) is a double, the product is rounded to float; :308-309: int step = std::pow(int, int))."""
    out = []
    for scale_index in range(number_of_scales):
        e = number_of_scales - 1 - scale_index
        d = np.float32(math.pow(float(np.float32(search_distance_increase_factor_per_scale)), e) * float(np.float32(max_correspondence_distance)))
        step = int(math.pow(downscale_step, e)) if scale_index < number_of_scales - 1 else 1
        out.append((d, max(1, step)))
    return out


def rotation_of(T):
    """Eigen::Transform::rotation() of an affine 4x4 (computeRotationScaling: the orthogonal polar factor of the linear part, with the
    determinant's sign moved into the last singular direction) — what the tool stores in `object_ptr->R` after every iteration (:352)."""
    A = np.asarray(T, np.float32)[:3, :3]
    U, _, Vt = np.linalg.svd(A.astype(np.float32))
    x = np.float32(1.0) if np.linalg.det((U @ Vt).astype(np.float64)) >= 0 else np.float32(-1.0)
    return (U @ np.diag(np.array([1, 1, x], np.float32)) @ Vt).astype(np.float64)


def _default_icp(max_correspondence_distance):
    from .icp import PointToPlaneICP
    return PointToPlaneICP(index_distance_hint=float(max_correspondence_distance))


def _default_normals(xyz, k):
    from .normals import estimate_normals
    return estimate_normals(xyz, k, (0.0, 0.0, 0.0))


def align_scans(scans, poses, fixed=None, max_correspondence_distance=0.10, max_num_iterations=50, convergence_threshold_max_movement=1e-6,
                normal_estimation_neighbor_count=32, number_of_scales=1, downscale_step=4, search_distance_increase_factor_per_scale=2.0,
                icp_factory=None, estimate_normals=None, print_progress=False):
    """scans: list of (n, 3) float32 point arrays in their own frames; poses: list of 4x4 global_T_cloud (double, as the tool's R | T);
    fixed[i]: the cloud's pose is not optimised (`!object_ptr->optimize_pose`).
    Returns (poses as a list of 4x4 float64, log) with one log entry per scale: distance, stride, points per cloud, iterations, converged."""
    n = len(scans)
    fixed = [False] * n if fixed is None else [bool(f) for f in fixed]
    icp_factory = icp_factory or _default_icp
    estimate_normals = estimate_normals or _default_normals
    R = [np.asarray(T, np.float64)[:3, :3].copy() for T in poses]
    t = [np.asarray(T, np.float64)[:3, 3].copy() for T in poses]
    log = []
    for d, step in scale_schedule(number_of_scales, max_correspondence_distance, downscale_step, search_distance_increase_factor_per_scale):
        icp = icp_factory(d)
        ids, counts = [], []
        for i in range(n):
            xyz = np.ascontiguousarray(np.asarray(scans[i], np.float32)[::step])           # at(0), at(step), ... (:311-313)
            nrm = np.asarray(estimate_normals(xyz, normal_estimation_neighbor_count), np.float32)[:, :3]
            T = np.eye(4, dtype=np.float64); T[:3, :3] = R[i]; T[:3, 3] = t[i]
            ids.append(icp.AddPointCloud(xyz, np.ascontiguousarray(nrm), T.astype(np.float32), fixed[i]))   # transform.cast<float>() (:336)
            counts.append(len(xyz))
        iterations, converged = 0, False
        for iteration in range(max_num_iterations):
            converged = bool(icp.Run(float(d), iteration, 1, convergence_threshold_max_movement, print_progress))
            iterations += 1
            for i in range(n):
                if fixed[i]:
                    continue
                G = np.asarray(icp.GetResultGlobalTCloud(ids[i]), np.float32)
                R[i] = rotation_of(G); t[i] = G[:3, 3].astype(np.float64)
            if converged:
                break
        if hasattr(icp, "close"):
            icp.close()
        log.append({"max_correspondence_distance": float(d), "stride": step, "points": counts, "iterations": iterations, "converged": converged})
    out = []
    for i in range(n):
        T = np.eye(4, dtype=np.float64); T[:3, :3] = R[i]; T[:3, 3] = t[i]
        out.append(T)
    return out, log
