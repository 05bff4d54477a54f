"""Synthetic inputs for benchmarks and parity tests. Data generators only — not on the product path."""
import ctypes as C
import math
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(os.path.dirname(_HERE), "_build", "libb2synth.so")
_lib = None


def build():
    subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB


def _load():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB):
            build()
        _lib = C.CDLL(_LIB)
        _lib.b2synth_scan.restype = C.c_size_t
        _lib.b2synth_scan.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_double), C.c_double, C.c_double, C.c_uint64,
                                      C.POINTER(C.c_float), C.POINTER(C.c_float)]
    return _lib


SCANNER_POSITIONS = [(x, y, 1.5) for y in (1.6, 4.0, 6.4) for x in (2.0, 5.0, 8.0) if not (x == 5.0 and y == 4.0)]


def rot_xyz(rx, ry, rz):
    cx, sx, cy, sy, cz, sz = math.cos(rx), math.sin(rx), math.cos(ry), math.sin(ry), math.cos(rz), math.sin(rz)
    Rx = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]]); Ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
    Rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]])
    return Rz @ Ry @ Rx


def room_scan(index, W, H, sigma_range=0.001, seed=10):
    """Scan `index` (0..7) of the config-2 scene: (xyz, normals) float32 (n,3) in the scanner frame + ground-truth 4x4 pose."""
    L = _load()
    pos = np.array(SCANNER_POSITIONS[index], np.float64)
    yaw = 0.35 * index
    xyz = np.empty((W * H, 3), np.float32); nrm = np.empty((W * H, 3), np.float32)
    n = L.b2synth_scan(W, H, pos.ctypes.data_as(C.POINTER(C.c_double)), yaw, sigma_range, seed + index,
                       xyz.ctypes.data_as(C.POINTER(C.c_float)), nrm.ctypes.data_as(C.POINTER(C.c_float)))
    T = np.eye(4)
    T[:3, :3] = rot_xyz(0, 0, yaw)
    T[:3, 3] = pos
    return xyz[:n], nrm[:n], T


def perturbed_poses(gt_poses, trans_mm=5.0, rot_deg=0.1, seed=99):
    """Initial poses = ground truth perturbed by U(-t,t) mm and U(-r,r) deg per axis."""
    rng = np.random.default_rng(seed)
    out = []
    for T in gt_poses:
        dt = rng.uniform(-trans_mm, trans_mm, 3) * 1e-3
        dr = np.deg2rad(rng.uniform(-rot_deg, rot_deg, 3))
        P = np.eye(4); P[:3, :3] = rot_xyz(*dr); P[:3, 3] = dt
        out.append((P @ T).astype(np.float32))
    return out


def room_scans(num_scans, W, H, sigma_range=0.001, seed=10, pose_seed=99):
    clouds, gts = [], []
    for i in range(num_scans):
        xyz, nrm, T = room_scan(i, W, H, sigma_range, seed)
        clouds.append((xyz, nrm)); gts.append(T)
    return clouds, perturbed_poses(gts, seed=pose_seed), gts


def relief_scans(n_points=50000, sigma=0.0005, seeds=(1, 2)):
    """Config 1: two scans of z = 2 + 0.03 sin(2 pi x) cos(2 pi y) over [-0.75,0.75]^2, scan B offset by
    t=(2,-1.5,1) mm and 0.1 deg about (1,1,1)/sqrt(3). Returns [(xyz, analytic normals)], initial poses."""
    clouds = []
    for s in seeds:
        rng = np.random.default_rng(s)
        x = rng.uniform(-0.75, 0.75, n_points); y = rng.uniform(-0.75, 0.75, n_points)
        z = 2 + 0.03 * np.sin(2 * np.pi * x) * np.cos(2 * np.pi * y)
        dzdx = 0.03 * 2 * np.pi * np.cos(2 * np.pi * x) * np.cos(2 * np.pi * y)
        dzdy = -0.03 * 2 * np.pi * np.sin(2 * np.pi * x) * np.sin(2 * np.pi * y)
        nrm = np.stack([dzdx, dzdy, -np.ones_like(x)], 1)      # towards the scanner at the origin (z decreasing)
        nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
        p = np.stack([x, y, z], 1)
        p += (p / np.linalg.norm(p, axis=1, keepdims=True)) * rng.normal(0, sigma, (n_points, 1))
        clouds.append((p.astype(np.float32), nrm.astype(np.float32)))
    a = np.deg2rad(0.1); k = np.ones(3) / math.sqrt(3)
    K = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    R = np.eye(3) + math.sin(a) * K + (1 - math.cos(a)) * K @ K
    T1 = np.eye(4); T1[:3, :3] = R; T1[:3, 3] = [0.002, -0.0015, 0.001]
    return clouds, [np.eye(4, dtype=np.float32), T1.astype(np.float32)]
