"""Re-creation of the procedurally generated inputs of the reference's own ICP tests
(/root/reference/src/opt/test/test_icp.cc:39-172) without Eigen/PCL.

std::mt19937(0) and std::uniform_real_distribution<double> (libstdc++ generate_canonical<double,53>, two 32-bit
draws per sample) are emulated exactly, so points / axes / angles / translations are the reference test's values.
The per-point normals of test_icp.cc:51-55 come from Eigen::Vector3f::Random() (glibc rand()), which cannot be
reproduced here; the test only needs *some* unit normals, so they are drawn from numpy default_rng(1) instead.
"""
import math

import numpy as np


class MT19937:
    def __init__(self, seed):
        self.mt = [0] * 624
        self.idx = 624
        self.mt[0] = seed & 0xFFFFFFFF
        for i in range(1, 624):
            self.mt[i] = (1812433253 * (self.mt[i - 1] ^ (self.mt[i - 1] >> 30)) + i) & 0xFFFFFFFF

    def _twist(self):
        mt = self.mt
        for i in range(624):
            y = (mt[i] & 0x80000000) | (mt[(i + 1) % 624] & 0x7FFFFFFF)
            mt[i] = mt[(i + 397) % 624] ^ (y >> 1) ^ (0x9908B0DF if y & 1 else 0)
        self.idx = 0

    def __call__(self):
        if self.idx >= 624:
            self._twist()
        y = self.mt[self.idx]
        self.idx += 1
        y ^= y >> 11
        y ^= (y << 7) & 0x9D2C5680
        y ^= (y << 15) & 0xEFC60000
        y ^= y >> 18
        return y & 0xFFFFFFFF


def canonical(gen):
    s = float(gen()) + float(gen()) * 4294967296.0
    r = s / 18446744073709551616.0
    return r if r < 1.0 else math.nextafter(1.0, 0.0)


def uniform(gen, a, b):
    return canonical(gen) * (b - a) + a


def angle_axis_matrix(angle, axis):
    """Eigen::AngleAxisf::toRotationMatrix in fp32."""
    f = np.float32
    angle = f(angle); axis = axis.astype(np.float32)
    s, c = f(math.sin(float(angle))), f(math.cos(float(angle)))
    sin_axis = s * axis
    cos1_axis = (f(1) - c) * axis
    R = np.zeros((3, 3), np.float32)
    tmp = cos1_axis[0] * axis[1]; R[0, 1] = tmp - sin_axis[2]; R[1, 0] = tmp + sin_axis[2]
    tmp = cos1_axis[0] * axis[2]; R[0, 2] = tmp + sin_axis[1]; R[2, 0] = tmp - sin_axis[1]
    tmp = cos1_axis[1] * axis[2]; R[1, 2] = tmp - sin_axis[0]; R[2, 1] = tmp + sin_axis[0]
    d = cos1_axis * axis + c
    R[0, 0], R[1, 1], R[2, 2] = d
    return R


def identical_cloud_alignment_inputs():
    """test_icp.cc:39-92: 50 random points, 20 randomly moved copies."""
    gen = MT19937(0)
    pts = np.zeros((50, 3), np.float32)
    for i in range(50):
        pts[i] = [uniform(gen, -1.0, 1.0) for _ in range(3)]
    nrm = np.random.default_rng(1).normal(size=(50, 3))
    nrm = (nrm / np.linalg.norm(nrm, axis=1, keepdims=True)).astype(np.float32)
    ta, tb = float(np.float32(-0.05)), float(np.float32(0.05))
    aa = math.pi / float(np.float32(180.0)) * float(np.float32(-10.0))
    ab = math.pi / float(np.float32(180.0)) * float(np.float32(10.0))
    poses = []
    for _ in range(20):
        while True:
            axis = np.array([uniform(gen, ta, tb) for _ in range(3)], np.float32)
            if float(np.linalg.norm(axis)) >= 1e-4:
                break
        axis = axis / np.float32(np.linalg.norm(axis))
        R = angle_axis_matrix(uniform(gen, aa, ab), axis)
        T = np.eye(4, dtype=np.float32)
        T[:3, :3] = R
        T[:3, 3] = [uniform(gen, ta, tb) for _ in range(3)]
        poses.append(T)
    return pts, nrm, poses


def plane_with_single_point_inputs():
    """test_icp.cc:111-158: 50x50 unit grid in z=0 + one point at (0,0,20); copy offset by (1,0,0)."""
    xs, ys = np.meshgrid(np.arange(50), np.arange(50), indexing="ij")
    pts = np.stack([xs.ravel(), ys.ravel(), np.zeros(2500)], 1).astype(np.float32)
    nrm = np.tile(np.array([[0, 0, 1]], np.float32), (2500, 1))
    pts = np.concatenate([pts, np.array([[0, 0, 20]], np.float32)])
    n1 = np.array([1, 0, 1], np.float32)
    n1 = n1 / np.float32(np.sqrt(np.float32(2.0)))
    nrm = np.concatenate([nrm, n1[None]])
    T0 = np.eye(4, dtype=np.float32)
    T1 = np.eye(4, dtype=np.float32); T1[0, 3] = 1.0
    return pts, nrm, [T0, T1]
