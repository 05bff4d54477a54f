"""Small triangle meshes shared by the Path B occlusion tests."""
import numpy as np


def _grid_mesh(x0, x1, y0, y1, z, n, tilt=0.0):
    xs = np.linspace(x0, x1, n + 1); ys = np.linspace(y0, y1, n + 1)
    X, Y = np.meshgrid(xs, ys, indexing="xy")
    V = np.stack([X.ravel(), Y.ravel(), z + tilt * X.ravel()], 1).astype(np.float32)
    F = []
    for j in range(n):
        for i in range(n):
            a = j * (n + 1) + i; b = a + 1; c = a + n + 1; d = c + 1
            F += [[a, b, d], [a, d, c]]
    return V, np.array(F, np.uint32)


def _box_mesh(c, h):
    x, y, z = c
    V = np.array([[x + sx * h, y + sy * h, z + sz * h] for sz in (-1, 1) for sy in (-1, 1) for sx in (-1, 1)], np.float32)
    F = np.array([[0, 1, 3], [0, 3, 2], [4, 7, 5], [4, 6, 7], [0, 5, 1], [0, 4, 5], [2, 3, 7], [2, 7, 6], [0, 2, 6], [0, 6, 4], [1, 5, 7], [1, 7, 3]], np.uint32)
    return V, F
