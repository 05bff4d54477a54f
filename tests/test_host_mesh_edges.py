"""Host side of b2_reg_set_mesh (csrc/b2_mesh_edges.h: counting-sort grouping of the half-edges, threaded classification) against the
oracle's edge list (oracle/orc_mesh.h: build_mesh_edges, a restatement of occlusion_geometry.cc:466-645). Pure host code: the header
is compiled into a small shared object with g++ and called through ctypes — no GPU, no CUDA runtime."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SHIM = r'''
#include <cstring>
#include "%s/dataset_pipeline_b200/csrc/b2_mesh_edges.h"
extern "C" size_t host_mesh_edges(const float* v, size_t nv, const uint32_t* f, size_t nf, uint32_t* out /* 5 per edge */, float* fn) {
  std::vector<float> normals; std::vector<b2::MeshEdgeHost> edges;
  b2::build_mesh_edges(v, nv, f, nf, &normals, &edges);
  if (out) std::memcpy(out, edges.data(), edges.size() * sizeof(b2::MeshEdgeHost));
  if (fn) std::memcpy(fn, normals.data(), normals.size() * sizeof(float));
  return edges.size();
}
'''


@pytest.fixture(scope="module")
def shim(tmp_path_factory):
    d = tmp_path_factory.mktemp("mesh_edges")
    src = d / "shim.cc"
    src.write_text(SHIM % ROOT)
    so = d / "libshim.so"
    subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-std=c++17", "-shared", "-fPIC", "-pthread", "-o", str(so), str(src)])
    L = C.CDLL(str(so))
    L.host_mesh_edges.restype = C.c_size_t
    L.host_mesh_edges.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]
    return L


def host_edges(L, v, f):
    v = np.ascontiguousarray(v, np.float32); f = np.ascontiguousarray(f, np.uint32)
    n = L.host_mesh_edges(v.ctypes.data, len(v), f.ctypes.data, len(f), None, None)
    out = np.zeros((n, 5), np.uint32); fn = np.zeros((len(f), 3), np.float32)
    assert L.host_mesh_edges(v.ctypes.data, len(v), f.ctypes.data, len(f), out.ctypes.data, fn.ctypes.data) == n
    return out, fn


def oracle_edges(oracle, v, f):
    reg = oracle.Registration()
    reg.set_mesh(v, f)
    v1, v2, f1, f2, fl = reg.mesh_edges()
    return np.stack([v1, v2, f1, f2, fl.astype(np.uint32)], axis=1)


def as_sorted(e):
    return e[np.lexsort((e[:, 1], e[:, 0]))]


def check(L, oracle, v, f):
    got, _ = host_edges(L, v, f)
    want = oracle_edges(oracle, v, f)
    assert len(got) == len(want)
    # the library's list is ordered by (v1, v2) (the reference iterates an unordered map: only the set is defined)
    assert np.array_equal(got, as_sorted(got))
    assert np.array_equal(got, as_sorted(want))
    return got


def box_mesh():
    v = np.array([[x, y, z] for z in (0, 1) for y in (0, 1) for x in (0, 1)], np.float32)
    q = [(0, 1, 3, 2), (4, 6, 7, 5), (0, 4, 5, 1), (2, 3, 7, 6), (0, 2, 6, 4), (1, 5, 7, 3)]
    f = []
    for a, b, c, d in q:
        f += [(a, b, c), (a, c, d)]
    return v, np.array(f, np.uint32)


def test_closed_box_has_its_twelve_creases_and_no_diagonals(shim, oracle):
    v, f = box_mesh()
    e = check(shim, oracle, v, f)
    assert len(e) == 12 and not (e[:, 4] & 1).any()          # the six face diagonals are coplanar pairs, nothing is open


def test_open_patch_fans_and_flipped_faces(shim, oracle):
    rng = np.random.default_rng(3)
    # a bumpy height field (open boundary), one extra fin sharing an interior edge (three faces on an edge) and a few flipped triangles
    n = 9
    xs, ys = np.meshgrid(np.arange(n), np.arange(n), indexing="xy")
    v = np.stack([xs.ravel(), ys.ravel(), rng.normal(0, 0.3, n * n)], axis=1).astype(np.float32)
    f = []
    for y in range(n - 1):
        for x in range(n - 1):
            a = y * n + x
            f += [(a, a + 1, a + n + 1), (a, a + n + 1, a + n)]
    f = np.array(f, np.uint32)
    flip = rng.choice(len(f), 12, replace=False)
    f[flip] = f[flip][:, ::-1]
    fin_top = len(v)
    v = np.vstack([v, [[4.5, 4.5, 3.0]]]).astype(np.float32)
    f = np.vstack([f, [[4 * n + 4, 5 * n + 5, fin_top]], [[5 * n + 5, 4 * n + 4, fin_top]]]).astype(np.uint32)   # two fins on one edge: 4 faces
    e = check(shim, oracle, v, f)
    assert (e[:, 4] & 1).sum() >= 4 * (n - 1)                 # the patch boundary is open
    assert (e[:, 4] & 2).any()                                # flipped neighbours: opposite normals


def test_room_mesh_matches_and_normals_are_unit(shim, oracle):
    from dataset_pipeline_b200.synth import room_views
    v, f = room_views.room_mesh(step=0.25)
    assert len(f) > 4000                                      # above the threading threshold of the header
    e = check(shim, oracle, v, f)
    assert len(e) > 0
    _, fn = host_edges(shim, v, f)
    assert np.allclose(np.linalg.norm(fn, axis=1), 1.0, atol=1e-5)
    # the normals are the reference's: normalised cross product of the first two edges, fp32, no contraction
    a = v[f[:, 1]] - v[f[:, 0]]; b = v[f[:, 2]] - v[f[:, 0]]
    c = np.stack([a[:, 1] * b[:, 2] - a[:, 2] * b[:, 1], a[:, 2] * b[:, 0] - a[:, 0] * b[:, 2], a[:, 0] * b[:, 1] - a[:, 1] * b[:, 0]], axis=1).astype(np.float32)
    nn = np.sqrt((c[:, 0] * c[:, 0] + (c[:, 1] * c[:, 1] + c[:, 2] * c[:, 2])).astype(np.float32))
    assert np.array_equal(fn, (c / nn[:, None]).astype(np.float32))
