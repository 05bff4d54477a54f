"""Port of the reference's renderer test, src/opt/test/test_renderer.cc:43-215 (TestRendererPixelAccuracy) with the thirteen cameras of
:217-315: a mesh with one vertex per 20th pixel at a random depth (std::mt19937(0); uniform_real_distribution<>(0.5, 20) + three colour
draws per vertex) is rendered from the identity pose with near / far 0.1 / 20.1, and at every grid pixel the rendered depth must equal
the vertex depth within 5e-2 while the vertex reprojects onto its pixel within 1e-2 px. The colour assertions need the GL colour pass
and are not ported (this library renders depth only). Shared by the oracle test (its software rasteriser) and the C-ABI test (K8)."""
import numpy as np

from tests.golden.ref_test_inputs import MT19937, canonical

W, H, STEP = 640, 480, 20
FX, FY, CX, CY = 250.0, 200.0, 319.5, 239.5
OMEGA, K1, K2, K3 = 1.0, 0.23, -0.66, 0.64
PT = [340.926, 341.124, 302.4, 201.6]


def cameras(orc):
    """(name, Type, GetParameters vector) — test_renderer.cc:233-315."""
    return [("Pinhole", orc.CAM_PINHOLE, [FX, FY, CX, CY]),
            ("SimplePinhole", orc.CAM_SIMPLE_PINHOLE, [FX, CX, CY]),
            ("Polynomial", orc.CAM_POLYNOMIAL, [FX, FY, CX, CY, K1, K2, K3]),
            ("Radial", orc.CAM_RADIAL, [FX, CX, CY, K1, -K2]),
            ("SimpleRadial", orc.CAM_SIMPLE_RADIAL, [FX, CX, CY, K1]),
            ("RadialFisheye", orc.CAM_RADIAL_FISHEYE, [FX, CX, CY, 0.221184, 0.128597]),
            ("SimpleRadialFisheye", orc.CAM_SIMPLE_RADIAL_FISHEYE, [0.5 * FX, CX, CY, K1]),
            ("FisheyeFOV", orc.CAM_FOV, [FX, FY, CX, CY, OMEGA]),
            ("PolynomialTangential", orc.CAM_POLYNOMIAL_TANGENTIAL, PT + [-0.101082, 0.0703954, 0.000438661, -0.000680887]),
            ("FisheyePolynomial4", orc.CAM_FISHEYE_POLYNOMIAL_4, PT + [0.221184, 0.128597, 0.0623079, 0.20419]),
            ("FullOpenCV", orc.CAM_FULL_OPENCV, PT + [-0.101082, 0.0703954, 0.00438661, -0.00680887, -0.00101082, .1, .001, -.001]),
            ("FisheyePolynomialTangential", orc.CAM_FISHEYE_POLYNOMIAL_TANGENTIAL, PT + [-0.101082, 0.0703954, 0.000438661, -0.000680887]),
            ("Benchmark", orc.CAM_BENCHMARK, PT + [-0.101082, 0.0703954, 0.000438661, -0.000680887, 0.002, 0.001, -0.003, 0.004])]


def base_intrinsics(orc, model, p):
    single = model in (orc.CAM_SIMPLE_PINHOLE, orc.CAM_RADIAL, orc.CAM_RADIAL_FISHEYE, orc.CAM_SIMPLE_RADIAL, orc.CAM_SIMPLE_RADIAL_FISHEYE)
    return (p[0], p[0], p[1], p[2]) if single else tuple(p[:4])


def build_mesh(orc, model, params):
    """The vertex grid of test_renderer.cc:61-92 and the faces of :98-122. ImageToNormalized goes through the undistortion lookup
    (camera_base_impl.h:183-204: the pixel is clamped to (w - 1.001, h - 1) and the four surrounding table entries are blended; a
    table entry = Undistort(f_inv * x + c_inv)); at the integer pixels used here that is one table entry, except on the far border."""
    fx, fy, cx, cy = [np.float32(v) for v in base_intrinsics(orc, model, params)]
    fxi, fyi = np.float32(1) / fx, np.float32(1) / fy
    cxi, cyi = -cx / fx, -cy / fy

    def lookup(ix, iy):
        d = np.stack([fxi * ix.astype(np.float32) + cxi, fyi * iy.astype(np.float32) + cyi], 1).astype(np.float32)
        return orc.cam_eval(model, W, H, params, "undistort", d)

    gen = MT19937(0)
    gw, gh = W // STEP + 1, H // STEP + 1
    px = np.array([(x, y) for y in range(0, H + 1, STEP) for x in range(0, W + 1, STEP)], np.float32)
    depth = np.zeros(len(px), np.float32)
    for i in range(len(px)):
        depth[i] = np.float32(0.5 + (20.0 - 0.5) * canonical(gen))       # libstdc++: uniform_real_distribution<double>(0.5f, 20.0f)
        gen(); gen(); gen()                                              # r, g, b: one engine output each (range 256 divides 2^32)
    cl = np.minimum(px, np.array([W - 1.001, H - 1.0], np.float32)).astype(np.float32)
    ip = cl.astype(np.int32); f = (cl - ip.astype(np.float32)).astype(np.float32)
    x1 = np.minimum(ip[:, 0] + 1, W - 1); y1 = np.minimum(ip[:, 1] + 1, H - 1)   # (weight 0 wherever the clamp bites)
    tl, tr = lookup(ip[:, 0], ip[:, 1]), lookup(x1, ip[:, 1])
    bl, br = lookup(ip[:, 0], y1), lookup(x1, y1)
    fxx, fyy = f[:, :1], f[:, 1:]
    nxy = ((1 - fyy) * ((1 - fxx) * tl + fxx * tr) + fyy * ((1 - fxx) * bl + fxx * br)).astype(np.float32)
    verts = np.concatenate([depth[:, None] * nxy, depth[:, None]], 1).astype(np.float32)
    bad = ~np.isfinite((nxy.astype(np.float64) ** 2).sum(1))
    verts[bad, 2] = -1                                                   # "not undistortable": z = -1, x and y stay depth * inf (:77-82)
    verts[bad, :2] = np.inf
    faces = []
    for y in range(gh - 1):
        for x in range(gw - 1):
            tl_, tr_, bl_, br_ = x + y * gw, x + 1 + y * gw, x + (y + 1) * gw, x + 1 + (y + 1) * gw
            faces.append((tl_, tr_, bl_)); faces.append((bl_, tr_, br_))
    return verts, np.array(faces, np.uint32), gw


def check(orc, model, params, depth_map, verts, gw, project):
    """The depth assertions of test_renderer.cc:166-203. Returns the fraction of grid pixels that were covered (depth > 0)."""
    covered = total = 0
    for y in range(0, H, STEP):
        for x in range(0, W, STEP):
            p = verts[x // STEP + (y // STEP) * gw]
            d = float(depth_map[y, x])
            if p[2] > 0:
                total += 1
                if d > 0:
                    u = project(np.array([[p[0] / p[2], p[1] / p[2]]], np.float32))[0]
                    assert abs(x - u[0]) <= 1e-2 and abs(y - u[1]) <= 1e-2, (x, y, u)
                    assert abs(d - p[2]) <= 5e-2, "depth %g vs vertex %g at pixel (%d, %d)" % (d, p[2], x, y)
                    covered += 1
            else:
                assert d == 0
    return covered / max(total, 1)
