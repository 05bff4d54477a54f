"""Synthetic inputs of BASELINE config 4 / 5: camera views of the config-2 room (the scene of synth/scene.c: a
10 x 8 x 3 m room with 12 boxes and 4 cylinders), its occlusion mesh tessellated at 2 cm, and the albedo the scan points and the
images share. Data generators only (harness: torch tensors on the GPU for the 24-megapixel renders) — not on the product path."""
import math

import numpy as np

ROOM = (10.0, 8.0, 3.0)
BOXES = [((1.0, 1.0, 0.0), (1.8, 2.2, 0.9)), ((3.0, 0.3, 0.0), (4.5, 0.9, 2.0)), ((6.0, 0.4, 0.0), (7.2, 1.2, 0.75)),
         ((8.5, 1.5, 0.0), (9.6, 3.0, 1.1)), ((0.3, 3.5, 0.0), (0.9, 5.0, 1.8)), ((2.5, 3.2, 0.0), (3.9, 4.4, 0.72)),
         ((6.2, 3.4, 0.0), (7.6, 4.6, 0.74)), ((8.8, 4.6, 0.0), (9.7, 6.0, 2.1)), ((1.2, 6.3, 0.0), (2.6, 7.5, 0.8)),
         ((4.2, 6.6, 0.0), (5.8, 7.7, 1.0)), ((7.0, 6.4, 0.0), (7.9, 7.3, 1.5)), ((4.6, 2.0, 0.0), (5.3, 2.7, 0.45))]
CYLS = [(2.2, 2.6, 0.25, 3.0), (7.8, 2.4, 0.25, 3.0), (2.4, 5.6, 0.3, 1.2), (5.0, 5.2, 0.35, 3.0)]       # cx, cy, r, h  (scene.c)

# ---- albedo: multi-octave solid value noise (gradients at every pyramid level: wavelengths 0.7 m ... 5 mm) ----------------------
OCTAVES = [(1.5, 1.0), (4.0, 0.7), (11.0, 0.55), (30.0, 0.45), (80.0, 0.35), (200.0, 0.3)]             # (cycles per metre, amplitude)


def _hash01(torch, ix, iy, iz, seed):
    h = (ix * 73856093) ^ (iy * 19349663) ^ (iz * 83492791) ^ (seed * 2654435761)
    h = (h ^ (h >> 13)) * 1274126177
    h = h ^ (h >> 16)
    return (h & 0xFFFFFF).to(torch.float64) * (1.0 / 16777215.0)


def albedo(points, seed=30):
    """Grey value in [~20, ~235] at world positions `points` (torch float64 (n,3) on any device)."""
    import torch
    acc = torch.zeros(points.shape[0], dtype=torch.float64, device=points.device)
    norm = 0.0
    for k, (f, a) in enumerate(OCTAVES):
        p = points * f
        i = torch.floor(p)
        t = p - i
        t = t * t * (3.0 - 2.0 * t)
        ix, iy, iz = i[:, 0].to(torch.int64), i[:, 1].to(torch.int64), i[:, 2].to(torch.int64)
        v = 0.0
        for dz in (0, 1):
            wz = t[:, 2] if dz else 1.0 - t[:, 2]
            for dy in (0, 1):
                wy = t[:, 1] if dy else 1.0 - t[:, 1]
                for dx in (0, 1):
                    wx = t[:, 0] if dx else 1.0 - t[:, 0]
                    v = v + wz * wy * wx * _hash01(torch, ix + dx, iy + dy, iz + dz, seed + k)
        acc = acc + a * (v - 0.5)
        norm += a
    return 128.0 + 215.0 * acc / norm


# ---- ray casting of the analytic scene (same primitives as scene.c) -----------------------------------------------------------------
def cast(o, d):
    """Nearest hit distance along rays o + t d (torch float64: o (3,), d (n,3)); inf where nothing is hit (cannot happen inside the room)."""
    import torch
    inf = float("inf")
    best = torch.full((d.shape[0],), inf, dtype=torch.float64, device=d.device)
    for a in range(3):
        b, c = (a + 1) % 3, (a + 2) % 3
        for plane in (0.0, ROOM[a]):
            t = (plane - o[a]) / d[:, a]
            pb = o[b] + t * d[:, b]; pc = o[c] + t * d[:, c]
            ok = (t > 1e-6) & (t < best) & (pb >= 0) & (pb <= ROOM[b]) & (pc >= 0) & (pc <= ROOM[c])
            best = torch.where(ok, t, best)
    for lo, hi in BOXES:
        t0 = torch.zeros_like(best); t1 = torch.full_like(best, 1e30)
        for a in range(3):
            inv = 1.0 / d[:, a]
            ta = (lo[a] - o[a]) * inv; tb = (hi[a] - o[a]) * inv
            t0 = torch.maximum(t0, torch.minimum(ta, tb)); t1 = torch.minimum(t1, torch.maximum(ta, tb))
        ok = (t0 <= t1) & (t0 > 1e-6) & (t0 < best)
        best = torch.where(ok, t0, best)
    for cx, cy, r, h in CYLS:
        ox, oy = o[0] - cx, o[1] - cy
        A = d[:, 0] ** 2 + d[:, 1] ** 2; B = 2 * (ox * d[:, 0] + oy * d[:, 1]); C = ox * ox + oy * oy - r * r
        disc = B * B - 4 * A * C
        t = (-B - torch.sqrt(torch.clamp(disc, min=0))) / (2 * A)
        z = o[2] + t * d[:, 2]
        ok = (disc > 0) & (A > 1e-14) & (t > 1e-6) & (t < best) & (z >= 0) & (z <= h)
        best = torch.where(ok, t, best)
        if h < ROOM[2]:
            t = (h - o[2]) / d[:, 2]
            x = ox + t * d[:, 0]; y = oy + t * d[:, 1]
            ok = (t > 1e-6) & (t < best) & (x * x + y * y <= r * r)
            best = torch.where(ok, t, best)
    return best


def render_view(width, height, fx, fy, cx, cy, R_wc, c, device, noise_sigma=1.0, seed=31, rows_per_chunk=250):
    """8-bit grey pinhole image of the room from camera centre c with camera-to-world rotation R_wc (numpy)."""
    import torch
    out = torch.empty((height, width), dtype=torch.uint8, device=device)
    R = torch.tensor(R_wc, dtype=torch.float64, device=device); o = torch.tensor(c, dtype=torch.float64, device=device)
    xs = (torch.arange(width, dtype=torch.float64, device=device) - cx) / fx
    g = torch.Generator(device=device); g.manual_seed(seed)
    for y0 in range(0, height, rows_per_chunk):
        y1 = min(height, y0 + rows_per_chunk)
        ys = (torch.arange(y0, y1, dtype=torch.float64, device=device) - cy) / fy
        dc = torch.stack([xs[None, :].expand(y1 - y0, width), ys[:, None].expand(y1 - y0, width), torch.ones((y1 - y0, width), dtype=torch.float64, device=device)], -1).reshape(-1, 3)
        d = dc @ R.T
        t = cast(o, d)
        p = o[None, :] + d * t[:, None]
        v = albedo(p) + noise_sigma * torch.randn(p.shape[0], dtype=torch.float64, device=device, generator=g)
        out[y0:y1] = torch.clamp(torch.round(v), 0, 255).to(torch.uint8).reshape(y1 - y0, width)
    return out


def point_colors(xyz, device, chunk=4_000_000):
    """Grey albedo of scan points (numpy float32 (n,3), global frame) -> numpy uint8 (n,3) rgb (r = g = b)."""
    import torch
    out = np.empty((xyz.shape[0], 3), np.uint8)
    for a in range(0, xyz.shape[0], chunk):
        p = torch.tensor(xyz[a:a + chunk], dtype=torch.float64, device=device)
        v = torch.clamp(torch.round(albedo(p)), 0, 255).to(torch.uint8).cpu().numpy()
        out[a:a + chunk] = v[:, None]
    return out


# ---- camera ring ----------------------------------------------------------------------------------------------------------------------
def _rot(ax, ay, az):
    cx, sx, cy, sy, cz, sz = math.cos(ax), math.sin(ax), math.cos(ay), math.sin(ay), math.cos(az), math.sin(az)
    Rx = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]]); Ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
    Rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]])
    return Rz @ Ry @ Rx


def quat_from_R(R):
    w = math.sqrt(max(0.0, 1 + R[0, 0] + R[1, 1] + R[2, 2])) / 2
    if w > 1e-6:
        return np.array([(R[2, 1] - R[1, 2]) / (4 * w), (R[0, 2] - R[2, 0]) / (4 * w), (R[1, 0] - R[0, 1]) / (4 * w), w])
    i = int(np.argmax(np.diag(R))); j, k = (i + 1) % 3, (i + 2) % 3
    s = math.sqrt(max(1e-30, 1 + R[i, i] - R[j, j] - R[k, k])) * 2
    q = np.zeros(4); q[i] = s / 4; q[j] = (R[j, i] + R[i, j]) / s; q[k] = (R[k, i] + R[i, k]) / s; q[3] = (R[k, j] - R[j, k]) / s
    return q


def view_poses(num_views, seed=31, radius=2.0, height=1.6, trans_mm=2.0, rot_deg=0.05):
    """`num_views` cameras on a ring around the room centre, looking outward / inward alternately with small pitch and roll.
    Returns (R_wc list, centres, poses_gt, poses_init): poses are image_T_global as (qx qy qz qw tx ty tz); init = gt perturbed by
    U(+-trans_mm) per axis and U(+-rot_deg) per axis. More than 20 views: a second ring (config 5)."""
    rng = np.random.default_rng(seed)
    Rs, cs, gt, init = [], [], [], []
    for i in range(num_views):
        ring = i // 20; k = i % 20
        a = 2 * math.pi * (k + 0.5 * ring) / 20
        c = np.array([5.0 + (radius - 0.5 * ring) * math.cos(a), 4.0 + (radius - 0.5 * ring) * math.sin(a), height - 0.3 * ring])
        look = np.array([math.cos(a), math.sin(a), 0.0]) * (1.0 if k % 2 == 0 else -1.0)
        y = np.array([0.0, 0.0, -1.0])                       # image y points down
        x = np.cross(y, look)
        R_wc = np.stack([x, y, look], 1) @ _rot(math.radians(8.0 * math.sin(1.7 * i)), 0.0, math.radians(4.0 * math.cos(2.3 * i)))
        R_cw = R_wc.T; t_cw = -R_cw @ c
        Rs.append(R_wc); cs.append(c)
        gt.append(np.concatenate([quat_from_R(R_cw), t_cw]).astype(np.float32))
        dR = _rot(*np.radians(rng.uniform(-rot_deg, rot_deg, 3))); dt = rng.uniform(-trans_mm, trans_mm, 3) * 1e-3
        init.append(np.concatenate([quat_from_R(dR @ R_cw), dR @ t_cw + dt]).astype(np.float32))
    return Rs, cs, gt, init


# ---- occlusion mesh: the analytic scene tessellated at `step` metres ------------------------------------------------------------------
def _patch(origin, u, v, nu, nv, verts, faces, flip=False):
    """Regular grid patch origin + i u / nu + j v / nv with shared vertices; two triangles per cell."""
    base = sum(len(x) for x in verts)
    i, j = np.meshgrid(np.arange(nu + 1), np.arange(nv + 1), indexing="xy")
    p = np.asarray(origin)[None, :] + (i.ravel()[:, None] / nu) * np.asarray(u)[None, :] + (j.ravel()[:, None] / nv) * np.asarray(v)[None, :]
    verts.append(p.astype(np.float32))
    ci, cj = np.meshgrid(np.arange(nu), np.arange(nv), indexing="xy")
    a = (cj * (nu + 1) + ci).ravel() + base
    b = a + 1; c = a + (nu + 1); d = c + 1
    f = np.concatenate([np.stack([a, b, c], 1), np.stack([c, b, d], 1)])
    faces.append(f[:, ::-1] if flip else f)


def room_mesh(step=0.02):
    """(vertices (nv,3) float32, faces (nf,3) uint32): room walls, box sides and tops, cylinder mantles and caps."""
    verts, faces = [], []
    n = lambda length: max(1, int(round(length / step)))
    X, Y, Z = ROOM
    _patch((0, 0, 0), (X, 0, 0), (0, Y, 0), n(X), n(Y), verts, faces)                    # floor
    _patch((0, 0, Z), (X, 0, 0), (0, Y, 0), n(X), n(Y), verts, faces, True)              # ceiling
    _patch((0, 0, 0), (X, 0, 0), (0, 0, Z), n(X), n(Z), verts, faces, True)              # y = 0
    _patch((0, Y, 0), (X, 0, 0), (0, 0, Z), n(X), n(Z), verts, faces)                    # y = Y
    _patch((0, 0, 0), (0, Y, 0), (0, 0, Z), n(Y), n(Z), verts, faces)                    # x = 0
    _patch((X, 0, 0), (0, Y, 0), (0, 0, Z), n(Y), n(Z), verts, faces, True)              # x = X
    for lo, hi in BOXES:
        dx, dy, dz = hi[0] - lo[0], hi[1] - lo[1], hi[2] - lo[2]
        _patch((lo[0], lo[1], hi[2]), (dx, 0, 0), (0, dy, 0), n(dx), n(dy), verts, faces)                  # top
        _patch((lo[0], lo[1], lo[2]), (dx, 0, 0), (0, 0, dz), n(dx), n(dz), verts, faces)                  # y = lo
        _patch((lo[0], hi[1], lo[2]), (dx, 0, 0), (0, 0, dz), n(dx), n(dz), verts, faces, True)            # y = hi
        _patch((lo[0], lo[1], lo[2]), (0, dy, 0), (0, 0, dz), n(dy), n(dz), verts, faces, True)            # x = lo
        _patch((hi[0], lo[1], lo[2]), (0, dy, 0), (0, 0, dz), n(dy), n(dz), verts, faces)                  # x = hi
    for cx, cy, r, h in CYLS:
        na, nh = n(2 * math.pi * r), n(h)
        base = sum(len(x) for x in verts)
        ang = 2 * math.pi * np.arange(na) / na
        i, j = np.meshgrid(np.arange(na), np.arange(nh + 1), indexing="xy")
        p = np.stack([cx + r * np.cos(ang[i.ravel()]), cy + r * np.sin(ang[i.ravel()]), h * j.ravel() / nh], 1)
        verts.append(p.astype(np.float32))
        ci, cj = np.meshgrid(np.arange(na), np.arange(nh), indexing="xy")
        a = (cj * na + ci).ravel() + base; b = (cj * na + (ci + 1) % na).ravel() + base
        c = a + na; d = b + na
        faces.append(np.concatenate([np.stack([a, b, c], 1), np.stack([c, b, d], 1)]))
        if h < Z:                                                                          # top cap: a fan of rings
            nr = n(r)
            base = sum(len(x) for x in verts)
            ri, ai = np.meshgrid(np.arange(1, nr + 1), np.arange(na), indexing="xy")
            rr = r * ri.ravel() / nr
            p = np.concatenate([[[cx, cy, h]], np.stack([cx + rr * np.cos(ang[ai.ravel()]), cy + rr * np.sin(ang[ai.ravel()]), np.full(rr.shape, h)], 1)])
            verts.append(p.astype(np.float32))
            idx = lambda ring, k: base + 1 + (k % na) * nr + (ring - 1)
            k = np.arange(na)
            fan = np.stack([np.full(na, base), idx(1, k), idx(1, k + 1)], 1)
            quads = []
            for ring in range(1, nr):
                a_, b_, c_, d_ = idx(ring, k), idx(ring, k + 1), idx(ring + 1, k), idx(ring + 1, k + 1)
                quads.append(np.stack([a_, c_, b_], 1)); quads.append(np.stack([b_, c_, d_], 1))
            faces.append(np.concatenate([fan] + quads) if quads else fan)
    return np.concatenate(verts), np.concatenate(faces).astype(np.uint32)
