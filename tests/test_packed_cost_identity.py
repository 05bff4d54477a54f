"""The identity the packed-fp32 cost path of K5 rests on (csrc/b2_icp_kernels.cuh: cost_record_packed), checked in IEEE fp32 on the CPU:
lane 0 / lane 1 of every packed operation carry the source / target side of the scalar evaluation (cost_record, the reference's order of
icp_point_to_plane_impl.h:240-266), products and sums rounded once each; the second residual comes out negated, which its square does not
see. numpy float32 arithmetic is round-to-nearest without contraction, like the kernel's fma(a, b, -0) / fma(a, 1, b) lanes."""
import numpy as np

f32 = np.float32


def sum3(a, b, c):
    return a + (b + c)          # Eigen's three-term order, as b2_common.cuh: sum3


def scalar_cost(ps0, ns0, pt0, nt0, Rs, ts, Rt, tt):
    ps = [sum3(Rs[3 * r] * ps0[0], Rs[3 * r + 1] * ps0[1], Rs[3 * r + 2] * ps0[2]) + ts[r] for r in range(3)]
    ns = [sum3(Rs[3 * r] * ns0[0], Rs[3 * r + 1] * ns0[1], Rs[3 * r + 2] * ns0[2]) for r in range(3)]
    pt = [sum3(Rt[3 * r] * pt0[0], Rt[3 * r + 1] * pt0[1], Rt[3 * r + 2] * pt0[2]) + tt[r] for r in range(3)]
    nt = [sum3(Rt[3 * r] * nt0[0], Rt[3 * r + 1] * nt0[1], Rt[3 * r + 2] * nt0[2]) for r in range(3)]
    r1 = sum3(ns[0] * (pt[0] - ps[0]), ns[1] * (pt[1] - ps[1]), ns[2] * (pt[2] - ps[2]))
    r2 = sum3(nt[0] * (ps[0] - pt[0]), nt[1] * (ps[1] - pt[1]), nt[2] * (ps[2] - pt[2]))
    return r1 * r1, r2 * r2, r2


def packed_cost(ps0, ns0, pt0, nt0, Rs, ts, Rt, tt):
    # every quantity is a (source lane, target lane) pair; arrays of shape (..., 2) stand for the f32x2 registers
    P = [np.stack([Rs[k], Rt[k]], -1) for k in range(9)] + [np.stack([ts[k], tt[k]], -1) for k in range(3)]
    x = [np.stack([ps0[k], pt0[k]], -1) for k in range(3)]
    n = [np.stack([ns0[k], nt0[k]], -1) for k in range(3)]
    p = [sum3(P[3 * r] * x[0], P[3 * r + 1] * x[1], P[3 * r + 2] * x[2]) + P[9 + r] for r in range(3)]
    nn = [sum3(P[3 * r] * n[0], P[3 * r + 1] * n[1], P[3 * r + 2] * n[2]) for r in range(3)]
    d = [(p[r][..., 1] - p[r][..., 0])[..., None] for r in range(3)]       # pt - ps, broadcast to both lanes
    rr = sum3(nn[0] * d[0], nn[1] * d[1], nn[2] * d[2])                     # (r1, -r2)
    sq = rr * rr
    return sq[..., 0], sq[..., 1], rr


def test_packed_lanes_reproduce_the_scalar_cost_bit_for_bit():
    rng = np.random.default_rng(11)
    m = 200000
    ps0 = [rng.uniform(-9, 9, m).astype(f32) for _ in range(3)]
    pt0 = [(ps0[k] + rng.normal(0, 0.004, m)).astype(f32) for k in range(3)]
    unit = lambda v: [c.astype(f32) for c in (v / np.linalg.norm(v, axis=0))]
    ns0, nt0 = unit(rng.normal(size=(3, m))), unit(rng.normal(size=(3, m)))

    def pose(scale):
        w = rng.normal(0, scale, 3)
        th = np.linalg.norm(w); k = w / th
        K = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
        R = np.eye(3) + np.sin(th) * K + (1 - np.cos(th)) * K @ K
        return [f32(v) for v in R.ravel()], [f32(v) for v in rng.normal(0, scale, 3)]

    for scale in (1e-4, 1e-2, 0.5):
        Rs, ts = pose(scale); Rt, tt = pose(scale)
        a1, a2, r2 = scalar_cost(ps0, ns0, pt0, nt0, Rs, ts, Rt, tt)
        b1, b2, rr = packed_cost(ps0, ns0, pt0, nt0, Rs, ts, Rt, tt)
        assert a1.dtype == np.float32 and b1.dtype == np.float32
        assert np.array_equal(a1.view(np.uint32), b1.view(np.uint32))
        assert np.array_equal(a2.view(np.uint32), b2.view(np.uint32))
        # the second lane is the NEGATED reference residual, bit for bit (not merely equal in square)
        assert np.array_equal((-r2).view(np.uint32), np.ascontiguousarray(rr[..., 1]).view(np.uint32))
