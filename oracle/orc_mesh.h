// ORACLE — TEST INFRASTRUCTURE ONLY. Not part of the product path (see orc_math.h header).
//
// Mesh occlusion geometry of Path B (rows B1/B2):
//   half-edge construction + edge filtering   /root/reference/src/opt/occlusion_geometry.cc:466-645
//   depth pass (the reference renders with OpenGL ES: /root/reference/src/opt/occlusion_geometry.cc:213-245,
//       /root/reference/src/opengl/renderer.cc:42-131,745-847,913-974): linear camera-z output, background 0, GL_LEQUAL,
//       no culling, near/far clip at min_depth/max_depth, perspective-correct interpolation of var_depth
//   MaskOutOcclusionBoundaries / DrawSplatsAtEdgeIfVisible / DrawEdgeSplatIfVisible   occlusion_geometry.cc:284-402
//
// A GL driver's rasteriser cannot be reproduced bit for bit (and no GL/EGL exists in this environment): PARITY WITH THE GL
// OUTPUT IS UNPINNED. This file DEFINES the software rasteriser both the oracle and the CUDA path implement, bit-identically:
//   * camera-space vertices p = R v + t (fp32, Eigen 3-term order); polygon clipped at z = min_depth (Sutherland-Hodgman,
//     intersection t = (zn - za) / (zb - za) in fp32, z set to zn), fan triangulation;
//   * window coordinates X = fx x/z + cx + 0.5, Y likewise (pixel i has its centre at i + 0.5, as GL's viewport transform of
//     renderer.cc:932-940 gives); a pixel is covered when its centre is inside, edge functions in double, ties resolved by the
//     rule "edge a->b owns its boundary iff dy < 0 or (dy == 0 and dx > 0)" on the positively oriented triangle;
//   * depth = 1 / sum(lambda_i / z_i) (perspective-correct), fragments with depth > max_depth dropped, minimum kept;
//   * pixels without fragment = 0 (the GL clear colour, renderer.cc:766).
//   * boundary masking: an edge is a silhouette edge by the face-normal sign test (:312-323); its splats are tested against and
//     written from the UNMASKED depth map (the reference reads the map while other OpenMP threads write -1 into it, which is a
//     race; testing the input map is the deterministic reading of the same rule).
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <limits>
#include <map>
#include <vector>

#include "orc_math.h"

namespace orc {

struct MeshEdge { uint32_t v1, v2; uint32_t f1, f2; bool open; bool opposite; };   // f2 unused when open

struct OccMesh {
  std::vector<float> v;          // nv * 3
  std::vector<uint32_t> f;       // nf * 3
  std::vector<float> fn;         // nf * 3 face normals (a x b).normalized()
  std::vector<MeshEdge> edges;   // after FilterEdgeList
};

static inline V3f v3(const float* p) { return V3f{p[0], p[1], p[2]}; }
static inline V3f normalized(const V3f& a) { const float n = std::sqrt(dot(a, a)); return V3f{a.x / n, a.y / n, a.z / n}; }

// ComputeEdgeNormalsList + FilterEdgeList (occlusion_geometry.cc:466-645). Edges are visited in sorted (v1,v2) order (the reference
// walks unordered_map buckets under OpenMP; the resulting SET of edges does not depend on the order).
static void build_mesh_edges(OccMesh* m) {
  const size_t nf = m->f.size() / 3;
  m->fn.resize(3 * nf);
  typedef std::pair<uint32_t, uint32_t> Key;
  typedef std::pair<uint32_t, bool> FaceSign;
  std::map<Key, std::vector<FaceSign>> he;
  auto add = [&](uint32_t a, uint32_t b, uint32_t face) {
    const bool swap = a > b;
    if (swap) std::swap(a, b);
    he[Key(a, b)].push_back(FaceSign(face, swap));
  };
  for (size_t fi = 0; fi < nf; ++fi) {
    const uint32_t* t = &m->f[3 * fi];
    const V3f a = sub(v3(&m->v[3 * t[1]]), v3(&m->v[3 * t[0]])), b = sub(v3(&m->v[3 * t[2]]), v3(&m->v[3 * t[0]]));
    const V3f n = normalized(cross(a, b));
    m->fn[3 * fi] = n.x; m->fn[3 * fi + 1] = n.y; m->fn[3 * fi + 2] = n.z;
    add(t[0], t[1], (uint32_t)fi); add(t[1], t[2], (uint32_t)fi); add(t[2], t[0], (uint32_t)fi);
  }
  m->edges.clear();
  for (const auto& kv : he) {
    const float kEpsilon = 1e-4f;
    const std::vector<FaceSign>& nv = kv.second;
    MeshEdge e; e.v1 = kv.first.first; e.v2 = kv.first.second; e.f1 = nv[0].first; e.f2 = 0; e.open = false; e.opposite = false;
    if (nv.size() == 1) { e.open = true; m->edges.push_back(e); continue; }
    const V3f edge = sub(v3(&m->v[3 * e.v2]), v3(&m->v[3 * e.v1]));
    float factor1 = nv[0].second ? -1.f : 1.f;
    const V3f n1v{m->fn[3 * e.f1] * factor1, m->fn[3 * e.f1 + 1] * factor1, m->fn[3 * e.f1 + 2] * factor1};
    uint32_t face2 = nv[1].first;
    float factor2 = nv[1].second ? -1.f : 1.f;
    const V3f n2v{m->fn[3 * face2] * factor2, m->fn[3 * face2 + 1] * factor2, m->fn[3 * face2 + 2] * factor2};
    e.f2 = face2;
    e.opposite = factor1 * factor2 > 0;
    float n1x = 1.f, n1y = 0.f;
    const V3f base_x = normalized(n1v);
    const V3f base_y = normalized(cross(base_x, edge));
    float n2x = dot(base_x, n2v), n2y = dot(base_y, n2v);
    if (n2x < 0 && std::abs(n2y) < kEpsilon) continue;          // coplanar faces: not an edge
    if (nv.size() == 2) { m->edges.push_back(e); continue; }
    const float cross_n1n2 = n2y;
    bool hemisphere = true;
    for (size_t k = 2; k < nv.size(); ++k) {
      const uint32_t f3 = nv[k].first;
      const float factor3 = nv[k].second ? -1.f : 1.f;
      const V3f cn{m->fn[3 * f3] * factor3, m->fn[3 * f3 + 1] * factor3, m->fn[3 * f3 + 2] * factor3};
      const float n3x = dot(base_x, cn), n3y = dot(base_y, cn);
      const float cross_n1n3 = n1x * n3y - n1y * n3x;
      const float cross_n2n3 = n2x * n3y - n2y * n3x;
      const bool sign1 = cross_n1n3 * cross_n1n2 > 0;
      const bool sign2 = cross_n2n3 * cross_n1n2 < 0;
      if (sign1 && !sign2) { n2x = n3x; n2y = n3y; e.f2 = f3; factor2 = factor3; e.opposite = factor1 * factor3 != 1; }
      else if (sign2 && !sign1) { n1x = n3x; n1y = n3y; e.f1 = f3; factor1 = factor3; e.opposite = factor3 * factor2 != 1; }
      else if (!sign2 && !sign2) { hemisphere = false; break; }   // sic (occlusion_geometry.cc:633)
    }
    if (hemisphere) m->edges.push_back(e);
  }
}

typedef Camera RasterCam;   // orc_camera.h (included before this header)

// What the reference's vertex shader does to a camera-space vertex before the pinhole projection matrix (opengl/renderer.cc:42-131,
// distortion snippets :581-583 pinhole, :630-653 benchmark): (x, y) <- z * distort(x/z, y/z); beyond the camera's own cut-off
// (never for the benchmark camera, whose radius_cutoff_squared() is +inf, :668-680) the vertex is pushed far out (x, y) *= 99.
// The thin-prism model has no renderer program of its own in the reference (its objects report Type::kBenchmark,
// camera_thin_prism.cc:37,47); here it gets the same snippet without the fisheye step.
// The other models (renderer.cc:154-560): every snippet evaluates z * Distort(x/z, y/z) with the (x, y) * 99 push-out beyond the cut-off;
// GLSL float arithmetic is driver-defined, so their individual operation orders are not restated: Camera::distort is used.
static inline void vertex_distort(const Camera& c, V3f* p) {
  if (c.dist == kDistNone && !c.fisheye) return;
  float nx = p->x / p->z, ny = p->y / p->z;
  float r2 = nx * nx + ny * ny;
  if (c.dist != kDistThinPrism) {
    float dx, dy;
    if (r2 <= c.cutoff2) c.distort(nx, ny, &dx, &dy);
    if (r2 <= c.cutoff2 && std::isfinite(dx) && std::isfinite(dy)) { p->x = p->z * dx; p->y = p->z * dy; }
    else { p->x = p->x * 99.0f; p->y = p->y * 99.0f; }
    return;
  }
  if (r2 <= c.cutoff2) {
    if (c.type == kCamBenchmark) {
      const float r = std::sqrt(r2);
      if (r > 1e-6f) { const float theta_by_r = atan_r1(r) / r; nx = theta_by_r * nx; ny = theta_by_r * ny; }
    }
    const float k1 = c.d[0], k2 = c.d[1], p1 = c.d[2], p2 = c.d[3], k3 = c.d[4], k4 = c.d[5], sx1 = c.d[6], sy1 = c.d[7];
    const float x2 = nx * nx, xy = nx * ny, y2 = ny * ny;
    r2 = x2 + y2;
    const float radial = 1.0f + r2 * (k1 + r2 * (k2 + r2 * (k3 + r2 * k4)));
    p->x = p->z * (radial * nx + 2.0f * p1 * xy + p2 * (r2 + 2.0f * x2) + sx1 * r2);
    p->y = p->z * (radial * ny + 2.0f * p2 * xy + p1 * (r2 + 2.0f * y2) + sy1 * r2);
  } else {
    p->x = p->x * 99.0f; p->y = p->y * 99.0f;
  }
}

// One clipped, camera-space triangle (all z >= zn > 0) into the min-depth buffer (inf = empty).
static inline void raster_triangle(const RasterCam& c, const V3f& a, const V3f& b, const V3f& cc, float max_depth, float* depth) {
  float X[3], Y[3]; const float Z[3] = {a.z, b.z, cc.z};
  const V3f P[3] = {a, b, cc};
  for (int i = 0; i < 3; ++i) { X[i] = c.fx * (P[i].x / P[i].z) + c.cx + 0.5f; Y[i] = c.fy * (P[i].y / P[i].z) + c.cy + 0.5f; }
  double x0 = X[0], y0 = Y[0], x1 = X[1], y1 = Y[1], x2 = X[2], y2 = Y[2];
  double z0 = Z[0], z1 = Z[1], z2 = Z[2];
  double area = (x1 - x0) * (y2 - y0) - (y1 - y0) * (x2 - x0);
  if (area == 0.0 || !(area == area)) return;
  if (area < 0.0) { std::swap(x1, x2); std::swap(y1, y2); std::swap(z1, z2); area = -area; }
  const double minx = std::min(x0, std::min(x1, x2)), maxx = std::max(x0, std::max(x1, x2));
  const double miny = std::min(y0, std::min(y1, y2)), maxy = std::max(y0, std::max(y1, y2));
  if (!(maxx >= 0.0 && maxy >= 0.0 && minx <= (double)c.w && miny <= (double)c.h)) return;
  const int ix0 = std::max(0, (int)std::floor(minx - 0.5)), ix1 = std::min(c.w - 1, (int)std::ceil(maxx - 0.5));
  const int iy0 = std::max(0, (int)std::floor(miny - 0.5)), iy1 = std::min(c.h - 1, (int)std::ceil(maxy - 0.5));
  auto owns = [](double dx, double dy) { return dy < 0.0 || (dy == 0.0 && dx > 0.0); };
  const bool o0 = owns(x2 - x1, y2 - y1), o1 = owns(x0 - x2, y0 - y2), o2 = owns(x1 - x0, y1 - y0);
  const double iz0 = 1.0 / z0, iz1 = 1.0 / z1, iz2 = 1.0 / z2;
  for (int iy = iy0; iy <= iy1; ++iy) {
    const double py = iy + 0.5;
    for (int ix = ix0; ix <= ix1; ++ix) {
      const double px = ix + 0.5;
      const double w0 = (x2 - x1) * (py - y1) - (y2 - y1) * (px - x1);
      const double w1 = (x0 - x2) * (py - y2) - (y0 - y2) * (px - x2);
      const double w2 = (x1 - x0) * (py - y0) - (y1 - y0) * (px - x0);
      if (!((w0 > 0.0 || (w0 == 0.0 && o0)) && (w1 > 0.0 || (w1 == 0.0 && o1)) && (w2 > 0.0 || (w2 == 0.0 && o2)))) continue;
      // perspective-correct depth 1 / sum(lambda_i / z_i), lambda_i = w_i / area, as ONE division per pixel: area / sum(w_i * (1 / z_i))
      const float z = (float)(area / (w0 * iz0 + w1 * iz1 + w2 * iz2));
      if (!(z <= max_depth)) continue;
      float& d = depth[(size_t)iy * c.w + ix];
      if (z < d) d = z;
    }
  }
}

// Depth pass over the whole mesh under image_T_global (R row-major, t). Output: linear z, 0 where nothing was drawn.
static void raster_mesh(const OccMesh& m, const RasterCam& c, const float R[9], const V3f& t, float min_depth, float max_depth, std::vector<float>* out) {
  out->assign((size_t)c.w * c.h, std::numeric_limits<float>::infinity());
  const M3f& M = *reinterpret_cast<const M3f*>(R);
  for (size_t fi = 0; fi < m.f.size() / 3; ++fi) {
    V3f p[3];
    bool finite = true;
    for (int k = 0; k < 3; ++k) {
      const V3f r = mul(M, v3(&m.v[3 * m.f[3 * fi + k]])); p[k] = V3f{r.x + t.x, r.y + t.y, r.z + t.z};
      finite = finite && std::isfinite(p[k].x) && std::isfinite(p[k].y) && std::isfinite(p[k].z);
      if (finite) vertex_distort(c, &p[k]);
    }
    // a triangle with a non-finite vertex draws nothing: the reference's renderer test feeds such vertices (pixels that cannot be
    // undistorted get x = y = depth * inf, test_renderer.cc:77-82) and requires depth 0 at their pixels (:204-206)
    if (!finite) continue;
    // clip against z >= min_depth (after the vertex stage, as GL clips)
    V3f poly[4]; int np = 0;
    for (int k = 0; k < 3; ++k) {
      const V3f& A = p[k]; const V3f& B = p[(k + 1) % 3];
      const bool ain = A.z >= min_depth, bin = B.z >= min_depth;
      if (ain) poly[np++] = A;
      if (ain != bin) {
        const float tt = (min_depth - A.z) / (B.z - A.z);
        poly[np++] = V3f{A.x + tt * (B.x - A.x), A.y + tt * (B.y - A.y), min_depth};
      }
    }
    if (np < 3) continue;
    raster_triangle(c, poly[0], poly[1], poly[2], max_depth, out->data());
    if (np == 4) raster_triangle(c, poly[0], poly[2], poly[3], max_depth, out->data());
  }
  for (float& d : *out) if (std::isinf(d)) d = 0.f;
}

// MaskOutOcclusionBoundaries (occlusion_geometry.cc:284-402), deterministic reading (tests and writes use the unmasked map `in`).
static void mask_boundaries(const OccMesh& m, const RasterCam& c, const float R[9], const V3f& t, const V3f& image_position, float splat_radius,
                            const std::vector<float>& in, std::vector<float>* out) {
  *out = in;
  const M3f& M = *reinterpret_cast<const M3f*>(R);
  const float kOcclusionDepthThreshold = 0.05f;
  for (const MeshEdge& e : m.edges) {
    const V3f e1 = v3(&m.v[3 * e.v1]), e2 = v3(&m.v[3 * e.v2]);
    if (!e.open) {
      const V3f to_image = sub(image_position, e1);
      const bool face1 = dot(v3(&m.fn[3 * e.f1]), to_image) > 0;
      const bool face2 = dot(v3(&m.fn[3 * e.f2]), to_image) > 0;
      if (!((e.opposite && (face1 == face2)) || (face1 != face2 && !e.opposite))) continue;
    }
    V3f a = mul(M, e1); a = V3f{a.x + t.x, a.y + t.y, a.z + t.z};
    if (a.z <= 0) continue;
    V3f b = mul(M, e2); b = V3f{b.x + t.x, b.y + t.y, b.z + t.z};
    if (b.z <= 0) continue;
    const V3f delta = sub(b, a);
    const int count = 1 + std::min(static_cast<int>(std::sqrt(dot(delta, delta)) / splat_radius + 0.5f), 150);
    for (int i = 0; i < count; ++i) {
      const float factor = i / (count - 1.0f);
      const V3f p{a.x + factor * delta.x, a.y + factor * delta.y, a.z + factor * delta.z};
      if (!(p.z > 0)) continue;
      const float nx = p.x / p.z, ny = p.y / p.z;
      float px, py; c.project(nx, ny, &px, &py);                      // camera.NormalizedToImage (occlusion_geometry.cc:377)
      const int ix = f2i(px + 0.5f), iy = f2i(py + 0.5f);
      if (!(px + 0.5f >= 0 && py + 0.5f >= 0 && ix >= 0 && iy >= 0 && ix < c.w && iy < c.h && in[(size_t)iy * c.w + ix] + kOcclusionDepthThreshold >= p.z)) continue;
      float dd[6]; c.d_by_world(p, dd);                                // camera.ImageDerivativeByWorld (:386)
      const float rx = std::sqrt(sum3(dd[0] * dd[0], dd[1] * dd[1], dd[2] * dd[2])) * splat_radius, ry = std::sqrt(sum3(dd[3] * dd[3], dd[4] * dd[4], dd[5] * dd[5])) * splat_radius;
      const int min_x = std::max(0, int(ix - rx + 0.5)), min_y = std::max(0, int(iy - ry + 0.5));
      const int end_x = std::min(c.w, int(ix + rx + 1.5)), end_y = std::min(c.h, int(iy + ry + 1.5));
      for (int y = min_y; y < end_y; ++y) for (int x = min_x; x < end_x; ++x) {
        const float old = in[(size_t)y * c.w + x];
        if (old == 0 || old + kOcclusionDepthThreshold > p.z) (*out)[(size_t)y * c.w + x] = -1;
      }
    }
  }
}

}  // namespace orc
