// Host side of b2_reg_set_mesh: face normals and the edge list of the occlusion mesh (OcclusionGeometry::ComputeEdgesIfNeeded,
// occlusion_geometry.cc:521-671: half-edges grouped by their sorted vertex pair, in ascending (v1, v2) order, the faces of a pair
// in ascending face order; coplanar pairs dropped, non-manifold fans reduced to their two outer faces).
// Grouping is a counting sort by the smaller vertex followed by a stable insertion sort of each vertex's handful of half-edges by
// the larger one: linear in the mesh, no node allocations, and the same order an ordered map would iterate in. The face normals and
// the per-vertex edge classification run on a few host threads (disjoint outputs, concatenated in vertex order: the result does not
// depend on the thread count). The 1.8 M-triangle room mesh of the benchmark takes ~0.1 s instead of the ~1.5 s of an ordered map
// of vectors.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <thread>
#include <utility>
#include <vector>

namespace b2 {

struct MeshEdgeHost { unsigned int v1, v2, f1, f2, flags; };   // layout of MeshEdgeDev (b2_reg_kernels.cuh); flags: bit0 open, bit1 opposite_normals

// fn(part, begin, end) over [0, n) cut into `parts` contiguous ranges, one host thread each.
template <typename F>
inline void host_parallel_ranges(size_t n, int parts, const F& fn) {
  if (parts <= 1 || n < 4096) { fn(0, (size_t)0, n); return; }
  std::vector<std::thread> th;
  for (int p = 1; p < parts; ++p) th.emplace_back([&, p] { fn(p, n * p / parts, n * (p + 1) / parts); });
  fn(0, (size_t)0, n / parts);
  for (auto& t : th) t.join();
}

inline void build_mesh_edges(const float* vertices, size_t nv, const uint32_t* faces, size_t nf, std::vector<float>* face_normals,
                             std::vector<MeshEdgeHost>* edges_out) {
  auto P = [&](uint32_t i, int c) { return vertices[3 * (size_t)i + c]; };
  auto nrm3 = [](float* a) { const float n = std::sqrt(a[0] * a[0] + (a[1] * a[1] + a[2] * a[2])); a[0] /= n; a[1] /= n; a[2] /= n; };
  auto dot3f = [](const float* a, const float* b) { return a[0] * b[0] + (a[1] * b[1] + a[2] * b[2]); };
  std::vector<float>& fn = *face_normals;
  fn.resize(3 * nf);
  const int threads = (int)std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
  host_parallel_ranges(nf, threads, [&](int, size_t f0, size_t f1) {
    for (size_t fi = f0; fi < f1; ++fi) {
      const uint32_t* t = faces + 3 * fi;
      const float a[3] = {P(t[1], 0) - P(t[0], 0), P(t[1], 1) - P(t[0], 1), P(t[1], 2) - P(t[0], 2)};
      const float b[3] = {P(t[2], 0) - P(t[0], 0), P(t[2], 1) - P(t[0], 1), P(t[2], 2) - P(t[0], 2)};
      float n[3] = {a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]};
      nrm3(n);
      fn[3 * fi] = n[0]; fn[3 * fi + 1] = n[1]; fn[3 * fi + 2] = n[2];
    }
  });
  std::vector<uint32_t> first(nv + 1, 0u);          // half-edges per smaller vertex -> bucket starts
  for (size_t fi = 0; fi < nf; ++fi) {
    const uint32_t* t = faces + 3 * fi;
    for (int k = 0; k < 3; ++k) ++first[std::min(t[k], t[(k + 1) % 3]) + 1];
  }
  for (size_t v = 0; v < nv; ++v) first[v + 1] += first[v];
  struct Half { uint32_t w, face; bool swapped; };
  std::vector<Half> half(3 * nf);
  {
    std::vector<uint32_t> at(first.begin(), first.end() - 1);
    for (size_t fi = 0; fi < nf; ++fi) {            // ascending face order within a bucket
      const uint32_t* t = faces + 3 * fi;
      for (int k = 0; k < 3; ++k) {
        uint32_t u = t[k], w = t[(k + 1) % 3];
        const bool swapped = u > w;
        if (swapped) std::swap(u, w);
        half[at[u]++] = Half{w, (uint32_t)fi, swapped};
      }
    }
  }
  std::vector<std::vector<MeshEdgeHost>> part_edges((size_t)threads);
  host_parallel_ranges(nv, threads, [&](int part, size_t v0, size_t v1) {
  std::vector<MeshEdgeHost>& edges = part_edges[(size_t)part];
  edges.reserve((size_t)(first[v1] - first[v0]) / 2 + 16);
  for (size_t v = v0; v < v1; ++v) {
    Half* hb = half.data() + first[v];
    const size_t cnt = first[v + 1] - first[v];
    for (size_t i = 1; i < cnt; ++i) {              // stable insertion sort by the larger vertex
      const Half x = hb[i];
      size_t j = i;
      while (j > 0 && hb[j - 1].w > x.w) { hb[j] = hb[j - 1]; --j; }
      hb[j] = x;
    }
    for (size_t g0 = 0; g0 < cnt;) {
      size_t g1 = g0 + 1;
      while (g1 < cnt && hb[g1].w == hb[g0].w) ++g1;
      const Half* fl = hb + g0;
      const size_t nfl = g1 - g0;
      g0 = g1;
      MeshEdgeHost e; e.v1 = (unsigned int)v; e.v2 = fl[0].w; e.f1 = fl[0].face; e.f2 = 0; e.flags = 0;
      if (nfl == 1) { e.flags = 1; edges.push_back(e); continue; }
      const float ed[3] = {P(e.v2, 0) - P(e.v1, 0), P(e.v2, 1) - P(e.v1, 1), P(e.v2, 2) - P(e.v1, 2)};
      float s1 = fl[0].swapped ? -1.f : 1.f, s2 = fl[1].swapped ? -1.f : 1.f;
      const float n1v[3] = {fn[3 * (size_t)e.f1] * s1, fn[3 * (size_t)e.f1 + 1] * s1, fn[3 * (size_t)e.f1 + 2] * s1};
      const uint32_t face2 = fl[1].face;
      const float n2v[3] = {fn[3 * (size_t)face2] * s2, fn[3 * (size_t)face2 + 1] * s2, fn[3 * (size_t)face2 + 2] * s2};
      e.f2 = face2;
      bool opposite = s1 * s2 > 0;
      float bx[3] = {n1v[0], n1v[1], n1v[2]}; nrm3(bx);
      float by[3] = {bx[1] * ed[2] - bx[2] * ed[1], bx[2] * ed[0] - bx[0] * ed[2], bx[0] * ed[1] - bx[1] * ed[0]}; nrm3(by);
      float n1x = 1.f, n1y = 0.f, n2x = dot3f(bx, n2v), n2y = dot3f(by, n2v);
      if (n2x < 0 && std::fabs(n2y) < 1e-4f) continue;                      // coplanar pair: not an edge (:571-575)
      bool keep = true;
      if (nfl > 2) {
        const float c12 = n2y;
        for (size_t k = 2; k < nfl; ++k) {
          const uint32_t f3 = fl[k].face; const float s3 = fl[k].swapped ? -1.f : 1.f;
          const float cn[3] = {fn[3 * (size_t)f3] * s3, fn[3 * (size_t)f3 + 1] * s3, fn[3 * (size_t)f3 + 2] * s3};
          const float n3x = dot3f(bx, cn), n3y = dot3f(by, cn);
          const float c13 = n1x * n3y - n1y * n3x, c23 = n2x * n3y - n2y * n3x;
          const bool sign1 = c13 * c12 > 0, sign2 = c23 * c12 < 0;
          if (sign1 && !sign2) { n2x = n3x; n2y = n3y; e.f2 = f3; s2 = s3; opposite = s1 * s3 != 1; }
          else if (sign2 && !sign1) { n1x = n3x; n1y = n3y; e.f1 = f3; s1 = s3; opposite = s3 * s2 != 1; }
          else if (!sign2) { keep = false; break; }                          // `!sign2 && !sign2` in the reference (:633)
        }
      }
      if (!keep) continue;
      e.flags = opposite ? 2u : 0u;
      edges.push_back(e);
    }
  }
  });
  std::vector<MeshEdgeHost>& edges = *edges_out;
  edges.clear();
  size_t total = 0;
  for (const auto& pe : part_edges) total += pe.size();
  edges.reserve(total);
  for (const auto& pe : part_edges) edges.insert(edges.end(), pe.begin(), pe.end());
}

}  // namespace b2
