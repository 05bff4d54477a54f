/* Synthetic laser-scan generator for the benchmark / parity inputs:
 * a 10 x 8 x 3 m room (6 planes) with 12 boxes and 4 cylinders, ray-cast from a scanner position on an
 * equirectangular W x H angular grid with per-ray jitter, Gaussian range noise, misses dropped.
 * Output: points in the SCANNER frame (scanner at the origin, yaw removed) and the analytic surface normal of the hit,
 * oriented towards the scanner (what NormalEstimator's viewpoint flip would give).
 * Data generator only — not on the product path. Deterministic (counter-based RNG), OpenMP over rows. */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <omp.h>

static inline uint64_t mix64(uint64_t z) {
  z += 0x9E3779B97F4A7C15ull; z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
static inline double u01(uint64_t seed, uint64_t i, uint64_t k) {
  return (double)(mix64(seed * 0x100000001B3ull + i * 4 + k) >> 11) * (1.0 / 9007199254740992.0);
}

typedef struct { double lo[3], hi[3]; } box_t;
typedef struct { double cx, cy, r, h; } cyl_t;

#define NBOX 12
#define NCYL 4
static const box_t kBoxes[NBOX] = {
  {{1.0, 1.0, 0.0}, {1.8, 2.2, 0.9}}, {{3.0, 0.3, 0.0}, {4.5, 0.9, 2.0}}, {{6.0, 0.4, 0.0}, {7.2, 1.2, 0.75}},
  {{8.5, 1.5, 0.0}, {9.6, 3.0, 1.1}}, {{0.3, 3.5, 0.0}, {0.9, 5.0, 1.8}}, {{2.5, 3.2, 0.0}, {3.9, 4.4, 0.72}},
  {{6.2, 3.4, 0.0}, {7.6, 4.6, 0.74}}, {{8.8, 4.6, 0.0}, {9.7, 6.0, 2.1}}, {{1.2, 6.3, 0.0}, {2.6, 7.5, 0.8}},
  {{4.2, 6.6, 0.0}, {5.8, 7.7, 1.0}}, {{7.0, 6.4, 0.0}, {7.9, 7.3, 1.5}}, {{4.6, 2.0, 0.0}, {5.3, 2.7, 0.45}}};
static const cyl_t kCyls[NCYL] = {{2.2, 2.6, 0.25, 3.0}, {7.8, 2.4, 0.25, 3.0}, {2.4, 5.6, 0.3, 1.2}, {5.0, 5.2, 0.35, 3.0}};
static const double kRoom[3] = {10.0, 8.0, 3.0};

/* Nearest hit of ray o + t d (t > 1e-6). Returns t (or -1) and the outward surface normal. */
static double cast(const double o[3], const double d[3], double n[3]) {
  double best = 1e30; n[0] = n[1] = n[2] = 0;
  /* room interior */
  for (int a = 0; a < 3; ++a) {
    if (fabs(d[a]) < 1e-12) continue;
    for (int s = 0; s < 2; ++s) {
      const double plane = s ? kRoom[a] : 0.0;
      const double t = (plane - o[a]) / d[a];
      if (t > 1e-6 && t < best) {
        const int b = (a + 1) % 3, c = (a + 2) % 3;
        const double pb = o[b] + t * d[b], pc = o[c] + t * d[c];
        if (pb >= 0 && pb <= kRoom[b] && pc >= 0 && pc <= kRoom[c]) { best = t; n[0] = n[1] = n[2] = 0; n[a] = s ? -1.0 : 1.0; }
      }
    }
  }
  for (int k = 0; k < NBOX; ++k) {
    double t0 = 0, t1 = 1e30; int ax = -1, sg = 0, ok = 1;
    for (int a = 0; a < 3 && ok; ++a) {
      if (fabs(d[a]) < 1e-12) { if (o[a] < kBoxes[k].lo[a] || o[a] > kBoxes[k].hi[a]) ok = 0; continue; }
      double ta = (kBoxes[k].lo[a] - o[a]) / d[a], tb = (kBoxes[k].hi[a] - o[a]) / d[a]; int s = -1;
      if (ta > tb) { double t = ta; ta = tb; tb = t; s = 1; }
      if (ta > t0) { t0 = ta; ax = a; sg = s; }
      if (tb < t1) t1 = tb;
      if (t0 > t1) ok = 0;
    }
    if (ok && ax >= 0 && t0 > 1e-6 && t0 < best) { best = t0; n[0] = n[1] = n[2] = 0; n[ax] = (double)sg; }
  }
  for (int k = 0; k < NCYL; ++k) {
    const double ox = o[0] - kCyls[k].cx, oy = o[1] - kCyls[k].cy;
    const double A = d[0] * d[0] + d[1] * d[1], B = 2 * (ox * d[0] + oy * d[1]), C = ox * ox + oy * oy - kCyls[k].r * kCyls[k].r;
    if (A > 1e-14) {
      const double disc = B * B - 4 * A * C;
      if (disc > 0) {
        const double t = (-B - sqrt(disc)) / (2 * A);
        const double z = o[2] + t * d[2];
        if (t > 1e-6 && t < best && z >= 0 && z <= kCyls[k].h) {
          best = t; n[0] = (ox + t * d[0]) / kCyls[k].r; n[1] = (oy + t * d[1]) / kCyls[k].r; n[2] = 0;
        }
      }
    }
    if (kCyls[k].h < kRoom[2] && fabs(d[2]) > 1e-12) {   /* top cap */
      const double t = (kCyls[k].h - o[2]) / d[2];
      const double x = ox + t * d[0], y = oy + t * d[1];
      if (t > 1e-6 && t < best && x * x + y * y <= kCyls[k].r * kCyls[k].r) { best = t; n[0] = n[1] = 0; n[2] = 1; }
    }
  }
  return best < 1e29 ? best : -1.0;
}

/* Generates up to W*H points. Returns the number written. xyz/nrm: capacity W*H*3 floats each. */
size_t b2synth_scan(int W, int H, const double scanner_pos[3], double yaw, double sigma_range, uint64_t seed,
                    float* xyz, float* nrm) {
  const double el0 = -60.0 * M_PI / 180.0, el1 = 90.0 * M_PI / 180.0;
  size_t* row_count = (size_t*)calloc((size_t)H + 1, sizeof(size_t));
  float* tmp_xyz = (float*)malloc((size_t)W * H * 3 * sizeof(float));
  float* tmp_nrm = (float*)malloc((size_t)W * H * 3 * sizeof(float));
  const double cy = cos(yaw), sy = sin(yaw);
#pragma omp parallel for schedule(dynamic, 8)
  for (int r = 0; r < H; ++r) {
    size_t cnt = 0;
    for (int c = 0; c < W; ++c) {
      const uint64_t i = (uint64_t)r * W + c;
      const double az = ((double)c + u01(seed, i, 0)) / W * 2.0 * M_PI;
      const double el = el0 + ((double)r + u01(seed, i, 1)) / H * (el1 - el0);
      /* direction in the scanner frame, then rotate by yaw into the room */
      const double dl[3] = {cos(el) * cos(az), cos(el) * sin(az), sin(el)};
      const double d[3] = {cy * dl[0] - sy * dl[1], sy * dl[0] + cy * dl[1], dl[2]};
      double n[3];
      double t = cast(scanner_pos, d, n);
      if (t < 0) continue;
      const double g = sqrt(-2.0 * log(1.0 - u01(seed, i, 2))) * cos(2.0 * M_PI * u01(seed, i, 3));
      t += sigma_range * g;
      /* orient the normal towards the scanner; express in the scanner frame */
      if (n[0] * d[0] + n[1] * d[1] + n[2] * d[2] > 0) { n[0] = -n[0]; n[1] = -n[1]; n[2] = -n[2]; }
      const double nl[3] = {cy * n[0] + sy * n[1], -sy * n[0] + cy * n[1], n[2]};
      float* px = tmp_xyz + ((size_t)r * W + cnt) * 3; float* pn = tmp_nrm + ((size_t)r * W + cnt) * 3;
      px[0] = (float)(t * dl[0]); px[1] = (float)(t * dl[1]); px[2] = (float)(t * dl[2]);
      pn[0] = (float)nl[0]; pn[1] = (float)nl[1]; pn[2] = (float)nl[2];
      ++cnt;
    }
    row_count[r + 1] = cnt;
  }
  for (int r = 0; r < H; ++r) row_count[r + 1] += row_count[r];
#pragma omp parallel for schedule(static)
  for (int r = 0; r < H; ++r) {
    const size_t cnt = row_count[r + 1] - row_count[r];
    memcpy(xyz + row_count[r] * 3, tmp_xyz + (size_t)r * W * 3, cnt * 3 * sizeof(float));
    memcpy(nrm + row_count[r] * 3, tmp_nrm + (size_t)r * W * 3, cnt * 3 * sizeof(float));
  }
  const size_t total = row_count[H];
  free(row_count); free(tmp_xyz); free(tmp_nrm);
  return total;
}
