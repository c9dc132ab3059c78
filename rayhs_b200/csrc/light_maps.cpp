// light_maps.cpp — "nearest possible occluder" cube maps, one per (point light, occluder mesh).
//
// shadowIntersection (RayHs.hs:74-87) only asks whether SOME triangle of a mesh lies between the shaded point p and
// the light L (inFrontOfLight, RayHs.hs:84-87).  Every such shadow ray ends in L, so the triangles it can meet are the
// ones that, seen from L, cover the direction of p - L, and a triangle T can only be in front of the light when
// dist(L, T) < |p - L|.  For every cell of a cube map around L this file stores a LOWER bound of dist(L, T) over all
// triangles T that cover any direction of the cell (+inf for a cell no triangle covers).  The shadow kernels look up
// the cell of p - L and skip the tree walk of that mesh when |p - L| is below the stored bound: no triangle the walk
// could find would count.  The map never says "occluded" — that is always decided by the exact double test.
//
// Conservative by construction:
//  * a triangle is the exact one the kernels intersect, (p0, p0 + e1, p0 + e2) of its rh_tri record (Mesh.hs:70-71);
//  * per cube face it is clipped to the face's pyramid widened by `delta` (two cells), projected (a projective map of a
//    convex polygon), and every cell whose neighbourhood of one cell on each side meets the polygon is marked
//    (scanline over three-row strips, one more column on each side).  The kernels compute the cell in float
//    (error ~1e-6 of a face against a cell of 2/R) — the one-cell rim absorbs that and a face chosen the other way
//    on a tie of two direction components;
//  * the stored value is the exact point-triangle distance times (1 - 1e-6), less 1e-9 of the scene's scale, rounded
//    down to float (an accepted hit may lie outside its triangle by the roundings of u and v); a triangle
//    closer to L than 1e-6 of the scene's scale turns the whole map off (the ray through such a point misses L by
//    roundings that are then no longer small against the distance).
#include <algorithm>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <limits>
#include <new>
#include <atomic>
#include <thread>
#include <vector>

#include "common.h"

#include "light_geom.h"

namespace rh {

using lg::P3;
using lg::add;
using lg::sub;

// ------------------------------------------------------------------ lit triangles
// A triangle T0 is "lit" by a light when no other triangle of its mesh meets K (common.h; for a directional light K is
// the prism over T0 along the light's vector instead of the hull with the light's position).  Then for every
// shaded point p of T0 the shadow query towards that light finds nothing in this mesh: a hit the reference accepts
// lies on a triangle, on the segment from p to L (inFrontOfLight, RayHs.hs:84-87), at least 2e-8 above T0's plane —
// inside K.  T0 itself cannot be hit (its plane is behind the ray origin: t < 0).  Margins: the base is taken at 1e-8
// instead of 2e-8, and the side planes of K are moved outward by 1e-9 (1 + |coordinates|): a shaded point may lie
// outside T0 by the roundings of its own hit test — at most ~1e-16 |o - p0| / cos(view angle), and Mesh.hs:73's
// |det| >= 1e-6 keeps that below 1e-10 for triangles up to unit size — plus a few ulps of p = o + t d.  (A neighbour
// across an edge that rises more steeply than ~84 degrees above T0's plane therefore counts as blocking: inside the
// 1e-9 rim it can reach the 1e-8 base.)  Coordinates of 1e6 and more get no flags.  The geometry itself is in
// light_geom.h, shared with the CUDA builder (setup_kernels.cu).
bool lit_query_make(const rh_tri& t0, const double L[3], bool directional, LitQuery* q) { return lg::lit_query_make(t0, L, directional, q); }
bool lit_query_box_outside(const LitQuery& q, const double* lo, const double* hi) { return lg::lit_query_box_outside(q, lo, hi); }
bool lit_query_tri_meets(const LitQuery& q, const rh_tri& t) { return lg::lit_query_tri_meets(q, t); }

// Fills out[6 * R * R] (face-major, rows of R cells) for the light at L and the triangles tris[slots[0 .. n)].
// Returns false when the map would be useless or unsafe — a triangle (nearly) touches the light, or fewer than
// `min_empty` of the cells stay empty (a light buried in a triangle soup) — and true otherwise, with the fraction
// of empty cells in *empty_fraction.
bool build_light_map(const double L[3], const rh_tri* tris, const uint32_t* slots, size_t n, int R, float* out,
                     double min_empty, double* empty_fraction) {
  const float inf = std::numeric_limits<float>::infinity();
  const size_t cells = (size_t)6 * R * R;
  std::fill(out, out + cells, inf);
  double scale_abs = 1.0;
  for (int k = 0; k < 3; k++) scale_abs = std::max(scale_abs, std::fabs(L[k]));
  const P3 lp = {L[0], L[1], L[2]};
  // Triangles go in a strided order so that an early look at the fill tells a soup from a shape.
  const size_t stride = n > 4096 ? 4093 : 1;  // prime: a permutation of 0..n-1 whenever n is not a multiple of it
  const bool permute = stride > 1 && n % stride != 0;
  std::atomic<bool> unsafe{false};
  // One cube face per thread (the faces are disjoint parts of `out`): every thread walks the same stretch of the
  // triangle sequence and rasterises only its own face.
  auto raster_range = [&](size_t begin, size_t end, int face) {
    const int k = face / 2;
    const double sgn = (face & 1) ? -1.0 : 1.0;
    for (size_t i = begin; i < end; i++) {
      const size_t at = permute ? (i * stride) % n : i;
      const rh_tri& t = tris[slots[at]];
      const P3 p0 = {t.p0[0], t.p0[1], t.p0[2]};
      const P3 a = sub(p0, lp);
      const P3 b = sub(add(p0, lg::mk(t.e1[0], t.e1[1], t.e1[2])), lp);
      const P3 c = sub(add(p0, lg::mk(t.e2[0], t.e2[1], t.e2[2])), lp);
      float val;
      if (!lg::light_map_value(a, b, c, scale_abs, &val)) { unsafe = true; return; }
      const P3 tri[3] = {a, b, c};
      float* cells_of_face = out + (size_t)face * R * R;
      auto mark = [&](int row, int c0, int c1) {
        float* r = cells_of_face + (size_t)row * R;
        for (int cc = c0; cc <= c1; cc++) r[cc] = std::min(r[cc], val);
      };
      lg::raster_face(R, tri, k, sgn, mark);
    }
  };
  const bool threaded = n >= 2048 && std::thread::hardware_concurrency() >= 4;
  size_t done = 0, next_check = 16384, prev_empty = 0;
  while (done < n) {
    const size_t end = std::min(n, next_check);
    if (threaded) {
      std::thread pool[6];
      for (int f = 0; f < 6; f++) pool[f] = std::thread(raster_range, done, end, f);
      for (int f = 0; f < 6; f++) pool[f].join();
    } else {
      for (int f = 0; f < 6; f++) raster_range(done, end, f);
    }
    if (unsafe) return false;
    done = end;
    if (done == next_check && done < n) {
      next_check *= 2;
      size_t empty = 0;
      for (size_t q = 0; q < cells; q++) empty += out[q] == inf;
      if ((double)empty < min_empty * (double)cells) return false;
      // A soup shows early.  The last done/2 triangles (a uniform sample of the mesh: strided order) took the fraction r
      // of the cells that were still empty before them; triangles scattered at random keep taking that fraction, a
      // shape's silhouette saturates (r -> 0).  When the empties predicted for all n triangles are far below the useful
      // minimum the map is given up now (this only decides whether a map is worth having, never what a map says).
      if (prev_empty > 0 && empty < prev_empty) {
        const double r = (double)(prev_empty - empty) / (double)prev_empty;
        const double blocks = (double)(n - done) / (0.5 * (double)done);
        if ((double)empty / (double)cells * std::pow(1.0 - r, blocks) < 1e-3 * min_empty) return false;
      }
      prev_empty = empty;
    }
  }
  size_t empty = 0;
  for (size_t q = 0; q < cells; q++) empty += out[q] == inf;
  if (empty_fraction) *empty_fraction = (double)empty / (double)cells;
  return (double)empty >= min_empty * (double)cells;
}

}  // namespace rh

extern "C" int rh_light_map_build(const double light_pos[3], const rh_tri* tris, uint32_t n_tris, int res, float* out, int* useful,
                                  double* empty_fraction) {
  if (!light_pos || (!tris && n_tris) || !out || res < 1 || res > 4096)
    return rh::set_error(RH_ERR_ARG, "rh_light_map_build: null argument or resolution outside 1..4096");
  try {
    std::vector<uint32_t> slots(n_tris);
    for (uint32_t i = 0; i < n_tris; i++) slots[i] = i;
    double empty = 0;
    const bool ok = rh::build_light_map(light_pos, tris, slots.data(), n_tris, res, out, 0.02, &empty);
    if (useful) *useful = ok ? 1 : 0;
    if (empty_fraction) *empty_fraction = empty;
  } catch (const std::bad_alloc&) {
    return rh::set_error(RH_ERR_OOM, "rh_light_map_build: out of host memory");
  }
  return RH_OK;
}

extern "C" int rh_lit_triangles(int light_kind, const double light_pos[3], const rh_tri* tris, uint32_t n_tris, uint8_t* out) {
  if (!light_pos || (!tris && n_tris) || (!out && n_tris)) return rh::set_error(RH_ERR_ARG, "rh_lit_triangles: null argument");
  if (light_kind != RH_LIGHT_POINT && light_kind != RH_LIGHT_DIRECTIONAL) return rh::set_error(RH_ERR_ARG, "rh_lit_triangles: unknown light kind");
  // groups of eight consecutive triangles with their bounding box, so that the box test is exercised too
  const uint32_t n_groups = (n_tris + 7) / 8;
  try {
    std::vector<double> lo((size_t)n_groups * 3, std::numeric_limits<double>::infinity()), hi((size_t)n_groups * 3, -std::numeric_limits<double>::infinity());
    for (uint32_t i = 0; i < n_tris; i++)
      for (int k = 0; k < 3; k++) {
        const double a = tris[i].p0[k], b = a + tris[i].e1[k], c = a + tris[i].e2[k];
        lo[(size_t)(i / 8) * 3 + k] = std::min(lo[(size_t)(i / 8) * 3 + k], std::min(a, std::min(b, c)));
        hi[(size_t)(i / 8) * 3 + k] = std::max(hi[(size_t)(i / 8) * 3 + k], std::max(a, std::max(b, c)));
      }
    for (uint32_t i = 0; i < n_tris; i++) {
      rh::LitQuery q;
      out[i] = 0;
      if (!rh::lit_query_make(tris[i], light_pos, light_kind == RH_LIGHT_DIRECTIONAL, &q)) continue;
      bool blocked = false;
      for (uint32_t g = 0; g < n_groups && !blocked; g++) {
        if (rh::lit_query_box_outside(q, &lo[(size_t)g * 3], &hi[(size_t)g * 3])) continue;
        for (uint32_t j = g * 8; j < std::min(n_tris, g * 8 + 8) && !blocked; j++) blocked = j != i && rh::lit_query_tri_meets(q, tris[j]);
      }
      out[i] = blocked ? 0 : 1;
    }
  } catch (const std::bad_alloc&) {
    return rh::set_error(RH_ERR_OOM, "rh_lit_triangles: out of host memory");
  }
  return RH_OK;
}
