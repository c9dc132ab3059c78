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

namespace rh {

namespace {

struct P3 {
  double x, y, z;
  double operator[](int k) const { return k == 0 ? x : (k == 1 ? y : z); }
};
inline P3 sub(P3 a, P3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline P3 add(P3 a, P3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline P3 scale(double s, P3 a) { return {s * a.x, s * a.y, s * a.z}; }
inline double dot(P3 a, P3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline double len(P3 a) { return std::sqrt(dot(a, a)); }

// Distance from the origin to the closest point of triangle (a, b, c): the Voronoi-region walk over vertices,
// edges and face.  Any point it returns lies on the triangle, so a rounding slip in the region choice costs
// O(ulp) of the distance; degenerate triangles fall back to a bound that needs no division.
double dist_origin_triangle(P3 a, P3 b, P3 c) {
  const double da = len(a), db = len(b), dc = len(c);
  const double min_vertex = std::min(da, std::min(db, dc));
  const P3 ab = sub(b, a), ac = sub(c, a), bc = sub(c, b);
  const double longest = std::max(len(ab), std::max(len(ac), len(bc)));
  const double loose = std::max(0.0, min_vertex - longest);  // every point of T is within `longest` of a vertex
  const P3 cr = {ab.y * ac.z - ab.z * ac.y, ab.z * ac.x - ab.x * ac.z, ab.x * ac.y - ab.y * ac.x};
  if (!(dot(cr, cr) > 1e-24 * dot(ab, ab) * dot(ac, ac))) return loose;  // sliver or point
  const P3 ap = scale(-1, a), bp = scale(-1, b), cp = scale(-1, c);
  const double d1 = dot(ab, ap), d2 = dot(ac, ap);
  double d;
  const double d3 = dot(ab, bp), d4 = dot(ac, bp);
  const double d5 = dot(ab, cp), d6 = dot(ac, cp);
  const double vc = d1 * d4 - d3 * d2, vb = d5 * d2 - d1 * d6, va = d3 * d6 - d5 * d4;
  if (d1 <= 0 && d2 <= 0) d = da;
  else if (d3 >= 0 && d4 <= d3) d = db;
  else if (vc <= 0 && d1 >= 0 && d3 <= 0) d = len(add(a, scale(d1 / (d1 - d3), ab)));
  else if (d6 >= 0 && d5 <= d6) d = dc;
  else if (vb <= 0 && d2 >= 0 && d6 <= 0) d = len(add(a, scale(d2 / (d2 - d6), ac)));
  else if (va <= 0 && (d4 - d3) >= 0 && (d5 - d6) >= 0) d = len(add(b, scale((d4 - d3) / ((d4 - d3) + (d5 - d6)), bc)));
  else {
    const double denom = 1.0 / (va + vb + vc);
    d = len(add(a, add(scale(vb * denom, ab), scale(vc * denom, ac))));
  }
  if (!std::isfinite(d)) return loose;
  return std::max(loose, std::min(d, min_vertex));
}

inline float round_down(double x) {
  float f = (float)x;
  if ((double)f > x) f = std::nextafterf(f, -std::numeric_limits<float>::infinity());
  return f;
}

struct Poly {
  P3 v[12];
  int n = 0;
};

// Sutherland-Hodgman step: keep the part of `in` with f(v) = cw * v[k] + ca * v[a] >= 0 (a plane through the light).
void clip_plane(const Poly& in, Poly& out, int k, double cw, int a, double ca) {
  out.n = 0;
  for (int i = 0; i < in.n; i++) {
    const P3 p = in.v[i], q = in.v[(i + 1) % in.n];
    const double fp = cw * p[k] + ca * p[a], fq = cw * q[k] + ca * q[a];
    if (fp >= 0) out.v[out.n++] = p;
    if ((fp >= 0) != (fq >= 0)) {
      const double t = fp / (fp - fq);
      out.v[out.n++] = add(p, scale(t, sub(q, p)));
    }
  }
}

void mark(float* face, int R, int row, int c0, int c1, float val) {
  c0 = std::max(c0, 0);
  c1 = std::min(c1, R - 1);
  float* r = face + (size_t)row * R;
  for (int c = c0; c <= c1; c++) r[c] = std::min(r[c], val);
}

// One triangle (relative to the light) onto cube face (k, sgn).  Face coordinates: (v[a], v[b]) / |v[k]| with
// (a, b) = (1, 2), (0, 2), (0, 1) for k = 0, 1, 2 — the kernels' light_map_cell uses the same convention.
void raster_face(float* face, int R, const P3 tri[3], int k, double sgn, float val) {
  if (sgn * tri[0][k] <= 0 && sgn * tri[1][k] <= 0 && sgn * tri[2][k] <= 0) return;
  const int a = (k == 0) ? 1 : 0, b = (k == 2) ? 1 : 2;
  const double widen = 1.0 + 4.0 / R;  // two cells beyond the face's own pyramid
  Poly p, q;
  p.n = 3;
  for (int i = 0; i < 3; i++) p.v[i] = tri[i];
  clip_plane(p, q, k, sgn * widen, a, -1.0);
  if (q.n < 3) return;
  clip_plane(q, p, k, sgn * widen, a, 1.0);
  if (p.n < 3) return;
  clip_plane(p, q, k, sgn * widen, b, -1.0);
  if (q.n < 3) return;
  clip_plane(q, p, k, sgn * widen, b, 1.0);
  if (p.n < 3) return;
  double x[12], y[12];
  double ymin = std::numeric_limits<double>::infinity(), ymax = -ymin;
  const double half = 0.5 * R;
  for (int i = 0; i < p.n; i++) {
    const double w = sgn * p.v[i][k];
    if (!(w > 1e-300)) {  // the polygon reaches the light itself: no projection; mark the whole face
      for (int row = 0; row < R; row++) mark(face, R, row, 0, R - 1, val);
      return;
    }
    x[i] = (p.v[i][a] / w + 1.0) * half;
    y[i] = (p.v[i][b] / w + 1.0) * half;
    ymin = std::min(ymin, y[i]);
    ymax = std::max(ymax, y[i]);
  }
  const int j0 = std::max(0, (int)std::floor(ymin) - 1), j1 = std::min(R - 1, (int)std::floor(ymax) + 1);
  for (int j = j0; j <= j1; j++) {
    const double y0 = j - 1.0, y1 = j + 2.0;  // the row and one row on each side
    double xmin = std::numeric_limits<double>::infinity(), xmax = -xmin;
    for (int i = 0; i < p.n; i++) {
      const int i2 = (i + 1) % p.n;
      const double ya = y[i], yb = y[i2];
      if ((ya < y0 && yb < y0) || (ya > y1 && yb > y1)) continue;
      double t0 = 0, t1 = 1;
      if (ya != yb) {
        double ta = (y0 - ya) / (yb - ya), tb = (y1 - ya) / (yb - ya);
        if (ta > tb) std::swap(ta, tb);
        t0 = std::max(t0, ta);
        t1 = std::min(t1, tb);
        if (t0 > t1) continue;
      }
      const double xa = x[i] + t0 * (x[i2] - x[i]), xb = x[i] + t1 * (x[i2] - x[i]);
      xmin = std::min(xmin, std::min(xa, xb));
      xmax = std::max(xmax, std::max(xa, xb));
    }
    if (xmin > xmax) continue;
    mark(face, R, j, (int)std::floor(xmin) - 1, (int)std::floor(xmax) + 1, val);
  }
}

}  // namespace

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
// 1e-9 rim it can reach the 1e-8 base.)  Coordinates of 1e6 and more get no flags.
bool lit_query_make(const rh_tri& t0, const double L[3], bool directional, LitQuery* q) {
  const P3 a = {t0.p0[0], t0.p0[1], t0.p0[2]};
  const P3 b = add(a, P3{t0.e1[0], t0.e1[1], t0.e1[2]}), c = add(a, P3{t0.e2[0], t0.e2[1], t0.e2[2]});
  const P3 l = {L[0], L[1], L[2]};  // the light's position, or its direction vector (Light.hs:8-9)
  double coord = 0;
  for (int k = 0; k < 3; k++) coord = std::max(coord, std::max(std::max(std::fabs(a[k]), std::fabs(b[k])), std::max(std::fabs(c[k]), std::fabs(l[k]))));
  if (!(coord < 1e6)) return false;
  const P3 ab = sub(b, a), ac = sub(c, a);
  P3 n0 = {ab.y * ac.z - ab.z * ac.y, ab.z * ac.x - ab.x * ac.z, ab.x * ac.y - ab.y * ac.x};
  const double area2 = len(n0);
  if (!(area2 > 1e-18)) return false;
  n0 = scale(1 / area2, n0);
  // Point light: height of the light above T0's plane, and |cos| >= 0.01 for every point of the (widened) triangle.
  // Directional light: the shadow ray is (p + 1e-6 d, d) with the light's own, un-normalised vector d (Light.hs:14,
  // RayHs.hs:93) and counts hits from t = 1e-6 on: at least 2e-6 |n0.d| above the plane, so |n0.d| >= 0.011 will do.
  double h = directional ? dot(n0, l) : dot(n0, sub(l, a));
  if (h < 0) {
    n0 = scale(-1, n0);
    h = -h;
  }
  if (directional) {
    if (!(h >= 0.011)) return false;
  } else {
    const double far = std::max(len(sub(l, a)), std::max(len(sub(l, b)), len(sub(l, c))));
    if (!(h >= 0.011 * (far + 1e-6))) return false;
  }
  // base: n0.(x - a) >= 1e-8
  for (int k = 0; k < 3; k++) q->n[0][k] = n0[k];
  q->d[0] = -dot(n0, a) - 1e-8;
  // sides: plane through an edge and the light (or along its direction), normal towards the third vertex, moved outward
  const double eps = 1e-9 * (1 + coord);
  const P3 v[3] = {a, b, c};
  for (int e = 0; e < 3; e++) {
    const P3 p = v[e], r = v[(e + 1) % 3], o = v[(e + 2) % 3];
    const P3 pr = sub(r, p), pl = directional ? l : sub(l, p);
    P3 m = {pr.y * pl.z - pr.z * pl.y, pr.z * pl.x - pr.x * pl.z, pr.x * pl.y - pr.y * pl.x};
    const double ml = len(m);
    if (!(ml > 1e-18)) return false;
    m = scale(1 / ml, m);
    if (dot(m, sub(o, p)) < 0) m = scale(-1, m);
    for (int k = 0; k < 3; k++) q->n[1 + e][k] = m[k];
    q->d[1 + e] = -dot(m, p) + eps;
  }
  const double inf = std::numeric_limits<double>::infinity();
  for (int k = 0; k < 3; k++) {
    q->lo[k] = std::min(a[k], std::min(b[k], c[k])) - 2 * eps;
    q->hi[k] = std::max(a[k], std::max(b[k], c[k])) + 2 * eps;
    if (directional) {  // the prism runs to infinity along d
      if (l[k] > 0) q->hi[k] = inf;
      if (l[k] < 0) q->lo[k] = -inf;
    } else {
      q->lo[k] = std::min(q->lo[k], l[k] - 2 * eps);
      q->hi[k] = std::max(q->hi[k], l[k] + 2 * eps);
    }
  }
  return true;
}

bool lit_query_box_outside(const LitQuery& q, const double* lo, const double* hi) {
  for (int k = 0; k < 3; k++)
    if (lo[k] > q.hi[k] || hi[k] < q.lo[k]) return true;
  for (int i = 0; i < 4; i++) {  // the corner of the box farthest along the plane normal
    double s = q.d[i];
    for (int k = 0; k < 3; k++) s += q.n[i][k] * (q.n[i][k] >= 0 ? hi[k] : lo[k]);
    if (s < 0) return true;
  }
  return false;
}

bool lit_query_tri_meets(const LitQuery& q, const rh_tri& t) {
  Poly p, r;
  p.n = 3;
  p.v[0] = {t.p0[0], t.p0[1], t.p0[2]};
  p.v[1] = add(p.v[0], P3{t.e1[0], t.e1[1], t.e1[2]});
  p.v[2] = add(p.v[0], P3{t.e2[0], t.e2[1], t.e2[2]});
  for (int i = 0; i < 4; i++) {  // clip to K, one half-space at a time
    r.n = 0;
    for (int j = 0; j < p.n; j++) {
      const P3 u = p.v[j], w = p.v[(j + 1) % p.n];
      const double fu = dot(P3{q.n[i][0], q.n[i][1], q.n[i][2]}, u) + q.d[i], fw = dot(P3{q.n[i][0], q.n[i][1], q.n[i][2]}, w) + q.d[i];
      if (fu >= 0) r.v[r.n++] = u;
      if ((fu >= 0) != (fw >= 0)) r.v[r.n++] = add(u, scale(fu / (fu - fw), sub(w, u)));
    }
    if (r.n == 0) return false;
    p = r;
  }
  return true;
}

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
      const P3 b = sub(add(p0, P3{t.e1[0], t.e1[1], t.e1[2]}), lp);
      const P3 c = sub(add(p0, P3{t.e2[0], t.e2[1], t.e2[2]}), lp);
      double coord = scale_abs;
      for (int q = 0; q < 3; q++) coord = std::max(coord, std::max(std::fabs(a[q]), std::max(std::fabs(b[q]), std::fabs(c[q]))));
      if (!std::isfinite(coord)) { unsafe = true; return; }
      const double d = dist_origin_triangle(a, b, c);
      if (!(d > 1e-6 * coord)) { unsafe = true; return; }
      const float val = round_down(d * (1.0 - 1e-6) - 1e-9 * coord);
      const P3 tri[3] = {a, b, c};
      raster_face(out + (size_t)face * R * R, R, tri, k, sgn, val);
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
