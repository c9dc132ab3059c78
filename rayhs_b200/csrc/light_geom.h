// light_geom.h — the geometry behind the light-space tables (cube maps of the nearest possible occluder distance, lit
// triangles), written once for both builders: the host one (light_maps.cpp; also what the CPU tests exercise through
// rh_light_map_build / rh_lit_triangles) and the CUDA one (setup_kernels.cu, what rh_scene_create runs).  Plain IEEE
// double arithmetic in a fixed order, no fused multiply-add on either side (-ffp-contract=off / -fmad=false), so both
// produce the same tables bit for bit.  See light_maps.cpp for why the tables are conservative.
#pragma once
#include <math.h>
#include <stdint.h>

#include "../../include/rayhs_b200.h"

#ifdef __CUDACC__
#define RH_HD __host__ __device__ __forceinline__
#else
#define RH_HD inline
#endif

namespace rh {

// K = { x : n[i].x + d[i] >= 0 for all i } with its bounding box (common.h explains what K is)
struct LitQuery {
  double n[4][3], d[4];
  double lo[3], hi[3];
};

namespace lg {

struct P3 {
  double x, y, z;
  RH_HD double operator[](int k) const { return k == 0 ? x : (k == 1 ? y : z); }
};
RH_HD P3 mk(double x, double y, double z) { P3 p; p.x = x; p.y = y; p.z = z; return p; }
RH_HD P3 sub(P3 a, P3 b) { return mk(a.x - b.x, a.y - b.y, a.z - b.z); }
RH_HD P3 add(P3 a, P3 b) { return mk(a.x + b.x, a.y + b.y, a.z + b.z); }
RH_HD P3 scale(double s, P3 a) { return mk(s * a.x, s * a.y, s * a.z); }
RH_HD double dot(P3 a, P3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
RH_HD double len(P3 a) { return sqrt(dot(a, a)); }
RH_HD P3 cross(P3 a, P3 b) { return mk(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
// std::min / std::max semantics (the first argument unless the second is strictly smaller / larger)
RH_HD double mn(double a, double b) { return b < a ? b : a; }
RH_HD double mx(double a, double b) { return a < b ? b : a; }
RH_HD double inf() { return (double)INFINITY; }

// Distance from the origin to the closest point of triangle (a, b, c): the Voronoi-region walk over vertices,
// edges and face.  Any point it returns lies on the triangle, so a rounding slip in the region choice costs
// O(ulp) of the distance; degenerate triangles fall back to a bound that needs no division.
RH_HD double dist_origin_triangle(P3 a, P3 b, P3 c) {
  const double da = len(a), db = len(b), dc = len(c);
  const double min_vertex = mn(da, mn(db, dc));
  const P3 ab = sub(b, a), ac = sub(c, a), bc = sub(c, b);
  const double longest = mx(len(ab), mx(len(ac), len(bc)));
  const double loose = mx(0.0, min_vertex - longest);  // every point of T is within `longest` of a vertex
  const P3 cr = cross(ab, ac);
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
  if (!isfinite(d)) return loose;
  return mx(loose, mn(d, min_vertex));
}

RH_HD float round_down(double x) {
  float f = (float)x;
  if ((double)f > x) f = nextafterf(f, -INFINITY);
  return f;
}

struct Poly {
  P3 v[12];
  int n;
};

// Sutherland-Hodgman step: keep the part of `in` with f(v) = cw * v[k] + ca * v[a] >= 0 (a plane through the light).
RH_HD void clip_plane(const Poly& in, Poly& out, int k, double cw, int a, double ca) {
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

// The triangle (a, b, c) relative to the light -> the value its cells get, or false when the map is unsafe (a triangle
// (nearly) touches the light, or a coordinate is not finite).  scale_abs = max(1, |L|_inf).
RH_HD bool light_map_value(P3 a, P3 b, P3 c, double scale_abs, float* val) {
  double coord = scale_abs;
  for (int q = 0; q < 3; q++) coord = mx(coord, mx(fabs(a[q]), mx(fabs(b[q]), fabs(c[q]))));
  if (!isfinite(coord)) return false;
  const double d = dist_origin_triangle(a, b, c);
  if (!(d > 1e-6 * coord)) return false;
  *val = round_down(d * (1.0 - 1e-6) - 1e-9 * coord);
  return true;
}

// One triangle (relative to the light) onto cube face (k, sgn).  Face coordinates: (v[a], v[b]) / |v[k]| with
// (a, b) = (1, 2), (0, 2), (0, 1) for k = 0, 1, 2 — the kernels' light_map_cell uses the same convention.
// mark(row, c0, c1): cells c0..c1 (already clamped to the face) of `row` take min(cell, val).
template <class Mark>
RH_HD void raster_face(int R, const P3 tri[3], int k, double sgn, Mark& mark) {
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
  double ymin = inf(), ymax = -inf();
  const double half = 0.5 * R;
  for (int i = 0; i < p.n; i++) {
    const double w = sgn * p.v[i][k];
    if (!(w > 1e-300)) {  // the polygon reaches the light itself: no projection; mark the whole face
      for (int row = 0; row < R; row++) mark(row, 0, R - 1);
      return;
    }
    x[i] = (p.v[i][a] / w + 1.0) * half;
    y[i] = (p.v[i][b] / w + 1.0) * half;
    ymin = mn(ymin, y[i]);
    ymax = mx(ymax, y[i]);
  }
  const int jlo = (int)floor(ymin) - 1, jhi = (int)floor(ymax) + 1;
  const int j0 = jlo > 0 ? jlo : 0, j1 = jhi < R - 1 ? jhi : R - 1;
  for (int j = j0; j <= j1; j++) {
    const double y0 = j - 1.0, y1 = j + 2.0;  // the row and one row on each side
    double xmin = inf(), xmax = -inf();
    for (int i = 0; i < p.n; i++) {
      const int i2 = (i + 1) % p.n;
      const double ya = y[i], yb = y[i2];
      if ((ya < y0 && yb < y0) || (ya > y1 && yb > y1)) continue;
      double t0 = 0, t1 = 1;
      if (ya != yb) {
        double ta = (y0 - ya) / (yb - ya), tb = (y1 - ya) / (yb - ya);
        if (ta > tb) { const double s = ta; ta = tb; tb = s; }
        t0 = mx(t0, ta);
        t1 = mn(t1, tb);
        if (t0 > t1) continue;
      }
      const double xa = x[i] + t0 * (x[i2] - x[i]), xb = x[i] + t1 * (x[i2] - x[i]);
      xmin = mn(xmin, mn(xa, xb));
      xmax = mx(xmax, mx(xa, xb));
    }
    if (xmin > xmax) continue;
    const int clo = (int)floor(xmin) - 1, chi = (int)floor(xmax) + 1;
    mark(j, clo > 0 ? clo : 0, chi < R - 1 ? chi : R - 1);
  }
}

// ------------------------------------------------------------------ lit triangles (see light_maps.cpp)
RH_HD bool lit_query_make(const rh_tri& t0, const double L[3], bool directional, LitQuery* q) {
  const P3 a = mk(t0.p0[0], t0.p0[1], t0.p0[2]);
  const P3 b = add(a, mk(t0.e1[0], t0.e1[1], t0.e1[2])), c = add(a, mk(t0.e2[0], t0.e2[1], t0.e2[2]));
  const P3 l = mk(L[0], L[1], L[2]);  // the light's position, or its direction vector (Light.hs:8-9)
  double coord = 0;
  for (int k = 0; k < 3; k++) coord = mx(coord, mx(mx(fabs(a[k]), fabs(b[k])), mx(fabs(c[k]), fabs(l[k]))));
  if (!(coord < 1e6)) return false;
  const P3 ab = sub(b, a), ac = sub(c, a);
  P3 n0 = cross(ab, ac);
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
    const double far = mx(len(sub(l, a)), mx(len(sub(l, b)), len(sub(l, c))));
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
    P3 m = cross(pr, pl);
    const double ml = len(m);
    if (!(ml > 1e-18)) return false;
    m = scale(1 / ml, m);
    if (dot(m, sub(o, p)) < 0) m = scale(-1, m);
    for (int k = 0; k < 3; k++) q->n[1 + e][k] = m[k];
    q->d[1 + e] = -dot(m, p) + eps;
  }
  for (int k = 0; k < 3; k++) {
    q->lo[k] = mn(a[k], mn(b[k], c[k])) - 2 * eps;
    q->hi[k] = mx(a[k], mx(b[k], c[k])) + 2 * eps;
    if (directional) {  // the prism runs to infinity along d
      if (l[k] > 0) q->hi[k] = inf();
      if (l[k] < 0) q->lo[k] = -inf();
    } else {
      q->lo[k] = mn(q->lo[k], l[k] - 2 * eps);
      q->hi[k] = mx(q->hi[k], l[k] + 2 * eps);
    }
  }
  return true;
}

RH_HD bool lit_query_box_outside(const LitQuery& q, const double* lo, const double* hi) {
  for (int k = 0; k < 3; k++)
    if (lo[k] > q.hi[k] || hi[k] < q.lo[k]) return true;
  for (int i = 0; i < 4; i++) {  // the corner of the box farthest along the plane normal
    double s = q.d[i];
    for (int k = 0; k < 3; k++) s += q.n[i][k] * (q.n[i][k] >= 0 ? hi[k] : lo[k]);
    if (s < 0) return true;
  }
  return false;
}

RH_HD bool lit_query_tri_meets(const LitQuery& q, const rh_tri& t) {
  P3 pv[12], rv[12];  // a triangle clipped by four half-spaces has at most 7 vertices
  int pn = 3;
  pv[0] = mk(t.p0[0], t.p0[1], t.p0[2]);
  pv[1] = add(pv[0], mk(t.e1[0], t.e1[1], t.e1[2]));
  pv[2] = add(pv[0], mk(t.e2[0], t.e2[1], t.e2[2]));
  for (int i = 0; i < 4; i++) {  // clip to K, one half-space at a time
    int rn = 0;
    const P3 ni = mk(q.n[i][0], q.n[i][1], q.n[i][2]);
    for (int j = 0; j < pn; j++) {
      const P3 u = pv[j], w = pv[(j + 1) % pn];
      const double fu = dot(ni, u) + q.d[i], fw = dot(ni, w) + q.d[i];
      if (rn > 10) return true;  // (vertices within rounding of a plane, signs alternating: call it a meeting)
      if (fu >= 0) rv[rn++] = u;
      if ((fu >= 0) != (fw >= 0)) rv[rn++] = add(u, scale(fu / (fu - fw), sub(w, u)));
    }
    if (rn == 0) return false;
    pn = rn;
    for (int j = 0; j < rn; j++) pv[j] = rv[j];
  }
  return true;
}

}  // namespace lg
}  // namespace rh
