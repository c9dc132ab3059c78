// oracle.cpp — CPU restatement of the RayHs ray-casting path.  TEST INFRASTRUCTURE ONLY.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
// legs may load this.  The product (rayhs_b200/) never links or calls it.
//
// PARITY UNPINNED: the reference ships no golden vectors, known-answer tests or
// fixtures for this path and GHC is absent, so the reference
// itself cannot be run here.  This file follows the Haskell source line by line
// (citations are /root/reference/src/<file>:<line>), evaluates in IEEE double with
// -ffp-contract=off (GHC emits no fused multiply-adds), and is pinned only by the
// hand-derived known answers of SURVEY.md App. D (tests/test_oracle.py) and by a second,
// independent restatement of the whole path in Python (tests/pyref.py,
// tests/test_pyref_pins_oracle.py: same colours, bytes and ray counts).
//
// It includes include/rayhs_b200.h for the POD *input* structs only (rh_raw_scene,
// rh_camera, rh_material, rh_light, rh_texture); it builds its own tree with the
// KDTree.hs rule and shares no code with the product.
#include "../include/rayhs_b200.h"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <memory>
#include <thread>
#include <vector>

namespace {

// ---------------------------------------------------------------- GHC Ord Double
// GHC.Classes class defaults: every comparison with NaN is false.
inline double hs_max(double x, double y) { return (x <= y) ? y : x; }
inline double hs_min(double x, double y) { return (x <= y) ? x : y; }

// ---------------------------------------------------------------- Vec.hs
struct Vec { double x, y, z; };
struct UV { double u, v; };
struct Color { double r, g, b; };

inline Vec operator+(Vec a, Vec b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }  // Vec.hs:36
inline Vec operator-(Vec a, Vec b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }  // Vec.hs:40
inline Vec neg(Vec a) { return {-a.x, -a.y, -a.z}; }                              // Vec.hs:42
inline Vec mul(double l, Vec a) { return {l * a.x, l * a.y, l * a.z}; }            // Vec.hs:69
inline UV mul(double l, UV a) { return {l * a.u, l * a.v}; }                        // Vec.hs:75
inline UV operator+(UV a, UV b) { return {a.u + b.u, a.v + b.v}; }                  // Vec.hs:46
inline Color mul(double l, Color c) { return {l * c.r, l * c.g, l * c.b}; }         // Color.hs:59-61
inline Color operator+(Color a, Color b) { return {a.r + b.r, a.g + b.g, a.b + b.b}; }  // Color.hs:21
inline Color operator*(Color a, Color b) { return {a.r * b.r, a.g * b.g, a.b * b.b}; }  // Color.hs:23
const Color black{0, 0, 0};                                                            // Color.hs:36

inline double dot(Vec a, Vec b) { return a.x * b.x + a.y * b.y + a.z * b.z; }  // Vec.hs:105 (left-assoc)
inline Vec cross(Vec a, Vec b) {                                                 // Vec.hs:108-110
  return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
inline double sqrLen(Vec v) { return dot(v, v); }                                // Vec.hs:114
inline double sqrDist(Vec v, Vec w) { return sqrLen(v - w); }                    // Vec.hs:118
inline double dist(Vec v, Vec w) { return std::sqrt(sqrDist(v, w)); }            // Vec.hs:122
inline Vec normalize(Vec v) { return mul(1 / std::sqrt(sqrLen(v)), v); }         // Vec.hs:126
inline Vec reflect(Vec v, Vec n) { return v - mul(2 * dot(v, n), n); }           // Vec.hs:130
// Vec.hs:132-140
inline bool refract(Vec i, Vec n, double n1, double n2, Vec* out) {
  double n1n2 = n1 / n2;
  double cos0 = -(dot(i, n));
  double sin20 = n1n2 * n1n2 * (1 - cos0 * cos0);
  if (sin20 > 1) return false;
  double coeff = n1n2 * cos0 - std::sqrt(1.0 - sin20);
  *out = mul(n1n2, i) + mul(coeff, n);
  return true;
}
inline double at(Vec v, int axis) { return axis == 0 ? v.x : axis == 1 ? v.y : axis == 2 ? v.z : 0; }  // Vec.hs:55-59
inline Vec minV(Vec a, Vec b) { return {hs_min(a.x, b.x), hs_min(a.y, b.y), hs_min(a.z, b.z)}; }  // Vec.hs:144
inline Vec maxV(Vec a, Vec b) { return {hs_max(a.x, b.x), hs_max(a.y, b.y), hs_max(a.z, b.z)}; }  // Vec.hs:148

const double piInv = 1 / M_PI;  // Math.hs:11-12
const double inf = std::numeric_limits<double>::infinity();  // Math.hs:17-18
const double eps = 0.000001;    // Geometry.hs:31-32

// ---------------------------------------------------------------- Geometry.hs
struct Ray { Vec o, d; };
inline Vec rayAt(const Ray& r, double t) { return r.o + mul(t, r.d); }  // Geometry.hs:29
inline Ray rayEps(Vec p, Vec n) { return {p + mul(eps, n), n}; }         // Geometry.hs:36

struct Hit {       // Geometry.hs:39
  Vec p, n;
  UV uv;
  double t;
  int tri;         // bookkeeping only: index in `triangles mesh`, -1 for shapes
};

// Data.List.minimumBy (compare `on` t): foldl1 keeping x unless compare x y == GT;
// compare on Double: LT if x<y, EQ if x==y, else GT (so NaN => GT).
inline bool keep_first(double tx, double ty) { return (tx < ty) || (tx == ty); }

// ---------------------------------------------------------------- Mesh.hs
struct Vertex { Vec p, n; UV uv; };               // Mesh.hs:32
struct Triangle { Vertex a, b, c; int id; };      // Mesh.hs:48

// Mesh.hs:59-82
inline bool triangleIntersection(const Ray& ray, const Triangle& tr, Hit* out) {
  const Vec o = ray.o, d = ray.d;
  const Vec p0 = tr.a.p, p1 = tr.b.p, p2 = tr.c.p;
  Vec e1 = p1 - p0;
  Vec e2 = p2 - p0;
  Vec p = cross(d, e2);
  double det = dot(e1, p);
  double idet = 1 / det;
  Vec t0 = o - p0;
  double u = idet * dot(t0, p);
  Vec q = cross(t0, e1);
  double v = idet * dot(d, q);
  double t = idet * dot(e2, q);
  if (std::fabs(det) < eps || u < 0 || u > 1 || v < 0 || (u + v) > 1 || t < eps) return false;
  double w = 1 - u - v;
  // barycentricInterp u n1 v n2 (1-u-v) n0 = mul a p + mul b q + mul c r  (Mesh.hs:57)
  Vec n = mul(u, tr.b.n) + mul(v, tr.c.n) + mul(w, tr.a.n);
  UV uv = mul(u, tr.b.uv) + mul(v, tr.c.uv) + mul(w, tr.a.uv);
  *out = {rayAt(ray, t), n, uv, t, tr.id};
  return true;
}

// ---------------------------------------------------------------- KDTree.hs
struct Box { Vec lower, upper; };  // KDTree.hs:13-14

struct KDTree {                    // KDTree.hs:59-61
  enum Kind { Leaf, Node, Empty } kind = Empty;
  Box box{};
  std::vector<Triangle> tris;      // Leaf
  std::unique_ptr<KDTree> left, right;
};

inline Box include(Box b, Vec p) { return {minV(b.lower, p), maxV(b.upper, p)}; }  // KDTree.hs:19-20
inline Box buildBoundingBox(const std::vector<Triangle>& ts) {                      // KDTree.hs:22-29
  Box b{{inf, inf, inf}, {-inf, -inf, -inf}};
  for (const Triangle& t : ts) b = include(include(include(b, t.a.p), t.b.p), t.c.p);
  return b;
}
inline Vec baryCenter(const Triangle& t) { return mul(1.0 / 3, t.a.p + t.b.p + t.c.p); }  // KDTree.hs:71-74

// KDTree.hs:79-90
std::unique_ptr<KDTree> buildNode(std::vector<Triangle> tris, int depth, int axis) {
  auto node = std::make_unique<KDTree>();
  if (tris.empty()) return node;  // Empty
  Box box = buildBoundingBox(tris);
  node->box = box;
  if (tris.size() < 20 || depth >= 100) {
    node->kind = KDTree::Leaf;
    node->tris = std::move(tris);
    return node;
  }
  double split = 0.5 * (at(box.upper, axis) + at(box.lower, axis));
  std::vector<Triangle> l, r;
  for (const Triangle& t : tris) {
    double c = at(baryCenter(t), axis);
    if (c < split) l.push_back(t);
    if (split <= c) r.push_back(t);
  }
  tris.clear();
  tris.shrink_to_fit();
  int nextAxis = (axis + 1) % 3;
  node->kind = KDTree::Node;
  node->left = buildNode(std::move(l), depth + 1, nextAxis);
  node->right = buildNode(std::move(r), depth + 1, nextAxis);
  return node;
}

struct Counters {
  uint64_t rays[5] = {0, 0, 0, 0, 0};  // primary, reflect, probe, exit, shadow
  uint64_t box_tests = 0, tri_tests = 0, prim_tests = 0;
};

// KDTree.hs:39-56
inline bool rayInterBox(const Ray& r, const Box& b) {
  double idx = 1 / r.d.x, idy = 1 / r.d.y, idz = 1 / r.d.z;
  double t1 = idx * (b.lower.x - r.o.x);
  double t2 = idx * (b.upper.x - r.o.x);
  double t3 = idy * (b.lower.y - r.o.y);
  double t4 = idy * (b.upper.y - r.o.y);
  double t5 = idz * (b.lower.z - r.o.z);
  double t6 = idz * (b.upper.z - r.o.z);
  double tmin = hs_max(hs_max(hs_min(t1, t2), hs_min(t3, t4)), hs_min(t5, t6));
  double tmax = hs_min(hs_min(hs_max(t1, t2), hs_max(t3, t4)), hs_max(t5, t6));
  return !(tmax < 0 || tmin > tmax);
}

// KDTree.hs:96-115
bool rayInter(const Ray& ray, const KDTree& k, Hit* out, Counters& c) {
  switch (k.kind) {
    case KDTree::Empty:
      return false;
    case KDTree::Leaf: {
      c.box_tests++;
      if (!rayInterBox(ray, k.box)) return false;
      bool have = false;  // closestHit, Geometry.hs:54-57
      Hit best{};
      for (const Triangle& t : k.tris) {
        Hit h;
        c.tri_tests++;
        if (!triangleIntersection(ray, t, &h)) continue;
        if (!have) { best = h; have = true; }
        else if (!keep_first(best.t, h.t)) best = h;
      }
      if (have) *out = best;
      return have;
    }
    case KDTree::Node: {
      c.box_tests++;
      if (!rayInterBox(ray, k.box)) return false;
      Hit l, r;
      bool hl = rayInter(ray, *k.left, &l, c);
      bool hr = rayInter(ray, *k.right, &r, c);
      if (!hl && !hr) return false;            // minMaybeHit, KDTree.hs:109-115
      if (hl && !hr) { *out = l; return true; }
      if (!hl && hr) { *out = r; return true; }
      *out = (l.t < r.t) ? l : r;
      return true;
    }
  }
  return false;
}

// ---------------------------------------------------------------- scene
struct Object {
  int kind;
  int material;
  Vec a, b, c;       // plane: point, normal, tangent; sphere: center, radius in b.x
  std::unique_ptr<KDTree> tree;
};

struct Bitmap { int w, h; const double* px; };

struct Scene {
  std::vector<Object> shapes;
  std::vector<rh_material> materials;
  std::vector<rh_light> lights;
  std::vector<Bitmap> textures;
  std::vector<double> texels;
};

// Geometry.hs:68-96
inline bool rayShapeIntersection(const Ray& ray, const Object& ob, Hit* out) {
  const Vec o = ray.o, d = ray.d;
  if (ob.kind == RH_OBJ_PLANE) {
    const Vec p = ob.a, n = ob.b, t = ob.c;
    double dDotn = dot(d, n);
    double time = dot(n, p - o) / dDotn;
    if (std::fabs(dDotn) > 0 && time > 0) {
      Vec pos = rayAt(ray, time);
      Vec b = cross(t, n);
      Vec rel = pos - p;
      *out = {pos, n, {dot(t, rel), dot(b, rel)}, time, -1};
      return true;
    }
    return false;
  }
  const Vec ct = ob.a;
  const double r = ob.b.x;
  double a = dot(d, d);
  double b = 2.0 * dot(d, o - ct);
  double c = sqrLen(o - ct) - r * r;
  double delta = b * b - 4.0 * a * c;
  if (delta < 0.0) return false;
  auto polar = [](Vec p) { return UV{piInv * std::atan(p.z / p.x), piInv * std::acos(p.y)}; };
  double t0 = 0.5 * ((-b) - std::sqrt(delta)) / a;
  if (t0 > 0) {
    Vec p0 = rayAt(ray, t0);
    Vec n0 = normalize(p0 - ct);
    *out = {p0, n0, polar(n0), t0, -1};
    return true;
  }
  double t1 = 0.5 * ((-b) + std::sqrt(delta)) / a;
  if (t1 > 0) {
    Vec p1 = rayAt(ray, t1);
    Vec n1 = normalize(p1 - ct);
    *out = {p1, n1, polar(n1), t1, -1};
    return true;
  }
  return false;
}

struct MatHit {  // RayHs.hs:52-56
  Vec p, n;
  UV uv;
  double t;
  int material;
  int object, tri;  // bookkeeping
};

// RayHs.hs:58-62
inline bool intersection(const Ray& ray, const Scene& sc, int oi, MatHit* out, Counters& c) {
  const Object& ob = sc.shapes[oi];
  Hit h;
  bool ok;
  if (ob.kind == RH_OBJ_MESH) ok = rayInter(ray, *ob.tree, &h, c);
  else { c.prim_tests++; ok = rayShapeIntersection(ray, ob, &h); }
  if (!ok) return false;
  *out = {h.p, h.n, h.uv, h.t, ob.material, oi, h.tri};
  return true;
}

// RayHs.hs:67-71
bool closestIntersection(const Scene& sc, const Ray& r, MatHit* out, Counters& c, int cls) {
  c.rays[cls]++;
  bool have = false;
  MatHit best{};
  for (int i = 0; i < (int)sc.shapes.size(); i++) {
    MatHit h;
    if (!intersection(r, sc, i, &h, c)) continue;
    if (!have) { best = h; have = true; }
    else if (!keep_first(best.t, h.t)) best = h;
  }
  if (have) *out = best;
  return have;
}

// RayHs.hs:74-87.  Only emptiness of the filtered list is ever inspected
// (RayHs.hs:94-96), and Haskell lists are lazy, so the scan stops at the first
// object that survives both filters; the value is the same either way.
bool shadowIntersection(const Scene& sc, const rh_light& light, const Ray& ray, Counters& c) {
  c.rays[4]++;
  for (int i = 0; i < (int)sc.shapes.size(); i++) {
    MatHit h;
    if (!intersection(ray, sc, i, &h, c)) continue;
    bool inFront;
    if (light.kind == RH_LIGHT_DIRECTIONAL) inFront = true;
    else {
      Vec lp{light.vec[0], light.vec[1], light.vec[2]};
      inFront = sqrDist(ray.o, lp) > sqrDist(ray.o, h.p);
    }
    if (!inFront) continue;
    if (sc.materials[h.material].kind == RH_MAT_EMMIT) continue;  // isOccluder
    return true;
  }
  return false;
}

// Light.hs:12-17
inline void lightAt(const rh_light& l, Vec p, Vec* ld, Color* lc) {
  Color c{l.color[0], l.color[1], l.color[2]};
  Vec v{l.vec[0], l.vec[1], l.vec[2]};
  if (l.kind == RH_LIGHT_DIRECTIONAL) { *ld = v; *lc = c; return; }
  double d = dist(v, p);
  double s = 1.0 + d / l.radius;
  double falloff = 1.0 / (s * s);
  *ld = mul(1 / d, v - p);
  *lc = mul(falloff, c);
}

// Material.hs:22-33
inline double r0(double n1, double n2) { double q = (n1 - n2) / (n1 + n2); return q * q; }
inline double fresnel(double ior, double cos0) {
  double r = r0(1.0, ior);
  double x = 1 - cos0;
  double x2 = x * x;
  double x5 = (x2 * x2) * x;  // GHC (^) square-and-multiply, SURVEY App. A-S7
  return r + (1 - r) * x5;
}
inline Color diffuse(Color cd, Color lc, Vec l, Vec n) { return mul(hs_max(dot(l, n), 0) * piInv, cd * lc); }

// Data.Fixed.mod' n d = n - fromInteger (floor (toRational n / toRational d)) * d : exact rational floor.
inline double hs_mod1(double n, double d) {
  if (!std::isfinite(n) || !std::isfinite(d) || d == 0) return std::numeric_limits<double>::quiet_NaN();
  double q = std::floor(n / d);
  if (std::fabs(q) >= 4503599627370496.0) return n - q * d;  // no exact fix-up possible / needed
  // fix q so that q*d <= n < (q+1)*d holds exactly (sign of a correctly rounded fma is exact)
  auto rem = [&](double k) { return std::fma(-k, d, n); };
  if (d > 0) {
    while (rem(q) < 0) q -= 1;
    while (rem(q + 1) >= 0) q += 1;
  } else {
    while (rem(q) > 0) q -= 1;
    while (rem(q + 1) <= 0) q += 1;
  }
  return n - q * d;
}

inline long long hs_mod_int(long long a, long long m) { long long r = a % m; return (r != 0 && ((r < 0) != (m < 0))) ? r + m : r; }

// ColorMap.hs:18-58
Color colorAt(const Scene& sc, const rh_material& m, UV uv) {
  Color c1{m.color1[0], m.color1[1], m.color1[2]};
  if (m.cmap_kind == RH_CMAP_FLAT) return c1;
  if (m.cmap_kind == RH_CMAP_CHECKER) {
    Color c2{m.color2[0], m.color2[1], m.color2[2]};
    double s = m.size;
    return ((hs_mod1(uv.u, s) - (0.5 * s)) * (hs_mod1(uv.v, s) - (0.5 * s)) < 0) ? c1 : c2;
  }
  const Bitmap& bm = sc.textures[m.texture];
  double u = hs_mod1(uv.u, 1) * (double)bm.w;  // toPixel / repeatUV
  double v = hs_mod1(uv.v, 1) * (double)bm.h;
  long long ui = (long long)std::nearbyint(u);  // round: half to even
  long long vi = (long long)std::nearbyint(v);
  long long x0 = hs_mod_int(ui - 1, bm.w), x1 = hs_mod_int(ui, bm.w);
  long long y0 = hs_mod_int(vi - 1, bm.h), y1 = hs_mod_int(vi, bm.h);
  double lx = u - (double)(ui - 1) - 0.5;
  double ly = v - (double)(vi - 1) - 0.5;
  auto px = [&](long long i, long long j) {  // Bitmap.hs:17-18
    const double* p = bm.px + 3 * (i + (long long)bm.w * j);
    return Color{p[0], p[1], p[2]};
  };
  Color c0 = px(x0, y0), c1t = px(x1, y0), c2 = px(x0, y1), c3 = px(x1, y1);
  Color cx0 = mul(lx, c1t) + mul(1 - lx, c0);  // bilinearInterp, ColorMap.hs:41-45
  Color cx1 = mul(lx, c3) + mul(1 - lx, c2);
  return mul(ly, cx1) + mul(1 - ly, cx0);
}

Color traceRay(const Scene& sc, int depth, int maxDepth, const Ray& ray, Counters& c, int cls, int* hit_obj, int* hit_tri);

// RayHs.hs:89-97
Color accumDiffuse(const Scene& sc, Vec p, Vec n, Color color, Counters& c) {
  Color acc = black;
  for (const rh_light& l : sc.lights) {
    Vec ld;
    Color lc;
    lightAt(l, p, &ld, &lc);
    bool shadowed = shadowIntersection(sc, l, rayEps(p, ld), c);
    acc = acc + (shadowed ? black : diffuse(color, lc, ld, n));
  }
  return acc;
}

// RayHs.hs:99-104
Color specular(const Scene& sc, int depth, int maxDepth, Vec v, Vec p, Vec n, Counters& c) {
  if (depth < maxDepth) {
    Vec rdir = reflect(v, n);
    return mul(dot(rdir, n), traceRay(sc, depth + 1, maxDepth, rayEps(p, rdir), c, 1, nullptr, nullptr));
  }
  return black;
}

// RayHs.hs:107-147
Color irradiance(int d, int maxDepth, const Scene& sc, const rh_material& m, Vec v, Vec p, Vec n, UV uv, Counters& c) {
  switch (m.kind) {
    case RH_MAT_DIFFUSE: {
      Color cd = colorAt(sc, m, uv);
      return mul(0.2, cd) + accumDiffuse(sc, p, n, cd, c);
    }
    case RH_MAT_PLASTIC: {
      Color cd = colorAt(sc, m, uv);
      Color diff = accumDiffuse(sc, p, n, cd, c);
      return diff + mul(fresnel(m.ior, dot(n, neg(v))), specular(sc, d, maxDepth, v, p, n, c));
    }
    case RH_MAT_MIRROR:
      return mul(fresnel(m.ior, dot(n, neg(v))), specular(sc, d, maxDepth, v, p, n, c));
    case RH_MAT_EMMIT:
      return {m.color1[0], m.color1[1], m.color1[2]};
    case RH_MAT_TRANSPARENT: {
      bool have = false;
      Color radiance = black;
      if (d != maxDepth) {
        Vec refDir;
        if (refract(v, n, 1.0, m.ior, &refDir)) {
          Ray refr = rayEps(p, refDir);
          MatHit out;
          if (closestIntersection(sc, refr, &out, c, 2)) {
            Vec outDir;
            if (refract(refDir, neg(out.n), m.ior, 1.0, &outDir)) {
              radiance = traceRay(sc, d + 1, maxDepth, rayEps(out.p, outDir), c, 3, nullptr, nullptr);
              have = true;
            }
          }
        }
      }
      Color spec = mul(fresnel(m.ior, dot(n, neg(v))), specular(sc, d, maxDepth, v, p, n, c));
      if (!have) return spec;
      return mul(1 - r0(m.ior, 1.0), radiance) + spec;
    }
    case RH_MAT_SHOWNORMAL:
      return {n.x, n.y, n.z};
    case RH_MAT_SHOWUV:
      return {uv.u, uv.v, 0};
  }
  return black;
}

// RayHs.hs:149-154
Color traceRay(const Scene& sc, int depth, int maxDepth, const Ray& ray, Counters& c, int cls, int* hit_obj, int* hit_tri) {
  MatHit h;
  if (closestIntersection(sc, ray, &h, c, cls)) {
    if (hit_obj) { *hit_obj = h.object; *hit_tri = h.tri; }
    return irradiance(depth, maxDepth, sc, sc.materials[h.material], ray.d, h.p, h.n, h.uv, c);
  }
  if (hit_obj) { *hit_obj = -1; *hit_tri = -1; }
  return black;
}

// ---------------------------------------------------------------- Projection.hs / Mat.hs
struct Mat3 { double a, b, c, d, e, f, g, h, i; };
inline Vec apply(const Mat3& m, Vec v) {  // Mat.hs:40-44
  return {m.a * v.x + m.b * v.y + m.c * v.z, m.d * v.x + m.e * v.y + m.f * v.z, m.g * v.x + m.h * v.y + m.i * v.z};
}
inline Mat3 fromColumns(Vec v1, Vec v2, Vec v3) {  // Mat.hs:83-87
  return {v1.x, v2.x, v3.x, v1.y, v2.y, v3.y, v1.z, v2.z, v3.z};
}
inline Mat3 lookAt(Vec pos, Vec target, Vec tup) {  // Mat.hs:89-93
  Vec forward = normalize(target - pos);
  Vec right = normalize(cross(tup, forward));
  Vec up = cross(forward, right);
  return fromColumns(right, up, forward);
}
inline void aspectSize(double w, double h, double* apw, double* aph) {  // Projection.hs:41-46
  double aspect = w / h;
  if (aspect > 1) { *apw = w; *aph = w / aspect; }
  else { *apw = aspect * h; *aph = h; }
}
// Projection.hs:22-39
inline Ray rayFromPixel(double w, double h, const rh_camera& cam, double px, double py) {
  Vec p{cam.position[0], cam.position[1], cam.position[2]};
  Vec t{cam.target[0], cam.target[1], cam.target[2]};
  Vec up{cam.up[0], cam.up[1], cam.up[2]};
  double apw, aph;
  aspectSize(w, h, &apw, &aph);
  Vec o, d;
  if (cam.projection == RH_PROJ_ORTHOGRAPHIC) {
    o = {apw * (px - (w / 2)) / w, aph * ((-py) + (h / 2)) / h, 0};
    d = {0, 0, 1};
  } else {
    double f = 0.5 * h / (std::tan(0.5) * cam.fovy);  // precedence quirk, SURVEY App. A-C1
    Vec viewPlanePos{apw * (px - (w / 2)) / w, aph * ((-py) + (h / 2)) / h, f};
    d = normalize(viewPlanePos);
    o = {0, 0, 0};
  }
  Mat3 mat = lookAt(p, t, up);
  return {o + p, apply(mat, d)};
}

// Image.hs:54-55: truncate (255 * min c 1)
inline long long toIntC(double c) {
  double v = 255 * hs_min(c, 1);  // hs_min NaN 1 = 1, so NaN prints 255
  if (v < -9.0e18) return INT64_MIN;
  return (long long)v;  // toward zero
}

struct OracleScene {
  Scene sc;
};

}  // namespace

// ================================================================== C interface
extern "C" {

struct orc_result_counts {
  uint64_t rays[5];
  uint64_t box_tests, tri_tests, prim_tests;
  double seconds;
  int32_t threads;
  int32_t pad_;
};

void* orc_scene_create(const rh_raw_scene* raw) {
  auto* os = new OracleScene();
  Scene& sc = os->sc;
  sc.materials.assign(raw->materials, raw->materials + raw->n_materials);
  sc.lights.assign(raw->lights, raw->lights + raw->n_lights);
  sc.texels.assign(raw->texels, raw->texels + 3 * raw->n_texels);
  for (uint32_t i = 0; i < raw->n_textures; i++)
    sc.textures.push_back({raw->textures[i].w, raw->textures[i].h, sc.texels.data() + 3 * raw->textures[i].offset});
  for (uint32_t i = 0; i < raw->n_objects; i++) {
    const rh_raw_object& ro = raw->objects[i];
    Object ob;
    ob.kind = ro.kind;
    ob.material = ro.material;
    ob.a = {ro.a[0], ro.a[1], ro.a[2]};
    ob.b = {ro.b[0], ro.b[1], ro.b[2]};
    ob.c = {ro.c[0], ro.c[1], ro.c[2]};
    if (ro.kind == RH_OBJ_MESH) {
      // Mesh.hs:105-109 `triangles`: indices grouped in threes
      std::vector<Triangle> tris;
      auto vert = [&](uint32_t k) {
        return Vertex{{ro.positions[3 * k], ro.positions[3 * k + 1], ro.positions[3 * k + 2]},
                      {ro.normals[3 * k], ro.normals[3 * k + 1], ro.normals[3 * k + 2]},
                      {ro.uvs[2 * k], ro.uvs[2 * k + 1]}};
      };
      for (uint32_t k = 0; k + 2 < ro.n_indices; k += 3)
        tris.push_back({vert(ro.indices[k]), vert(ro.indices[k + 1]), vert(ro.indices[k + 2]), (int)(k / 3)});
      ob.tree = buildNode(std::move(tris), 0, 0);  // KDTree.hs:68-69
    }
    sc.shapes.push_back(std::move(ob));
  }
  return os;
}

void orc_scene_destroy(void* s) { delete (OracleScene*)s; }

static void tree_stats(const KDTree& k, int depth, uint32_t* inner, uint32_t* leaves, uint32_t* empties, uint32_t* maxdepth,
                       uint32_t* maxleaf) {
  if (k.kind == KDTree::Empty) { (*empties)++; return; }
  if ((uint32_t)depth > *maxdepth) *maxdepth = depth;
  if (k.kind == KDTree::Leaf) {
    (*leaves)++;
    if (k.tris.size() > *maxleaf) *maxleaf = (uint32_t)k.tris.size();
    return;
  }
  (*inner)++;
  tree_stats(*k.left, depth + 1, inner, leaves, empties, maxdepth, maxleaf);
  tree_stats(*k.right, depth + 1, inner, leaves, empties, maxdepth, maxleaf);
}

// out[5] = inner, leaves, empty children, max depth, max leaf size for object `obj`
int orc_tree_stats(void* s, int obj, uint32_t* out) {
  Scene& sc = ((OracleScene*)s)->sc;
  if (obj < 0 || obj >= (int)sc.shapes.size() || !sc.shapes[obj].tree) return -1;
  out[0] = out[1] = out[2] = out[3] = out[4] = 0;
  tree_stats(*sc.shapes[obj].tree, 0, &out[0], &out[1], &out[2], &out[3], &out[4]);
  return 0;
}

// Camera ray for tests (Projection.hs:22-25): out[6] = origin, direction
void orc_ray_from_pixel(const rh_camera* cam, double w, double h, double px, double py, double* out) {
  Ray r = rayFromPixel(w, h, *cam, px, py);
  out[0] = r.o.x; out[1] = r.o.y; out[2] = r.o.z;
  out[3] = r.d.x; out[4] = r.d.y; out[5] = r.d.z;
}

// One closestIntersection for tests: returns 1 on hit; out = p(3) n(3) uv(2) t(1); ids = object, tri
int orc_closest(void* s, const double* o, const double* d, double* out, int32_t* ids) {
  Scene& sc = ((OracleScene*)s)->sc;
  Counters c;
  MatHit h;
  Ray r{{o[0], o[1], o[2]}, {d[0], d[1], d[2]}};
  if (!closestIntersection(sc, r, &h, c, 0)) return 0;
  out[0] = h.p.x; out[1] = h.p.y; out[2] = h.p.z;
  out[3] = h.n.x; out[4] = h.n.y; out[5] = h.n.z;
  out[6] = h.uv.u; out[7] = h.uv.v; out[8] = h.t;
  ids[0] = h.object; ids[1] = h.tri;
  return 1;
}

// colorAt for tests
void orc_color_at(void* s, int material, double u, double v, double* out) {
  Scene& sc = ((OracleScene*)s)->sc;
  Color c = colorAt(sc, sc.materials[material], {u, v});
  out[0] = c.r; out[1] = c.g; out[2] = c.b;
}

double orc_mod1(double n, double d) { return hs_mod1(n, d); }

// rayTrace (RayHs.hs:161-166) / distributedRayTrace (RayHs.hs:190-195) over the rows
// row_begin, row_begin+row_step, ... < row_end (a bounded sample for the CPU baseline;
// 0, h, 1 = the whole frame).  Pixel order and the 10-pixel chunking follow
// Image.hs:31-36 (the extra pixel w*h is computed by the reference and then dropped
// by the writer, Image.hs:63-65; it is not rendered here and not counted).
// offsets: double[w*h][spp][2] (already x-0.5, y-0.5; RayHs.hs:185-188) or NULL for the
// 1-sample path.  Outputs (any may be NULL) are indexed by the FULL frame:
//   rgb_f64[w*h*3] raw colour, rgb_u8[w*h*3] clamped toIntC, rgb_int[w*h*3] raw toIntC,
//   hit_ids[w*h*spp*2] (object, tri).
// Value k (0-based) of the SplitMix64 stream rh_sample_offsets_f64(seed, ...) writes, as a double in [0, 1): the
// generator's state after k + 1 steps is seed + (k + 1) * gamma, so any value can be produced on its own.
static double splitmix01_at(uint64_t seed, uint64_t k) {
  uint64_t z = seed + (k + 1) * 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z = z ^ (z >> 31);
  return (double)(z >> 11) * (1.0 / 9007199254740992.0);
}

// rayTrace (RayHs.hs:161-166) / distributedRayTrace (RayHs.hs:190-195) over the pixels (row, col) with
// row in row_begin:row_end:row_step and col in col_begin:col_end:col_step.  Sample offsets: `offsets` (full-frame array,
// pixel-major) or, when it is null and spp_seeded != 0, the SplitMix64 stream of `seed` evaluated in place — the frames
// of configs[4] have 34 GB of offsets.  Output arrays are full-frame; only the selected pixels are written.
int orc_render2(void* s, const rh_camera* cam, int w, int h, int maxDepth, int spp, const double* offsets, int spp_seeded,
                uint64_t seed, int row_begin, int row_end, int row_step, int col_begin, int col_end, int col_step, int compact_out, int n_threads,
                double* rgb_f64, uint8_t* rgb_u8, int32_t* rgb_int, int32_t* hit_ids, orc_result_counts* counts) {
  Scene& sc = ((OracleScene*)s)->sc;
  if (n_threads <= 0) n_threads = (int)std::thread::hardware_concurrency();
  if (n_threads <= 0) n_threads = 1;
  if (row_step <= 0) row_step = 1;
  if (col_step <= 0) col_step = 1;
  std::vector<int> rows, cols;
  for (int y = row_begin; y < row_end && y < h; y += row_step) rows.push_back(y);
  for (int x = col_begin; x < col_end && x < w; x += col_step) cols.push_back(x);
  const long long nc = (long long)cols.size();
  const long long npix = (long long)rows.size() * nc;
  const long long nchunks = (npix + 9) / 10;  // parListChunk 10, Image.hs:36
  std::atomic<long long> next{0};
  std::vector<Counters> tc(n_threads);
  const double dw = (double)w, dh = (double)h;
  const bool sampled = offsets != nullptr || spp_seeded != 0;
  auto t_begin = std::chrono::steady_clock::now();
  auto worker = [&](int tid) {
    Counters& c = tc[tid];
    for (;;) {
      long long ch = next.fetch_add(1);
      if (ch >= nchunks) break;
      for (long long k = ch * 10; k < std::min(npix, ch * 10 + 10); k++) {
        int y = rows[k / nc], x = cols[k % nc];
        const long long i = (long long)y * w + x;  // pixelCoord, Image.hs:31-32
        const long long io = compact_out ? k : i;     // where the pixel's results go
        double pi_ = (double)x, pj = (double)y;
        Color col;
        if (!sampled) {
          int ho, ht;
          col = traceRay(sc, 0, maxDepth, rayFromPixel(dw, dh, *cam, pi_, pj), c, 0, &ho, &ht);  // tracePixel, RayHs.hs:156-159
          if (hit_ids) { hit_ids[2 * io] = ho; hit_ids[2 * io + 1] = ht; }
        } else {
          Color sumc = black;  // average, RayHs.hs:169-171
          for (int sidx = 0; sidx < spp; sidx++) {
            double of[2];
            const uint64_t g = (uint64_t)i * (uint64_t)spp + (uint64_t)sidx;
            if (offsets) {
              of[0] = offsets[2 * g];
              of[1] = offsets[2 * g + 1];
            } else {  // (x - 0.5, y - 0.5), x drawn before y (RayHs.hs:185-188)
              of[0] = splitmix01_at(seed, 2 * g) - 0.5;
              of[1] = splitmix01_at(seed, 2 * g + 1) - 0.5;
            }
            int ho, ht;
            Color cs = traceRay(sc, 0, maxDepth, rayFromPixel(dw, dh, *cam, pi_ + of[0], pj + of[1]), c, 0, &ho, &ht);
            if (hit_ids) { hit_ids[2 * (io * spp + sidx)] = ho; hit_ids[2 * (io * spp + sidx) + 1] = ht; }
            sumc = sumc + cs;
          }
          col = mul(1.0 / (double)spp, sumc);
        }
        if (rgb_f64) { rgb_f64[3 * io] = col.r; rgb_f64[3 * io + 1] = col.g; rgb_f64[3 * io + 2] = col.b; }
        long long q[3] = {toIntC(col.r), toIntC(col.g), toIntC(col.b)};
        for (int k2 = 0; k2 < 3; k2++) {
          if (rgb_int) rgb_int[3 * io + k2] = (int32_t)std::max<long long>(INT32_MIN, std::min<long long>(INT32_MAX, q[k2]));
          if (rgb_u8) rgb_u8[3 * io + k2] = (uint8_t)std::max<long long>(0, std::min<long long>(255, q[k2]));
        }
      }
    }
  };
  std::vector<std::thread> th;
  for (int t = 1; t < n_threads; t++) th.emplace_back(worker, t);
  worker(0);
  for (auto& t : th) t.join();
  auto t_end = std::chrono::steady_clock::now();
  if (counts) {
    memset(counts, 0, sizeof(*counts));
    for (const Counters& c : tc) {
      for (int k = 0; k < 5; k++) counts->rays[k] += c.rays[k];
      counts->box_tests += c.box_tests;
      counts->tri_tests += c.tri_tests;
      counts->prim_tests += c.prim_tests;
    }
    counts->seconds = std::chrono::duration<double>(t_end - t_begin).count();
    counts->threads = n_threads;
  }
  return 0;
}

int orc_render(void* s, const rh_camera* cam, int w, int h, int maxDepth, int spp, const double* offsets, int row_begin,
               int row_end, int row_step, int n_threads, double* rgb_f64, uint8_t* rgb_u8, int32_t* rgb_int,
               int32_t* hit_ids, orc_result_counts* counts) {
  return orc_render2(s, cam, w, h, maxDepth, spp, offsets, 0, 0, row_begin, row_end, row_step, 0, w, 1, 0, n_threads, rgb_f64, rgb_u8,
                     rgb_int, hit_ids, counts);
}

}  // extern "C"
