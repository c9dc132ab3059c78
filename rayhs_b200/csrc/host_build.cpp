// host_build.cpp — host-side tree build + flattening (C++; GHC is not in this image).
// Restates `buildKDTree` / `buildNode` (KDTree.hs:68-90) on index arrays and lays the
// result out as the flat arrays of include/rayhs_b200.h (rh_node / rh_tri / rh_tri_shade).
// A Haskell host would produce the same arrays from its own `KDTree` value
// (INTEGRATION.md); this file is what the GHC-less harness and the synthetic scene use.
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <memory>
#include <new>
#include <thread>
#include <vector>

#include "common.h"

namespace {

template <class T>
struct AlignedArray {  // 32-byte aligned, as the ABI asks for node / triangle arrays
  T* p = nullptr;
  size_t n = 0;
  ~AlignedArray() { free(p); }
  void assign(const std::vector<T>& v) {
    free(p);
    p = nullptr;
    n = v.size();
    if (!n) return;
    size_t bytes = ((n * sizeof(T) + 31) / 32) * 32;
    p = (T*)aligned_alloc(32, bytes);
    if (!p) throw std::bad_alloc();
    memcpy(p, v.data(), n * sizeof(T));
  }
};

struct MeshView {
  const double* pos;
  const double* nrm;
  const double* uv;
  const uint32_t* idx;
};

// Runs fn(begin, end) over [0, n) on up to `threads` host threads (contiguous ranges).  fn must not throw.
template <class F>
void parallel_ranges(size_t n, unsigned threads, F fn) {
  threads = (unsigned)std::max<size_t>(1, std::min<size_t>(threads, n / 65536 + 1));
  if (threads == 1) {
    fn((size_t)0, n);
    return;
  }
  std::vector<std::thread> pool;
  for (unsigned t = 0; t < threads; t++) pool.emplace_back([=]() { fn(n * t / threads, n * (t + 1) / threads); });
  for (std::thread& t : pool) t.join();
}

// One subtree in its own arrays: nodes in preorder with LOCAL indices (children, first triangle slot, leaf number),
// triangles as ids in left-to-right leaf order.
struct Sub {
  std::vector<rh_node> nodes;
  std::vector<uint32_t> order;
  uint32_t n_leaves = 0, max_depth = 0;
};

struct Builder {
  MeshView m;
  std::vector<double> centroid;  // [tri][3]
  std::vector<double> tbox;      // [tri][6]: lo.xyz, hi.xyz of the triangle's three corners
  bool has_nan = false;          // a NaN coordinate: GHC's min / max depend on the fold order, keep the literal fold
  std::atomic<bool> failed{false};

  const double* P(uint32_t tri, int corner) const { return m.pos + 3 * (size_t)m.idx[3 * (size_t)tri + corner]; }

  void prepare(uint32_t nt, unsigned threads) {
    centroid.resize((size_t)nt * 3);
    tbox.resize((size_t)nt * 6);
    std::atomic<bool> nan{false};
    parallel_ranges(nt, threads, [&](size_t b, size_t e) {
      bool bad = false;
      for (size_t t = b; t < e; t++) {
        const double *p0 = P((uint32_t)t, 0), *p1 = P((uint32_t)t, 1), *p2 = P((uint32_t)t, 2);
        for (int k = 0; k < 3; k++) {
          centroid[3 * t + k] = (1.0 / 3) * ((p0[k] + p1[k]) + p2[k]);  // baryCenter, KDTree.hs:71-74: mul (1/3) (a + b + c)
          tbox[6 * t + k] = std::min(p0[k], std::min(p1[k], p2[k]));
          tbox[6 * t + 3 + k] = std::max(p0[k], std::max(p1[k], p2[k]));
          bad |= (p0[k] != p0[k]) | (p1[k] != p1[k]) | (p2[k] != p2[k]);
        }
      }
      if (bad) nan = true;
    });
    has_nan = nan.load();
  }

  // buildBoundingBox, KDTree.hs:22-29: foldl include over every corner of every triangle, GHC min / max.  Without
  // NaNs the fold is a plain minimum / maximum, taken here over the per-triangle boxes.
  void node_box(const std::vector<uint32_t>& ids, rh_node& nd) const {
    const double inf = std::numeric_limits<double>::infinity();
    for (int k = 0; k < 3; k++) { nd.lo[k] = inf; nd.hi[k] = -inf; }
    if (has_nan) {
      for (uint32_t t : ids)
        for (int c = 0; c < 3; c++) {
          const double* p = P(t, c);
          for (int k = 0; k < 3; k++) {
            nd.lo[k] = (nd.lo[k] <= p[k]) ? nd.lo[k] : p[k];
            nd.hi[k] = (nd.hi[k] <= p[k]) ? p[k] : nd.hi[k];
          }
        }
      return;
    }
    double lo0 = inf, lo1 = inf, lo2 = inf, hi0 = -inf, hi1 = -inf, hi2 = -inf;
    for (uint32_t t : ids) {
      const double* b = &tbox[6 * (size_t)t];
      lo0 = b[0] < lo0 ? b[0] : lo0; lo1 = b[1] < lo1 ? b[1] : lo1; lo2 = b[2] < lo2 ? b[2] : lo2;
      hi0 = b[3] > hi0 ? b[3] : hi0; hi1 = b[4] > hi1 ? b[4] : hi1; hi2 = b[5] > hi2 ? b[5] : hi2;
    }
    nd.lo[0] = lo0; nd.lo[1] = lo1; nd.lo[2] = lo2;
    nd.hi[0] = hi0; nd.hi[1] = hi1; nd.hi[2] = hi2;
  }

  // Appends subtree `in` to `out`; returns its root's index in `out` (RH_NO_NODE for `Empty`).
  static uint32_t append(Sub& out, const Sub& in, uint32_t depth_of_root) {
    if (in.nodes.empty()) return RH_NO_NODE;
    const uint32_t node_off = (uint32_t)out.nodes.size(), tri_off = (uint32_t)out.order.size(), leaf_off = out.n_leaves;
    for (rh_node nd : in.nodes) {
      if (nd.is_leaf) {
        nd.left += tri_off;
        nd.leaf_index += leaf_off;
      } else {
        if (nd.left != RH_NO_NODE) nd.left += node_off;
        if (nd.right != RH_NO_NODE) nd.right += node_off;
      }
      out.nodes.push_back(nd);
    }
    out.order.insert(out.order.end(), in.order.begin(), in.order.end());
    out.n_leaves += in.n_leaves;
    out.max_depth = std::max(out.max_depth, in.max_depth);
    (void)depth_of_root;
    return node_off;
  }

  // KDTree.hs:79-90.  Returns the node's index in `out` or RH_NO_NODE for `Empty`.  The two children of a large node
  // are built by two threads while `fork_levels` > 0 (into their own arrays, spliced in left-to-right order: the
  // result is the array the sequential build writes).
  uint32_t build(std::vector<uint32_t>& ids, int depth, int axis, Sub& out, int fork_levels) {
    if (ids.empty() || failed.load(std::memory_order_relaxed)) return RH_NO_NODE;
    if ((uint32_t)depth > out.max_depth) out.max_depth = depth;
    rh_node nd{};
    node_box(ids, nd);
    const uint32_t self = (uint32_t)out.nodes.size();
    out.nodes.push_back(nd);
    if (ids.size() < 20 || depth >= 100) {
      out.nodes[self].is_leaf = 1;
      out.nodes[self].left = (uint32_t)out.order.size();
      out.nodes[self].right = (uint32_t)ids.size();
      out.nodes[self].leaf_index = out.n_leaves++;
      out.order.insert(out.order.end(), ids.begin(), ids.end());
      return self;
    }
    const double split = 0.5 * (nd.hi[axis] + nd.lo[axis]);
    std::vector<uint32_t> l, r;
    l.reserve(ids.size() / 2 + 16);
    r.reserve(ids.size() / 2 + 16);
    for (uint32_t t : ids) {
      const double c = centroid[3 * (size_t)t + axis];
      if (c < split) l.push_back(t);
      if (split <= c) r.push_back(t);
    }
    std::vector<uint32_t>().swap(ids);
    const int next = (axis + 1) % 3;
    uint32_t li, ri;
    if (fork_levels > 0 && l.size() + r.size() >= (1u << 15)) {
      Sub L, R;
      L.max_depth = R.max_depth = 0;
      std::thread th([&]() {
        try {
          build(l, depth + 1, next, L, fork_levels - 1);
        } catch (...) {
          failed = true;  // (no exception may leave a thread; the caller reports out-of-memory)
        }
      });
      try {
        build(r, depth + 1, next, R, fork_levels - 1);
      } catch (...) {
        failed = true;
      }
      th.join();
      if (failed) return RH_NO_NODE;
      li = append(out, L, depth + 1);
      ri = append(out, R, depth + 1);
    } else {
      li = build(l, depth + 1, next, out, 0);
      ri = build(r, depth + 1, next, out, 0);
    }
    out.nodes[self].left = li;
    out.nodes[self].right = ri;
    return self;
  }

  void emit(uint32_t t, rh_tri& tr, rh_tri_shade& sh) const {
    tr = rh_tri{};
    sh = rh_tri_shade{};
    const double *p0 = P(t, 0), *p1 = P(t, 1), *p2 = P(t, 2);
    for (int k = 0; k < 3; k++) {
      tr.p0[k] = p0[k];
      tr.e1[k] = p1[k] - p0[k];  // Mesh.hs:70
      tr.e2[k] = p2[k] - p0[k];  // Mesh.hs:71
    }
    tr.tri_id = t;
    const uint32_t* ix = m.idx + 3 * (size_t)t;
    for (int k = 0; k < 3; k++) {
      sh.n0[k] = m.nrm[3 * (size_t)ix[0] + k];
      sh.n1[k] = m.nrm[3 * (size_t)ix[1] + k];
      sh.n2[k] = m.nrm[3 * (size_t)ix[2] + k];
    }
    for (int k = 0; k < 2; k++) {
      sh.uv0[k] = m.uv[2 * (size_t)ix[0] + k];
      sh.uv1[k] = m.uv[2 * (size_t)ix[1] + k];
      sh.uv2[k] = m.uv[2 * (size_t)ix[2] + k];
    }
  }
};

}  // namespace

struct rh_flat_scene {
  std::vector<rh_object> objects;
  std::vector<rh_material> materials;
  std::vector<rh_light> lights;
  std::vector<rh_texture> textures;
  std::vector<double> texels;
  AlignedArray<rh_node> nodes;
  AlignedArray<rh_tri> tris;
  AlignedArray<rh_tri_shade> shade;
  rh_scene_desc desc{};
};

extern "C" {

int rh_flatten(const rh_raw_scene* raw, rh_flat_scene** out) {
  if (!raw || !out) return rh::set_error(RH_ERR_ARG, "rh_flatten: null argument");
  try {
    auto F = std::make_unique<rh_flat_scene>();
    std::vector<rh_node> nodes;
    std::vector<rh_tri> tris;
    std::vector<rh_tri_shade> shade;
    F->materials.assign(raw->materials, raw->materials + raw->n_materials);
    F->lights.assign(raw->lights, raw->lights + raw->n_lights);
    F->textures.assign(raw->textures, raw->textures + raw->n_textures);
    F->texels.assign(raw->texels, raw->texels + 3 * raw->n_texels);
    const unsigned n_threads = std::max(1u, std::min(32u, std::thread::hardware_concurrency()));
    for (uint32_t i = 0; i < raw->n_objects; i++) {
      const rh_raw_object& ro = raw->objects[i];
      rh_object o{};
      o.kind = ro.kind;
      o.material = ro.material;
      memcpy(o.a, ro.a, sizeof o.a);
      memcpy(o.b, ro.b, sizeof o.b);
      memcpy(o.c, ro.c, sizeof o.c);
      o.root = RH_NO_NODE;
      if (ro.material < 0 || (uint32_t)ro.material >= raw->n_materials)
        return rh::set_error(RH_ERR_ARG, "rh_flatten: object material index out of range");
      if (ro.kind == RH_OBJ_MESH) {
        uint32_t nt = ro.n_indices / 3;  // Mesh.hs:105-109
        for (uint32_t k = 0; k < ro.n_indices; k++)
          if (ro.indices[k] >= ro.n_verts) return rh::set_error(RH_ERR_ARG, "rh_flatten: vertex index out of range");
        Builder b;
        b.m = MeshView{ro.positions, ro.normals, ro.uvs, ro.indices};
        b.prepare(nt, n_threads);
        std::vector<uint32_t> ids(nt);
        for (uint32_t t = 0; t < nt; t++) ids[t] = t;
        Sub sub;
        int fork_levels = 0;
        while ((1u << fork_levels) < n_threads) fork_levels++;
        const uint32_t root = b.build(ids, 0, 0, sub, nt >= (1u << 16) ? fork_levels + 1 : 0);
        if (b.failed) throw std::bad_alloc();
        // splice the mesh's arrays into the scene's: node indices and triangle slots move by the sizes so far
        const uint32_t node_off = (uint32_t)nodes.size(), tri_off = (uint32_t)tris.size();
        if ((uint64_t)nodes.size() + sub.nodes.size() >= 0xFFFFFFF0ull || (uint64_t)tris.size() + sub.order.size() >= 0xFFFFFFF0ull)
          return rh::set_error(RH_ERR_ARG, "rh_flatten: too many nodes or triangles");
        for (rh_node nd : sub.nodes) {
          if (nd.is_leaf) nd.left += tri_off;
          else {
            if (nd.left != RH_NO_NODE) nd.left += node_off;
            if (nd.right != RH_NO_NODE) nd.right += node_off;
          }
          nodes.push_back(nd);
        }
        tris.resize((size_t)tri_off + sub.order.size());
        shade.resize((size_t)tri_off + sub.order.size());
        parallel_ranges(sub.order.size(), n_threads, [&](size_t b0, size_t e0) {
          for (size_t k = b0; k < e0; k++) b.emit(sub.order[k], tris[tri_off + k], shade[tri_off + k]);
        });
        o.root = root == RH_NO_NODE ? RH_NO_NODE : root + node_off;
        o.n_leaves = sub.n_leaves;
        o.depth = sub.max_depth;
      } else if (ro.kind != RH_OBJ_PLANE && ro.kind != RH_OBJ_SPHERE) {
        return rh::set_error(RH_ERR_ARG, "rh_flatten: unknown object kind");
      }
      F->objects.push_back(o);
    }
    F->nodes.assign(nodes);
    F->tris.assign(tris);
    F->shade.assign(shade);
    rh_scene_desc& d = F->desc;
    d.n_objects = (uint32_t)F->objects.size();
    d.n_materials = (uint32_t)F->materials.size();
    d.n_lights = (uint32_t)F->lights.size();
    d.n_textures = (uint32_t)F->textures.size();
    d.n_nodes = (uint32_t)F->nodes.n;
    d.n_tris = (uint32_t)F->tris.n;
    d.objects = F->objects.data();
    d.materials = F->materials.data();
    d.lights = F->lights.data();
    d.textures = F->textures.data();
    d.texels = F->texels.data();
    d.n_texels = F->texels.size() / 3;
    d.nodes = F->nodes.p;
    d.tris = F->tris.p;
    d.tri_shade = F->shade.p;
    *out = F.release();
  } catch (const std::bad_alloc&) {
    return rh::set_error(RH_ERR_OOM, "rh_flatten: out of host memory");
  }
  return RH_OK;
}

const rh_scene_desc* rh_flat_desc(const rh_flat_scene* f) { return f ? &f->desc : nullptr; }
void rh_flat_destroy(rh_flat_scene* f) { delete f; }

}  // extern "C"
