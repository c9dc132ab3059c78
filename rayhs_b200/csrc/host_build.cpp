// host_build.cpp — host-side tree build + flattening (C++; GHC is not in this image).
// Restates `buildKDTree` / `buildNode` (KDTree.hs:68-90) on index arrays and lays the
// result out as the flat arrays of include/rayhs_b200.h (rh_node / rh_tri / rh_tri_shade).
// A Haskell host would produce the same arrays from its own `KDTree` value
// (INTEGRATION.md); this file is what the GHC-less harness and the synthetic scene use.
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <memory>
#include <new>
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

struct Builder {
  MeshView m;
  std::vector<double> centroid;  // [tri][3]
  std::vector<rh_node>& nodes;
  std::vector<rh_tri>& tris;
  std::vector<rh_tri_shade>& shade;
  uint32_t n_leaves = 0;
  uint32_t max_depth = 0;

  const double* P(uint32_t tri, int corner) const { return m.pos + 3 * (size_t)m.idx[3 * (size_t)tri + corner]; }

  // KDTree.hs:79-90.  Returns the node index or RH_NO_NODE for `Empty`.
  uint32_t build(std::vector<uint32_t>& ids, int depth, int axis) {
    if (ids.empty()) return RH_NO_NODE;
    if ((uint32_t)depth > max_depth) max_depth = depth;
    rh_node nd{};
    const double inf = std::numeric_limits<double>::infinity();
    for (int k = 0; k < 3; k++) { nd.lo[k] = inf; nd.hi[k] = -inf; }
    for (uint32_t t : ids)  // buildBoundingBox, KDTree.hs:22-29
      for (int c = 0; c < 3; c++) {
        const double* p = P(t, c);
        for (int k = 0; k < 3; k++) {
          nd.lo[k] = (nd.lo[k] <= p[k]) ? nd.lo[k] : p[k];
          nd.hi[k] = (nd.hi[k] <= p[k]) ? p[k] : nd.hi[k];
        }
      }
    uint32_t self = (uint32_t)nodes.size();
    nodes.push_back(nd);
    if (ids.size() < 20 || depth >= 100) {
      nodes[self].is_leaf = 1;
      nodes[self].left = (uint32_t)tris.size();
      nodes[self].right = (uint32_t)ids.size();
      nodes[self].leaf_index = n_leaves++;
      for (uint32_t t : ids) emit(t);
      return self;
    }
    double split = 0.5 * (nd.hi[axis] + nd.lo[axis]);
    std::vector<uint32_t> l, r;
    for (uint32_t t : ids) {
      double c = centroid[3 * (size_t)t + axis];
      if (c < split) l.push_back(t);
      if (split <= c) r.push_back(t);
    }
    std::vector<uint32_t>().swap(ids);
    int next = (axis + 1) % 3;
    uint32_t li = build(l, depth + 1, next);
    uint32_t ri = build(r, depth + 1, next);
    nodes[self].left = li;
    nodes[self].right = ri;
    return self;
  }

  void emit(uint32_t t) {
    rh_tri tr{};
    rh_tri_shade sh{};
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
    tris.push_back(tr);
    shade.push_back(sh);
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
        Builder b{{ro.positions, ro.normals, ro.uvs, ro.indices}, {}, nodes, tris, shade};
        b.centroid.resize((size_t)nt * 3);
        for (uint32_t t = 0; t < nt; t++)  // baryCenter, KDTree.hs:71-74: mul (1/3) (a + b + c)
          for (int k = 0; k < 3; k++) b.centroid[3 * (size_t)t + k] = (1.0 / 3) * ((b.P(t, 0)[k] + b.P(t, 1)[k]) + b.P(t, 2)[k]);
        std::vector<uint32_t> ids(nt);
        for (uint32_t t = 0; t < nt; t++) ids[t] = t;
        o.root = b.build(ids, 0, 0);
        o.n_leaves = b.n_leaves;
        o.depth = b.max_depth;
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
