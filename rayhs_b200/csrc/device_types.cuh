// device_types.cuh — HBM layouts and kernel parameter blocks of the ray-casting path.
// See DESIGN.md "Data layout in HBM".
#pragma once
#include <cstdint>

#include "../../include/rayhs_b200.h"

namespace rhd {

constexpr uint32_t kEmpty = 0xFFFFFFFFu;    // KDTree.hs:61 `Empty`
constexpr uint32_t kLeafBit = 0x80000000u;  // child reference is a leaf: low 30 bits = triangle (or sphere) count
constexpr uint32_t kSphereLeafBit = 0x40000000u;  // ... a leaf of the sphere tree: `first` indexes sphere_refs
constexpr uint32_t kCountMask = 0x3FFFFFFFu;
constexpr int kMaxDepth = 30;               // ray-tree depth limit accepted by rh_render (reference scenes use 3)
constexpr int kMaxPasses = 2 * kMaxDepth + 2;  // a Transparent hit inserts one probe pass per level (RayHs.hs:136-143)
constexpr int kStack = 112;                 // tree depth is <= 100 by construction (KDTree.hs:76-77, 82) + leaf refinement levels
#ifndef RH_SMEM_NODES
#define RH_SMEM_NODES 448
#endif
constexpr int kSmemNodes = RH_SMEM_NODES;   // top wide (fp32) nodes staged in shared memory (64 B each)
constexpr int kSmemObjects = 64;            // object / material tables staged when the scene has at most this many
constexpr int kSmemLights = 16;
constexpr int kBlock = 128;                // resolve kernel
constexpr int kMaxPeers = 16;
#ifndef RH_TRACE_BLOCK
#define RH_TRACE_BLOCK 1024
#define RH_TRACE_MINB 1
#define RH_SHADOW_BLOCK 768
#define RH_SHADOW_MINB 1
#endif
#ifndef RH_SHADOW_PAIRS
#define RH_SHADOW_PAIRS 9  // (hit, light) pairs per lane whose terms the pooled fast shadow kernel keeps in shared memory: 3 hits per lane per batch with 3 lights (measured: 3 pairs 37.3 ms, 6 33.9, 9 33.7, 12 34.7 ms of shadow work per bench frame)
#endif
#ifndef RH_SHADOW_T
#define RH_SHADOW_T 4   // shaded hits per lane per warp batch in the pooled shadow kernel
#endif
#ifndef RH_SHADOW_POOL
#define RH_SHADOW_POOL 1
#endif
#ifndef RH_WALK_BLOCK
#define RH_WALK_BLOCK 640  // threads per block of the shadow walk kernel (one block per SM): 640 -> up to 102 registers (measured best of 512, 640, 768)
#endif
#ifndef RH_WALK_UNROLL
#define RH_WALK_UNROLL 2
#endif
#ifndef RH_ISECT_BLOCK
#define RH_ISECT_BLOCK 1024  // threads per block of intersect_kernel (one block per SM; measured on the synthetic scene: 512 84 ms, 640 74 ms, 768 71 ms, 1024 69 ms)
#endif
#ifndef RH_REFILL_MIN
#define RH_REFILL_MIN 16
#endif
#ifndef RH_SHADOW_SPLIT
#define RH_SHADOW_SPLIT 1  // 0: one pooled shadow kernel per pass instead of classify -> walk -> fold
#endif
#ifndef RH_SHADOW_FAST
#define RH_SHADOW_FAST 1  // 0: always use the general pooled kernel (A/B and validation builds)
#endif
constexpr int kTraceBlock = RH_TRACE_BLOCK, kTraceMinBlocks = RH_TRACE_MINB;      // one block per SM, tables staged once per SM.  trace: 1024 threads at 64 registers (bench frame: equal to 768 at 80; incoherent synthetic scene: 57 vs 64 ms); shadow: 768 (measured best)
constexpr int kShadowBlock = RH_SHADOW_BLOCK, kShadowMinBlocks = RH_SHADOW_MINB;

// 128-byte "wide" node: one record per inner tree node holding BOTH child boxes, so a
// visit is one 128-byte line (4 sectors) and every box is still tested exactly once,
// as in KDTree.hs:96-107 (box test at every Node and Leaf).  A synthetic super-root per
// mesh carries the root's own box in slot 0.  Wide nodes are stored in level order over
// all meshes, so the first kSmemNodes records are the top levels of every tree.
struct __align__(16) WideNode {
  double box[12];     // child0 lo.xyz hi.xyz, child1 lo.xyz hi.xyz
  uint32_t child[2];  // kEmpty | kLeafBit|count | wide-node index
  uint32_t first[2];  // leaf child: first triangle slot
  uint32_t refine;    // bit c: child c's box is a culling refinement inside a reference leaf (not a reference box)
  uint32_t pad_[3];
};
#ifndef RH_STREAM_CHUNK_MI
#define RH_STREAM_CHUNK_MI 16  // chunk size (Mi samples) when the sample offsets stream in from the host
#define RH_STREAM_FIRST_MI 4   // ... and of the first chunk, whose upload nothing overlaps
#endif
#ifndef RH_LANES
#define RH_LANES 2  // chunks in flight (streams with their own queues); 1 = strictly one chunk after the other
#endif
#ifndef RH_SUBLEAF
#define RH_SUBLEAF 4  // triangles per leaf of the float path's cull tree (measured: 3-4 best of 2, 3, 4, 6, 8)
#endif
constexpr uint32_t kSubLeaf = RH_SUBLEAF;
static_assert(sizeof(WideNode) == 128, "WideNode must be 128 bytes");

// 64-byte culling copy of a WideNode: the same two child boxes rounded OUTWARD to float.  The fp32
// slab test on it is conservative (never rejects a box the double test of KDTree.hs:39-56 accepts,
// see kernels.cu), so it only decides which triangles get the exact double test.
struct __align__(16) WideNode32 {
  float box[12];
  uint32_t child[2];
  uint32_t first[2];
};
static_assert(sizeof(WideNode32) == 64, "WideNode32 must be 64 bytes");

// Object table entry (device copy of rh_object with the wide super-root index).
struct __align__(16) DObject {
  double a[3], b[3], c[3];
  int32_t kind;
  int32_t material;
  uint32_t root;  // wide-node index of the mesh's super-root or kEmpty
  uint32_t is_emitter;
};
static_assert(sizeof(DObject) == 96, "DObject must be 96 bytes");

// Occluder tables of the fast shadow kernel (shadowIntersection, RayHs.hs:74-82): the non-emitter planes and the
// non-emitter spheres that are not in the sphere tree, as compact records staged in shared memory, and the
// super-roots of the non-emitter, non-empty meshes.
struct OccPlane { double p[3], n[3]; };   // Geometry.hs:62
struct OccSphere { double c[3], r; };     // Geometry.hs:63
constexpr int kOccPlanes = 16, kOccSpheres = 16, kOccMeshes = 8, kFastLights = 12;

struct SceneView {
  const WideNode* wide;      // exact double boxes: rays with a zero direction component, RH_FLAG_EXACT_BOXES
  const WideNode32* wide32;  // conservative float boxes: everything else
  const rh_tri* tris;
  const rh_tri_shade* shade;
  const DObject* objects;
  const rh_material* materials;
  const rh_light* lights;
  const rh_texture* textures;
  const double* texels;
  const uint32_t* lin_objs;     // objects scanned linearly in scene order (RayHs.hs:64-71): all of them, or all but the spheres
  const uint32_t* sphere_refs;  // sphere tree leaves: object indices
  const uint32_t* exact_index;  // exact walk: slot of the reference tree's leaves -> slot in tris / shade (null: identical)
  uint32_t n_lin;
  uint32_t sphere_root;         // super-root of the sphere tree or kEmpty (spheres are in lin_objs then)
  const OccPlane* occ_planes;
  const OccSphere* occ_spheres;
  const uint32_t* occ_meshes;
  uint32_t n_occ_planes, n_occ_spheres, n_occ_meshes;
  uint32_t shadow_fast;         // the occluder tables and the lights fit the fast shadow kernel's shared-memory tables
  uint32_t n_wide, n_tris, n_objects, n_materials, n_lights, n_textures;
  uint32_t n_smem_nodes;    // min(n_wide, kSmemNodes)
  uint32_t tables_in_smem;  // objects/materials/lights fit the staged tables
  float abs_max;            // largest |coordinate - center| of any tree box (pads the fp32 slab test)
  uint32_t light_map_res;   // cells per edge of a cube-map face
  double center[3];         // the float boxes of wide32 are stored relative to this point
  // Per (point light, occluder mesh) cube maps of the nearest possible occluder distance (light_maps.cpp): map k is
  // light_maps[k * 6 * res * res ..], face-major; light_map_index[light * kOccMeshes + mesh] = k or kEmpty.
  // Null when the scene has none (no point light, no mesh, maps useless or switched off).
  const float* light_maps;
  const uint32_t* light_map_index;
  // Lit triangles (light_maps.cpp), per triangle slot: bit l = nothing of the triangle's own mesh can shadow a point
  // of it from light l (l < 12, point or directional); bits 12-15 = the mesh's index in occ_meshes.  Null when the
  // scene has none.
  const uint16_t* lit_flags;
};

// Camera with everything `rayFromPixel` recomputes per pixel hoisted to the host
// (Projection.hs:22-46, Mat.hs:89-93): the values are identical, they are just computed once.
struct CameraParams {
  double m[9];  // lookAt matrix rows (Mat.hs:83-93)
  double pos[3];
  double w, h, half_w, half_h, apw, aph, f;
  int32_t projection;
  int32_t pad_;
};

// Per-chunk control block in device memory (zeroed before each chunk).
struct ChunkCtl {
  uint32_t ray_count[kMaxPasses + 2];     // entries in the ray queue consumed by pass k
  uint32_t shadow_count[kMaxPasses + 2];  // shadow tasks produced by pass k
  uint32_t trace_cursor[kMaxPasses + 2];  // persistent-warp work cursors
  uint32_t shadow_cursor[kMaxPasses + 2];
  // split shadow pipeline: low word = deferred hits, high word = (hit, light) pairs queued for a tree walk
  unsigned long long deferred_walk_count[kMaxPasses + 2];
  uint32_t walk_cursor[kMaxPasses + 2];
  uint32_t isect_cursor[kMaxPasses + 2];  // split trace schedule: intersect_kernel's work cursor
  uint32_t overflow;
  uint32_t pad_[3];
};

// Frame-level counters (zeroed per rh_render).
struct KernelCounters {  // RH_FLAG_COUNT only
  unsigned long long box_tests, tri_tests, prim_tests, node_visits, shade_fetches, texel_fetches;
};
struct FrameCounters {
  unsigned long long rays_reflect, rays_probe, rays_exit, negative_channels;
  unsigned long long shadow_culled;  // (hit, light) pairs with l.n <= 0: Lambert term is exactly 0, query skipped
  unsigned long long exact_walks;    // RH_FLAG_COUNT: shadow rays that took the exact (double box) walk
  unsigned long long max_walk_nodes; // RH_FLAG_COUNT: most node records one shadow ray visited
  unsigned long long exact_closest;  // RH_FLAG_COUNT: closest-hit rays that took the exact walk
  unsigned long long max_closest_nodes;
  KernelCounters k[2];  // 0 = trace_kernel, 1 = shadow_kernel
};

// SoA-of-16-byte planes so that a warp's compacted pushes are fully coalesced.
// Ray queue: 4 planes (ox,oy) (oz,dx) (dy,dz) (weight, bits).
// bits = sample | depth << 32 | kind << 40 | material << 48.
struct RayQueue {
  double2* plane;  // plane k at plane + k*capacity
  uint32_t capacity;
};
// Shadow queue: 5 planes (px,py) (pz,nx) (ny,nz) (cr,cg) (cb,w) + sample ids (bit 31 = ambient term) + lit flags.
struct ShadowQueue {
  double2* plane;
  uint32_t* sample;
  uint32_t* lit;  // lit_flags of the hit triangle, 0 for other hits
  uint32_t capacity;
};

enum { kRayNormal = 0, kRayProbe = 1 };
enum { kOffIndexLocal = 0, kOffIndexGlobal = 1 };

struct ChunkParams {
  uint32_t first_row;     // first local (shard-compact) row of the chunk
  uint32_t n_rows;        // rows in the chunk
  uint32_t n_samples;     // n_rows * width * spp
  uint32_t spp;
  uint32_t width, height;
  uint32_t shard_index, shard_count, band_height;
  int32_t max_depth;
  int32_t offset_mode;    // RH_OFFSETS_*
  int32_t offset_index;   // kOffIndex*: per-pixel offsets indexed by chunk-local or full-frame pixel
  int32_t offset_tile;
  int32_t pass;           // wavefront pass (0 = primary)
  int32_t exact_boxes;    // RH_FLAG_EXACT_BOXES: double slab test for every ray (validation)
  int32_t no_light_maps;  // RH_FLAG_NO_LIGHT_MAPS: shadow rays ignore the lights' cube maps (validation, A/B)
  const void* offsets;    // device
  unsigned long long offset_seed;  // RH_OFFSETS_SPLITMIX64
  double* accum;          // 3 planes of accum_stride (r,g,b), chunk-local sample order
  uint32_t accum_stride;
  uint32_t pad_;
  int2* hit_ids;          // shard-compact [pixel][spp] (object, tri) or null
  uint8_t* rgb;           // shard-compact framebuffer (or null with peer frames)
  uint8_t* peer[kMaxPeers];  // RH_FLAG_PEER_FRAMES: full frames of all shards
  uint32_t n_peers;
  uint32_t pad4_;
  ChunkCtl* ctl;
  FrameCounters* counters;
  RayQueue q_in, q_out;
  ShadowQueue q_shadow;
  // split shadow pipeline (classify -> walk -> fold)
  uint2* walk_q;          // (shadow-queue item, light) pairs whose ray enters a tree's root box
  uint32_t* deferred_q;   // shadow-queue items with at least one queued pair
  uint8_t* pair_flags;    // [item * n_lights + light]: 1 = the pair adds nothing (l.n <= 0 or occluded)
  uint32_t walk_capacity;
  uint32_t pad3_;
  double4* hits;          // split trace schedule: (t, u, v, slot | obj << 32) per work item of the pass
};

// Launchers (kernels.cu).  `count` selects the instrumented instantiation (box/tri counters).
// split: intersect_kernel (closest hits with per-lane refill -> P.hits) then trace_kernel<.., true> (shade from the records)
void launch_trace(const SceneView& S, const CameraParams& cam, const ChunkParams& P, bool count, bool split, int grid, void* stream);
// split: classify -> walk -> fold (3 launches) instead of one pooled kernel; only when shadow_split_possible(S)
void launch_shadow(const SceneView& S, const ChunkParams& P, bool count, bool split, int grid, void* stream);
bool shadow_split_possible(const SceneView& S);
void launch_resolve(const ChunkParams& P, void* stream);
void launch_deinterleave(const uint8_t* gathered, uint8_t* out, int width, int height, int shard_count, int band_height,
                         void* stream);
int configure_kernels();  // opt in to > 48 KB dynamic shared memory; returns a cudaError_t
int trace_blocks_per_sm(bool count);
int shadow_blocks_per_sm(bool count);
// micro-benchmarks
void launch_gather_bench(const double2* buf, uint64_t n_records, uint32_t loads_per_thread, double2* sink, int grid, int block,
                         void* stream);
void launch_dfma_bench(double* sink, int iters, int grid, int block, void* stream);

}  // namespace rhd
