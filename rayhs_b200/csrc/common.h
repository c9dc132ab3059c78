// common.h — shared by the host-side translation units of librayhs_b200.
#pragma once
#include <cstddef>
#include <string>

#include "../../include/rayhs_b200.h"

namespace rh {
// Records the message for rh_last_error() (thread-local) and returns `code`.
int set_error(int code, const std::string& msg);
// light_maps.cpp: cube map (out[6 * R * R]) of a lower bound of the distance from the point light at L to the
// triangles tris[slots[..]] covering each direction; false when the map would be useless or unsafe.
bool build_light_map(const double L[3], const rh_tri* tris, const unsigned* slots, size_t n, int R, float* out, double min_empty,
                     double* empty_fraction);

// light_maps.cpp, "lit triangles": can any other triangle of a mesh shadow a point of triangle T0 from the point light
// at L?  The query region K is the hull of T0 (widened a little) and L, minus the slab within 1e-8 of T0's plane: a
// shadow ray from a point p of T0 starts 1e-6 along the unit light direction (rayEps, Geometry.hs:36) and only counts
// hits at t >= 1e-6 (Mesh.hs:76), i.e. at least 2e-6 |cos| above the plane, and the query is only made for |cos| >= 0.01.
struct LitQuery;  // light_geom.h: K = { x : n[i].x + d[i] >= 0 for all i } with its bounding box
bool lit_query_make(const rh_tri& t0, const double L[3], bool directional, LitQuery* q);  // false: grazing light or degenerate triangle
// (directional: L is the light's vector, K the prism over T0 along it)
bool lit_query_box_outside(const LitQuery& q, const double* lo, const double* hi);  // the box cannot meet K
bool lit_query_tri_meets(const LitQuery& q, const rh_tri& t);                  // the triangle meets K
}  // namespace rh
