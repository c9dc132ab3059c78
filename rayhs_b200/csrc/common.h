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
}  // namespace rh
