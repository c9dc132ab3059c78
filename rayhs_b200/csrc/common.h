// common.h — shared by the host-side translation units of librayhs_b200.
#pragma once
#include <string>

#include "../../include/rayhs_b200.h"

namespace rh {
// Records the message for rh_last_error() (thread-local) and returns `code`.
int set_error(int code, const std::string& msg);
}  // namespace rh
