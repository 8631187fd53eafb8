// GPU index construction kernels (K3/K4) — see DESIGN.md "Build".
#pragma once
#include "common.cuh"

namespace hb {
}  // namespace hb
