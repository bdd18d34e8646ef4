// kernels_fast.cu -- tier 2 (vectorised) conversions.  Filled in below the generic tier.
#include "acgpu_internal.h"
namespace acgpu {
bool convert_fast(const ConvertArgs &) { return false; }
}  // namespace acgpu
