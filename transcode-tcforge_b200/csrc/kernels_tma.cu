// kernels_tma.cu -- tier 3 (TMA / bulk-copy staged) conversions.
#include "acgpu_internal.h"
namespace acgpu {
bool convert_tma(const ConvertArgs &) { return false; }
}  // namespace acgpu
