// mas_tmap.cuh -- host-side construction of TMA tensor maps (CUtensorMap) without linking
// libcuda: cuTensorMapEncodeTiled is fetched through the runtime's driver entry point.
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace mas {

// 3-D fp32 tensor map.  dims / box innermost first; strides in bytes for dims 1 and 2
// (multiples of 16).  swizzle128: box[0] must be 32 floats (128 bytes).
// Returns false when the driver rejects the description (caller falls back to plain loads/stores).
bool make_tmap_f32_3d(CUtensorMap *out, const void *base, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t stride1_bytes,
                      uint64_t stride2_bytes, uint32_t b0, uint32_t b1, uint32_t b2, bool swizzle128);

}  // namespace mas
