// mas_fused.cuh -- host interface of the fused contraction + MAS kernels (mas_fused.cu), used by mas_api.cu
#pragma once

#include "mas_common.cuh"

namespace mas {

// fused_launch() could not run in this context (cooperative grid does not fit, no tensor maps, no kernel for the
// shape): not an error, the caller runs contraction and DP as separate launches
constexpr int kFusedFallback = -1;

bool fused_supported(int B, int D, int T, int S);
bool fused_noise_supported(int B, int D, int T, int S);
size_t fused_flags_bytes(int B, int T);
size_t fused_plane_bytes(int B, int T, int S);
// noise == nullptr: contraction and DP run concurrently; otherwise contraction + statistics, grid barrier, then
// noise appliers next to the DP (stats: two doubles of device scratch).  `plane`: private cost plane of
// fused_plane_bytes().  path_out may be nullptr (compact outputs only).
int fused_launch(const float *z_p, const float *m_p, const float *logs_p, const int32_t *t_ys, const int32_t *t_xs,
                 const float *noise, float noise_scale, double *stats, float *plane, void *path_out, int path_dtype,
                 int32_t *dur_out, int32_t *idx_out, int32_t *status_out, void *cost_ws, size_t cost_ws_bytes, void *dp_ws,
                 size_t dp_ws_bytes, uint32_t *flags, int B, int D, int T, int S, cudaStream_t stream);

}  // namespace mas
