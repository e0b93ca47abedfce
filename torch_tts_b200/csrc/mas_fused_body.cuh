// mas_fused_body.cuh -- the no-noise fused kernel (contraction role, then DP or zero-fill role) as a template, shared
// by the translation units that instantiate it: mas_fused.cu (S <= 256, the value / origin warp split) and
// mas_fused_wide{2,4}.cu (256 < S <= 1024: two / four single-role DP warps, several column blocks per mel tile).
#pragma once

#include "mas_cost_tc.cuh"
#include "mas_dp.cuh"

namespace mas {

struct FusedParams {
    TcParams tc;
    DpParams dp;
    int n_dp;              // the first n_dp CTAs turn into DP CTAs after their share of the contraction
    uint32_t *grid_bar;    // noise kernel: grid barrier counter (cleared with the flags)
    int feed_pairs;        // noise kernel: > 0 = that many CTA pairs run {DP CTA, noise feeder CTA} after the barrier
    const float *noise;    // noise kernel with feeders: the draw [B][T][S] (16-byte rows), its scale, the statistics
    float noise_scale;
    const double *stats;
};

constexpr int kNoiseHelpWarps = 8;   // helper warps of a DP CTA in the noise kernel (dp_role, kHelp)

template <int C, int R, int W, bool kVK, int kHelp = 0>
__device__ __forceinline__ void fused_dp_ctas(const FusedParams &fp, unsigned char *smem)
{
    // DP CTA j aligns utterances j, j + n_dp, ... on its first dp_threads(W) (+ helper) threads
    if ((int)threadIdx.x >= dp_threads(W, kVK) + 32 * kHelp) return;
    const int j = (int)blockIdx.x;
    uint32_t g_base = 0;
    dp_role_init(fp.dp, smem, threadIdx.x, kDpBar);
    for (int b = j; b < fp.dp.B; b += fp.n_dp)
        dp_role<C, R, W, true, false, kVK, kHelp>(fp.dp, smem, b, j, g_base, threadIdx.x, kDpBar);
}

template <int C, int R, int W, bool kPair, bool kVK>
__device__ __forceinline__ void fused_body(const FusedParams &fp, const CUtensorMap *tm_z, const CUtensorMap *tm_out,
                                           unsigned char *smem)
{
    // every CTA starts in the contraction role; the first n_dp CTAs ("hybrid") leave it after seq_k rounds of
    // units and become the DP CTAs, the others finish the remaining units (unit_index() in cost_tc_role)
    if (MAS_TR(fp.tc) && threadIdx.x == 0) fp.tc.trace[49152 + blockIdx.x] = globaltimer_ns();  // CTA entry
    if (kPair)
        cost_tc_role<false, true>(fp.tc, tm_z, tm_out, smem, blockIdx.x >> 1, gridDim.x >> 1);
    else
        cost_tc_role<false, false>(fp.tc, tm_z, tm_out, smem, blockIdx.x, gridDim.x);
    if (MAS_TR(fp.tc) && threadIdx.x == 0) fp.tc.trace[49152 + 256 + blockIdx.x] = globaltimer_ns();  // contraction role left
    if ((int)blockIdx.x >= fp.n_dp) {
        // out of tiles: zero-fill the dense path planes while the DP CTAs are still busy
        if (fp.dp.zero_flags) zero_fill_role(fp.dp, smem);
        if (MAS_TR(fp.tc) && threadIdx.x == 0) fp.tc.trace[49152 + 512 + blockIdx.x] = globaltimer_ns();  // zero-fill done
        return;
    }
    fused_dp_ctas<C, R, W, kVK>(fp, smem);
}

// contraction CTAs in pairs (clusters of 2, n_gemm even); the DP CTAs ignore their cluster
template <int C, int R, int W, bool kVK>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kTcThreads, 1)
    mas_fused_pair_kernel(const __grid_constant__ FusedParams fp, const __grid_constant__ CUtensorMap tm_z,
                          const __grid_constant__ CUtensorMap tm_out)
{
    extern __shared__ __align__(128) unsigned char smem[];
    fused_body<C, R, W, true, kVK>(fp, &tm_z, &tm_out, smem);
}

// kernels for 256 < S <= 1024 (mas_fused_wide2.cu: two DP warps, chunk heights 16 / 32; mas_fused_wide4.cu: four DP
// warps, chunk heights 8 / 16); nullptr when (C, R) is not instantiated
const void *fused_pair_kernel_wide2(int C, int R);
const void *fused_pair_kernel_wide4(int C, int R);

}  // namespace mas
