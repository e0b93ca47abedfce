// mas_dp_launch.cuh -- the standalone MAS kernel and its launch templates, shared by the three translation units
// that instantiate it (mas_dp.cu: two DP warps per team, chunk heights 32 / 16; mas_dp_wide.cu: four warps;
// mas_dp_tall.cu: two warps, tall chunks): ~70 kernel variants compile in parallel instead of as one 2m40 file.
#pragma once

#include "mas_dp.cuh"

namespace mas {

template <int C, int R, int W, bool kVec, bool kNoise, bool kVK>
__global__ void __launch_bounds__(dp_threads(W, kVK), 1) mas_dp_kernel(const DpParams p)
{
    extern __shared__ __align__(128) unsigned char smem[];
    const int b = p.order ? p.order[blockIdx.x] : (int)blockIdx.x;
    uint32_t g_base = 0;
    dp_role_init(p, smem, threadIdx.x, kDpBar);
    dp_role<C, R, W, kVec, kNoise, kVK>(p, smem, b, blockIdx.x, g_base, threadIdx.x, kDpBar);
}

template <int C, int R, int W, bool kVec, bool kNoise, bool kVK = false>
static int launch_dp_cv(const DpPlan &pl, cudaStream_t stream)
{
    static thread_local int configured_dev = -1;
    int dev = 0;
    MAS_CUDA_TRY(cudaGetDevice(&dev));
    if (dev != configured_dev) {
        MAS_CUDA_TRY(cudaFuncSetAttribute(mas_dp_kernel<C, R, W, kVec, kNoise, kVK>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)kSmemBudget));
        configured_dev = dev;
    }
    mas_dp_kernel<C, R, W, kVec, kNoise, kVK><<<pl.p.B, dp_threads(W, kVK), pl.smem_bytes, stream>>>(pl.p);
    note_launch();
    MAS_CUDA_TRY(cudaGetLastError());
    return MAS_OK;
}

template <int C, int R, int W>
static int launch_dp_c(const DpPlan &pl, cudaStream_t stream)
{
    // vector cost loads need every tile row 16-byte aligned in shared memory
    const bool vec = (pl.p.ld % 4 == 0) && ((reinterpret_cast<uintptr_t>(pl.p.neg_cent) & 15) == 0);
    if (pl.p.noise) {
        // noise applied while the cost streams in: vector path only (dp_noise_supported)
        if (!vec || (reinterpret_cast<uintptr_t>(pl.p.noise) & 15)) return MAS_ERR_UNSUPPORTED_SHAPE;
        return launch_dp_cv<C, R, W, true, true>(pl, stream);
    }
    if (pl.p.vk) {
        if (!vec) return MAS_ERR_UNSUPPORTED_SHAPE;
        if constexpr (((W == 2 && C <= 4) || (W == 4 && C == 2)) && R == 32)
            return launch_dp_cv<C, R, W, true, false, true>(pl, stream);
        else
            return MAS_ERR_UNSUPPORTED_SHAPE;
    }
    return vec ? launch_dp_cv<C, R, W, true, false>(pl, stream) : launch_dp_cv<C, R, W, false, false>(pl, stream);
}

// returns kDpNoCase when (C, R, W) is not one of this translation unit's cases
constexpr int kDpNoCase = -2;
int dp_dispatch_narrow(const DpPlan &pl, int C, cudaStream_t stream);   // mas_dp.cu
int dp_dispatch_wide(const DpPlan &pl, int C, cudaStream_t stream);     // mas_dp_wide.cu
int dp_dispatch_tall(const DpPlan &pl, int C, cudaStream_t stream);     // mas_dp_tall.cu

}  // namespace mas
