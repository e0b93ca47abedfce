// mas_fused.cu -- neg_cent contraction and MAS in ONE kernel (no-noise alignment path,
// reference vits2/models.py:1224-1256 with mas_noise_scale None).
//
// One CTA per SM, roles by block index: the first n_gemm CTAs run the tcgen05 contraction
// (mas_cost_tc.cuh) over the tile list -- utterance groups of n_dp, mel-tile-major inside a
// group -- and publish every finished 128-row cost tile with a release store to a flag; the
// last n_dp CTAs run the forward DP + backtrack (mas_dp.cuh), one utterance at a time, and
// acquire the flag of a tile before the TMA engine streams its rows out of L2.  The cost
// plane round-trips through L2 only; the DP trails the contraction by a few tiles instead of
// waiting for the whole batch.  Producers never wait on consumers, and the launch is
// cooperative so all CTAs are co-resident.
#include "mas_cost_tc.cuh"
#include "mas_dp.cuh"

namespace mas {

struct FusedParams {
    TcParams tc;
    DpParams dp;
    int n_gemm, n_dp;
};

template <int C, bool kPair>
__device__ __forceinline__ void fused_body(const FusedParams &fp, const CUtensorMap *tm_z, const CUtensorMap *tm_out,
                                           unsigned char *smem)
{
    if ((int)blockIdx.x < fp.n_gemm) {
        if (kPair)
            cost_tc_role<false, true>(fp.tc, tm_z, tm_out, smem, blockIdx.x >> 1, fp.n_gemm >> 1);
        else
            cost_tc_role<false, false>(fp.tc, tm_z, tm_out, smem, blockIdx.x, fp.n_gemm);
    } else {
        if (threadIdx.x >= kThreads) return;
        const int j = (int)blockIdx.x - fp.n_gemm;
        uint32_t g_base = 0;
        dp_role_init(fp.dp, smem);
        for (int b = j; b < fp.dp.B; b += fp.n_dp) dp_role<C, true>(fp.dp, smem, b, j, g_base);
    }
}

template <int C>
__global__ void __launch_bounds__(kTcThreads, 1) mas_fused_kernel(const __grid_constant__ FusedParams fp,
                                                                  const __grid_constant__ CUtensorMap tm_z,
                                                                  const __grid_constant__ CUtensorMap tm_out)
{
    extern __shared__ __align__(128) unsigned char smem[];
    fused_body<C, false>(fp, &tm_z, &tm_out, smem);
}

// contraction CTAs in pairs (clusters of 2, n_gemm even); the DP CTAs ignore their cluster
template <int C>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kTcThreads, 1)
    mas_fused_pair_kernel(const __grid_constant__ FusedParams fp, const __grid_constant__ CUtensorMap tm_z,
                          const __grid_constant__ CUtensorMap tm_out)
{
    extern __shared__ __align__(128) unsigned char smem[];
    fused_body<C, true>(fp, &tm_z, &tm_out, smem);
}

static int env_int(const char *name, int dflt)
{
    const char *s = getenv(name);
    return (s && *s) ? atoi(s) : dflt;
}

bool fused_supported(int B, int D, int T, int S)
{
    if (env_int("MAS_NO_FUSED", 0)) return false;
    // tensor-map stores / vector cost loads need 16-byte rows; the contraction takes S <= 256
    return cost_tc_supported(B, D, T, S) && (S % 4 == 0) && (T % 4 == 0) && S <= 2 * kDpThreads;
}

size_t fused_flags_bytes(int B, int T) { return align_up((size_t)B * ((T + kBM - 1) / kBM) * 4, 256); }

int fused_launch(const float *z_p, const float *m_p, const float *logs_p, const int32_t *t_ys, const int32_t *t_xs,
                 float *neg_cent, bool skip_dead_tiles, void *path_out, int path_dtype, int32_t *dur_out,
                 int32_t *idx_out, int32_t *status_out, void *cost_ws, size_t cost_ws_bytes, void *dp_ws,
                 size_t dp_ws_bytes, uint32_t *flags, int B, int D, int T, int S, cudaStream_t stream)
{
    int dev = 0, sms = 148;
    MAS_CUDA_TRY(cudaGetDevice(&dev));
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int m_tiles = (T + kBM - 1) / kBM;
    TcPlan tc;
    int rc = cost_tc_prepare(tc, z_p, m_p, logs_p, neg_cent, nullptr, skip_dead_tiles ? t_ys : nullptr, cost_ws,
                             cost_ws_bytes, B, D, T, S, flags, B * m_tiles, stream);
    if (rc) return rc;
    if (!tc.p.z_tma || !tc.p.out_tma) return MAS_ERR_UNSUPPORTED_SHAPE;
    DpPlan dp;
    rc = dp_prepare(dp, neg_cent, t_ys, t_xs, path_out, path_dtype, dur_out, idx_out, status_out, dp_ws, dp_ws_bytes, B, T,
                    S, nullptr);
    if (rc) return rc;
    FusedParams fp;
    fp.tc = tc.p;
    fp.dp = dp.p;
    // split of the SMs between the two roles (MAS_FUSED_DP_CTAS overrides)
    int n_dp = env_int("MAS_FUSED_DP_CTAS", 0);
    if (n_dp <= 0) n_dp = B < 64 ? B : 64;
    if (n_dp > B) n_dp = B;
    if (n_dp > sms - 8) n_dp = sms - 8;
    const bool pair = cost_tc_pair_enabled();
    const int units = B * (pair ? (m_tiles + 1) / 2 : m_tiles);
    fp.n_dp = n_dp;
    fp.n_gemm = sms - n_dp;
    if (pair) {
        if (fp.n_gemm > 2 * units) fp.n_gemm = 2 * units;
        fp.n_gemm &= ~1;
        if ((fp.n_gemm + fp.n_dp) & 1) fp.n_dp -= 1;  // whole clusters only
        if (fp.n_dp < 1) return MAS_ERR_UNSUPPORTED_SHAPE;
    } else if (fp.n_gemm > units) {
        fp.n_gemm = units;
    }
    fp.tc.wave = fp.n_dp;
    fp.tc.flags = flags;
    fp.dp.flags = flags;
    fp.tc.trace = trace_buffer();
    fp.dp.trace = fp.tc.trace;
    fp.dp.flag_tiles = m_tiles;
    const size_t smem = dp.smem_bytes > kTcSmem ? dp.smem_bytes : (size_t)kTcSmem;

    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(fp.n_gemm + fp.n_dp);
    cfg.blockDim = dim3(kTcThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeCooperative;
    attr[0].val.cooperative = 1;
    cfg.attrs = attr;
    cfg.numAttrs = env_int("MAS_FUSED_COOP", 1) ? 1 : 0;
    static thread_local int configured_dev = -1;
    if (dev != configured_dev) {
        MAS_CUDA_TRY(cudaFuncSetAttribute(mas_fused_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget));
        MAS_CUDA_TRY(cudaFuncSetAttribute(mas_fused_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget));
        MAS_CUDA_TRY(cudaFuncSetAttribute(mas_fused_pair_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget));
        MAS_CUDA_TRY(cudaFuncSetAttribute(mas_fused_pair_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget));
        configured_dev = dev;
    }
    cudaError_t e;
    if (dp.C == 1)
        e = pair ? cudaLaunchKernelEx(&cfg, mas_fused_pair_kernel<1>, fp, tc.tm_z, tc.tm_out)
                 : cudaLaunchKernelEx(&cfg, mas_fused_kernel<1>, fp, tc.tm_z, tc.tm_out);
    else if (dp.C == 2)
        e = pair ? cudaLaunchKernelEx(&cfg, mas_fused_pair_kernel<2>, fp, tc.tm_z, tc.tm_out)
                 : cudaLaunchKernelEx(&cfg, mas_fused_kernel<2>, fp, tc.tm_z, tc.tm_out);
    else
        return MAS_ERR_UNSUPPORTED_SHAPE;
    note_launch();
    if (e != cudaSuccess) return note_cuda_error(e, "cudaLaunchKernelEx(mas_fused_kernel)");
    return MAS_OK;
}

}  // namespace mas
