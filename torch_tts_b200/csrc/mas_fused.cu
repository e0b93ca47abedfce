// mas_fused.cu -- neg_cent contraction and MAS in ONE kernel (no-noise alignment path,
// reference vits2/models.py:1224-1256 with mas_noise_scale None).
//
// One CTA per SM.  Every CTA starts in the tcgen05 contraction role (mas_cost_tc.cuh) over the unit list --
// utterance groups of n_dp, mel-tile-major inside a group -- and publishes every finished 128-row cost tile with
// a release store to a flag.  After seq_k rounds the first n_dp CTAs leave the contraction and run the forward
// DP + backtrack (mas_dp.cuh), one utterance at a time, acquiring the flag of a tile before the TMA engine
// streams its rows out of L2; the other CTAs finish the units and then drain the zero-fill queue of the path
// planes.  The cost plane round-trips through L2 only; the DP trails the contraction by a few tiles instead of
// waiting for the whole batch.  Producers never wait on consumers; the launch is cooperative so that all CTAs
// are co-resident, and a programmatic dependent of the prior-images kernel (fused_launch).
#include <atomic>

#include "mas_cost_tc.cuh"
#include "mas_dp.cuh"

namespace mas {

struct FusedParams {
    TcParams tc;
    DpParams dp;
    int n_dp;              // the first n_dp CTAs turn into DP CTAs after their share of the contraction
};

template <int C, int R, int W, bool kPair, bool kVK>
__device__ __forceinline__ void fused_body(const FusedParams &fp, const CUtensorMap *tm_z, const CUtensorMap *tm_out,
                                           unsigned char *smem)
{
    // every CTA starts in the contraction role; the first n_dp CTAs ("hybrid") leave it after seq_k rounds of
    // units and become the DP CTAs, the others finish the remaining units (unit_index() in cost_tc_role)
    if (fp.tc.trace && threadIdx.x == 0) fp.tc.trace[49152 + blockIdx.x] = globaltimer_ns();  // CTA entry
    if (kPair)
        cost_tc_role<false, true>(fp.tc, tm_z, tm_out, smem, blockIdx.x >> 1, gridDim.x >> 1);
    else
        cost_tc_role<false, false>(fp.tc, tm_z, tm_out, smem, blockIdx.x, gridDim.x);
    if (fp.tc.trace && threadIdx.x == 0) fp.tc.trace[49152 + 256 + blockIdx.x] = globaltimer_ns();  // contraction role left
    if ((int)blockIdx.x >= fp.n_dp) {
        // out of tiles: zero-fill the dense path planes while the DP CTAs are still busy
        if (fp.dp.zero_flags) zero_fill_role(fp.dp, smem);
        if (fp.tc.trace && threadIdx.x == 0) fp.tc.trace[49152 + 512 + blockIdx.x] = globaltimer_ns();  // zero-fill done
        return;
    }
    // DP CTA j aligns utterances j, j + n_dp, ... on its first dp_threads(W) threads
    if ((int)threadIdx.x >= dp_threads(W, kVK)) return;
    const int j = (int)blockIdx.x;
    uint32_t g_base = 0;
    dp_role_init(fp.dp, smem, threadIdx.x, kDpBar);
    for (int b = j; b < fp.dp.B; b += fp.n_dp)
        dp_role<C, R, W, true, false, kVK>(fp.dp, smem, b, j, g_base, threadIdx.x, kDpBar);
}

template <int C, int R, int W, bool kVK>
__global__ void __launch_bounds__(kTcThreads, 1) mas_fused_kernel(const __grid_constant__ FusedParams fp,
                                                                  const __grid_constant__ CUtensorMap tm_z,
                                                                  const __grid_constant__ CUtensorMap tm_out)
{
    extern __shared__ __align__(128) unsigned char smem[];
    fused_body<C, R, W, false, kVK>(fp, &tm_z, &tm_out, smem);
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

static int env_int(const char *name, int dflt)
{
    const char *s = getenv(name);
    return (s && *s) ? atoi(s) : dflt;
}

bool fused_supported(int B, int D, int T, int S)
{
    if (env_int("MAS_NO_FUSED", 0)) return false;
    // tensor-map stores / vector cost loads need 16-byte rows; the contraction takes S <= 256
    return cost_tc_supported(B, D, T, S) && (S % 4 == 0) && (T % 4 == 0) && S <= kNMax;
}

// tile flags [B][m_tiles], the zero-fill flags [B], the zero-fill queue counter
size_t fused_flags_bytes(int B, int T) { return align_up(((size_t)B * ((T + kBM - 1) / kBM + 1) + 1) * 4, 256); }

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
                             cost_ws_bytes, B, D, T, S, flags, B * (m_tiles + 1) + 1, stream);
    if (rc) return rc;
    if (!tc.p.z_tma || !tc.p.out_tma) return MAS_ERR_UNSUPPORTED_SHAPE;
    DpPlan dp;
    rc = dp_prepare(dp, neg_cent, t_ys, t_xs, path_out, path_dtype, dur_out, idx_out, status_out, dp_ws, dp_ws_bytes, B, T,
                    S, nullptr, 0, 0);
    if (rc) return rc;
    FusedParams fp;
    fp.tc = tc.p;
    fp.dp = dp.p;
    // DP CTAs (MAS_FUSED_DP_CTAS overrides): one per utterance; when the batch is larger than the GPU, nearly
    // every CTA becomes a DP CTA after the contraction and aligns several utterances in turn (the standalone
    // MAS kernel reaches the HBM roofline that way: the DP wants every SM's memory pipe)
    const bool pair = cost_tc_pair_enabled();
    const int grid = pair ? (sms & ~1) : sms;
    const int per = pair ? 2 : 1;
    const int units = B * (pair ? (m_tiles + 1) / 2 : m_tiles);
    // Unit schedule: all CTAs take seq_k rounds of units, then the DP CTAs leave.  A small cost model (unit
    // ~7 us, DP ~40 ns per mel row + 8 us per utterance) picks seq_k and, for batches larger than 64, between
    // 64 DP CTAs and nearly all of them: the contraction must not end long after the DP could, and the DP
    // must not start long before its tiles exist.
    auto model = [&](int n_dp_try, int &k_out) {
        const int P = grid / per, P_pure = (grid - n_dp_try) / per;
        const int utts = (B + n_dp_try - 1) / n_dp_try;
        double best = 1e30;
        for (int k = 0; k <= 64; ++k) {
            const int rest = units - k * P > 0 ? units - k * P : 0;
            const double t_u = 7.0;
            const double gemm_end = (k + (rest + P_pure - 1) / P_pure) * t_u;
            const double dp_end = k * t_u + utts * (T * 0.040 + 8.0);
            const double t = (gemm_end + 8.0 > dp_end) ? gemm_end + 8.0 : dp_end;
            if (t < best - 1e-9) best = t, k_out = k;
            if (rest == 0) break;
        }
        return best;
    };
    auto legal = [&](int n) {
        if (n > B) n = B;
        if (pair) n = (n + 1) & ~1;  // whole pairs
        if (n > grid - 8) n = (grid - 8) & ~1;
        return n;
    };
    int n_dp = env_int("MAS_FUSED_DP_CTAS", 0), best_k = 0;
    if (n_dp > 0) {
        n_dp = legal(n_dp);
        model(n_dp, best_k);
    } else {
        int k_a = 0, k_b = 0;
        const int n_a = legal(B < 64 ? B : 64), n_b = legal(B);
        const double t_a = model(n_a, k_a), t_b = model(n_b, k_b);
        n_dp = (t_b < t_a) ? n_b : n_a;
        best_k = (t_b < t_a) ? k_b : k_a;
    }
    if (n_dp < 1) return MAS_ERR_UNSUPPORTED_SHAPE;
    fp.n_dp = n_dp;
    const int utts_per_cta = (B + n_dp - 1) / n_dp;
    fp.tc.seq_k = env_int("MAS_FUSED_ROUNDS", -1) >= 0 ? env_int("MAS_FUSED_ROUNDS", -1) : best_k;
    fp.tc.seq_pure0 = n_dp / per;
    fp.tc.wave = fp.n_dp;
    fp.tc.flags = flags;
    fp.dp.flags = flags;
    fp.tc.trace = trace_buffer();
    fp.dp.trace = fp.tc.trace;
    fp.dp.flag_tiles = m_tiles;
    // the contraction-only CTAs zero-fill the path planes when there are enough of them to do it in time
    const bool offload = env_int("MAS_FUSED_ZERO_OFFLOAD", 1) && (grid - n_dp) * 2 >= n_dp && utts_per_cta == 1;
    fp.dp.zero_flags = offload ? flags + (size_t)B * m_tiles : nullptr;
    fp.dp.zero_queue = flags + (size_t)B * (m_tiles + 1);
    size_t smem = dp.smem_bytes;
    if (smem < kTcSmem) smem = kTcSmem;
    if (smem < kZeroFillBuf) smem = kZeroFillBuf;

    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(kTcThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    // Launch attributes.  Cooperative: the DP CTAs spin on flags the other CTAs raise, so all of them must be
    // resident.  Programmatic stream serialization (MAS_FUSED_PDL=0 turns it off): the grid may start while the
    // prior-images kernel ahead of it in the stream is still running, and overlaps its prologue (barriers, TMEM,
    // tensor maps) with that kernel; cost_tc_role waits for it before the first dependent read.  Tried once
    // outside stream capture; a driver that refuses the combination gets plain launches from then on.
    static std::atomic<int> pdl_state{-1};  // -1 untested, 0 refused, 1 works (calls may come from several threads)
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    cudaStreamIsCapturing(stream, &cap);
    bool use_pdl = env_int("MAS_FUSED_PDL", 1) && (pdl_state == 1 || (pdl_state < 0 && cap == cudaStreamCaptureStatusNone));
    cudaLaunchAttribute attr[2];
    int n_attr = 0;
    if (env_int("MAS_FUSED_COOP", 1)) {
        attr[n_attr].id = cudaLaunchAttributeCooperative;
        attr[n_attr].val.cooperative = 1;
        ++n_attr;
    }
    if (use_pdl) {
        attr[n_attr].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[n_attr].val.programmaticStreamSerializationAllowed = 1;
        ++n_attr;
    }
    cfg.attrs = attr;
    cfg.numAttrs = n_attr;
    fp.tc.pdl = 1;  // (the wait is a no-op under a plain launch)
    cudaError_t e = cudaErrorInvalidValue;
#define MAS_FUSED_CASE(CC, WW, VK)                                                                               \
    if (dp.C == CC && dp.p.R == 32 && dp.p.W == WW && (dp.p.vk != 0) == VK) {                                    \
        static thread_local int cfg_dev = -1;                                                                    \
        if (dev != cfg_dev) {                                                                                    \
            MAS_CUDA_TRY(cudaFuncSetAttribute(mas_fused_kernel<CC, 32, WW, VK>,                                  \
                                              cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget));        \
            MAS_CUDA_TRY(cudaFuncSetAttribute(mas_fused_pair_kernel<CC, 32, WW, VK>,                             \
                                              cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget));        \
            cfg_dev = dev;                                                                                       \
        }                                                                                                        \
        e = pair ? cudaLaunchKernelEx(&cfg, mas_fused_pair_kernel<CC, 32, WW, VK>, fp, tc.tm_z, tc.tm_out)       \
                 : cudaLaunchKernelEx(&cfg, mas_fused_kernel<CC, 32, WW, VK>, fp, tc.tm_z, tc.tm_out);           \
    }
    // S <= 256 (the contraction's limit): C = ceil(S / 64) columns per thread with 2 DP warps (value /
    // bookkeeping split by default), or ceil(S / 128) with 4 (MAS_DP_WARPS=4)
    for (int attempt = 0; attempt < 2; ++attempt) {
    MAS_FUSED_CASE(1, 2, true)
    else MAS_FUSED_CASE(2, 2, true)
    else MAS_FUSED_CASE(3, 2, true)
    else MAS_FUSED_CASE(4, 2, true)
    else MAS_FUSED_CASE(1, 2, false)
    else MAS_FUSED_CASE(2, 2, false)
    else MAS_FUSED_CASE(3, 2, false)
    else MAS_FUSED_CASE(4, 2, false)
    else MAS_FUSED_CASE(2, 4, false)
    else return MAS_ERR_UNSUPPORTED_SHAPE;
        if (!use_pdl) break;
        if (e == cudaSuccess) {
            pdl_state = 1;
            break;
        }
        if (pdl_state == 1) break;  // worked before: a real error
        (void)cudaGetLastError();
        pdl_state = 0;
        use_pdl = false;
        cfg.numAttrs = n_attr - 1;  // the programmatic attribute is the last one
    }
#undef MAS_FUSED_CASE
    note_launch();
    if (e != cudaSuccess) return note_cuda_error(e, "cudaLaunchKernelEx(mas_fused_kernel)");
    return MAS_OK;
}

}  // namespace mas
