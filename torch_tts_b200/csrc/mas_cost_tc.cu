// mas_cost_tc.cu -- standalone launch of the tcgen05 neg_cent contraction (role code in
// mas_cost_tc.cuh) and the prior preparation kernel that feeds it.
#include "mas_cost_tc.cuh"
#include "mas_tmap.cuh"

namespace mas {

// ---------------------------------------------------------------------------
// prior preparation: B tile images (bf16 hi/lo, pre-swizzled) + per-K-block bias partials
//   image (b, kb, part) = kBPart bytes: rows s in [0, 256), k in [0, 32):
//     k < 16: r[kb*16 + k][s]      k >= 16: (m r)[kb*16 + k - 16][s]       (0 past S or D)
//   bias_part[b][kb][s] = sum over the 16 channels of the block of
//     (-0.5 log 2pi - logs_p) + (-0.5 m^2 r);  the contraction's epilogue adds the n_kb
//     partials in a fixed order (deterministic, no atomics).
// One CTA per (kb, b, half of the column block), one thread per text column: every global load is coalesced along s
// and all 32 loads of a thread are in flight before the first use.  The grid also zeroes the
// tile flags of the fused kernel.
// ---------------------------------------------------------------------------
constexpr int kPriorThreads = 128;              // two CTAs per image: 1536 small CTAs at config 2 fill the GPU's
constexpr int kPriorParts = kNMax / kPriorThreads;  // second wave where 768 CTAs of 256 threads left it 70 % idle
__global__ void __launch_bounds__(kPriorThreads) mas_prior_images_kernel(const float *__restrict__ m_p,
                                                                const float *__restrict__ logs_p,
                                                                unsigned char *__restrict__ images,
                                                                float *__restrict__ bias_part, int D, int S, int n_kb,
                                                                uint32_t *flags_to_clear, int n_flags)
{
    // let a programmatic dependent (the fused kernel) start its prologue while this grid runs; it waits for
    // this grid's completion before it reads anything written here
    asm volatile("griddepcontrol.launch_dependents;");
    const int kb = blockIdx.x, b = blockIdx.y;                    // K block, utterance
    const int nb = blockIdx.z / kPriorParts, n_blocks = gridDim.z / kPriorParts;   // column block (of kNMax columns)
    const int s_img = (blockIdx.z % kPriorParts) * kPriorThreads + threadIdx.x;    // row of the image
    const int s = nb * kNMax + s_img;                             // text column
    if (flags_to_clear) {
        const int n_cta = gridDim.x * gridDim.y * gridDim.z, cta = (blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
        for (int i = cta * kPriorThreads + threadIdx.x; i < n_flags; i += n_cta * kPriorThreads) flags_to_clear[i] = 0u;
    }
    const bool live = s < S;
    const size_t base = (size_t)b * D * S + (live ? s : 0);
    const float c0 = -0.91893853320467274178f;  // -0.5*log(2*pi)
    float l[kDPerKb], m[kDPerKb];
#pragma unroll
    for (int j = 0; j < kDPerKb; ++j) {
        const int d = kb * kDPerKb + j;
        const bool ok = live && d < D;
        l[j] = ok ? __ldg(logs_p + base + (size_t)d * S) : 0.f;
        m[j] = ok ? __ldg(m_p + base + (size_t)d * S) : 0.f;
    }
    float acc1 = 0.f, acc4 = 0.f;
    float r[kDPerKb];
#pragma unroll
    for (int j = 0; j < kDPerKb; ++j) {
        const int d = kb * kDPerKb + j;
        const bool ok = live && d < D;
        const float rr = ok ? expf(-2.0f * l[j]) : 0.f;   // models.py:1226
        r[j] = rr;
        acc1 += ok ? c0 - l[j] : 0.f;                      // :1227-1229
        acc4 += -0.5f * (m[j] * m[j]) * rr;                // :1236-1238 (0 when !ok)
        m[j] = m[j] * rr;                                  // m r, :1234
    }
    unsigned char *img_hi = images + (((size_t)(b * n_blocks + nb) * n_kb + kb) * 2 + 0) * kBPart;
    unsigned char *img_lo = images + (((size_t)(b * n_blocks + nb) * n_kb + kb) * 2 + 1) * kBPart;
#pragma unroll
    for (int half = 0; half < 2; ++half) {      // 0: r, 1: m r
#pragma unroll
        for (int c = 0; c < 2; ++c) {            // 16-byte chunks of 8 k each
            uint32_t hi[4], lo[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int k = c * 8 + j * 2;
                split2(half ? m[k] : r[k], half ? m[k + 1] : r[k + 1], hi[j], lo[j]);
            }
            const uint32_t off = sw64_offset(s_img, half * 16 + c * 8);
            *reinterpret_cast<uint4 *>(img_hi + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
            *reinterpret_cast<uint4 *>(img_lo + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        }
    }
    if (live) bias_part[((size_t)b * n_kb + kb) * S + s] = acc1 + acc4;
}

#ifdef MAS_TRACE
// single-CTA contraction (cta_group::1): trace build only, for A/B runs against the CTA-pair kernel (MAS_TC_PAIR=0)
template <bool kStats>
__global__ void __launch_bounds__(kTcThreads, 1) mas_cost_tc_kernel(const TcParams p,
                                                                    const __grid_constant__ CUtensorMap tm_z,
                                                                    const __grid_constant__ CUtensorMap tm_out)
{
    extern __shared__ unsigned char smem_raw[];
    cost_tc_role<kStats, false>(p, &tm_z, &tm_out, smem_raw, blockIdx.x, gridDim.x);
}
#endif

// CTA-pair version: clusters of 2, one M = 256 cta_group::2 MMA per pair
template <bool kStats>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kTcThreads, 1)
    mas_cost_tc_pair_kernel(const TcParams p, const __grid_constant__ CUtensorMap tm_z,
                            const __grid_constant__ CUtensorMap tm_out)
{
    extern __shared__ unsigned char smem_raw[];
    cost_tc_role<kStats, true>(p, &tm_z, &tm_out, smem_raw, blockIdx.x >> 1, gridDim.x >> 1);
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
// trace build: MAS_TC_PAIR=0 selects the single-CTA (cta_group::1) contraction
bool cost_tc_pair_enabled() { return !kTrace || config().tc_pair != 0; }

bool cost_tc_supported(int B, int D, int T, int S)
{
    (void)B;
    (void)T;
    return S <= 4 * kNMax && D >= 1;
}

size_t cost_tc_workspace_bytes(int B, int D, int T, int S)
{
    (void)T;
    const int n_kb = (D + kDPerKb - 1) / kDPerKb, n_blocks = (S + kNMax - 1) / kNMax;
    return align_up((size_t)B * n_blocks * n_kb * 2 * kBPart, 256) + align_up((size_t)B * n_kb * S * 4, 256);
}

int cost_tc_prepare(TcPlan &plan, const float *z_p, const float *m_p, const float *logs_p, float *neg_cent_out,
                    double *stats_out, const int32_t *t_ys, void *workspace, size_t workspace_bytes, int B, int D, int T,
                    int S, uint32_t *flags_to_clear, int n_flags, cudaStream_t stream, int ld)
{
    if (ld <= 0) ld = S;
    if (!workspace || workspace_bytes < cost_tc_workspace_bytes(B, D, T, S)) return MAS_ERR_WORKSPACE;
    const int n_kb = (D + kDPerKb - 1) / kDPerKb, n_blocks = (S + kNMax - 1) / kNMax;
    unsigned char *images = static_cast<unsigned char *>(workspace);
    float *bias = reinterpret_cast<float *>(images + align_up((size_t)B * n_blocks * n_kb * 2 * kBPart, 256));
    if (stats_out) MAS_CUDA_TRY(cudaMemsetAsync(stats_out, 0, 2 * sizeof(double), stream));
    if (config().stage != 2) {   // MAS_STAGE=2: reuse the images already in the workspace (timing experiments)
        mas_prior_images_kernel<<<dim3(n_kb, B, n_blocks * kPriorParts), kPriorThreads, 0, stream>>>(m_p, logs_p, images, bias, D, S, n_kb,
                                                                     flags_to_clear, n_flags);
        note_launch();
        MAS_CUDA_TRY(cudaGetLastError());
    }

    TcParams &p = plan.p;
    p = TcParams{};
    p.z_p = z_p;
    p.images = images;
    p.bias_part = bias;
    p.out = neg_cent_out;
    p.stats = stats_out;
    p.t_ys = stats_out ? nullptr : t_ys;  // the noise std covers every cell, padding included (models.py:1243)
    p.B = B;
    p.D = D;
    p.T = T;
    p.S = S;
    p.n_kb = n_kb;
    p.n_blocks = n_blocks;
    p.m_tiles = (T + kBM - 1) / kBM;
    p.wave = B;
    p.seq_k = 1 << 28;
    p.seq_pure0 = 0;
    p.ld = ld;
    p.trace = trace_buffer();
    p.debug = config().tc_debug;
    // MAS_TC_NO_TMA: bit 1 plain z loads, bit 2 plain output stores -- the fallbacks for pointers / pitches a tensor map
    // cannot describe (or a driver without cuTensorMapEncodeTiled), forced by tests/test_gpu_round2.py
    const int no_tma = config().tc_no_tma;
    // tensor maps: z_p as [B][D][T] with a [1][16][128] box, neg_cent as [B][T][ld] with a swizzled [1][32][32] box
    p.z_tma = !(no_tma & 1) && make_tmap_f32_3d(&plan.tm_z, z_p, (uint64_t)T, (uint64_t)D, (uint64_t)B, (uint64_t)T * 4,
                                                (uint64_t)D * T * 4, kBM, kDPerKb, 1, false);
    p.out_tma = !(no_tma & 2) && make_tmap_f32_3d(&plan.tm_out, neg_cent_out, (uint64_t)ld, (uint64_t)T, (uint64_t)B,
                                                  (uint64_t)ld * 4, (uint64_t)T * ld * 4, 32, 32, 1, true);
    return MAS_OK;
}

int cost_tc_launch(const float *z_p, const float *m_p, const float *logs_p, float *neg_cent_out, double *stats_out,
                   const int32_t *t_ys, void *workspace, size_t workspace_bytes, int B, int D, int T, int S,
                   cudaStream_t stream, int ld)
{
    TcPlan plan;
    int rc = cost_tc_prepare(plan, z_p, m_p, logs_p, neg_cent_out, stats_out, t_ys, workspace, workspace_bytes, B, D, T,
                             S, nullptr, 0, stream, ld);
    if (rc) return rc;
    static thread_local int configured_dev = -1;
    int dev = 0, sms = 148;
    MAS_CUDA_TRY(cudaGetDevice(&dev));
    if (dev != configured_dev) {
#ifdef MAS_TRACE
        MAS_CUDA_TRY(cudaFuncSetAttribute(mas_cost_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)kTcSmem));
        MAS_CUDA_TRY(cudaFuncSetAttribute(mas_cost_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)kTcSmem));
#endif
        MAS_CUDA_TRY(cudaFuncSetAttribute(mas_cost_tc_pair_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)kTcSmem));
        MAS_CUDA_TRY(cudaFuncSetAttribute(mas_cost_tc_pair_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)kTcSmem));
        configured_dev = dev;
    }
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (config().tc_grid > 0 && config().tc_grid < sms) sms = config().tc_grid;  // trace build: fewer SMs
    if (config().stage == 1) return MAS_OK;  // MAS_STAGE=1: prior preparation only (bench.py times it alone)
    if (cost_tc_pair_enabled()) {
        const int n_units = B * ((plan.p.m_tiles + 1) / 2) * plan.p.n_blocks;
        int grid = 2 * n_units < sms ? 2 * n_units : (sms & ~1);
        if (stats_out)
            mas_cost_tc_pair_kernel<true><<<grid, kTcThreads, kTcSmem, stream>>>(plan.p, plan.tm_z, plan.tm_out);
        else
            mas_cost_tc_pair_kernel<false><<<grid, kTcThreads, kTcSmem, stream>>>(plan.p, plan.tm_z, plan.tm_out);
    } else {
#ifdef MAS_TRACE
        const int n_tiles = B * plan.p.m_tiles * plan.p.n_blocks;
        const int grid = n_tiles < sms ? n_tiles : sms;
        if (stats_out)
            mas_cost_tc_kernel<true><<<grid, kTcThreads, kTcSmem, stream>>>(plan.p, plan.tm_z, plan.tm_out);
        else
            mas_cost_tc_kernel<false><<<grid, kTcThreads, kTcSmem, stream>>>(plan.p, plan.tm_z, plan.tm_out);
#endif
    }
    note_launch();
    MAS_CUDA_TRY(cudaGetLastError());
    return MAS_OK;
}

}  // namespace mas
