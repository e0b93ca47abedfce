// mas_cost_tc.cu -- the neg_cent contraction on the 5th-gen tensor cores (tcgen05 / TMEM).
//
// Reference: vits2/models.py:1226-1239.  With r = exp(-2 logs_p):
//   neg_cent[b,t,s] = bias[b,s] + sum_d (-0.5 z^2)[d,t] r[d,s] + z[d,t] (m r)[d,s]
// i.e. one K = 2D contraction  A[t,k] . B[s,k]  per utterance with
//   A = [-0.5 z^2 | z]   (from z_p, converted on the fly, never materialised in HBM)
//   B = [ r       | m r] (from m_p/logs_p, prepared once per utterance)
//
// Precision: fp32 operands are split into bf16 hi + bf16 lo (x ~ hi + lo, 16 mantissa
// bits) and the product is evaluated as A_hi B_hi + A_lo B_hi + A_hi B_lo with fp32
// accumulation in TMEM: three kind::f16 MMAs per K step, ~2^-16 relative per term,
// ~1e-6 relative on neg_cent (the reference's own fp32 sgemm noise level), at the
// cost of 3 bf16 passes = 1.5 TF32 passes instead of the 3 a 3xTF32 split needs.
//
// Kernel layout (persistent, one CTA per SM, 12 warps):
//   warp 0      B producer: 1-D TMA bulk copies of pre-swizzled B tile images -> smem
//   warp 1      MMA issuer: one thread issues tcgen05.mma, commits to mbarriers
//   warp 2      TMEM allocator
//   warps 4-7   epilogue: tcgen05.ld accumulator -> + bias -> global (and noise statistics)
//   warps 8-11  A converters: z_p (global, coalesced along t) -> -0.5 z^2, z -> bf16 hi/lo
//               -> K-major SWIZZLE_128B tiles in smem
// Tile = 128 mel rows x N text columns (N = S rounded up to 16, <= 256), K blocks of
// 64 bf16 (32 prior channels), 2 smem stages of 96 KB, 2 TMEM accumulators of 256 columns.
#include "mas_common.cuh"

namespace mas {

constexpr int kTcThreads = 384;
constexpr int kBM = 128;            // mel rows per tile (UMMA M)
constexpr int kBK = 64;             // bf16 K elements per block (= one 128-byte swizzle row)
constexpr int kDPerKb = kBK / 2;    // prior channels per K block
constexpr int kNMax = 256;          // text columns per tile (UMMA N max)
constexpr int kTcStages = 2;
constexpr uint32_t kAPart = kBM * 128;      // 16 KB: one split part of A per stage
constexpr uint32_t kBPart = kNMax * 128;    // 32 KB: one split part of B per stage / per image
constexpr uint32_t kStageBytes = 2 * kAPart + 2 * kBPart;  // 96 KB
constexpr uint32_t kTcSmem = kTcStages * kStageBytes + 1024 /*bias*/ + 256 /*barriers*/ + 1024 /*align*/;

// ---- PTX: tcgen05 ------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t *dst_smem, uint32_t cols)
{
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols)
{
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem] . B[smem]^T, bf16 x bf16 -> fp32
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrive on an mbarrier when all previously issued MMAs of this thread have completed
__device__ __forceinline__ void umma_commit(uint64_t *bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t addr, uint32_t (&r)[32])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(addr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major SWIZZLE_128B shared-memory matrix descriptor: rows of 128 bytes, 8-row atoms of
// 1024 bytes (stride byte offset), version 1 (sm_100), layout type 2.
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t smem_addr)
{
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;            // leading byte offset (unused for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;  // stride byte offset between 8-row groups
    d |= (uint64_t)1 << 46;            // descriptor version
    d |= (uint64_t)2 << 61;            // SWIZZLE_128B
    return d;
}

// kind::f16 instruction descriptor: D fp32, A/B bf16, both K-major, M = 128, N = n
__host__ __device__ __forceinline__ uint32_t make_idesc_bf16(int n)
{
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(kBM >> 4) << 24);
}

// byte offset of element (row, k) inside a K-major SWIZZLE_128B bf16 tile (row pitch 128 B)
__host__ __device__ __forceinline__ uint32_t sw128_offset(int row, int k)
{
    return (uint32_t)((row >> 3) * 1024 + (row & 7) * 128 + ((((k >> 3) ^ (row & 7)) & 7) << 4) + (k & 7) * 2);
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi)
{
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));  // first source -> upper half
    return r;
}

// split two floats into packed bf16 hi parts and packed bf16 lo parts (x ~ hi + lo)
__device__ __forceinline__ void split2(float x0, float x1, uint32_t &hi, uint32_t &lo)
{
    hi = pack_bf16x2(x0, x1);
    const float h0 = __uint_as_float(hi << 16);
    const float h1 = __uint_as_float(hi & 0xffff0000u);
    lo = pack_bf16x2(x0 - h0, x1 - h1);
}

// ---------------------------------------------------------------------------
// prior preparation: B tile images (bf16 hi/lo, pre-swizzled) + bias
//   image (b, kb, part) = kBPart bytes: rows s in [0, 256), k in [0, 64):
//     k < 32: r[kb*32 + k][s]      k >= 32: (m r)[kb*32 + k - 32][s]       (0 past S or D)
// ---------------------------------------------------------------------------
__global__ void mas_prior_images_kernel(const float *__restrict__ m_p, const float *__restrict__ logs_p,
                                        unsigned char *__restrict__ images, float *__restrict__ bias_out, int D,
                                        int S, int n_kb)
{
    const int b = blockIdx.y;
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= kNMax) return;
    const bool live = s < S;
    const size_t base = (size_t)b * D * S + s;
    const float c0 = -0.91893853320467274178f;  // -0.5*log(2*pi)
    float acc1 = 0.f, acc4 = 0.f;
    for (int kb = 0; kb < n_kb; ++kb) {
        unsigned char *img_hi = images + ((size_t)(b * n_kb + kb) * 2 + 0) * kBPart;
        unsigned char *img_lo = images + ((size_t)(b * n_kb + kb) * 2 + 1) * kBPart;
#pragma unroll
        for (int half = 0; half < 2; ++half) {      // 0: r, 1: m r
#pragma unroll
            for (int c = 0; c < 4; ++c) {            // 16-byte chunks of 8 k each
                uint32_t hi[4], lo[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    float x[2];
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const int d = kb * kDPerKb + c * 8 + j * 2 + e;
                        float val = 0.f;
                        if (live && d < D) {
                            const float l = logs_p[base + (size_t)d * S];
                            const float m = m_p[base + (size_t)d * S];
                            const float r = expf(-2.0f * l);
                            val = half ? m * r : r;
                            if (half == 0) {
                                acc1 += c0 - l;
                                acc4 += -0.5f * (m * m) * r;
                            }
                        }
                        x[e] = val;
                    }
                    split2(x[0], x[1], hi[j], lo[j]);
                }
                const uint32_t off = sw128_offset(s, half * 32 + c * 8);
                *reinterpret_cast<uint4 *>(img_hi + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                *reinterpret_cast<uint4 *>(img_lo + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
            }
        }
    }
    if (live) bias_out[(size_t)b * S + s] = acc1 + acc4;
}

struct TcParams {
    const float *z_p;
    const unsigned char *images;
    const float *bias;
    float *out;
    double *stats;          // nullable
    const int32_t *t_ys;    // nullable: skip mel tiles entirely past t_y (no noise statistics then)
    int B, D, T, S;
    int n_kb;               // K blocks = ceil(D / 32)
    int n_cols;             // UMMA N = S rounded up to 16
    int m_tiles;            // ceil(T / 128)
};

__global__ void __launch_bounds__(kTcThreads, 1) mas_cost_tc_kernel(const TcParams p)
{
    extern __shared__ unsigned char smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    unsigned char *smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);  // SWIZZLE_128B needs 1024-byte alignment
    float *bias_s = reinterpret_cast<float *>(smem + kTcStages * kStageBytes);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + kTcStages * kStageBytes + 1024);
    uint64_t *full = bars;            // [2] stage filled: B bytes landed + 4 converter warps arrived
    uint64_t *empty = bars + 2;       // [2] stage consumed by the MMAs
    uint64_t *acc_full = bars + 4;    // [2] accumulator complete
    uint64_t *acc_empty = bars + 6;   // [2] accumulator drained by the epilogue
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 8);

    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    const int lane = tid & 31;
    const int n_tiles = p.B * p.m_tiles;

    if (tid == 0) {
        for (int i = 0; i < 2; ++i) {
            mbar_init(&full[i], 1 + 4);
            mbar_init(&empty[i], 1);
            mbar_init(&acc_full[i], 1);
            mbar_init(&acc_empty[i], 4);
        }
        fence_mbar_init();
    }
    if (warp == 2) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    auto tile_live = [&](int tile) {
        if (!p.t_ys) return true;
        const int b = tile / p.m_tiles, mt = tile - b * p.m_tiles;
        return mt * kBM < p.t_ys[b];
    };

    if (warp == 0) {
        // ======================= B producer =======================
        if (lane == 0) {
            uint32_t it = 0;
            const uint32_t bytes = (uint32_t)p.n_cols * 128u;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                if (!tile_live(tile)) continue;
                const int b = tile / p.m_tiles;
                for (int kb = 0; kb < p.n_kb; ++kb, ++it) {
                    const uint32_t s = it % kTcStages, ph = (it / kTcStages) & 1u;
                    mbar_wait(&empty[s], ph ^ 1u);
                    unsigned char *stage = smem + s * kStageBytes;
                    const unsigned char *img = p.images + (size_t)(b * p.n_kb + kb) * 2 * kBPart;
                    mbar_arrive_expect_tx(&full[s], 2 * bytes);
                    bulk_g2s(stage + 2 * kAPart, img, bytes, &full[s]);
                    bulk_g2s(stage + 2 * kAPart + kBPart, img + kBPart, bytes, &full[s]);
                }
            }
        }
    } else if (warp == 1) {
        // ======================= MMA issuer =======================
        if (lane == 0) {
            const uint32_t idesc = make_idesc_bf16(p.n_cols);
            uint32_t it = 0, nt = 0;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                if (!tile_live(tile)) continue;
                const uint32_t a = nt & 1u, aph = (nt >> 1) & 1u;
                mbar_wait(&acc_empty[a], aph ^ 1u);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + a * kNMax;
                for (int kb = 0; kb < p.n_kb; ++kb, ++it) {
                    const uint32_t s = it % kTcStages, ph = (it / kTcStages) & 1u;
                    mbar_wait(&full[s], ph);
                    tc_fence_after();
                    const uint32_t st = smem_u32(smem + s * kStageBytes);
                    const uint64_t a_hi = make_desc_sw128(st), a_lo = make_desc_sw128(st + kAPart);
                    const uint64_t b_hi = make_desc_sw128(st + 2 * kAPart), b_lo = make_desc_sw128(st + 2 * kAPart + kBPart);
#pragma unroll
                    for (int k = 0; k < kBK / 16; ++k) {
                        const uint64_t adv = (uint64_t)((k * 32) >> 4);  // 16 bf16 = 32 bytes along K
                        umma_bf16(tmem_d, a_hi + adv, b_hi + adv, idesc, (kb | k) ? 1u : 0u);
                        umma_bf16(tmem_d, a_lo + adv, b_hi + adv, idesc, 1u);
                        umma_bf16(tmem_d, a_hi + adv, b_lo + adv, idesc, 1u);
                    }
                    umma_commit(&empty[s]);  // frees the stage when these MMAs have read it
                }
                umma_commit(&acc_full[a]);
                ++nt;
            }
        }
    } else if (warp >= 4 && warp < 8) {
        // ======================= epilogue =======================
        const int wq = warp & 3;  // TMEM lane quarter this warp may read
        const int row = wq * 32 + lane;
        uint32_t nt = 0;
        double ssum = 0.0, ssq = 0.0;
        int bias_b = -1;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            if (!tile_live(tile)) continue;
            const int b = tile / p.m_tiles, mt = tile - b * p.m_tiles;
            if (b != bias_b) {
                // bias of this utterance -> smem (only the 4 epilogue warps sync here)
                asm volatile("bar.sync 1, 128;" ::: "memory");
                for (int s = tid - 128; s < p.n_cols; s += 128) bias_s[s] = (s < p.S) ? p.bias[(size_t)b * p.S + s] : 0.f;
                asm volatile("bar.sync 1, 128;" ::: "memory");
                bias_b = b;
            }
            const uint32_t a = nt & 1u, aph = (nt >> 1) & 1u;
            mbar_wait(&acc_full[a], aph);
            tc_fence_after();
            const int t = mt * kBM + row;
            float *orow = p.out + ((size_t)b * p.T + t) * p.S;
            const uint32_t taddr = tmem_base + a * kNMax + ((uint32_t)(wq * 32) << 16);
            const bool vec_ok = (p.S & 3) == 0;
            for (int c0 = 0; c0 < p.n_cols; c0 += 32) {
                uint32_t r[32];
                tmem_ld32(taddr + c0, r);
                tmem_ld_wait();
                if (t < p.T) {
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        float v4[4];
#pragma unroll
                        for (int e = 0; e < 4; ++e) v4[e] = __uint_as_float(r[j + e]) + bias_s[c0 + j + e];
                        const int s = c0 + j;
                        if (vec_ok && s + 3 < p.S) {
                            *reinterpret_cast<float4 *>(orow + s) = make_float4(v4[0], v4[1], v4[2], v4[3]);
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                ssum += (double)v4[e];
                                ssq += (double)v4[e] * (double)v4[e];
                            }
                        } else {
#pragma unroll
                            for (int e = 0; e < 4; ++e)
                                if (s + e < p.S) {
                                    orow[s + e] = v4[e];
                                    ssum += (double)v4[e];
                                    ssq += (double)v4[e] * (double)v4[e];
                                }
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[a]);
            ++nt;
        }
        if (p.stats) {
            for (int o = 16; o > 0; o >>= 1) {
                ssum += __shfl_xor_sync(kFullMask, ssum, o);
                ssq += __shfl_xor_sync(kFullMask, ssq, o);
            }
            if (lane == 0) {
                atomicAdd(&p.stats[0], ssum);
                atomicAdd(&p.stats[1], ssq);
            }
        }
    } else if (warp >= 8) {
        // ======================= A converters =======================
        const int row = tid - 256;  // 0..127: mel row of the tile handled by this thread
        uint32_t it = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            if (!tile_live(tile)) continue;
            const int b = tile / p.m_tiles, mt = tile - b * p.m_tiles;
            const int t = mt * kBM + row;
            const bool live = t < p.T;
            const float *zb = p.z_p + (size_t)b * p.D * p.T + (live ? t : 0);
            float zn[kDPerKb];
            // prefetch K block 0
#pragma unroll
            for (int d = 0; d < kDPerKb; ++d) zn[d] = (live && d < p.D) ? zb[(size_t)d * p.T] : 0.f;
            for (int kb = 0; kb < p.n_kb; ++kb, ++it) {
                float zc[kDPerKb];
#pragma unroll
                for (int d = 0; d < kDPerKb; ++d) zc[d] = zn[d];
                if (kb + 1 < p.n_kb) {
                    const int d0 = (kb + 1) * kDPerKb;
#pragma unroll
                    for (int d = 0; d < kDPerKb; ++d)
                        zn[d] = (live && d0 + d < p.D) ? zb[(size_t)(d0 + d) * p.T] : 0.f;
                }
                const uint32_t s = it % kTcStages, ph = (it / kTcStages) & 1u;
                mbar_wait(&empty[s], ph ^ 1u);
                unsigned char *a_hi = smem + s * kStageBytes;
                unsigned char *a_lo = a_hi + kAPart;
#pragma unroll
                for (int c = 0; c < 8; ++c) {  // chunk c: k in [8c, 8c+8); c < 4: -0.5 z^2, c >= 4: z
                    uint32_t hi[4], lo[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int d = (c & 3) * 8 + j * 2;
                        float x0 = zc[d], x1 = zc[d + 1];
                        if (c < 4) {
                            x0 = -0.5f * (x0 * x0);
                            x1 = -0.5f * (x1 * x1);
                        }
                        split2(x0, x1, hi[j], lo[j]);
                    }
                    const uint32_t off = sw128_offset(row, c * 8);
                    *reinterpret_cast<uint4 *>(a_hi + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                    *reinterpret_cast<uint4 *>(a_lo + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
                }
                fence_proxy_async();  // generic-proxy stores -> visible to the tensor core (async proxy)
                __syncwarp();
                if (lane == 0) mbar_arrive(&full[s]);
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, 512);
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
bool cost_tc_supported(int B, int D, int T, int S)
{
    (void)B;
    (void)T;
    return S <= kNMax && D >= 1;
}

size_t cost_tc_workspace_bytes(int B, int D, int T, int S)
{
    (void)T;
    const int n_kb = (D + kDPerKb - 1) / kDPerKb;
    return align_up((size_t)B * n_kb * 2 * kBPart, 256) + align_up((size_t)B * S * 4, 256);
}

int cost_tc_launch(const float *z_p, const float *m_p, const float *logs_p, float *neg_cent_out, double *stats_out,
                   const int32_t *t_ys, void *workspace, size_t workspace_bytes, int B, int D, int T, int S,
                   cudaStream_t stream)
{
    if (!workspace || workspace_bytes < cost_tc_workspace_bytes(B, D, T, S)) return MAS_ERR_WORKSPACE;
    const int n_kb = (D + kDPerKb - 1) / kDPerKb;
    unsigned char *images = static_cast<unsigned char *>(workspace);
    float *bias = reinterpret_cast<float *>(images + align_up((size_t)B * n_kb * 2 * kBPart, 256));
    if (stats_out) MAS_CUDA_TRY(cudaMemsetAsync(stats_out, 0, 2 * sizeof(double), stream));
    mas_prior_images_kernel<<<dim3(kNMax / 128, B), 128, 0, stream>>>(m_p, logs_p, images, bias, D, S, n_kb);
    note_launch();
    MAS_CUDA_TRY(cudaGetLastError());

    TcParams p{};
    p.z_p = z_p;
    p.images = images;
    p.bias = bias;
    p.out = neg_cent_out;
    p.stats = stats_out;
    p.t_ys = stats_out ? nullptr : t_ys;  // the noise std covers every cell, padding included (models.py:1243)
    p.B = B;
    p.D = D;
    p.T = T;
    p.S = S;
    p.n_kb = n_kb;
    p.n_cols = (S + 15) / 16 * 16;
    p.m_tiles = (T + kBM - 1) / kBM;
    static thread_local int configured_dev = -1;
    int dev = 0, sms = 148;
    MAS_CUDA_TRY(cudaGetDevice(&dev));
    if (dev != configured_dev) {
        MAS_CUDA_TRY(cudaFuncSetAttribute(mas_cost_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTcSmem));
        configured_dev = dev;
    }
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int n_tiles = B * p.m_tiles;
    const int grid = n_tiles < sms ? n_tiles : sms;
    mas_cost_tc_kernel<<<grid, kTcThreads, kTcSmem, stream>>>(p);
    note_launch();
    MAS_CUDA_TRY(cudaGetLastError());
    return MAS_OK;
}

}  // namespace mas
