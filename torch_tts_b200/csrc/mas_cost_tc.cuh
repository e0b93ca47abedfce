// mas_cost_tc.cuh -- the neg_cent contraction on the 5th-gen tensor cores (tcgen05 / TMEM),
// as a device-side "role" that both the standalone cost kernel (mas_cost_tc.cu) and the
// fused cost+DP kernel (mas_fused.cu) run.
//
// Reference: vits2/models.py:1226-1239.  With r = exp(-2 logs_p):
//   neg_cent[b,t,s] = bias[b,s] + sum_d (-0.5 z^2)[d,t] r[d,s] + z[d,t] (m r)[d,s]
// i.e. one K = 2D contraction  A[t,k] . B[s,k]  per utterance with
//   A = [-0.5 z^2 | z]   (from z_p, converted on the fly, never materialised in HBM)
//   B = [ r       | m r] (from m_p/logs_p, prepared once per utterance as pre-swizzled images)
//
// Precision: fp32 operands are split into bf16 hi + bf16 lo (x ~ hi + lo, 16 mantissa
// bits) and the product is evaluated as A_hi B_hi + A_lo B_hi + A_hi B_lo with fp32
// accumulation in TMEM: three kind::f16 MMAs per K step, ~1e-6 relative on neg_cent (the
// reference's own fp32 sgemm noise level), at the cost of 3 bf16 passes = 1.5 TF32 passes
// instead of the 3 a 3xTF32 split needs.
//
// CTA layout (12 warps, one CTA per SM, persistent over a static tile list):
//   warp 0      B producer: 1-D TMA bulk copies of the pre-swizzled B images -> smem
//   warp 3      z producer: TMA tensor loads of raw z_p tiles [16 ch x 128 mel] -> smem
//   warp 1      MMA issuer: one thread issues tcgen05.mma, commits to mbarriers
//   warp 2      TMEM allocator
//   warps 4-7   epilogue: tcgen05.ld accumulator -> + bias -> swizzled smem -> TMA tensor store
//               (and the noise statistics); publishes the tile flag in the fused kernel
//   warps 8-11  A converters: raw z (smem) -> -0.5 z^2, z -> bf16 hi/lo -> K-major
//               SWIZZLE_64B operand tiles in smem
// Tile = 128 mel rows x N text columns (N = S rounded up to 16, <= 256); K blocks of 32 bf16
// (16 prior channels): 3 operand stages of 48 KB, 4 raw-z stages of 8 KB, 2 TMEM accumulators.
#pragma once

#include <cuda.h>

#include "mas_common.cuh"

namespace mas {

constexpr int kTcThreads = 384;
constexpr int kBM = 128;            // mel rows per tile (UMMA M)
constexpr int kBK = 32;             // bf16 K elements per block (= one 64-byte swizzle row)
constexpr int kDPerKb = kBK / 2;    // prior channels per K block
constexpr int kNMax = 256;          // text columns per tile (UMMA N max)
constexpr int kTcStages = 3;        // operand stages
constexpr int kZStages = 4;         // raw z stages
constexpr uint32_t kRowBytes = kBK * 2;               // 64
constexpr uint32_t kAPart = kBM * kRowBytes;          // 8 KB: one split part of A per stage
constexpr uint32_t kBPart = kNMax * kRowBytes;        // 16 KB: one split part of B per stage / per image
constexpr uint32_t kStageBytes = 2 * kAPart + 2 * kBPart;   // 48 KB
constexpr uint32_t kZStageBytes = kDPerKb * kBM * 4;        // 8 KB
constexpr uint32_t kEpiBufBytes = 32 * 128;                 // 32 rows x 32 fp32, SWIZZLE_128B
constexpr uint32_t kTcOffZ = kTcStages * kStageBytes;
constexpr uint32_t kTcOffEpi = kTcOffZ + kZStages * kZStageBytes;
constexpr uint32_t kTcOffBias = kTcOffEpi + 8 * kEpiBufBytes;
constexpr uint32_t kTcOffBar = kTcOffBias + kNMax * 4;
constexpr uint32_t kTcSmemUsed = kTcOffBar + 256;
constexpr uint32_t kTcSmem = kTcSmemUsed + 1024 /*alignment slack*/;

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

// K-major SWIZZLE_64B shared-memory matrix descriptor: rows of 64 bytes, 8-row atoms of
// 512 bytes (stride byte offset), version 1 (sm_100), layout type 4.
__device__ __forceinline__ uint64_t make_desc_sw64(uint32_t smem_addr)
{
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;           // leading byte offset (unused for swizzled K-major)
    d |= (uint64_t)(512 >> 4) << 32;  // stride byte offset between 8-row groups
    d |= (uint64_t)1 << 46;           // descriptor version
    d |= (uint64_t)4 << 61;           // SWIZZLE_64B
    return d;
}

// kind::f16 instruction descriptor: D fp32, A/B bf16, both K-major, M = 128, N = n
__host__ __device__ __forceinline__ uint32_t make_idesc_bf16(int n)
{
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(kBM >> 4) << 24);
}

// byte offset of element (row, k) inside a K-major SWIZZLE_64B bf16 tile (row pitch 64 B):
// Swizzle<2,4,3>: 16-byte chunk index ^= address bits [7,9) = (row >> 1) & 3
__host__ __device__ __forceinline__ uint32_t sw64_offset(int row, int k)
{
    return (uint32_t)(row * 64 + ((((k >> 3) ^ (row >> 1)) & 3) << 4) + (k & 7) * 2);
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

struct TcParams {
    const float *z_p;
    const unsigned char *images;   // [B][n_kb][2 parts][kBPart]
    const float *bias_part;        // [B][n_kb][S]
    float *out;                    // [B][T][S]
    double *stats;                 // nullable
    const int32_t *t_ys;           // nullable: skip mel tiles entirely past t_y (no noise statistics then)
    uint32_t *flags;               // nullable: [B][m_tiles], set to 1 (release) when a tile is in memory
    int B, D, T, S;
    int n_kb;                      // K blocks = ceil(D / 16)
    int n_cols;                    // UMMA N = S rounded up to 16
    int m_tiles;                   // ceil(T / 128)
    int wave;                      // tile order: utterances in groups of `wave`, mel-tile-major inside a group
    int z_tma, out_tma;            // tensor maps usable (T % 4 == 0 / S % 4 == 0)
    int debug;                     // MAS_TC_DEBUG bit mask (profiling experiments): 1 no A stores, 2 no epilogue stores, 4 no MMA
};

// tile order index -> (b, mt); utterance groups of p.wave, mel-tile-major inside a group, so that in
// the fused kernel every utterance of a group receives its first tiles early
__device__ __forceinline__ void tc_tile_coords(const TcParams &p, int i, int &b, int &mt)
{
    const int per_wave = p.wave * p.m_tiles;
    const int w = i / per_wave;
    const int r = i - w * per_wave;
    const int base = w * p.wave;
    const int wc = min(p.wave, p.B - base);
    mt = r / wc;
    b = base + (r - mt * wc);
}

__device__ __forceinline__ bool tc_tile_live(const TcParams &p, int b, int mt)
{
    if (!p.t_ys) return true;
    const int t_y = p.t_ys[b];
    return t_y >= 1 && t_y <= p.T && mt * kBM < t_y;
}

// The contraction role.  Runs on all kTcThreads threads of the CTA (dynamic smem `smem_raw`,
// at least kTcSmem bytes); processes tiles first, first + step, ... of the tile order.
template <bool kStats>
__device__ __forceinline__ void cost_tc_role(const TcParams &p, const CUtensorMap *tm_z, const CUtensorMap *tm_out,
                                             unsigned char *smem_raw, int first, int step)
{
    const uint32_t raw = smem_u32(smem_raw);
    unsigned char *smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);  // swizzled tiles need 1024-byte alignment
    float *bias_s = reinterpret_cast<float *>(smem + kTcOffBias);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + kTcOffBar);
    uint64_t *full = bars;                 // [3] operand stage filled: B bytes landed + 4 converter warps arrived
    uint64_t *empty = bars + 3;            // [3] operand stage consumed by the MMAs
    uint64_t *zfull = bars + 6;            // [4] raw z stage landed
    uint64_t *zempty = bars + 10;          // [4] raw z stage read by the 4 converter warps
    uint64_t *acc_full = bars + 14;        // [2] accumulator complete
    uint64_t *acc_empty = bars + 16;       // [2] accumulator drained by the epilogue
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 18);

    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    const int lane = tid & 31;
    const int n_tiles = p.B * p.m_tiles;

    if (tid == 0) {
        for (int i = 0; i < kTcStages; ++i) {
            mbar_init(&full[i], 1 + 4);
            mbar_init(&empty[i], 1);
        }
        for (int i = 0; i < kZStages; ++i) {
            mbar_init(&zfull[i], 1);
            mbar_init(&zempty[i], 4);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&acc_full[i], 1);
            mbar_init(&acc_empty[i], 4);
        }
        fence_mbar_init();
        if (p.z_tma) tma_prefetch_desc(tm_z);
        if (p.out_tma) tma_prefetch_desc(tm_out);
    }
    if (warp == 2) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    bar_sync(1, kTcThreads);
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ======================= producer =======================
        if (lane == 0) {
            uint32_t it = 0;
            const uint32_t b_bytes = (uint32_t)p.n_cols * kRowBytes;
            for (int i = first; i < n_tiles; i += step) {
                int b, mt;
                tc_tile_coords(p, i, b, mt);
                if (!tc_tile_live(p, b, mt)) continue;
                for (int kb = 0; kb < p.n_kb; ++kb, ++it) {
                    const uint32_t s = it % kTcStages, ph = (it / kTcStages) & 1u;
                    mbar_wait(&empty[s], ph ^ 1u);
                    unsigned char *stage = smem + s * kStageBytes;
                    const unsigned char *img = p.images + (size_t)(b * p.n_kb + kb) * 2 * kBPart;
                    mbar_arrive_expect_tx(&full[s], 2 * b_bytes);
                    bulk_g2s(stage + 2 * kAPart, img, b_bytes, &full[s]);
                    bulk_g2s(stage + 2 * kAPart + kBPart, img + kBPart, b_bytes, &full[s]);
                }
            }
        }
    } else if (warp == 3) {
        // ======================= raw z producer (runs ahead of the operand ring) =======================
        if (lane == 0 && p.z_tma) {
            uint32_t it = 0;
            for (int i = first; i < n_tiles; i += step) {
                int b, mt;
                tc_tile_coords(p, i, b, mt);
                if (!tc_tile_live(p, b, mt)) continue;
                for (int kb = 0; kb < p.n_kb; ++kb, ++it) {
                    const uint32_t zs = it % kZStages, zph = (it / kZStages) & 1u;
                    mbar_wait(&zempty[zs], zph ^ 1u);
                    mbar_arrive_expect_tx(&zfull[zs], kZStageBytes);
                    tma_load_3d(smem + kTcOffZ + zs * kZStageBytes, tm_z, mt * kBM, kb * kDPerKb, b, &zfull[zs]);
                }
            }
        }
    } else if (warp == 1) {
        // ======================= MMA issuer =======================
        if (lane == 0) {
            const uint32_t idesc = make_idesc_bf16(p.n_cols);
            uint32_t it = 0, nt = 0;
            for (int i = first; i < n_tiles; i += step) {
                int b, mt;
                tc_tile_coords(p, i, b, mt);
                if (!tc_tile_live(p, b, mt)) continue;
                const uint32_t a = nt & 1u, aph = (nt >> 1) & 1u;
                mbar_wait(&acc_empty[a], aph ^ 1u);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + a * kNMax;
                for (int kb = 0; kb < p.n_kb; ++kb, ++it) {
                    const uint32_t s = it % kTcStages, ph = (it / kTcStages) & 1u;
                    mbar_wait(&full[s], ph);
                    tc_fence_after();
                    const uint32_t st = smem_u32(smem + s * kStageBytes);
                    const uint64_t a_hi = make_desc_sw64(st), a_lo = make_desc_sw64(st + kAPart);
                    const uint64_t b_hi = make_desc_sw64(st + 2 * kAPart), b_lo = make_desc_sw64(st + 2 * kAPart + kBPart);
#pragma unroll
                    for (int k = 0; k < kBK / 16 && !(p.debug & 4); ++k) {
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
        unsigned char *ebuf = smem + kTcOffEpi + wq * 2 * kEpiBufBytes;
        uint32_t nt = 0, nst = 0;
        double ssum = 0.0, ssq = 0.0;
        for (int i = first; i < n_tiles; i += step) {
            int b, mt;
            tc_tile_coords(p, i, b, mt);
            if (!tc_tile_live(p, b, mt)) continue;
            // bias of this utterance -> smem (only the 4 epilogue warps sync here)
            bar_sync(2, 128);
            for (int s = tid - 128; s < p.n_cols; s += 128) {
                float acc = 0.f;
                if (s < p.S)
                    for (int kb = 0; kb < p.n_kb; ++kb) acc += p.bias_part[((size_t)b * p.n_kb + kb) * p.S + s];
                bias_s[s] = acc;
            }
            bar_sync(2, 128);
            const uint32_t a = nt & 1u, aph = (nt >> 1) & 1u;
            mbar_wait(&acc_full[a], aph);
            tc_fence_after();
            const int t = mt * kBM + row;
            float *orow = p.out + ((size_t)b * p.T + t) * p.S;
            const uint32_t taddr = tmem_base + a * kNMax + ((uint32_t)(wq * 32) << 16);
            for (int c0 = 0; c0 < p.n_cols; c0 += 32) {
                uint32_t r[32];
                tmem_ld32(taddr + c0, r);
                tmem_ld_wait();
                float v[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]) + bias_s[c0 + j];
                if (kStats) {
                    if (t < p.T) {
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            if (c0 + j < p.S) {
                                ssum += (double)v[j];
                                ssq += (double)v[j] * (double)v[j];
                            }
                    }
                }
                if (p.debug & 2) continue;
                if (p.out_tma) {
                    unsigned char *buf = ebuf + (nst & 1u) * kEpiBufBytes;
                    if (lane == 0) bulk_wait_read<1>();  // the store that last read this buffer is done with it
                    __syncwarp();
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        *reinterpret_cast<float4 *>(buf + lane * 128 + ((j ^ (lane & 7)) << 4)) =
                            make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                    fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) {
                        tma_store_3d(tm_out, c0, mt * kBM + wq * 32, b, buf);  // rows >= T / cols >= S are clipped
                        bulk_commit();
                    }
                    ++nst;
                } else if (t < p.T) {
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (c0 + j < p.S) orow[c0 + j] = v[j];
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[a]);
            ++nt;
            if (p.flags) {
                // publish the tile: all stores of the 4 epilogue warps are complete, then release
                if (p.out_tma) {
                    if (lane == 0) bulk_wait_all();
                } else {
                    __threadfence();
                }
                bar_sync(2, 128);
                if (tid == 128) {
                    fence_proxy_async_all();
                    __threadfence();
                    st_release_gpu(p.flags + (size_t)b * p.m_tiles + mt, 1u);
                }
            }
        }
        if (p.out_tma && lane == 0) bulk_wait_all();
        if (kStats) {
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
        for (int i = first; i < n_tiles; i += step) {
            int b, mt;
            tc_tile_coords(p, i, b, mt);
            if (!tc_tile_live(p, b, mt)) continue;
            const int t = mt * kBM + row;
            const bool live = t < p.T;
            const float *zb = p.z_p + (size_t)b * p.D * p.T + (live ? t : 0);
            for (int kb = 0; kb < p.n_kb; ++kb, ++it) {
                float zc[kDPerKb];
                if (p.z_tma) {
                    const uint32_t zs = it % kZStages, zph = (it / kZStages) & 1u;
                    mbar_wait(&zfull[zs], zph);
                    const float *zr = reinterpret_cast<const float *>(smem + kTcOffZ + zs * kZStageBytes) + row;
#pragma unroll
                    for (int d = 0; d < kDPerKb; ++d) zc[d] = zr[d * kBM];  // zero-filled past T / D by the TMA
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&zempty[zs]);
                } else {
                    const int d0 = kb * kDPerKb;
#pragma unroll
                    for (int d = 0; d < kDPerKb; ++d) zc[d] = (live && d0 + d < p.D) ? zb[(size_t)(d0 + d) * p.T] : 0.f;
                }
                const uint32_t s = it % kTcStages, ph = (it / kTcStages) & 1u;
                mbar_wait(&empty[s], ph ^ 1u);
                unsigned char *a_hi = smem + s * kStageBytes;
                unsigned char *a_lo = a_hi + kAPart;
#pragma unroll
                for (int c = 0; c < 4; ++c) {  // chunk c: k in [8c, 8c+8); c < 2: -0.5 z^2, c >= 2: z
                    uint32_t hi[4], lo[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int d = (c & 1) * 8 + j * 2;
                        float x0 = zc[d], x1 = zc[d + 1];
                        if (c < 2) {
                            x0 = -0.5f * (x0 * x0);
                            x1 = -0.5f * (x1 * x1);
                        }
                        split2(x0, x1, hi[j], lo[j]);
                    }
                    const uint32_t off = sw64_offset(row, c * 8);
                    if (!(p.debug & 1)) {
                        *reinterpret_cast<uint4 *>(a_hi + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                        *reinterpret_cast<uint4 *>(a_lo + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
                    }
                }
                fence_proxy_async();  // generic-proxy stores -> visible to the tensor core (async proxy)
                __syncwarp();
                if (lane == 0) mbar_arrive(&full[s]);
            }
        }
    }

    tc_fence_before();
    bar_sync(1, kTcThreads);
    if (warp == 2) tmem_dealloc(tmem_base, 512);
}

// host side (mas_cost_tc.cu)
struct TcPlan {
    TcParams p;
    CUtensorMap tm_z, tm_out;
};
bool cost_tc_supported(int B, int D, int T, int S);
size_t cost_tc_workspace_bytes(int B, int D, int T, int S);
// launches the prior-image preparation (which also zeroes `flags_to_clear`, n_flags words, if given)
// and fills `plan` for the contraction
int cost_tc_prepare(TcPlan &plan, const float *z_p, const float *m_p, const float *logs_p, float *neg_cent_out,
                    double *stats_out, const int32_t *t_ys, void *workspace, size_t workspace_bytes, int B, int D, int T,
                    int S, uint32_t *flags_to_clear, int n_flags, cudaStream_t stream);

}  // namespace mas
