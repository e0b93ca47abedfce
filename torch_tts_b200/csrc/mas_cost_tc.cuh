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
// CTA layout (16 warps, one CTA per SM, persistent over a static tile list):
//   warp 0      B producer: 1-D TMA bulk copies of the pre-swizzled B images -> smem
//   warp 3      z producer: TMA tensor loads of raw z_p tiles [16 ch x 128 mel] -> smem
//   warp 1      MMA issuer: one thread issues tcgen05.mma, commits to mbarriers
//   warp 2      TMEM allocator
//   warps 4-7   epilogue: tcgen05.ld accumulator -> + bias -> swizzled smem -> TMA tensor store
//               (and the noise statistics); publishes the tile flag in the fused kernel
//   warps 8-15  A converters, two groups of 4 warps taking alternate K blocks (a K block is a
//               latency chain: wait -> LDS -> split -> STS -> fence -> arrive): raw z (smem) ->
//               -0.5 z^2, z -> bf16 hi/lo -> K-major SWIZZLE_64B operand tiles in smem
// Tile = 128 mel rows x N text columns (N = S rounded up to 16, <= 256); K blocks of 32 bf16
// (16 prior channels): 3 operand stages of 48 KB, 4 raw-z stages of 8 KB, 2 TMEM accumulators.
#pragma once

#include <cuda.h>

#include "mas_common.cuh"

namespace mas {

constexpr int kTcThreads = 512;
constexpr int kBM = 128;            // mel rows per tile (UMMA M)
constexpr int kBK = 32;             // bf16 K elements per block (= one 64-byte swizzle row)
constexpr int kDPerKb = kBK / 2;    // prior channels per K block
constexpr int kNMax = 256;          // text columns per tile (UMMA N max)
constexpr uint32_t kRowBytes = kBK * 2;               // 64
constexpr uint32_t kAPart = kBM * kRowBytes;          // 8 KB: one split part of A per stage
constexpr uint32_t kBPart = kNMax * kRowBytes;        // 16 KB: one split part of B per image
constexpr uint32_t kZStageBytes = kDPerKb * kBM * 4;        // 8 KB
constexpr uint32_t kEpiBufBytes = 32 * 128;                 // 32 rows x 32 fp32, SWIZZLE_128B

// shared-memory layout of the role.  kPair: two CTAs of a cluster run one M = 256 tcgen05.mma
// (cta_group::2); each holds its own 128 A rows and HALF of the B rows, so the B bytes every SM
// pulls out of L2 -- the bound of the single-CTA version -- are halved.
template <bool kPair>
struct TcCfg {
    static constexpr int kStages = kPair ? 4 : 3;                          // operand stages
    static constexpr int kZStages = kPair ? 6 : 5;                         // raw z stages (bytes in flight towards HBM)
    static constexpr uint32_t kBPartS = kPair ? kBPart / 2 : kBPart;       // bytes of one B part in a stage
    static constexpr uint32_t kStage = 2 * kAPart + 2 * kBPartS;           // 32 KB / 48 KB
    static constexpr uint32_t kOffZ = kStages * kStage;
    static constexpr uint32_t kOffEpi = kOffZ + kZStages * kZStageBytes;
    static constexpr uint32_t kOffBias = kOffEpi + 8 * kEpiBufBytes;
    static constexpr uint32_t kOffBar = kOffBias + 2 * kNMax * 4;  // two bias buffers
    static constexpr uint32_t kSmem = kOffBar + 512 + 1024 /*alignment slack*/;
};
constexpr uint32_t kTcSmem = TcCfg<false>::kSmem > TcCfg<true>::kSmem ? TcCfg<false>::kSmem : TcCfg<true>::kSmem;

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
// ---- cta_group::2 (CTA pair) variants -------------------------------------------
__device__ __forceinline__ void tmem_alloc2(uint32_t *dst_smem, uint32_t cols)
{
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t addr, uint32_t cols)
{
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
__device__ __forceinline__ void umma_bf16_2cta(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                               uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrive on the mbarrier at the same offset in BOTH CTAs of the pair once the MMAs issued so far are done
__device__ __forceinline__ void umma_commit_2cta(uint64_t *bar)
{
    asm volatile(
        "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
        ::"r"(smem_u32(bar)), "h"((uint16_t)3)
        : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank()
{
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all()
{
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `local` (a shared::cta address) in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_u32(uint32_t local, uint32_t rank)
{
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr)
{
    // default semantics (.release.cta): an explicit .release.cluster costs a MEMBAR.ALL.GPU per arrive.  The
    // data this orders stays in the arriving SM's own shared memory (its tensor core reads it), as in
    // CUTLASS's ClusterBarrier::arrive(cta_id).
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// wait on a local mbarrier that is also arrived on by the peer CTA
__device__ __forceinline__ void mbar_wait_cluster(uint64_t *bar, uint32_t parity)
{
    uint32_t ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!ok);
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

// kind::f16 instruction descriptor: D fp32, A/B bf16, both K-major, M = m (128, or 256 for a CTA pair), N = n
__host__ __device__ __forceinline__ uint32_t make_idesc_bf16(int m, int n)
{
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
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

// packed fp32 pairs (Blackwell f32x2 arithmetic: one instruction per two cells, each half rounded like the scalar op)
__device__ __forceinline__ uint64_t f2_pack(float a, float b)
{
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ void f2_unpack(uint64_t v, float &a, float &b)
{
    asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
}
__device__ __forceinline__ uint64_t f2_add(uint64_t a, uint64_t b)
{
    uint64_t r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ uint64_t f2_fma(uint64_t a, uint64_t b, uint64_t c)
{
    uint64_t r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}

struct TcParams {
    const float *z_p;
    const unsigned char *images;   // [B][n_blocks][n_kb][2 parts][kBPart]
    const float *bias_part;        // [B][n_kb][S]
    float *out;                    // [B][T][ld]
    int ld;                        // row stride of `out` in floats: S, or S rounded up to 4 (private plane; pad columns get 0)
    double *stats;                 // nullable
    const int32_t *t_ys;           // nullable: skip mel tiles entirely past t_y (no noise statistics then)
    uint32_t *flags;               // nullable: [B][m_tiles], set to 1 (release) when a tile is in memory
    int B, D, T, S;
    int n_kb;                      // K blocks = ceil(D / 16)
    int n_blocks;                  // column blocks of kNMax text columns per utterance (1 for S <= 256)
    int m_tiles;                   // ceil(T / 128)
    int wave;                      // tile order: utterances in groups of `wave`, mel-tile-major inside a group
    int pdl;                       // launched as a programmatic dependent of the prior-images kernel
    int seq_k, seq_pure0;          // unit schedule: see unit_index() in cost_tc_role (standalone: seq_k huge)
    int z_tma, out_tma;            // tensor maps usable (T % 4 == 0 / S % 4 == 0)
    unsigned long long *trace;     // nullable diagnostics buffer: [cta][64] publication times
    int debug;                     // MAS_TC_DEBUG bit mask (profiling experiments): 1 no A stores, 2 no epilogue stores, 4 no MMA,
                                   // 32 no z loads, 64 no B loads (8 / 16 on the host: prior only / contraction only)
};

// work-unit order index -> (b, mu); utterance groups of p.wave, mel-major inside a group, so that in
// the fused kernel every utterance of a group receives its first tiles early.  A unit is one mel
// tile (single CTA) or a pair of adjacent mel tiles (CTA pair): m_units per utterance.
__device__ __forceinline__ void tc_unit_coords(const TcParams &p, int m_units, int i, int &b, int &mu)
{
    const int per_wave = p.wave * m_units;
    const int w = i / per_wave;
    const int r = i - w * per_wave;
    const int base = w * p.wave;
    const int wc = min(p.wave, p.B - base);
    mu = r / wc;
    b = base + (r - mu * wc);
}

__device__ __forceinline__ bool tc_tile_live(const TcParams &p, int b, int mt)
{
    if (!p.t_ys) return true;
    const int t_y = p.t_ys[b];
    return t_y >= 1 && t_y <= p.T && mt * kBM < t_y;
}

// The contraction role.  Runs on all kTcThreads threads of the CTA (dynamic smem `smem_raw`,
// at least kTcSmem bytes); processes work units first, first + step, ... of the unit order.
// kPair: the CTA is one of a 2-CTA cluster; `first`/`step` count pairs; rank r of the pair owns mel
// tile 2 * mu + r, holds rows [r * N/2, (r+1) * N/2) of every B stage, and rank 0 issues the MMAs.
template <bool kStats, bool kPair>
__device__ __forceinline__ void cost_tc_role(const TcParams &p, const CUtensorMap *tm_z, const CUtensorMap *tm_out,
                                             unsigned char *smem_raw, int first, int step)
{
    using Cfg = TcCfg<kPair>;
    constexpr int kStages = Cfg::kStages;
    constexpr int kZStages = Cfg::kZStages;
    const uint32_t raw = smem_u32(smem_raw);
    unsigned char *smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);  // swizzled tiles need 1024-byte alignment
    float *bias_s = reinterpret_cast<float *>(smem + Cfg::kOffBias);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + Cfg::kOffBar);
    uint64_t *full = bars;                 // [4] operand stage filled (pair: rank 0's copy collects both CTAs)
    uint64_t *empty = bars + 4;            // [4] operand stage consumed by the MMAs
    uint64_t *bfull = bars + 8;            // [4] pair, rank 1: own B bytes landed (forwarded to rank 0's full)
    uint64_t *zfull = bars + 12;           // [8] raw z stage landed
    uint64_t *zempty = bars + 20;          // [8] raw z stage read by the 4 converter warps of a group
    uint64_t *acc_full = bars + 28;        // [2] accumulator complete
    uint64_t *acc_empty = bars + 30;       // [2] accumulator drained by the epilogue (pair: of both CTAs)
    uint64_t *bias_full = bars + 32;       // [2] bias buffer written by warp 2
    uint64_t *bias_empty = bars + 34;      // [2] bias buffer released by the 4 epilogue warps
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 36);
    uint32_t *pub_cnt = tmem_slot + 1;     // [4] epilogue warps whose stores of a unit have completed, slot = unit number % 4:
                                           // a fast epilogue warp can be publishing unit k + 2 while a slow one has not yet
                                           // published unit k (they release the accumulator before they publish), so two
                                           // slots alias -- seen with the 16-column units of S = 257, whose tiles went out
                                           // with a quarter of their rows still in flight

    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    const int lane = tid & 31;
    const uint32_t rank = kPair ? cluster_ctarank() : 0u;
    const int m_units = kPair ? (p.m_tiles + 1) / 2 : p.m_tiles;
    const int n_units = p.B * m_units * p.n_blocks;

    if (tid == 0) {
        for (int i = 0; i < kStages; ++i) {
            mbar_init(&full[i], kPair ? 10 : 5);  // B producer + 4 converter warps (+ the peer's 4 + its forwarder)
            mbar_init(&empty[i], 1);
            mbar_init(&bfull[i], 1);
        }
        for (int i = 0; i < kZStages; ++i) {
            mbar_init(&zfull[i], 1);
            mbar_init(&zempty[i], 4);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&acc_full[i], 1);
            mbar_init(&acc_empty[i], kPair ? 8 : 4);
            mbar_init(&bias_full[i], 1);
            mbar_init(&bias_empty[i], 4);
            pub_cnt[i] = 0u;
            pub_cnt[i + 2] = 0u;
        }
        fence_mbar_init();
        if (p.z_tma) tma_prefetch_desc(tm_z);
        if (p.out_tma) tma_prefetch_desc(tm_out);
    }
    if (warp == 2) {
        if (kPair)
            tmem_alloc2(tmem_slot, 512);
        else
            tmem_alloc(tmem_slot, 512);
    }
    tc_fence_before();
    bar_sync(1, kTcThreads);
    if (kPair) cluster_sync_all();  // the peer's barriers exist before anything arrives on them
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // Programmatic dependent launch: this grid may have started while the kernel that writes the operand
    // images, the bias partial sums and the cleared flags was still running; everything above overlapped
    // with it, everything below reads its output.  (Without such a launch the wait returns at once.)
    // Only the roles that touch that kernel's output wait: the B producer (images), the bias warp (partial sums) and
    // the epilogue (tile flags, cleared there).  The raw-z producer and the A converters start on the first K blocks
    // right away, so the operand ring is already filling when the images arrive.  (Every thread of the CTA passes
    // the role's final barrier after the B producer's wait has returned, i.e. after that grid has completed.)
    if (p.pdl && (warp == 0 || warp == 2 || (warp >= 4 && warp < 8))) {
        asm volatile("griddepcontrol.wait;" ::: "memory");
        fence_proxy_async_all();  // the images are read through the async proxy (bulk copies)
    }

    // n-th work unit of this CTA (pair), or -1 when it has none left.  Rounds 0 .. seq_k-1 stride all
    // CTAs (pairs); from round seq_k on only the "pure" contraction CTAs (first >= seq_pure0) continue and
    // share the remaining units among themselves -- the others move on to the DP role (fused kernel).
    auto unit_index = [&](int n) {
        int i;
        if (n < p.seq_k)
            i = first + n * step;
        else if (first >= p.seq_pure0)
            i = p.seq_k * step + (first - p.seq_pure0) + (n - p.seq_k) * (step - p.seq_pure0);
        else
            return -1;
        return i < n_units ? i : -1;
    };
    // this CTA's mel tile of unit (b, mu), and whether the unit is processed at all
    // (a unit is one column block nb of one mel tile / tile pair; S <= 256 has a single block)
    auto unit_tile = [&](int i, int &b, int &mt, int &nb) {
        int mu;
        nb = i % p.n_blocks;
        tc_unit_coords(p, m_units, i / p.n_blocks, b, mu);
        mt = kPair ? 2 * mu + (int)rank : mu;
        return tc_tile_live(p, b, kPair ? 2 * mu : mu);
    };
    // text columns of block nb, rounded up to the UMMA N granularity
    auto block_cols = [&](int nb) {
        const int left = p.S - nb * kNMax;
        return ((left < kNMax ? left : kNMax) + 15) & ~15;
    };
    // diagnostics: trace[16384 + cta * 64 + role * 16 + 2 * unit + {0: begin, 1: end}], roles 0 MMA, 1 epilogue, 2 converter, 3 z
    auto tr_mark = [&](int role, uint32_t unit, int which) {
        if (MAS_TR(p) && unit < 8) p.trace[16384 + (size_t)blockIdx.x * 64 + role * 16 + 2 * unit + which] = globaltimer_ns();
    };

    if (warp == 0) {
        // ======================= B producer =======================
        if (lane == 0) {
            uint32_t it = 0;
            for (int un = 0, i = unit_index(0); i >= 0; i = unit_index(++un)) {
                int b, mt, nb;
                if (!unit_tile(i, b, mt, nb)) continue;
                for (int kb = 0; kb < p.n_kb; ++kb, ++it) {
                    const uint32_t s = it % kStages, ph = (it / kStages) & 1u;
                    mbar_wait(&empty[s], ph ^ 1u);
                    unsigned char *stage = smem + s * Cfg::kStage;
                    // rows of the B image this CTA stages (a pair: half each), and their bytes
                    const uint32_t b_bytes = (uint32_t)block_cols(nb) / (kPair ? 2u : 1u) * kRowBytes;
                    const unsigned char *img = p.images + ((size_t)(b * p.n_blocks + nb) * p.n_kb + kb) * 2 * kBPart +
                                               (kPair ? rank * b_bytes : 0u);
                    uint64_t *bar = (kPair && rank) ? &bfull[s] : &full[s];
                    if MAS_DBG(p, 64) {  // experiment: no B traffic
                        mbar_arrive(bar);
                        continue;
                    }
                    mbar_arrive_expect_tx(bar, 2 * b_bytes);
                    bulk_g2s(stage + 2 * kAPart, img, b_bytes, bar);
                    bulk_g2s(stage + 2 * kAPart + Cfg::kBPartS, img + kBPart, b_bytes, bar);
                }
            }
        }
    } else if (warp == 3) {
        // ======================= raw z producer (runs ahead of the operand ring) =======================
        if (lane == 0 && p.z_tma && !MAS_DBG(p, 32)) {
            uint32_t it = 0;
            for (int un = 0, i = unit_index(0); i >= 0; i = unit_index(++un)) {
                int b, mt, nb;
                if (!unit_tile(i, b, mt, nb)) continue;
                for (int kb = 0; kb < p.n_kb; ++kb, ++it) {
                    const uint32_t zs = it % kZStages, zph = (it / kZStages) & 1u;
                    mbar_wait(&zempty[zs], zph ^ 1u);
                    mbar_arrive_expect_tx(&zfull[zs], kZStageBytes);
                    tma_load_3d(smem + Cfg::kOffZ + zs * kZStageBytes, tm_z, mt * kBM, kb * kDPerKb, b, &zfull[zs]);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && rank == 0) {
            // ======================= MMA issuer =======================
            uint32_t it = 0, nt = 0;
            for (int un = 0, i = unit_index(0); i >= 0; i = unit_index(++un)) {
                int b, mt, nb;
                if (!unit_tile(i, b, mt, nb)) continue;
                const uint32_t a = nt & 1u, aph = (nt >> 1) & 1u;
                if (kPair)
                    mbar_wait_cluster(&acc_empty[a], aph ^ 1u);
                else
                    mbar_wait(&acc_empty[a], aph ^ 1u);
                tc_fence_after();
                tr_mark(0, nt, 0);
                const uint32_t idesc = make_idesc_bf16(kPair ? 2 * kBM : kBM, block_cols(nb));
                const uint32_t tmem_d = tmem_base + a * kNMax;
                for (int kb = 0; kb < p.n_kb; ++kb, ++it) {
                    const uint32_t s = it % kStages, ph = (it / kStages) & 1u;
                    if (kPair)
                        mbar_wait_cluster(&full[s], ph);
                    else
                        mbar_wait(&full[s], ph);
                    tc_fence_after();
                    const uint32_t st = smem_u32(smem + s * Cfg::kStage);
                    const uint64_t a_hi = make_desc_sw64(st), a_lo = make_desc_sw64(st + kAPart);
                    const uint64_t b_hi = make_desc_sw64(st + 2 * kAPart),
                                   b_lo = make_desc_sw64(st + 2 * kAPart + Cfg::kBPartS);
#pragma unroll
                    for (int k = 0; k < kBK / 16 && !MAS_DBG(p, 4); ++k) {
                        const uint64_t adv = (uint64_t)((k * 32) >> 4);  // 16 bf16 = 32 bytes along K
                        if (kPair) {
                            umma_bf16_2cta(tmem_d, a_hi + adv, b_hi + adv, idesc, (kb | k) ? 1u : 0u);
                            umma_bf16_2cta(tmem_d, a_lo + adv, b_hi + adv, idesc, 1u);
                            umma_bf16_2cta(tmem_d, a_hi + adv, b_lo + adv, idesc, 1u);
                        } else {
                            umma_bf16(tmem_d, a_hi + adv, b_hi + adv, idesc, (kb | k) ? 1u : 0u);
                            umma_bf16(tmem_d, a_lo + adv, b_hi + adv, idesc, 1u);
                            umma_bf16(tmem_d, a_hi + adv, b_lo + adv, idesc, 1u);
                        }
                    }
                    // frees the stage (in both CTAs of a pair) when these MMAs have read it
                    if (kPair)
                        umma_commit_2cta(&empty[s]);
                    else
                        umma_commit(&empty[s]);
                }
                if (kPair)
                    umma_commit_2cta(&acc_full[a]);
                else
                    umma_commit(&acc_full[a]);
                tr_mark(0, nt, 1);
                ++nt;
            }
        } else if (kPair && lane == 0 && rank == 1) {
            // ======================= forwarder: own B half landed -> rank 0's full barrier =======================
            uint32_t it = 0;
            for (int un = 0, i = unit_index(0); i >= 0; i = unit_index(++un)) {
                int b, mt, nb;
                if (!unit_tile(i, b, mt, nb)) continue;
                for (int kb = 0; kb < p.n_kb; ++kb, ++it) {
                    const uint32_t s = it % kStages, ph = (it / kStages) & 1u;
                    mbar_wait(&bfull[s], ph);
                    mbar_arrive_cluster(mapa_u32(smem_u32(&full[s]), 0));
                }
            }
        }
    } else if (warp == 2) {
        // ======================= bias: sum of the K-block partials of the unit's utterance =======================
        uint32_t n = 0;
        for (int un = 0, i = unit_index(0); i >= 0; i = unit_index(++un)) {
            int b, mt, nb;
            if (!unit_tile(i, b, mt, nb)) continue;
            const uint32_t buf = n & 1u, ph = (n >> 1) & 1u;
            mbar_wait(&bias_empty[buf], ph ^ 1u);
            const float *bp = p.bias_part + (size_t)b * p.n_kb * p.S + nb * kNMax;
            const int ncols = block_cols(nb), s_left = p.S - nb * kNMax;
            if ((p.S & 3) == 0) {
                // 4 columns per lane and pass, up to 12 partials in flight per lane; summed in K-block order
                for (int s = 4 * lane; s < ncols; s += 128) {
                    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (s < s_left) {
                        for (int kb0 = 0; kb0 < p.n_kb; kb0 += 12) {
                            float4 part[12];
#pragma unroll
                            for (int j = 0; j < 12; ++j)
                                part[j] = (kb0 + j < p.n_kb)
                                              ? __ldg(reinterpret_cast<const float4 *>(bp + (size_t)(kb0 + j) * p.S + s))
                                              : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                            for (int j = 0; j < 12; ++j) {
                                acc.x += part[j].x, acc.y += part[j].y, acc.z += part[j].z, acc.w += part[j].w;
                            }
                        }
                    }
                    *reinterpret_cast<float4 *>(bias_s + buf * kNMax + s) = acc;
                }
            } else {
                for (int s = lane; s < ncols; s += 32) {
                    float acc = 0.f;
                    if (s < s_left)
                        for (int kb = 0; kb < p.n_kb; ++kb) acc += bp[(size_t)kb * p.S + s];
                    bias_s[buf * kNMax + s] = acc;
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&bias_full[buf]);
            ++n;
        }
    } else if (warp >= 4 && warp < 8) {
        // ======================= epilogue =======================
        const int wq = warp & 3;  // TMEM lane quarter this warp may read
        const int row = wq * 32 + lane;
        unsigned char *ebuf = smem + Cfg::kOffEpi + wq * 2 * kEpiBufBytes;
        uint32_t nt = 0, nst = 0;
        double ssum = 0.0, ssq = 0.0;
        // tile flag (fused kernel): goes out once all four warps' stores of the unit have completed
        // (MAS_TC_DEBUG=64: lazily, checked after the first store of the next unit)
        uint32_t *pend_flag = nullptr;
        uint32_t pend_par = 0;
        auto publish = [&](uint32_t *flag, uint32_t par) {
            // lane 0 only; this warp's stores of the unit are complete
            fence_proxy_async_all();
            __threadfence_block();
            if (atomicAdd(&pub_cnt[par], 1u) == 3u) {
                pub_cnt[par] = 0u;
                __threadfence();
                if (p.n_blocks > 1)
                    atomicAdd(flag, 1u);          // one publication per column block of the tile
                else
                    st_release_gpu(flag, 1u);
                if (MAS_TR(p)) {
                    unsigned long long *tr = p.trace + (size_t)blockIdx.x * 64;
                    const unsigned long long n = tr[0] + 1;   // word 0: count, then one time per published tile
                    tr[0] = n;
                    if (n < 63) tr[n] = globaltimer_ns();
                    if (n == 1) tr[63] = ((unsigned long long)(flag - p.flags));
                }
            }
        };
        for (int un = 0, i = unit_index(0); i >= 0; i = unit_index(++un)) {
            int b, mt, nb;
            if (!unit_tile(i, b, mt, nb)) continue;
            const uint32_t a = nt & 1u, aph = (nt >> 1) & 1u;
            mbar_wait(&bias_full[a], aph);
            const float *bias_u = bias_s + a * kNMax;
            mbar_wait(&acc_full[a], aph);
            tc_fence_after();
            if (tid == 128) tr_mark(1, nt, 0);
            const int t = mt * kBM + row;
            const int c_base = nb * kNMax, ncols = block_cols(nb), s_left = p.S - c_base;
            float *orow = p.out + ((size_t)b * p.T + t) * p.ld + c_base;
            const uint32_t taddr = tmem_base + a * kNMax + ((uint32_t)(wq * 32) << 16);
            uint64_t us0 = 0ull, us1 = 0ull, uq0 = 0ull, uq1 = 0ull;   // packed fp32 {sum, sum} / {sum sq, sum sq} of the unit
            for (int c0 = 0; c0 < ncols; c0 += 32) {
                uint32_t r[32];
                // tcgen05.ld is .sync.aligned: every lane of the warp arrives here together (the plain-store path
                // below predicates its stores on the lane's row)
                __syncwarp();
                tmem_ld32(taddr + c0, r);
                tmem_ld_wait();
                float v[32];
#pragma unroll
                for (int j = 0; j < 32; j += 4) {  // broadcast 128-bit loads of the bias
                    const float4 b4 = *reinterpret_cast<const float4 *>(bias_u + c0 + j);
                    v[j] = __uint_as_float(r[j]) + b4.x, v[j + 1] = __uint_as_float(r[j + 1]) + b4.y;
                    v[j + 2] = __uint_as_float(r[j + 2]) + b4.z, v[j + 3] = __uint_as_float(r[j + 3]) + b4.w;
                }
                if (kStats) {
                    // noise statistics: sum and sum of squares of every real cell (models.py:1243 takes the std over
                    // ALL cells, padding included).  fp32 per thread and unit (<= 256 cells of one mel row: the
                    // rounding errors are unbiased and average out over the 10^7 row sums; measured against the
                    // fp64 statistics in tests/test_gpu_round2.py), packed two cells per instruction, folded into
                    // fp64 once per unit.  (A per-block fp32 pivot with fp64 folds cost 3 + operations per cell and
                    // made the epilogue the slowest role: 10 us per unit against 6.5 us of MMA time.)
                    if (t < p.T && s_left > c0) {
                        if (s_left - c0 >= 32) {
#pragma unroll
                            for (int j = 0; j < 32; j += 4) {
                                const uint64_t a = f2_pack(v[j], v[j + 1]), b2 = f2_pack(v[j + 2], v[j + 3]);
                                us0 = f2_add(us0, a), us1 = f2_add(us1, b2);
                                uq0 = f2_fma(a, a, uq0), uq1 = f2_fma(b2, b2, uq1);
                            }
                        } else {
                            float ts = 0.f, tq = 0.f;
#pragma unroll
                            for (int j = 0; j < 32; ++j)
                                if (c0 + j < s_left) {
                                    ts += v[j];
                                    tq = fmaf(v[j], v[j], tq);
                                }
                            us0 = f2_add(us0, f2_pack(ts, 0.f));
                            uq0 = f2_add(uq0, f2_pack(tq, 0.f));
                        }
                    }
                }
                if MAS_DBG(p, 2) continue;
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
                    // (Tried in round 2: reading the block back and storing it with plain coalesced 16-byte stores so
                    // that the tile flag need not wait for the tensor-store engine -- a tile takes 5-7 us from the end
                    // of the epilogue to its publication.  The epilogue then took 6.5 us per unit instead of 3.5 and
                    // held up the accumulators: 82 -> 93 us per step at config 2.)
                    if (lane == 0) {
                        tma_store_3d(tm_out, c_base + c0, mt * kBM + wq * 32, b, buf);  // rows >= T / cols >= S are clipped
                        bulk_commit();
                        if (c0 == 0 && pend_flag) {
                            bulk_wait_group<1>();  // everything older than the store just committed has landed
                            publish(pend_flag, pend_par);
                            pend_flag = nullptr;
                        }
                    }
                    ++nst;
                } else if (t < p.T) {
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (c0 + j < s_left) orow[c0 + j] = v[j];
                }
            }
            if (kStats) {
                float a0, a1, a2, a3, q0, q1, q2, q3;
                f2_unpack(us0, a0, a1), f2_unpack(us1, a2, a3), f2_unpack(uq0, q0, q1), f2_unpack(uq1, q2, q3);
                ssum += ((double)a0 + (double)a1) + ((double)a2 + (double)a3);
                ssq += ((double)q0 + (double)q1) + ((double)q2 + (double)q3);
            }
            tc_fence_before();
            if (tid == 128) tr_mark(1, nt, 1);
            if (!p.out_tma && p.flags) __threadfence();  // plain stores of every lane, before lane 0 publishes
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(&bias_empty[a]);
                if (kPair && rank)
                    mbar_arrive_cluster(mapa_u32(smem_u32(&acc_empty[a]), 0));
                else
                    mbar_arrive(&acc_empty[a]);
                if (p.flags && mt < p.m_tiles) {
                    uint32_t *flag = p.flags + (size_t)b * p.m_tiles + mt;
                    if (p.out_tma && !MAS_DBG(p, 2)) {
                        if MAS_DBG(p, 64) {
                            pend_flag = flag;  // experiment: publish after the first store of the next unit
                            pend_par = nt & 3u;
                        } else {
                            // nothing else to do until the next accumulator fills: wait for the stores to land
                            // and hand the tile to the DP now rather than one unit later
                            bulk_wait_all();
                            publish(flag, nt & 3u);
                        }
                    } else {
                        __threadfence();
                        publish(flag, nt & 3u);
                    }
                }
            }
            ++nt;
        }
        if (lane == 0 && pend_flag) {
            bulk_wait_all();
            publish(pend_flag, pend_par);
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
        const int grp = (warp - 8) >> 2;           // converter group: takes K blocks with (it & 1) == grp
        const int row = (tid - 256) & (kBM - 1);    // 0..127: mel row of the tile handled by this thread
        uint32_t it = 0;
        // where this warp announces a converted stage: rank 0's full barriers (a linear window)
        const uint32_t full0 = kPair ? mapa_u32(smem_u32(&full[0]), 0) : smem_u32(&full[0]);
        long long ph_acc[5] = {0, 0, 0, 0, 0};  // diagnostics: cycles in wait-z, LDS, wait-empty, convert+STS, fence+arrive
        for (int un = 0, i = unit_index(0); i >= 0; i = unit_index(++un)) {
            int b, mt, nb;
            if (!unit_tile(i, b, mt, nb)) continue;
            const int t = mt * kBM + row;
            const bool live = t < p.T && mt < p.m_tiles;
            const float *zb = p.z_p + (size_t)b * p.D * p.T + (live ? t : 0);
            if (tid == 256) tr_mark(2, it / (uint32_t)p.n_kb, 0);
            for (int kb = 0; kb < p.n_kb; ++kb, ++it) {
                if (tid == 256 && kb == p.n_kb - 2) tr_mark(2, it / (uint32_t)p.n_kb, 1);
                if ((int)(it & 1u) != grp) continue;
                const long long c0k = MAS_TR(p) ? clock64() : 0;
                long long c1k = 0, c2k = 0, c3k = 0, c4k = 0;
                float zc[kDPerKb];
                if MAS_DBG(p, 32) {  // experiment: no z traffic
#pragma unroll
                    for (int d = 0; d < kDPerKb; ++d) zc[d] = 0.f;
                } else if (p.z_tma) {
                    const uint32_t zs = it % kZStages, zph = (it / kZStages) & 1u;
                    mbar_wait(&zfull[zs], zph);
                    if (MAS_TR(p)) c1k = clock64();
                    const float *zr = reinterpret_cast<const float *>(smem + Cfg::kOffZ + zs * kZStageBytes) + row;
#pragma unroll
                    for (int d = 0; d < kDPerKb; ++d) zc[d] = zr[d * kBM];  // zero-filled past T / D by the TMA
                    if (MAS_TR(p)) c2k = clock64() + (long long)(__float_as_int(zc[0]) & 0);
                    // The stage is handed back further down, once the values have been USED.  Arriving on zempty right
                    // here (16 LDS issued, none consumed) let the arrive overtake the loads: SYNCS.ARRIVE does not wait
                    // in the shared-memory load queue, so when that queue is long (the plain-store epilogue's scattered
                    // STGs) the producer's next TMA fill of the stage -- K block kb + 6 -- landed before the last LDS had
                    // read it, and a warp's 32 rows got a few channels of the wrong K block (found in round 2 as
                    // run-dependent 32-row groups of neg_cent with rank-2..12 errors, tools/debug_plain.py).
                } else {
                    const int d0 = kb * kDPerKb;
#pragma unroll
                    for (int d = 0; d < kDPerKb; ++d) zc[d] = (live && d0 + d < p.D) ? zb[(size_t)(d0 + d) * p.T] : 0.f;
                }
                const uint32_t s = it % kStages, ph = (it / kStages) & 1u;
                mbar_wait(&empty[s], ph ^ 1u);
                if (MAS_TR(p)) c3k = clock64();
                unsigned char *a_hi = smem + s * Cfg::kStage;
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
                    if (!MAS_DBG(p, 1)) {
                        *reinterpret_cast<uint4 *>(a_hi + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                        *reinterpret_cast<uint4 *>(a_lo + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
                    }
                }
                if (MAS_TR(p)) c4k = clock64();
                fence_proxy_async();  // generic-proxy stores -> visible to the tensor core (async proxy)
                __syncwarp();
                if (lane == 0) {
                    // every lane's stores above took the z values as operands, so the loads are done: free the z stage
                    if (p.z_tma && !MAS_DBG(p, 32)) mbar_arrive(&zempty[it % kZStages]);
                    if (kPair)
                        mbar_arrive_cluster(full0 + s * 8u);
                    else
                        mbar_arrive(&full[s]);
                }
                if (MAS_TR(p)) {
                    const long long c5k = clock64();
                    ph_acc[0] += c1k - c0k, ph_acc[1] += c2k - c1k, ph_acc[2] += c3k - c2k, ph_acc[3] += c4k - c3k,
                        ph_acc[4] += c5k - c4k;
                }
            }
        }
        if (MAS_TR(p) && (tid == 256 || tid == 384))
            for (int j = 0; j < 5; ++j)
                p.trace[32768 + (size_t)blockIdx.x * 16 + (tid == 384 ? 8 : 0) + j] = (unsigned long long)ph_acc[j];
    }

    tc_fence_before();
    bar_sync(1, kTcThreads);
    if (kPair) cluster_sync_all();  // nothing of the peer is in flight towards this CTA any more
    // The fused kernels reuse this shared memory for the DP role.  Barrier words must be invalidated before they
    // become plain data: without it the DP role's zero page (which overlays them) kept stale barrier words in the
    // bulk-store engine's view and the path planes came out with 16-byte holes (config 3, ragged B = 128).
    if (tid == 0) {
        for (int i = 0; i < kStages; ++i) mbar_inval(&full[i]), mbar_inval(&empty[i]), mbar_inval(&bfull[i]);
        for (int i = 0; i < kZStages; ++i) mbar_inval(&zfull[i]), mbar_inval(&zempty[i]);
        for (int i = 0; i < 2; ++i)
            mbar_inval(&acc_full[i]), mbar_inval(&acc_empty[i]), mbar_inval(&bias_full[i]), mbar_inval(&bias_empty[i]);
    }
    bar_sync(1, kTcThreads);
    if (warp == 2) {
        if (kPair)
            tmem_dealloc2(tmem_base, 512);
        else
            tmem_dealloc(tmem_base, 512);
    }
}

// host side (mas_cost_tc.cu)
struct TcPlan {
    TcParams p;
    CUtensorMap tm_z, tm_out;
};
bool cost_tc_supported(int B, int D, int T, int S);
bool cost_tc_pair_enabled();
size_t cost_tc_workspace_bytes(int B, int D, int T, int S);
// launches the prior-image preparation (which also zeroes `flags_to_clear`, n_flags words, if given)
// and fills `plan` for the contraction
int cost_tc_prepare(TcPlan &plan, const float *z_p, const float *m_p, const float *logs_p, float *neg_cent_out,
                    double *stats_out, const int32_t *t_ys, void *workspace, size_t workspace_bytes, int B, int D, int T,
                    int S, uint32_t *flags_to_clear, int n_flags, cudaStream_t stream, int ld = 0);

}  // namespace mas
