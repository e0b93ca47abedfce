// mas_fused.cu -- neg_cent contraction and MAS in ONE kernel (reference vits2/models.py:1224-1256).
//
// No noise (mas_noise_scale None) -- mas_fused_pair_kernel:
// One CTA per SM.  Every CTA starts in the tcgen05 contraction role (mas_cost_tc.cuh) over the unit list --
// utterance groups of n_dp, mel-tile-major inside a group -- and publishes every finished 128-row cost tile with
// a release store to a flag.  After seq_k rounds the first n_dp CTAs leave the contraction and run the forward
// DP + backtrack (mas_dp.cuh), one utterance at a time, acquiring the flag of a tile before the TMA engine
// streams its rows out of L2; the other CTAs finish the units and then drain the zero-fill queue of the path
// planes.  The cost plane round-trips through L2 only; the DP trails the contraction by a few tiles instead of
// waiting for the whole batch.  Producers never wait on consumers; the launch is cooperative so that all CTAs
// are co-resident, and a programmatic dependent of the prior-images kernel (fused_launch).
//
// VITS2 noise-scaled MAS (models.py:1241-1247) -- mas_fused_noise_kernel:
// the standard deviation over ALL cost cells has to exist before the first DP row, so the kernel has two phases
// around a grid barrier: (1) every CTA contracts, the epilogue also accumulates sum / sum of squares; (2) up to
// one DP CTA per utterance runs the DP as above.  Eight helper warps of each DP CTA add (std * noise) * scale to
// the cost tiles in shared memory one chunk step ahead of the value warps (dp_role, kHelp), which therefore run
// the plain body; the draw is read once from HBM, by the SM that needs it.  (A first version applied the noise
// from the idle CTAs to the L2-resident plane: 393 KB through one SM per 128-row tile, ~7 us each -- the DP
// waited for its tiles and the step took 141 us at config 2.)
//
// The private cost plane has its own row stride (S rounded up to 4 floats), so S and T are arbitrary.
#include <atomic>

#include "mas_fused_body.cuh"
#include "mas_fused.cuh"

namespace mas {

#ifdef MAS_TRACE
// single-CTA contraction (cta_group::1), kept for A/B runs in the trace build (MAS_TC_PAIR=0)
template <int C, int R, int W, bool kVK>
__global__ void __launch_bounds__(kTcThreads, 1) mas_fused_kernel(const __grid_constant__ FusedParams fp,
                                                                  const __grid_constant__ CUtensorMap tm_z,
                                                                  const __grid_constant__ CUtensorMap tm_out)
{
    extern __shared__ __align__(128) unsigned char smem[];
    fused_body<C, R, W, false, kVK>(fp, &tm_z, &tm_out, smem);
}
#endif

// ---------------------------------------------------------------------------
// noise-scaled alignment: contraction + statistics | grid barrier | DP + noise
// ---------------------------------------------------------------------------
// Noise feeder (batches of at most one utterance per CTA pair, the VITS2 default of 64 per GPU included): after the
// grid barrier the even CTA of a pair runs the PLAIN DP role and the odd CTA -- idle otherwise -- feeds it: it
// streams the cost tile (out of L2) and the same cells of the draw (out of HBM) through its own bulk-copy engine
// into its own shared memory, adds (std * noise) * scale with all its warps and writes the finished tile straight
// into the DP CTA's stage ring through distributed shared memory, then arrives on that stage's mbarrier.  The DP
// CTA's shared-memory bandwidth, issue slots and bulk-copy engine carry nothing but the DP (with the helper warps
// inside the DP CTA a 32-row step took 3400 cycles instead of 2170); its producer warp only returns one credit per
// consumed stage (and, its bulk-copy engine being free, zero-fills its own path plane as it goes).
constexpr int kFeedStages = 3;                 // feeder ring: {cost tile, noise tile} per stage
constexpr int kFeedWarps = 15;                 // applier warps (warp 15 issues the bulk copies)
constexpr int kFeedAhead = 6;                  // chunks of the draw kept ahead in L2
constexpr uint32_t kFeedOffEmpty = 32, kFeedOffCredit = 64, kFeedOffVerdict = 128, kFeedOffStage = 256;
__host__ __device__ inline uint32_t feed_tile_bytes(int R, int ld) { return (uint32_t)(((size_t)R * ld * 4 + 127) & ~(size_t)127); }
__host__ __device__ inline uint32_t feed_smem_bytes(int R, int ld) { return kFeedOffStage + kFeedStages * 2 * feed_tile_bytes(R, ld); }

__device__ __forceinline__ void noise_feeder_init(unsigned char *smem)
{
    uint64_t *ffull = reinterpret_cast<uint64_t *>(smem);
    uint64_t *fempty = reinterpret_cast<uint64_t *>(smem + kFeedOffEmpty);
    uint64_t *credit = reinterpret_cast<uint64_t *>(smem + kFeedOffCredit);
    uint64_t *verdict = reinterpret_cast<uint64_t *>(smem + kFeedOffVerdict);
    if (threadIdx.x == 0) {
        for (int i = 0; i < kFeedStages; ++i) {
            mbar_init(&ffull[i], 1);
            mbar_init(&fempty[i], kFeedWarps);
        }
        for (int i = 0; i < kMaxStages; ++i) mbar_init(&credit[i], 1);
        mbar_init(verdict, 1);
        fence_mbar_init();
    }
}

template <int R>
__device__ __forceinline__ void noise_feeder_role(const FusedParams &fp, unsigned char *smem, int first, int stride)
{
    const DpParams &p = fp.dp;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    constexpr int HT = kFeedWarps * 32;
    uint64_t *ffull = reinterpret_cast<uint64_t *>(smem);
    uint64_t *fempty = reinterpret_cast<uint64_t *>(smem + kFeedOffEmpty);
    uint64_t *credit = reinterpret_cast<uint64_t *>(smem + kFeedOffCredit);
    uint64_t *verdict = reinterpret_cast<uint64_t *>(smem + kFeedOffVerdict);
    volatile uint32_t *verdict_val = reinterpret_cast<volatile uint32_t *>(smem + kFeedOffVerdict + 8);
    const int T = p.T, S = p.S, ld = p.ld, ld4 = ld >> 2;     // (ld == S: checked by the launcher)
    const size_t plane = (size_t)T * S;
    const uint32_t tile_bytes = feed_tile_bytes(R, ld);
    const uint32_t n_stages = (uint32_t)p.stages;
    // unbiased std over ALL cells, padding included (torch.std default, models.py:1243), from fp64 sums
    float sd;
    {
        const double n = (double)p.B * (double)plane;
        const double s0 = __ldcg(fp.stats), s1 = __ldcg(fp.stats + 1);
        const double mean = s0 / n;
        double var = (s1 - s0 * mean) / (n > 1.0 ? n - 1.0 : 1.0);
        if (var < 0) var = 0;
        sd = (float)sqrt(var);
    }
    const float scale = fp.noise_scale;
    // the partner's stage ring and its full barriers, as shared::cluster addresses of cluster rank 0
    const uint32_t r_stage0 = dsm_map(smem_u32(smem + p.off_stage), 0u);
    const uint32_t r_full0 = dsm_map(smem_u32(smem + p.off_bar), 0u);
    uint32_t g_base = 0, n_verdicts = 0;
    for (int b = first; b < p.B; b += stride) {
        const int t_y = p.t_ys[b], t_x = p.t_xs[b];
        if (!(t_x >= 1 && t_x <= t_y && t_y <= T && t_x <= S)) continue;   // the DP role skips it the same way
        const int n_chunks = (t_y + R - 1) / R;
        const unsigned char *cost_b = reinterpret_cast<const unsigned char *>(p.neg_cent + (size_t)b * T * ld);
        const unsigned char *nz_b = reinterpret_cast<const unsigned char *>(fp.noise + (size_t)b * plane);
        int passes = 0;
        for (int pass = 0; pass < 2; ++pass) {
            const uint32_t g0 = g_base + (uint32_t)pass * n_chunks;
            if (warp == kFeedWarps) {
                // ---- bulk-copy issuer: the tile loads ----
                if (lane == 0) {
                    for (int c = 0; c < n_chunks; ++c) {
                        const uint32_t g = g0 + (uint32_t)c, s = g % kFeedStages, u = g / kFeedStages;
                        const uint32_t bytes = (uint32_t)min(R, t_y - c * R) * ld * 4;
                        mbar_wait(&fempty[s], (u & 1u) ^ 1u);
                        unsigned char *dst = smem + kFeedOffStage + s * 2 * tile_bytes;
                        mbar_arrive_expect_tx(&ffull[s], 2 * bytes);
                        bulk_g2s(dst, cost_b + (size_t)c * R * ld * 4, bytes, &ffull[s]);
                        bulk_g2s(dst + tile_bytes, nz_b + (size_t)c * R * ld * 4, bytes, &ffull[s]);
                    }
                }
                __syncwarp();
            } else {
                // ---- appliers: cost + (std * noise) * scale, rounded after every operation (models.py:1242-1247),
                //      from this CTA's stage straight into the partner's ----
                const int gl = warp * 32 + lane;
                long long facc[4] = {0, 0, 0, 0};   // diagnostics: cycles waiting for the loads, reading, waiting for the credit, storing
                // the first chunks of the draw: into L2 right away (one 128-byte line per thread and chunk)
                for (int c = 0; c < kFeedAhead && c < n_chunks; ++c) {
                    const int pbytes = min(R, t_y - c * R) * ld * 4;
                    if (gl * 128 < pbytes) prefetch_l2(nz_b + (size_t)c * R * ld * 4 + gl * 128);
                }
                for (int c = 0; c < n_chunks; ++c) {
                    const uint32_t g = g0 + (uint32_t)c, s = g % kFeedStages, u = g / kFeedStages;
                    const uint32_t st = g % n_stages, ud = g / n_stages;
                    const int n4 = min(R, t_y - c * R) * ld4;
                    const float4 *c4 = reinterpret_cast<const float4 *>(smem + kFeedOffStage + s * 2 * tile_bytes) + gl;
                    const float4 *z4 = reinterpret_cast<const float4 *>(smem + kFeedOffStage + s * 2 * tile_bytes + tile_bytes) + gl;
                    // the draw a few chunks ahead: into L2, one 128-byte line per thread
                    if (c + kFeedAhead < n_chunks) {
                        const int pbytes = min(R, t_y - (c + kFeedAhead) * R) * ld * 4;
                        if (gl * 128 < pbytes) prefetch_l2(nz_b + (size_t)(c + kFeedAhead) * R * ld * 4 + gl * 128);
                    }
                    const long long f0 = MAS_TR(fp.tc) ? clock64() : 0;
                    mbar_wait(&ffull[s], u & 1u);
                    const long long f1 = MAS_TR(fp.tc) ? clock64() : 0;
                    constexpr int KQ = (R * 64 + HT - 1) / HT;   // items per thread and chunk at the widest plane (256 floats)
                    float4 cv[KQ], nv[KQ];
#pragma unroll
                    for (int k = 0; k < KQ; ++k)
                        if (gl + k * HT < n4) cv[k] = c4[k * HT], nv[k] = z4[k * HT];
                    // (the stage is NOT handed back here: an arrive issued right after the loads does not queue behind
                    // them in the shared-memory pipe and the producer's refill can land first; see cost_tc_role's
                    // converters.  It goes back after the stores below.)
#pragma unroll
                    for (int k = 0; k < KQ; ++k)
                        if (gl + k * HT < n4) {
                            cv[k].x = __fadd_rn(cv[k].x, __fmul_rn(__fmul_rn(sd, nv[k].x), scale));
                            cv[k].y = __fadd_rn(cv[k].y, __fmul_rn(__fmul_rn(sd, nv[k].y), scale));
                            cv[k].z = __fadd_rn(cv[k].z, __fmul_rn(__fmul_rn(sd, nv[k].z), scale));
                            cv[k].w = __fadd_rn(cv[k].w, __fmul_rn(__fmul_rn(sd, nv[k].w), scale));
                        }
                    const long long f2 = MAS_TR(fp.tc) ? clock64() : 0;
                    if (lane == 0) mbar_wait_acq_cluster(&credit[st], (ud & 1u) ^ 1u);   // the partner has consumed the stage's previous tile
                    __syncwarp();
                    const long long f3 = MAS_TR(fp.tc) ? clock64() : 0;
                    const uint32_t dst = r_stage0 + st * p.stage_bytes + (uint32_t)gl * 16u;
#pragma unroll
                    for (int k = 0; k < KQ; ++k)
                        if (gl + k * HT < n4) {
                            // asynchronous store: counted on the partner's full barrier of the stage, which its
                            // producer lane has armed with the tile's byte count (no fence, no arrive here: a
                            // release-arrive per warp and tile cost the feeder 1.9 us per chunk)
                            dsm_st_async_v4(dst + (uint32_t)(k * HT) * 16u, cv[k], r_full0 + st * 8u);
                        }
                    // The stage goes back to the producer only now: the stores above (memory operations, which the
                    // release-arrive cannot pass) took every loaded value as an operand, so the loads are done.
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&fempty[s]);
                    if (MAS_TR(fp.tc)) {
                        const long long f4 = clock64();
                        facc[0] += f1 - f0, facc[1] += f2 - f1, facc[2] += f3 - f2, facc[3] += f4 - f3;
                    }
                }
                if (MAS_TR(fp.tc) && gl == 0)
                    for (int j = 0; j < 4; ++j) fp.tc.trace[40960 + (size_t)b * 16 + 12 + j] = (unsigned long long)facc[j];
            }
            // does the DP want the tiles once more (exact pass after a non-finite cost)?
            mbar_wait_acq_cluster(verdict, n_verdicts & 1u);
            const uint32_t again = *verdict_val;
            ++n_verdicts;
            passes = pass + 1;
            if (!again) break;
        }
        g_base += (uint32_t)passes * n_chunks;
    }
}

template <int C, int R, int W, bool kVK>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kTcThreads, 1)
    mas_fused_noise_kernel(const __grid_constant__ FusedParams fp, const __grid_constant__ CUtensorMap tm_z,
                           const __grid_constant__ CUtensorMap tm_out)
{
    extern __shared__ __align__(128) unsigned char smem[];
    // phase 1: the whole batch's cost plane + its statistics (no tile flags: nothing may be aligned yet)
    if (MAS_TR(fp.tc) && threadIdx.x == 0) fp.tc.trace[49152 + blockIdx.x] = globaltimer_ns();  // CTA entry
    cost_tc_role<true, true>(fp.tc, &tm_z, &tm_out, smem, blockIdx.x >> 1, gridDim.x >> 1);
    if (MAS_TR(fp.tc) && threadIdx.x == 0) fp.tc.trace[49152 + 256 + blockIdx.x] = globaltimer_ns();  // contraction done
    // with feeders: both CTAs of a pair set their barriers up BEFORE the grid barrier, so that neither can arrive on
    // a barrier of the other that does not exist yet
    const bool in_pair = kVK && fp.feed_pairs > 0 && (int)blockIdx.x < 2 * fp.feed_pairs;
    const bool is_feeder = in_pair && (blockIdx.x & 1u);
    if (in_pair) {
        if (is_feeder)
            noise_feeder_init(smem);
        else if ((int)threadIdx.x < dp_threads(W, kVK))
            dp_role_init(fp.dp, smem, threadIdx.x, kDpBar);
    }
    // grid barrier: every CTA's tiles are in memory (L2) and its partial sums are in stats[]
    fence_proxy_async_all();
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(fp.grid_bar, 1u);
        while (ld_acquire_gpu(fp.grid_bar) < gridDim.x) __nanosleep(32);
        __threadfence();
        fence_proxy_async_all();   // the DP's bulk copies (async proxy) read what other SMs' bulk stores wrote
    }
    __syncthreads();
    if (MAS_TR(fp.tc) && threadIdx.x == 0) fp.tc.trace[49152 + 512 + blockIdx.x] = globaltimer_ns();  // barrier passed
    if (kVK && fp.feed_pairs > 0) {
        if (!in_pair) return;
        if (is_feeder) {
            noise_feeder_role<R>(fp, smem, (int)(blockIdx.x >> 1), fp.feed_pairs);
            if (MAS_TR(fp.tc) && threadIdx.x == 0) fp.tc.trace[49152 + 768 + blockIdx.x] = globaltimer_ns();  // feeder done
            return;
        }
        if ((int)threadIdx.x >= dp_threads(W, kVK)) return;
        uint32_t g_base = 0;
        for (int b = (int)(blockIdx.x >> 1); b < fp.dp.B; b += fp.feed_pairs)
            dp_role<C, R, W, true, false, kVK, 0, true>(fp.dp, smem, b, (int)(blockIdx.x >> 1), g_base, threadIdx.x, kDpBar);
        return;
    }
    if ((int)blockIdx.x >= fp.n_dp) {
        if (fp.dp.zero_flags) zero_fill_role(fp.dp, smem);
        if (MAS_TR(fp.tc) && threadIdx.x == 0) fp.tc.trace[49152 + 1024 + blockIdx.x] = globaltimer_ns();  // zero fill done
        return;
    }
    // phase 2 without feeders (more utterances than CTA pairs): the helper warps of every DP CTA add
    // (std * noise) * scale to its cost tiles in shared memory; the draw is read exactly once
    fused_dp_ctas<C, R, W, kVK, kNoiseHelpWarps>(fp, smem);
}

// the private cost plane keeps 16-byte rows whatever S is
static int padded_ld(int S) { return (S + 3) & ~3; }

bool fused_supported(int B, int D, int T, int S)
{
    if (config().no_fused) return false;
    // up to four column blocks of 256 text columns (a mel tile is complete when all its blocks have been published)
    return cost_tc_supported(B, D, T, S);
}

// noise-scaled alignment in one kernel (any batch size: the DP CTAs apply the noise themselves)
bool fused_noise_supported(int B, int D, int T, int S)
{
    // (one column block: the noise kernel's DP roles are the warp-split ones)
    return config().noise_fused && fused_supported(B, D, T, S) && S <= kNMax && config().dp_vk;
}

// tile flags [B][m_tiles], the zero-fill flags [B], the zero-fill queue counter, the grid barrier counter
size_t fused_flags_bytes(int B, int T) { return align_up(((size_t)B * ((T + kBM - 1) / kBM + 1) + 2) * 4, 256); }
// bytes of the private cost plane inside the fused workspace
size_t fused_plane_bytes(int B, int T, int S) { return align_up((size_t)B * T * padded_ld(S) * 4, 256); }

// Launches the prior preparation and one fused kernel.  Returns MAS_OK, an error, or kFusedFallback when the
// cooperative launch is not possible in this context (MPS / MIG / green-context limits, no tensor maps): the
// caller then runs contraction and DP as separate launches.
int fused_launch(const float *z_p, const float *m_p, const float *logs_p, const int32_t *t_ys, const int32_t *t_xs,
                 const float *noise, float noise_scale, double *stats, float *plane, void *path_out, int path_dtype,
                 int32_t *dur_out, int32_t *idx_out, int32_t *status_out, void *cost_ws, size_t cost_ws_bytes, void *dp_ws,
                 size_t dp_ws_bytes, uint32_t *flags, int B, int D, int T, int S, cudaStream_t stream)
{
    int dev = 0, sms = 148;
    MAS_CUDA_TRY(cudaGetDevice(&dev));
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const Config &cf = config();
    if (noise && (reinterpret_cast<uintptr_t>(noise) & 3) != 0) return kFusedFallback;
    const int m_tiles = (T + kBM - 1) / kBM;
    const int ld = padded_ld(S);
    const int n_flags = B * (m_tiles + 1) + 2;
    // Noise: one utterance per CTA pair at most (and 16-byte rows of the draw) -> {DP CTA, noise feeder CTA} pairs;
    // larger batches -> every CTA a DP CTA with its own noise-helper warps
    const int grid_ctas = sms & ~1;
    const bool feed = noise && cf.noise_feed && 2 * B <= grid_ctas && (S & 3) == 0 &&
                      (reinterpret_cast<uintptr_t>(noise) & 15) == 0;
    // the DP plan first: nothing has been launched yet if it turns out not to fit
    DpPlan dp;
    int rc = dp_prepare(dp, plane, t_ys, t_xs, path_out, path_dtype, dur_out, idx_out, status_out, dp_ws, dp_ws_bytes, B, T,
                        S, nullptr, 0, 0, false, ld, (noise && !feed) ? kNoiseHelpWarps : 0);
    if (rc) return noise ? kFusedFallback : rc;
    TcPlan tc;
    // mel tiles wholly past t_y are skipped (the plane is private scratch) unless the noise statistics need them
    rc = cost_tc_prepare(tc, z_p, m_p, logs_p, plane, noise ? stats : nullptr, noise ? nullptr : t_ys, cost_ws,
                         cost_ws_bytes, B, D, T, S, flags, n_flags, stream, ld);
    if (rc) return rc;
    if (cf.stage == 1) return MAS_OK;   // prior preparation only (bench.py times it alone)
    if (!tc.p.out_tma) return kFusedFallback;
    FusedParams fp{};
    fp.tc = tc.p;
    fp.dp = dp.p;
    const bool pair = noise ? true : cost_tc_pair_enabled();
    const int per = pair ? 2 : 1;

    // S <= 256 (one column block): C = ceil(S / 64) columns per thread with 2 DP warps, or 2 columns with 4
    // (value / origin warp split unless MAS_DP_VK=0).  The noise kernel exists for the warp-split teams only.
    const void *kernel = nullptr;
#ifdef MAS_TRACE
#define MAS_FUSED_PLAIN(CC, WW, VK) \
    (pair ? (const void *)mas_fused_pair_kernel<CC, 32, WW, VK> : (const void *)mas_fused_kernel<CC, 32, WW, VK>)
#else
#define MAS_FUSED_PLAIN(CC, WW, VK) ((const void *)mas_fused_pair_kernel<CC, 32, WW, VK>)
#endif
#define MAS_FUSED_CASE(CC, WW, VK) \
    if (!noise && dp.C == CC && dp.p.R == 32 && dp.p.W == WW && (dp.p.vk != 0) == VK) kernel = MAS_FUSED_PLAIN(CC, WW, VK);
#define MAS_NOISE_CASE(CC) \
    if (noise && dp.C == CC && dp.p.R == 32 && dp.p.W == 2 && dp.p.vk) kernel = (const void *)mas_fused_noise_kernel<CC, 32, 2, true>;
    MAS_FUSED_CASE(1, 2, true)
    MAS_FUSED_CASE(2, 2, true)
    MAS_FUSED_CASE(3, 2, true)
    MAS_FUSED_CASE(4, 2, true)
#ifdef MAS_TRACE
    // A/B variants, trace build only (both measured slower): four value warps with two columns
    // each (MAS_DP_WARPS=4), single-role DP warps (MAS_DP_VK=0).  In the product build those knobs make the call
    // fall back to separate launches.
    MAS_FUSED_CASE(2, 4, true)
    MAS_FUSED_CASE(1, 2, false)
    MAS_FUSED_CASE(2, 2, false)
    MAS_FUSED_CASE(3, 2, false)
    MAS_FUSED_CASE(4, 2, false)
#endif
    MAS_NOISE_CASE(1)
    MAS_NOISE_CASE(2)
    MAS_NOISE_CASE(3)
    MAS_NOISE_CASE(4)
#undef MAS_FUSED_CASE
#undef MAS_NOISE_CASE
#undef MAS_FUSED_PLAIN
    if (!noise && S > kNMax && pair && !dp.p.vk && dp.p.W == 2) kernel = fused_pair_kernel_wide2(dp.C, dp.p.R);
    if (!noise && S > kNMax && pair && !dp.p.vk && dp.p.W == 4) kernel = fused_pair_kernel_wide4(dp.C, dp.p.R);
    if (!kernel) return kFusedFallback;

    size_t smem = dp.smem_bytes;
    if (smem < kTcSmem) smem = kTcSmem;
    if (smem < kZeroFillBuf) smem = kZeroFillBuf;
    if (feed && smem < feed_smem_bytes(dp.p.R, ld)) smem = feed_smem_bytes(dp.p.R, ld);
    if (smem > (size_t)kSmemBudget) return kFusedFallback;
    // all CTAs must be co-resident (the DP CTAs spin on flags the others raise): check what this context can
    // actually hold, one CTA per SM at most
    MAS_CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget));
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kTcThreads, smem) != cudaSuccess || per_sm < 1) {
        (void)cudaGetLastError();
        return kFusedFallback;
    }
    const int grid = pair ? (sms & ~1) : sms;
    const int units = B * (pair ? (m_tiles + 1) / 2 : m_tiles) * tc.p.n_blocks;

    // Unit schedule: all CTAs take seq_k rounds of units, then the DP CTAs leave.  A small cost model picks
    // seq_k and, for batches larger than 64, between 64 DP CTAs and nearly all of them: the contraction must
    // not end long after the DP could, and the DP must not start long before its tiles exist.
    //   unit: the bytes one SM moves per unit through its TMA path (~60 GB/s) or the MMA time, whichever is longer
    //   DP:   ~ (16 + 10 C) cycles per mel row with the value / origin split, + 8 us per utterance (backtrack, outputs)
    const int n_kb = tc.p.n_kb, n_cols = S > kNMax ? kNMax : ((S + 15) & ~15);   // (columns of one unit = one column block)
    const double unit_bytes = (double)kBM * D * 4 + (double)n_kb * 2 * (n_cols / per) * kRowBytes + (double)kBM * n_cols * 4;
    const double t_mma = 4.7 * (n_cols / 256.0) * (D / 192.0);
    const double t_u = unit_bytes / 60e3 > t_mma ? unit_bytes / 60e3 : t_mma;
    const double t_row = 0.040 * (16.0 + 10.0 * dp.C) / 56.0;
    auto model = [&](int n_dp_try, int &k_out) {
        const int P = grid / per, P_pure = (grid - n_dp_try) / per;
        const int utts = (B + n_dp_try - 1) / n_dp_try;
        double best = 1e30;
        for (int k = 0; k <= 64; ++k) {
            const int rest = units - k * P > 0 ? units - k * P : 0;
            const double gemm_end = (k + (rest + P_pure - 1) / P_pure) * t_u;
            const double dp_end = k * t_u + utts * (T * t_row + 8.0);
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
    int n_dp = cf.fused_dp_ctas, best_k = 0;
    if (noise) {
        n_dp = B < grid ? B : grid;  // after the barrier every CTA is free: one utterance per CTA as far as they go
        best_k = 1 << 28;            // nobody leaves the contraction before the barrier
    } else if (n_dp > 0) {
        n_dp = legal(n_dp);
        model(n_dp, best_k);
    } else {
        int k_a = 0, k_b = 0;
        const int n_a = legal(B < 64 ? B : 64), n_b = legal(B);
        const double t_a = model(n_a, k_a), t_b = model(n_b, k_b);
        n_dp = (t_b < t_a) ? n_b : n_a;
        best_k = (t_b < t_a) ? k_b : k_a;
    }
    if (n_dp < 1) return kFusedFallback;
    fp.n_dp = n_dp;
    const int utts_per_cta = (B + n_dp - 1) / n_dp;
    fp.tc.seq_k = (!noise && cf.fused_rounds >= 0) ? cf.fused_rounds : best_k;
    fp.tc.seq_pure0 = noise ? 0 : n_dp / per;
    fp.tc.wave = noise ? B : fp.n_dp;
    fp.tc.flags = noise ? nullptr : flags;    // noise: the whole plane exists before the first DP row (grid barrier)
    fp.dp.flags = noise ? nullptr : flags;
    if (noise) {
        if (!dp.p.vk) return kFusedFallback;   // the helper warps / the feeder come with the value / origin split
        fp.dp.noise = noise;
        fp.dp.stats = stats;
        fp.dp.noise_scale = noise_scale;
        fp.noise = noise;
        fp.stats = stats;
        fp.noise_scale = noise_scale;
        if (feed) {
            fp.feed_pairs = B;
            fp.dp.help = 0;
            fp.dp.fed = 1;
            fp.dp.fed_credit_off = kFeedOffCredit;
            fp.dp.fed_verdict_off = kFeedOffVerdict;
        } else {
            fp.dp.help = kNoiseHelpWarps;
        }
    }
    fp.tc.trace = trace_buffer();
    fp.dp.trace = fp.tc.trace;
    fp.dp.flag_tiles = m_tiles;
    fp.dp.flag_need = tc.p.n_blocks;          // every column block of a mel tile publishes it once
    // the contraction-only CTAs zero-fill the path planes when there are enough of them to do it in time
    // (with feeders the DP CTA zero-fills its own plane: its bulk-copy engine has nothing else to do.  Zero-filling
    // from the feeder cost 13 us at config 2 through its bulk-copy engine, which carries the tile loads, and 9 us
    // with plain stores from its applier warps; sharing zero_fill_role's work queue between the 20 CTAs that are in
    // no pair at config 2 and the DP CTAs' producer lanes 9 us -- and 4 us even with the queue unused, from what the
    // extra code did to the DP role's register allocation.)
    const bool offload = path_out && !feed && cf.fused_zero_offload && (grid - n_dp) * 2 >= n_dp && utts_per_cta == 1;
    fp.dp.zero_flags = offload ? flags + (size_t)B * m_tiles : nullptr;
    fp.dp.zero_queue = flags + (size_t)B * (m_tiles + 1);
    fp.grid_bar = flags + (size_t)B * (m_tiles + 1) + 1;

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
    bool use_pdl = cf.fused_pdl && (pdl_state == 1 || (pdl_state < 0 && cap == cudaStreamCaptureStatusNone));
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeCooperative;
    attr[0].val.cooperative = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    fp.tc.pdl = 1;  // (the wait is a no-op under a plain launch)
    void *args[3] = {&fp, &tc.tm_z, &tc.tm_out};
    cudaError_t e = cudaErrorInvalidValue;
    for (int attempt = 0; attempt < 2; ++attempt) {
        cfg.numAttrs = use_pdl ? 2 : 1;
        e = cudaLaunchKernelExC(&cfg, kernel, args);
        if (!use_pdl) break;
        if (e == cudaSuccess) {
            pdl_state = 1;
            break;
        }
        if (pdl_state == 1) break;  // worked before: a real error
        (void)cudaGetLastError();
        pdl_state = 0;
        use_pdl = false;
    }
    if (e == cudaErrorCooperativeLaunchTooLarge || e == cudaErrorLaunchOutOfResources) {
        // this context cannot hold the whole grid (MPS thread percentage, MIG slice, green context): not an error
        (void)cudaGetLastError();
        return kFusedFallback;
    }
    note_launch();
    if (e != cudaSuccess) return note_cuda_error(e, "cudaLaunchKernelEx(mas_fused_kernel)");
    return MAS_OK;
}

}  // namespace mas
