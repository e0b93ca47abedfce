// mas_dp.cu -- standalone launch of the MAS forward DP + backtrack (role code in mas_dp.cuh),
// the shared-memory plan, and the small helper kernels (lengths from mask, launch order, expand).
#include "mas_dp_launch.cuh"

namespace mas {

// ---------------------------------------------------------------------------
// host: shared-memory plan
// ---------------------------------------------------------------------------
bool dp_plan_try(DpPlan &pl, int T, int S, int ld, int W, int C, int R, int stages, bool bits_smem, bool hop_smem,
                 size_t budget, bool with_noise, bool vk, int help = 0)
{
    const int S_pad = W * 32 * C;
    const int n_blk = (T + kCheck - 1) / kCheck;  // 32-row blocks of decision words
    const int hop_rows = T / kCheck + 2;
    size_t off = 0;
    DpParams &p = pl.p;
    p.R = R;
    p.W = W;
    p.vk = vk ? 1 : 0;
    p.stages = stages;
    p.off_bar = (uint32_t)off;
    off += 128;
    p.off_bits = (uint32_t)off;
    if (bits_smem) off += align_up((size_t)n_blk * (S_pad + kBitsPad) * 4, 16);
    p.off_hop = (uint32_t)off;
    if (hop_smem) off += align_up((size_t)hop_rows * S_pad, 16);
    p.stage_bytes = (uint32_t)align_up((size_t)R * ld * 4 + 16 + 64, 128);  // tile + misalignment + zeroed pad
    p.noise_off = with_noise ? p.stage_bytes : 0;                          // the noise tile follows the cost tile
    if (with_noise) p.stage_bytes *= 2;
    off = align_up(off, 128);
    p.off_stage = (uint32_t)off;
    off += (size_t)stages * p.stage_bytes;
    p.off_nz = (uint32_t)off;                                  // helper warps: one chunk of the noise draw
    if (help > 0) off += align_up((size_t)R * ld * 4 + 64, 128);   // two halves, each with room for a misaligned start
    p.help = help;
    p.off_bnd_v = (uint32_t)off;
    off += (size_t)(W + 1) * 2 * R * 4;
    p.off_bnd_o = (uint32_t)off;
    off += (size_t)(W + 1) * 2 * R * 4;
    p.off_zero = (uint32_t)off;
    off += kZeroBytes;
    p.off_idx = (uint32_t)off;
    off += align_up(T <= 2048 ? (size_t)32 * ((T >> 5) + 2) * 2 : (size_t)T * 2 + 2, 16);   // idx_s (dp_role: ix())
    p.off_end = (uint32_t)off;
    off += (size_t)S_pad * 4;
    p.off_entry = (uint32_t)off;
    off += align_up((size_t)hop_rows * 2, 16);
    p.off_misc = (uint32_t)off;
    off += 16;
    p.bits_in_smem = bits_smem;
    p.hop_in_smem = hop_smem;
    p.bits_words_per_cta = (unsigned long long)n_blk * (S_pad + kBitsPad);
    p.hop_bytes_per_cta = (unsigned long long)align_up((size_t)hop_rows * S_pad, 16);
    pl.smem_bytes = off;
    pl.C = C;
    return off <= budget;
}

// Deepest cost-tile ring that fits next to on-chip decision bits / hops; long utterances spill the
// bits (then the hops) to the workspace.  stages_hint > 0 forces the ring depth (MAS_DP_STAGES).
// DP warps per team for S text columns: 2 (C = ceil(S / 64) <= 8 columns per thread) up to S = 512, else 4.
// MAS_DP_WARPS overrides (2 or 4) where the shape allows it.
int dp_team_warps(int S)
{
    int W = S <= 512 ? 2 : 4;
    if (config().dp_warps == 4 && S > 128 && S <= 256) W = 4;  // 4 warps x 2 columns per thread
    return W;
}

bool dp_make_plan(DpPlan &pl, int B, int T, int S, int ld, int stages_hint, int R_hint, size_t budget, bool with_noise,
                  bool vk, int help = 0)
{
    const int W = dp_team_warps(S);
    const int C = (S + W * 32 - 1) / (W * 32);
    if (C < 1 || C > 8) return false;
    // Chunk height: every chunk step costs a fixed ~700 cycles (barrier, mbarrier wait, bookkeeping), so take
    // the tallest chunk whose ring still fits: dp_chunk_rows(S) with bits / hops on chip, else up to twice
    // that once they are spilled anyway (long utterances).
    const int R0 = R_hint > 0 ? R_hint : dp_chunk_rows(S);
    bool ok = false;
    if (vk) {
        // warp split: everything on chip, one tile of prefetch distance at least, or not at all
        // (with helper warps one more chunk is in use at any time: the one being noised)
        const int min_st = W + 1 + (help > 0 ? 1 : 0);
        for (int st = 5; st >= min_st && !ok; --st)
            ok = dp_plan_try(pl, T, S, ld, W, C, R0, st, true, true, budget, false, true, help);
        if (!ok) return false;
        pl.ws_bits_bytes = pl.ws_hop_bytes = 0;
        return true;
    }
    for (int mode = 0; mode < 3 && !ok; ++mode) {
        const bool bits_smem = (mode == 0), hop_smem = (mode <= 1);
        // the W DP warps work on W consecutive chunks at once, so the ring needs at least W stages (then
        // without prefetch distance); on-chip bits / hops are only worth it with two stages to spare
        const int min_stages = (mode == 2) ? W + 1 : W + 2;
        for (int R = (mode == 2 && R_hint <= 0 && R0 < 32) ? 2 * R0 : R0; R >= R0 && !ok; R /= 2)
            for (int st = (stages_hint > 0 ? stages_hint : 6); st >= min_stages && !ok; --st) {
                ok = dp_plan_try(pl, T, S, ld, W, C, R, st, bits_smem, hop_smem, budget, with_noise, false);
                if (stages_hint > 0) break;
            }
    }
    if (ok && stages_hint <= 0 && R_hint <= 0 && !pl.p.bits_in_smem && pl.p.hop_in_smem && pl.p.R < 32) {
        // Long utterances of wide text: the decision words are spilled anyway, and a chunk step costs ~700 cycles
        // whatever its height -- at S = 600, T = 4000 that is 500 steps of 8 rows = 180 us of a 365 us DP.  Chunks
        // twice as tall with the hop bytes spilled too win as long as the backtrack can bring the hops back into the
        // idle tile ring in one sweep (dp_role): 365 -> 300 us.
        DpPlan alt = pl;
        bool alt_ok = false;
        for (int st = 6; st >= W + 1 && !alt_ok; --st)
            alt_ok = dp_plan_try(alt, T, S, ld, W, C, 2 * pl.p.R, st, false, false, budget, with_noise, false);
        if (alt_ok && (size_t)(T / kCheck + 2) * alt.p.W * 32 * C <= (size_t)alt.p.stages * alt.p.stage_bytes) pl = alt;
    }
    if (ok && stages_hint <= 0 && !pl.p.hop_in_smem) {
        // The search ended on the fully spilled layout.  If the same ring also fits next to on-chip hops, take
        // that: the backtrack's T/32 dependent hop reads cost ~0.6 us each from L2 and ~30 ns from shared memory
        // (noise-scaled MAS at S = 256: three 66 KB stages leave room for the 9 KB of hop bytes, not for the
        // 32 KB of decision words).
        DpPlan alt = pl;
        if (dp_plan_try(alt, T, S, ld, W, C, pl.p.R, pl.p.stages, false, true, budget, with_noise, false)) pl = alt;
    }
    if (!ok) {
        // last resort: no prefetch distance at all
        ok = dp_plan_try(pl, T, S, ld, W, C, R0, W, false, false, budget, with_noise, false);
    }
    if (!ok) return false;
    pl.ws_bits_bytes = pl.p.bits_in_smem ? 0 : (size_t)B * pl.p.bits_words_per_cta * 4;
    pl.ws_hop_bytes = pl.p.hop_in_smem ? 0 : (size_t)B * pl.p.hop_bytes_per_cta;
    return true;
}

// ---------------------------------------------------------------------------
// lengths from the dense mask (reference __init__.py:16-17)
// ---------------------------------------------------------------------------
__global__ void mas_lengths_kernel(const float *__restrict__ mask, int32_t *t_ys, int32_t *t_xs, int T, int S)
{
    const int b = blockIdx.x;
    const float *m = mask + (size_t)b * T * S;
    float sy = 0.f, sx = 0.f;
    for (int y = threadIdx.x; y < T; y += blockDim.x) sy += m[(size_t)y * S];
    for (int x = threadIdx.x; x < S; x += blockDim.x) sx += m[x];
    __shared__ float red[2][32];
    for (int o = 16; o > 0; o >>= 1) {
        sy += __shfl_xor_sync(kFullMask, sy, o);
        sx += __shfl_xor_sync(kFullMask, sx, o);
    }
    if ((threadIdx.x & 31) == 0) {
        red[0][threadIdx.x >> 5] = sy;
        red[1][threadIdx.x >> 5] = sx;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        float ty = 0.f, tx = 0.f;
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) {
            ty += red[0][i];
            tx += red[1][i];
        }
        t_ys[b] = (int32_t)ty;  // masks are 0/1, sums are exact integers
        t_xs[b] = (int32_t)tx;
    }
}

// longest-first launch order (length bucketing): rank by t_y descending, stable
__global__ void mas_order_kernel(const int32_t *__restrict__ t_ys, int32_t *order, int B)
{
    for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < B; b += gridDim.x * blockDim.x) {
        const int mine = t_ys[b];
        int rank = 0;
        for (int o = 0; o < B; ++o) {
            const int other = t_ys[o];
            rank += (other > mine) || (other == mine && o < b);
        }
        order[rank] = b;
    }
}

// compact idx [B,T] -> dense path (multi-GPU re-expansion)
__global__ void mas_expand_kernel(const int32_t *__restrict__ idx, unsigned char *path, int path_dtype, int T, int S)
{
    const size_t row = blockIdx.x;  // b*T + y
    const int c = idx[row];
    const int esize = path_elem_size(path_dtype);
    unsigned char *prow = path + row * (size_t)S * esize;
    const uint32_t one = path_one_bits(path_dtype);
    if (esize == 4) {
        uint32_t *q = reinterpret_cast<uint32_t *>(prow);
        for (int x = threadIdx.x; x < S; x += blockDim.x) q[x] = (x == c) ? one : 0u;
    } else {
        uint16_t *q = reinterpret_cast<uint16_t *>(prow);
        for (int x = threadIdx.x; x < S; x += blockDim.x) q[x] = (x == c) ? (uint16_t)one : (uint16_t)0;
    }
}

// ---------------------------------------------------------------------------
// host entry points used by mas_api.cu
// ---------------------------------------------------------------------------

size_t dp_workspace_bytes(int B, int T, int S)
{
    // the larger of the plans without / with a noise tile per stage (fewer stages may move the bits off chip)
    size_t bits = 0, hop = 0;
    for (int variant = 0; variant < 4; ++variant) {
        const bool noise = (variant & 1) != 0;
        const int ld = (variant & 2) ? ((S + 3) & ~3) : S;   // caller's plane / the fused kernel's padded private plane
        DpPlan pl{};
        if (!dp_make_plan(pl, B, T, S, ld, config().dp_stages, 0, kSmemBudget, noise, false)) {
            if (variant == 0) return 0;
            continue;
        }
        bits = pl.ws_bits_bytes > bits ? pl.ws_bits_bytes : bits;
        hop = pl.ws_hop_bytes > hop ? pl.ws_hop_bytes : hop;
    }
    return align_up((size_t)B * 4, 256) + align_up(bits, 256) + align_up(hop, 256);
}

// fills pl (shared-memory plan + parameters) without launching; `order_out` receives the workspace slot
// of the launch-order array
int dp_prepare(DpPlan &pl, const float *neg_cent, const int32_t *t_ys, const int32_t *t_xs, void *path_out,
               int path_dtype, int32_t *dur_out, int32_t *idx_out, int32_t *status_out, void *workspace,
               size_t workspace_bytes, int B, int T, int S, int32_t **order_out, int R, size_t smem_budget,
               bool with_noise, int ld, int help)
{
    pl = DpPlan{};
    if (ld <= 0) ld = S;   // cost rows packed like the caller's [B,T,S] tensor
    // value / origin warp split (MAS_DP_VK=0 turns it off): W value warps run the recursion and leave the decision
    // words, W origin warps one step behind replay them into origins / hops / checkpoints.  S <= 256 with 16-byte
    // rows, no noise, everything on chip; other shapes keep the single-role warps.
    bool vk = config().dp_vk && !with_noise && R == 0 && S <= 256 && ld % 4 == 0 &&
              (reinterpret_cast<uintptr_t>(neg_cent) & 15) == 0 && dp_chunk_rows(S) == 32 &&
              (dp_team_warps(S) == 2 || (S + 127) / 128 == 2);   // 4 value warps: C = 2 columns per thread only
    if (vk) vk = dp_make_plan(pl, B, T, S, ld, 0, 0, smem_budget ? smem_budget : (size_t)kSmemBudget, false, true, help);
    if (help > 0 && !vk) return MAS_ERR_UNSUPPORTED_SHAPE;   // the helper warps come with the value / origin split
    if (!vk && !dp_make_plan(pl, B, T, S, ld, config().dp_stages, R,
                             smem_budget ? smem_budget : (size_t)kSmemBudget, with_noise, false))
        return MAS_ERR_UNSUPPORTED_SHAPE;
    const size_t need = dp_workspace_bytes(B, T, S);
    if (need && (!workspace || workspace_bytes < need)) return MAS_ERR_WORKSPACE;
    unsigned char *ws = static_cast<unsigned char *>(workspace);
    DpParams &p = pl.p;
    p.neg_cent = neg_cent;
    p.noise = nullptr;
    p.stats = nullptr;
    p.noise_scale = 0.f;
    p.t_ys = t_ys;
    p.t_xs = t_xs;
    p.path = static_cast<unsigned char *>(path_out);
    p.dur = dur_out;
    p.idx = idx_out;
    p.status = status_out;
    p.flags = nullptr;
    p.flag_tiles = 0;
    p.flag_need = 1;
    p.zero_flags = nullptr;
    p.zero_queue = nullptr;
    p.trace = trace_buffer();
    p.debug = config().dp_debug;
    p.order = nullptr;
    p.B = B;
    p.T = T;
    p.S = S;
    p.ld = ld;
    p.path_dtype = path_dtype;
    if (order_out) *order_out = reinterpret_cast<int32_t *>(ws);
    ws += align_up((size_t)B * 4, 256);
    p.bits_ws = reinterpret_cast<uint32_t *>(ws);
    ws += align_up(pl.ws_bits_bytes, 256);
    p.hop_ws = ws;
    return MAS_OK;
}

// noise applied inside the DP needs 16-byte rows and pointers (vector cost loads)
bool dp_noise_supported(const float *neg_cent, const float *noise, int S)
{
    return (S % 4 == 0) && ((reinterpret_cast<uintptr_t>(neg_cent) | reinterpret_cast<uintptr_t>(noise)) & 15) == 0;
}

// noise != nullptr: align neg_cent + (std * noise) * noise_scale without materialising it; stats = {sum, sum of
// squares} of all cost cells on the device (the contraction's epilogue wrote them)
int dp_launch(const float *neg_cent, const int32_t *t_ys, const int32_t *t_xs, void *path_out, int path_dtype,
              int32_t *dur_out, int32_t *idx_out, int32_t *status_out, void *workspace, size_t workspace_bytes, int B,
              int T, int S, cudaStream_t stream, const float *noise, const double *stats, float noise_scale, int ld)
{
    DpPlan pl;
    int32_t *order = nullptr;
    if (noise && ld > 0 && ld != S) return MAS_ERR_UNSUPPORTED_SHAPE;   // the noise rows are packed
    int rc = dp_prepare(pl, neg_cent, t_ys, t_xs, path_out, path_dtype, dur_out, idx_out, status_out, workspace,
                        workspace_bytes, B, T, S, &order, 0, 0, noise != nullptr, ld, 0);
    if (rc) return rc;
    DpParams &p = pl.p;
    p.noise = noise;
    p.stats = stats;
    p.noise_scale = noise_scale;
    // Length bucketing: when the batch is more than one wave of CTAs, launch the
    // longest utterances first so short ones fill in behind them.
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (B > sms) {
        mas_order_kernel<<<(B + 255) / 256, 256, 0, stream>>>(t_ys, order, B);
        note_launch();
        MAS_CUDA_TRY(cudaGetLastError());
        p.order = order;
    } else {
        p.order = nullptr;
    }
    // (columns per thread, chunk rows, DP warps): W = 2 covers S <= 512, W = 4 the rest
    rc = dp_dispatch_narrow(pl, pl.C, stream);
    if (rc == kDpNoCase) rc = dp_dispatch_wide(pl, pl.C, stream);
    if (rc == kDpNoCase) rc = dp_dispatch_tall(pl, pl.C, stream);
    return rc == kDpNoCase ? MAS_ERR_UNSUPPORTED_SHAPE : rc;
}

int dp_dispatch_narrow(const DpPlan &pl, int C, cudaStream_t stream)
{
    const DpParams &p = pl.p;
#define MAS_DP_CASE(CC, RR, WW) \
    if (C == CC && p.R == RR && p.W == WW) return launch_dp_c<CC, RR, WW>(pl, stream);
    MAS_DP_CASE(1, 32, 2) MAS_DP_CASE(2, 32, 2) MAS_DP_CASE(3, 32, 2) MAS_DP_CASE(4, 32, 2)
    MAS_DP_CASE(5, 16, 2) MAS_DP_CASE(6, 16, 2) MAS_DP_CASE(7, 16, 2) MAS_DP_CASE(8, 16, 2)
#undef MAS_DP_CASE
    return kDpNoCase;
}

int lengths_launch(const float *mask, int32_t *t_ys, int32_t *t_xs, int B, int T, int S, cudaStream_t stream)
{
    mas_lengths_kernel<<<B, 256, 0, stream>>>(mask, t_ys, t_xs, T, S);
    note_launch();
    MAS_CUDA_TRY(cudaGetLastError());
    return MAS_OK;
}

int expand_launch(const int32_t *idx, void *path_out, int path_dtype, int B, int T, int S, cudaStream_t stream)
{
    mas_expand_kernel<<<(unsigned)((size_t)B * T), 128, 0, stream>>>(idx, static_cast<unsigned char *>(path_out),
                                                                      path_dtype, T, S);
    note_launch();
    MAS_CUDA_TRY(cudaGetLastError());
    return MAS_OK;
}

}  // namespace mas
