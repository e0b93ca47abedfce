// mas_dp.cuh -- monotonic alignment search (forward DP + backtrack) for sm_100a, as a device-side
// "role" run by the standalone kernel (mas_dp.cu) and by the DP CTAs of the fused kernel (mas_fused.cu).
//
// Replaces maximum_path_each / maximum_path_c of the reference
// (vits2/monotonic_align/core.pyx:7-42) and the zero-init / dtype cast of its
// Python wrapper (vits2/monotonic_align/__init__.py:14,19).  Bit-exact: per
// cell one fp32 compare-select and one fp32 add, in the reference's operand
// order; no FMA, no reassociation.
//
// Mapping (one CTA per utterance, 4 DP warps + 1 producer warp):
//   * text columns are blocked over the 128 DP threads, C = ceil(S/128)
//     consecutive columns per thread; the running DP row lives in registers.
//   * the left neighbour crosses lanes with one __shfl_up per row and crosses
//     warps through a small shared-memory ring; warp w runs one chunk of R mel
//     rows behind warp w-1 (wavefront), one named barrier per chunk step.  Only every
//     C-th link of the dependency chain crosses a lane, so a row costs about
//     (shuffle + C * (max + add)) / C cycles of latency.
//   * the producer warp streams the cost plane through a ring of shared-memory
//     tiles with 1-D TMA bulk copies (cp.async.bulk + mbarrier) and zero-fills
//     the dense path output with TMA bulk stores while the DP runs.
//   * 1 bit per cell ("took the diagonal") is packed by each thread into one
//     32-bit word per column per 32 rows (a predicated OR with an immediate) and
//     kept on chip so the backtrack never re-reads the cost; every 32 rows the
//     column each cell backtracks to ("hop") is checkpointed, so the backtrack is
//     T/32 dependent hops followed by T/32 independent 32-row walks, one per thread.
//   * long utterances whose bits/hops exceed shared memory spill them to the
//     caller's workspace (L2-resident).
#pragma once

#include "mas_common.cuh"

namespace mas {

// A DP team = W DP warps + 1 producer warp.  W = 2 (4 x more columns per thread than lanes need, i.e.
// C = ceil(S / 64)) is the default: the C cells of a row are independent of each other, so a wider
// thread has the instruction-level parallelism to hide the max -> add latency, and only every C-th
// link of the dependency chain pays for a shuffle.  W = 4 serves S > 512.
constexpr int kMaxDpWarps = 4;
__host__ __device__ constexpr int dp_threads(int W, bool vk = false) { return ((vk ? 2 * W : W) + 1) * 32; }  // + producer warp
constexpr int kMaxStages = 8;
constexpr int kCheck = 32;  // checkpoint interval (rows)
constexpr int kSmemBudget = 227 * 1024;
constexpr int kDpBar = 3;     // named barrier of the DP role (kThreads threads)
constexpr int kZeroBytes = 8192;  // zeroed shared buffer the path zero-fill bulk-stores from
constexpr uint32_t kZeroFillBuf = 32768;  // zero page of the zero-fill role (contraction CTAs that ran out of tiles)
constexpr int kZeroParts = 8;             // slices per path plane handed out by the zero-fill role
constexpr int kHelpAhead = 4;             // noise chunks kept ahead in L2 by the helpers

// mel rows per chunk (= per TMA tile) for S text columns: a stage stays <= 32 KB
constexpr int kBitsPad = 4;   // words between decision-word rows beyond S_pad (keeps 16-byte alignment)
__host__ __device__ constexpr int dp_chunk_rows(int S) { return S <= 256 ? 32 : (S <= 512 ? 16 : 8); }

struct DpParams {
    const float *neg_cent;
    const float *noise;      // nullable: VITS2 noise draw [B,T,S]; the DP then aligns neg_cent + (std * noise) * noise_scale
    const double *stats;     // with noise: {sum, sum of squares} of all B*T*S cost cells (device)
    float noise_scale;
    uint32_t noise_off;      // with noise: byte offset of the noise tile inside a stage
    const int32_t *t_ys;
    const int32_t *t_xs;
    const int32_t *order;  // nullable: CTA -> utterance (longest first)
    unsigned char *path;    // nullable: compact outputs only (no zero fill, no scatter of ones)
    int32_t *dur;
    int32_t *idx;
    int32_t *status;
    uint32_t *flags;        // nullable: [B][flag_tiles] cost-tile-ready flags of the fused kernel (consumed and reset here)
    int flag_tiles;         // mel tiles of 128 rows per utterance
    int flag_need;          // publications that complete a tile (column blocks of the contraction: 1 for S <= 256)
    uint32_t *zero_queue;   // fused kernel: work counter of the zero-fill role (cleared with the flags)
    uint32_t *zero_flags;   // nullable: [B] set by whoever zero-fills the path plane of an utterance (fused kernel: the
                            // contraction CTAs, once they run out of tiles); consumed and reset here
    unsigned long long *trace;  // nullable diagnostics buffer: [12288 + utterance * 32]: start, tile acquire times, ends
    uint32_t *bits_ws;      // global spill (per CTA region), used when !bits_in_smem
    unsigned char *hop_ws;  // global spill (per CTA region), used when !hop_in_smem
    int B, T, S;
    int ld;      // row stride of neg_cent in floats (S, or S rounded up to 4 for the fused kernel's private plane)
    int R;       // mel rows per chunk (template parameter of the role; 32, 16 or 8)
    int W;       // DP warps per team (template parameter of the role; 2 or 4)
    int vk;      // value / origin warp split (W value warps + W origin warps + producer)
    int fed;     // > 0: the cost tiles are pushed into the stage ring by the partner CTA of the cluster (noise feeder,
                 // mas_fused.cu) with asynchronous stores counted on the stage's full barrier; the producer warp only
                 // arms the barriers with the tiles' byte counts and returns one credit per consumed stage
    uint32_t fed_credit_off;   // byte offset of the feeder's credit barriers [kMaxStages] in ITS dynamic shared memory
    uint32_t fed_verdict_off;  // ... of its {verdict barrier (8 bytes), verdict word}
    int help;    // noise-helper warps behind the producer warp (fused noise kernel): they add (std * noise) * scale to
                 // every cost tile in shared memory one chunk step ahead of the value warps
    int stages;  // cost-tile ring depth (2..kMaxStages)
    int path_dtype;
    int debug;   // MAS_DP_DEBUG bit mask (timing experiments): 1 no zero fill, 2 no forward compute, 4 no cost loads,
                 // 16 / 32 warp split: bookkeeping / value warps idle
    int bits_in_smem, hop_in_smem;
    uint32_t off_nz;   // helper warps: the one-chunk noise buffer (behind the cost stages)
    uint32_t off_bits, off_hop, off_stage, stage_bytes, off_bnd_v, off_bnd_o, off_idx, off_end, off_entry, off_bar, off_zero, off_misc;
    unsigned long long bits_words_per_cta, hop_bytes_per_cta;
};

struct DpPlan {
    DpParams p;
    int C;
    size_t smem_bytes;
    size_t ws_bits_bytes, ws_hop_bytes;
};

// ---------------------------------------------------------------------------
// device
// ---------------------------------------------------------------------------
__device__ __forceinline__ void zero_bytes_warp(unsigned char *ptr, size_t bytes, int lane)
{
    // generic zero fill by one warp: 2-byte head/tail, 16-byte body
    size_t head = (16 - (reinterpret_cast<uintptr_t>(ptr) & 15)) & 15;
    if (head > bytes) head = bytes;
    for (size_t i = lane * 2; i < head; i += 64) *reinterpret_cast<uint16_t *>(ptr + i) = 0;
    size_t body = (bytes - head) & ~(size_t)15;
    for (size_t i = (size_t)lane * 16; i < body; i += 512) st_global_v4_zero(ptr + head + i);
    for (size_t i = head + body + lane * 2; i < bytes; i += 64) *reinterpret_cast<uint16_t *>(ptr + i) = 0;
}

__device__ __forceinline__ void store_one(unsigned char *path_b, size_t cell, int path_dtype)
{
    if (path_dtype == MAS_PATH_F32)
        reinterpret_cast<float *>(path_b)[cell] = 1.0f;
    else if (path_dtype == MAS_PATH_I32)
        reinterpret_cast<int32_t *>(path_b)[cell] = 1;
    else
        reinterpret_cast<uint16_t *>(path_b)[cell] = (uint16_t)path_one_bits(path_dtype);
}

// NR mel rows (4 in the main loop, 1 for the tail) of the forward DP for this
// thread's C consecutive columns.
//   v/org   running DP row and checkpoint origin per column (registers)
//   fin     turns NaN/Inf as soon as one cost was not finite (fast path only)
//   wl      decision bits of this chunk, bit (bit0 + i) = row i of this call
//   trow    &tile[r * S + x0]: this thread's first cost of the first row
//   carry   value/origin of column x0-1 after the previous row (lane 0 only matters)
//   bin_*   boundary ring written by the warp to the left, slot of row y
//   bout_*  this warp's boundary ring, slot of row y (lane 31 writes)
//   kExact  false: max via FMNMX (short dependency chain); exact while every cost is
//           finite, which `fin` tracks.
//           true: the reference's compare-select, bit-for-bit also for NaN/Inf input.
//   kNoise  the stage also holds the noise tile (nz_off floats behind the cost tile): the cost of a cell is
//           nc + (sd * noise) * scale, rounded after every operation like the reference (models.py:1241-1247)
struct DpNoise {
    int nz_off;
    float sd, scale;
};
//   kOrg    false: no origin tracking here (warp split: the origin warps rebuild it from the decision words)
template <int C, int NR, bool kEdge, bool kVec, bool kExact, bool kNoise, bool kOrg = true>
__device__ __forceinline__ void dp_rows(float (&v)[C], int (&org)[C], float &fin, uint32_t (&wl)[C], const float *trow,
                                        int ld, int bit0, float &carry_v, int &carry_o, const float *bin_v,
                                        const int *bin_o, float *bout_v, int *bout_o, int y, int x0, bool lane0,
                                        bool lane31, const DpNoise &nz)
{
    float cost[NR][C];
    float lv[NR + 1];
    int lo[NR + 1];
    lv[0] = carry_v;
    lo[0] = carry_o;
    // ---- all loads for the NR rows up front
#pragma unroll
    for (int i = 0; i < NR; ++i) {
        const float *src = trow + (size_t)i * ld;
        if (kVec && (C % 4 == 0)) {
#pragma unroll
            for (int k = 0; k < C; k += 4) {
                const float4 t = *reinterpret_cast<const float4 *>(src + k);
                cost[i][k] = t.x, cost[i][k + 1] = t.y, cost[i][k + 2] = t.z, cost[i][k + 3] = t.w;
            }
        } else if (kVec && (C % 2 == 0)) {
#pragma unroll
            for (int k = 0; k < C; k += 2) {
                const float2 t = *reinterpret_cast<const float2 *>(src + k);
                cost[i][k] = t.x, cost[i][k + 1] = t.y;
            }
        } else {
#pragma unroll
            for (int k = 0; k < C; ++k) cost[i][k] = src[k];
        }
        if (kNoise) {
            const float *nsrc = src + nz.nz_off;
            float n_[C];
            if (kVec && C % 4 == 0) {
#pragma unroll
                for (int k = 0; k < C; k += 4) {
                    const float4 t = *reinterpret_cast<const float4 *>(nsrc + k);
                    n_[k] = t.x, n_[k + 1] = t.y, n_[k + 2] = t.z, n_[k + 3] = t.w;
                }
            } else if (kVec && C % 2 == 0) {
#pragma unroll
                for (int k = 0; k < C; k += 2) {
                    const float2 t = *reinterpret_cast<const float2 *>(nsrc + k);
                    n_[k] = t.x, n_[k + 1] = t.y;
                }
            } else {
#pragma unroll
                for (int k = 0; k < C; ++k) n_[k] = nsrc[k];
            }
#pragma unroll
            for (int k = 0; k < C; ++k)
                cost[i][k] = __fadd_rn(cost[i][k], __fmul_rn(__fmul_rn(nz.sd, n_[k]), nz.scale));
        }
    }
    if (NR == 4) {
        const float4 t = *reinterpret_cast<const float4 *>(bin_v);
        lv[1] = t.x, lv[2] = t.y, lv[3] = t.z, lv[4] = t.w;
        if (kOrg) {
            const int4 u = *reinterpret_cast<const int4 *>(bin_o);
            lo[1] = u.x, lo[2] = u.y, lo[3] = u.z, lo[4] = u.w;
        }
    } else {
#pragma unroll
        for (int i = 0; i < NR; ++i) {
            lv[i + 1] = bin_v[i];
            if (kOrg) lo[i + 1] = bin_o[i];
        }
    }
    float ov[NR];
    int oo[NR];
#pragma unroll
    for (int i = 0; i < NR; ++i) {
        float up_v = __shfl_up_sync(kFullMask, v[C - 1], 1);
        int up_o = 0;
        if (kOrg) up_o = __shfl_up_sync(kFullMask, org[C - 1], 1);
        if (lane0) {
            up_v = lv[i];
            if (kOrg) up_o = lo[i];
        }
        if (!kExact) {
            // non-finite detection: a product / multiply-add is NaN or Inf as soon as a factor is
            // (a finite overflow is a harmless false positive: it only selects the exact pass)
#pragma unroll
            for (int k = 0; k + 1 < C; k += 2) fin = fmaf(cost[i][k], cost[i][k + 1], fin);
            if (C & 1) fin = fmaf(cost[i][C - 1], 0.0f, fin);
        }
#pragma unroll
        for (int k = C - 1; k >= 0; --k) {
            const float v_prev = (k == 0) ? up_v : v[k - 1];  // value[y-1, x-1]  (core.pyx:21-27)
            const int o_prev = kOrg ? ((k == 0) ? up_o : org[k - 1]) : 0;
            const float v_cur = v[k];                          // value[y-1, x]    (core.pyx:17-20)
            // Cython's max(v_prev, v_cur) is (v_cur > v_prev) ? v_cur : v_prev; with no NaN in
            // flight that is fmaxf (one FMNMX instead of FSETP -> FSEL on the dependency chain)
            float m;
            if (kExact)
                m = (v_cur > v_prev) ? v_cur : v_prev;
            else
                m = fmaxf(v_prev, v_cur);
            // backtrack rule, core.pyx:32: index == y or value[y-1,x] < value[y-1,x-1]
            // (the "index != 0" guard is applied by the backtrack itself)
            bool diag = v_cur < v_prev;
            if (kEdge) diag = diag || (x0 + k == y + i);
            float nv = cost[i][k] + m;  // core.pyx:28
            int no = 0;
            if (kOrg) no = diag ? o_prev : org[k];
            if (kEdge) {
                const bool in_band = (x0 + k <= y + i);  // upper band edge, core.pyx:16
                nv = in_band ? nv : v_cur;
                if (kOrg) no = in_band ? no : org[k];
            }
            v[k] = nv;
            if (kOrg) org[k] = no;
            if (diag) wl[k] |= 1u << (bit0 + i);
        }
        ov[i] = v[C - 1];
        oo[i] = kOrg ? org[C - 1] : 0;
    }
    if (lane31) {
        if (NR == 4) {
            *reinterpret_cast<float4 *>(bout_v) = make_float4(ov[0], ov[1], ov[2], ov[3]);
            if (kOrg) *reinterpret_cast<int4 *>(bout_o) = make_int4(oo[0], oo[1], oo[2], oo[3]);
        } else {
#pragma unroll
            for (int i = 0; i < NR; ++i) {
                bout_v[i] = ov[i];
                if (kOrg) bout_o[i] = oo[i];
            }
        }
    }
    carry_v = lv[NR];
    if (kOrg) carry_o = lo[NR];
}

// one chunk (<= R rows) of this warp's columns; bin/bout point at the ring slot of the chunk's first row
template <int C, int R, bool kEdge, bool kVec, bool kExact, bool kNoise, bool kOrg = true>
__device__ __forceinline__ void dp_chunk(float (&v)[C], int (&org)[C], float &fin, uint32_t (&wl)[C], const float *tile,
                                         int S, int ld, int rows, int row0, float &carry_v, int &carry_o, const float *bin_v,
                                         const int *bin_o, float *bout_v, int *bout_o, int x0, bool lane0, bool lane31,
                                         const DpNoise &nz)
{
    // Threads whose columns lie past the row (ld = S, or S rounded up to 4 with zeros in the pad columns) re-read
    // its last C columns instead of whatever follows: their results are never used, but they must not invent
    // NaNs.  (A thread straddling ld reads at most C-1 floats of the next row, or of the zeroed pad the
    // producer keeps behind every tile.)
    const float *trow = tile + (x0 < ld ? x0 : ld - C);
    // 8 rows per trip: small enough to stay in the instruction cache (a fully unrolled 32-row body is
    // ~10 KB per variant and its cold fetch cost more than the rows themselves), large enough that the
    // decision bits are still set with immediates; the 8-bit group is merged into the chunk word per trip.
    int r = 0;
#pragma unroll 1
    for (; r + 8 <= rows; r += 8) {
        uint32_t w8[C];
#pragma unroll
        for (int k = 0; k < C; ++k) w8[k] = 0u;
        dp_rows<C, 4, kEdge, kVec, kExact, kNoise, kOrg>(v, org, fin, w8, trow + (size_t)r * ld, ld, 0, carry_v, carry_o, bin_v + r,
                                           bin_o + r, bout_v + r, bout_o + r, row0 + r, x0, lane0, lane31, nz);
        dp_rows<C, 4, kEdge, kVec, kExact, kNoise, kOrg>(v, org, fin, w8, trow + (size_t)(r + 4) * ld, ld, 4, carry_v, carry_o,
                                           bin_v + r + 4, bin_o + r + 4, bout_v + r + 4, bout_o + r + 4, row0 + r + 4, x0,
                                           lane0, lane31, nz);
#pragma unroll
        for (int k = 0; k < C; ++k) wl[k] |= w8[k] << r;
    }
#pragma unroll 1
    for (; r < rows; ++r) {
        uint32_t w8[C];
#pragma unroll
        for (int k = 0; k < C; ++k) w8[k] = 0u;
        dp_rows<C, 1, kEdge, kVec, kExact, kNoise, kOrg>(v, org, fin, w8, trow + (size_t)r * ld, ld, 0, carry_v, carry_o, bin_v + r,
                                           bin_o + r, bout_v + r, bout_o + r, row0 + r, x0, lane0, lane31, nz);
#pragma unroll
        for (int k = 0; k < C; ++k) wl[k] |= w8[k] << r;
    }
}

// ---------------------------------------------------------------------------
// value / origin warp split (kVK): the value warps run dp_rows without origins; the origin warps replay
// the decision words (bit r of wd[k] = row row0 + r took the diagonal, core.pyx:32) into the origins.
// Same boundary protocol as dp_rows: bin_o / bout_o slots per row, carry_o across calls.
// ---------------------------------------------------------------------------
template <int C, int NR, bool kEdge>
__device__ __forceinline__ void org_rows(int (&org)[C], const uint32_t (&w8)[C], int bit0, int &carry_o, const int *bin_o,
                                         int *bout_o, int y, int x0, bool lane0, bool lane31)
{
    int lo[NR + 1];
    lo[0] = carry_o;
    if (NR == 4) {
        const int4 u = *reinterpret_cast<const int4 *>(bin_o);
        lo[1] = u.x, lo[2] = u.y, lo[3] = u.z, lo[4] = u.w;
    } else {
#pragma unroll
        for (int i = 0; i < NR; ++i) lo[i + 1] = bin_o[i];
    }
    int oo[NR];
#pragma unroll
    for (int i = 0; i < NR; ++i) {
        int up_o = __shfl_up_sync(kFullMask, org[C - 1], 1);
        if (lane0) up_o = lo[i];
#pragma unroll
        for (int k = C - 1; k >= 0; --k) {
            const int o_prev = (k == 0) ? up_o : org[k - 1];
            bool diag = (w8[k] >> (bit0 + i)) & 1u;
            if (kEdge) diag = diag && (x0 + k <= y + i);  // above the band nothing moves (core.pyx:16)
            org[k] = diag ? o_prev : org[k];
        }
        oo[i] = org[C - 1];
    }
    if (lane31) {
        if (NR == 4) {
            *reinterpret_cast<int4 *>(bout_o) = make_int4(oo[0], oo[1], oo[2], oo[3]);
        } else {
#pragma unroll
            for (int i = 0; i < NR; ++i) bout_o[i] = oo[i];
        }
    }
    carry_o = lo[NR];
}

template <int C, bool kEdge>
__device__ __forceinline__ void org_chunk(int (&org)[C], const uint32_t (&wd)[C], int rows, int row0, int &carry_o,
                                          const int *bin_o, int *bout_o, int x0, bool lane0, bool lane31)
{
    int r = 0;
#pragma unroll 1
    for (; r + 8 <= rows; r += 8) {
        uint32_t w8[C];
#pragma unroll
        for (int k = 0; k < C; ++k) w8[k] = wd[k] >> r;
        org_rows<C, 4, kEdge>(org, w8, 0, carry_o, bin_o + r, bout_o + r, row0 + r, x0, lane0, lane31);
        org_rows<C, 4, kEdge>(org, w8, 4, carry_o, bin_o + r + 4, bout_o + r + 4, row0 + r + 4, x0, lane0, lane31);
    }
#pragma unroll 1
    for (; r < rows; ++r) {
        uint32_t w8[C];
#pragma unroll
        for (int k = 0; k < C; ++k) w8[k] = wd[k] >> r;
        org_rows<C, 1, kEdge>(org, w8, 0, carry_o, bin_o + r, bout_o + r, row0 + r, x0, lane0, lane31);
    }
}

// checkpoint row c_j: c_0 = 0, c_j = 32 j - 1
__device__ __forceinline__ int check_row(int j) { return j == 0 ? 0 : kCheck * j - 1; }

// one-time setup of a DP team (mbarriers, the constant boundary ring, the zero page).  A team is
// dp_threads(W) consecutive threads (tid counts inside the team) with its own shared-memory region and
// named barrier `bar`; the standalone kernel runs one team per CTA, the fused kernel up to two.
__device__ __forceinline__ void dp_role_init(const DpParams &p, unsigned char *smem, int tid, int bar)
{
    const int nthr = dp_threads(p.W, p.vk != 0) + 32 * p.help;
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + p.off_bar);
    float *bnd_v = reinterpret_cast<float *>(smem + p.off_bnd_v);
    int *bnd_o = reinterpret_cast<int *>(smem + p.off_bnd_o);
    unsigned char *zero_s = smem + p.off_zero;
    if (tid == 0) {
        for (int s = 0; s < kMaxStages; ++s) mbar_init(&full[s], 1);
        if (p.help > 1)
            for (int h = 0; h < 2; ++h) {
                mbar_init(&full[kMaxStages + h], 1);                 // nfull: noise half landed
                mbar_init(&full[kMaxStages + 2 + h], p.help - 1);    // nfree: read by every applier warp
            }
        fence_mbar_init();
    }
    // ring 0 stands in for "the warp left of warp 0": column -1 is the -1e9 sentinel (core.pyx:24)
    for (int i = tid; i < 2 * p.R; i += nthr) {
        bnd_v[i] = kNeg;
        bnd_o[i] = 0;
    }
    for (int i = tid; i < kZeroBytes / 16; i += nthr) reinterpret_cast<uint4 *>(zero_s)[i] = make_uint4(0, 0, 0, 0);
    fence_proxy_async();  // zero_s is read by the bulk-store engine
    bar_sync(bar, nthr);
}

// The noise-helper warps of a DP team (fused noise kernel; dp_role with kHelp > 0).  VITS2 aligns
// neg_cent + (std * randn) * scale (models.py:1241-1247); the helpers add that term to every cost tile IN PLACE in
// shared memory one chunk step ahead of the value warps, which therefore run the plain body.
//   * The last helper warp is the noise producer: it streams the draw into a one-chunk shared buffer in two halves
//     of R/2 rows (bulk copies, nfull[h]); the other warps add (std * noise) * scale to the cost tile and hand each
//     half back (nfree[h]) as soon as they have read it, so the next chunk's half is on its way while this step
//     is still running.  (Its own warp: on the cost-tile producer's lane the two waits sat behind ~1300 cycles of
//     tile issue and every step paid for both.)
//   * Loads and stores are batched by hand (all loads of a half, then the arithmetic, then the stores): the
//     compiler must assume that the tile and the noise buffer alias and would otherwise serialise item by item.
//   * Tried and dropped: the draw straight from global memory into registers (no shared buffer).  One chunk ahead
//     with double-buffered registers serialised on the scoreboard; two warp groups on alternate chunks still saw
//     the loads arrive late; both ran 4600-5100 cycles per chunk step against 3100 for this version.
//   * A separate function on purpose: its registers must not weigh on the allocation of the value warps' loop.
template <int R, int kHelp>
__device__ __noinline__ void dp_noise_helper(const DpParams &p, unsigned char *smem, int b, int hw, int lane, uint32_t g0,
                                             int n_chunks, int n_steps, int t_y, int bar, int nthreads)
{
    static_assert(kHelp >= 2, "one producer warp + at least one applier warp");
    constexpr int HT = 32 * (kHelp - 1);                      // applier threads
    constexpr int KQ = ((R / 2) * 64 + HT - 1) / HT;          // 16-byte items per applier and half at the widest plane (256 floats)
    const int ld = p.ld, ld4 = ld >> 2;                       // (ld == S: 16-byte rows, checked by the launcher)
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + p.off_bar);
    uint64_t *nfull = full + kMaxStages;      // [2] noise half landed
    uint64_t *nfree = full + kMaxStages + 2;  // [2] noise half read by all applier warps
    const size_t plane = (size_t)p.T * p.S;
    const uint32_t n_stages = (uint32_t)p.stages;
    const uint32_t half_stride = (uint32_t)(R / 2) * ld * 4 + 32;   // room for the misaligned start and the 16-byte round-up
    const bool nvec = (p.S & 3) == 0 && (reinterpret_cast<uintptr_t>(p.noise) & 15) == 0;   // then ld == S and every half starts aligned

    if (hw == kHelp - 1) {
        // ---- noise producer ----
        const uintptr_t nz_b = reinterpret_cast<uintptr_t>(p.noise + (size_t)b * plane);
        const uintptr_t nz_end = reinterpret_cast<uintptr_t>(p.noise + (size_t)p.B * plane);
        auto issue = [&](int c) {
            if (c >= n_chunks) return;
            const uint32_t g = g0 + (uint32_t)c;
            const int rows = min(R, t_y - c * R);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int rows_h = h == 0 ? min(rows, R / 2) : max(rows - R / 2, 0);
                // the half's rows are rows_h * S consecutive floats of the draw; bulk copies move 16-byte units, so
                // copy the aligned superset (the appliers skip the `mis` leading bytes) and finish a tail that would
                // run past the end of the tensor by hand
                const uintptr_t addr = nz_b + ((size_t)c * R + (size_t)h * (R / 2)) * p.S * 4;
                const uint32_t mis = (uint32_t)(addr & 15);
                const uintptr_t src0 = addr - mis;
                const uint32_t want = rows_h ? mis + (uint32_t)rows_h * p.S * 4 : 0u;
                uint32_t bulk = (want + 15u) & ~15u;
                unsigned char *dst = smem + p.off_nz + h * half_stride;
                mbar_wait(&nfree[h], (g & 1u) ^ 1u);              // every applier warp has read the previous contents
                if (src0 + bulk > nz_end) {
                    bulk = want & ~15u;
                    for (uint32_t o = bulk; o < want; o += 4)
                        *reinterpret_cast<float *>(dst + o) = *reinterpret_cast<const float *>(src0 + o);
                }
                mbar_arrive_expect_tx(&nfull[h], bulk);           // (an empty half still completes its phase)
                if (bulk) bulk_g2s(dst, reinterpret_cast<const void *>(src0), bulk, &nfull[h]);
            }
        };
        if (lane == 0) issue(0);
        __syncwarp();
        for (int step = 0; step < n_steps; ++step) {
            if (lane == 0) issue(step + 1);   // follows the appliers through chunk `step`
            __syncwarp();
            bar_sync(bar, nthreads);
        }
        return;
    }

    // ---- appliers ----
    const int gl = hw * 32 + lane;
    // unbiased std over ALL cells, padding included (torch.std default, models.py:1243), from fp64 sums
    float sd;
    {
        const double n = (double)p.B * (double)plane;
        const double s0 = __ldcg(p.stats), s1 = __ldcg(p.stats + 1);
        const double mean = s0 / n;
        double var = (s1 - s0 * mean) / (n > 1.0 ? n - 1.0 : 1.0);
        if (var < 0) var = 0;
        sd = (float)sqrt(var);
    }
    const float scale = p.noise_scale;
    // the draw a few chunks ahead: into L2, one 128-byte line per thread and pass
    const char *nz_c = reinterpret_cast<const char *>(p.noise + (size_t)b * plane);
    auto prefetch_chunk = [&](int c) {
        if (c >= n_chunks) return;
        const char *pb = nz_c + (size_t)c * R * ld * 4;
        const int pbytes = min(R, t_y - c * R) * ld * 4;
        for (int o = gl * 128; o < pbytes; o += HT * 128) prefetch_l2(pb + o);
    };
    for (int c = 1; c <= kHelpAhead; ++c) prefetch_chunk(c);
    long long hacc[3] = {0, 0, 0};  // diagnostics: cycles waiting for the cost tile, for the noise halves, applying
    for (int step = 0; step < n_steps; ++step) {
        const int c = step;
        if (c < n_chunks) {
            prefetch_chunk(c + 1 + kHelpAhead);
            const int rows = min(R, t_y - c * R);
            const uint32_t g = g0 + (uint32_t)c;
            const uint32_t st = g % n_stages, st_par = (g / n_stages) & 1u, npar = g & 1u;
            float4 *tile4 = reinterpret_cast<float4 *>(smem + p.off_stage + (size_t)st * p.stage_bytes) + gl;
            const long long h0 = MAS_TR(p) ? clock64() : 0;
            mbar_wait(&full[st], st_par);
            long long h1 = MAS_TR(p) ? clock64() : 0;
            if (MAS_TR(p)) hacc[0] += h1 - h0;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int rows_h = h == 0 ? min(rows, R / 2) : max(rows - R / 2, 0);
                const int n4 = rows_h * ld4;
                const unsigned char *nzh = smem + p.off_nz + h * half_stride;
                float4 *t4 = tile4 + (size_t)h * (R / 2) * ld4;
                mbar_wait(&nfull[h], npar);
                const long long h2 = MAS_TR(p) ? clock64() : 0;
                if (nvec) {
                    const float4 *nz4 = reinterpret_cast<const float4 *>(nzh) + gl;
                    float4 cv[KQ], nv[KQ];
#pragma unroll
                    for (int k = 0; k < KQ; ++k)
                        if (gl + k * HT < n4) cv[k] = t4[k * HT], nv[k] = nz4[k * HT];
#pragma unroll
                    for (int k = 0; k < KQ; ++k)
                        if (gl + k * HT < n4) {
                            cv[k].x = __fadd_rn(cv[k].x, __fmul_rn(__fmul_rn(sd, nv[k].x), scale));
                            cv[k].y = __fadd_rn(cv[k].y, __fmul_rn(__fmul_rn(sd, nv[k].y), scale));
                            cv[k].z = __fadd_rn(cv[k].z, __fmul_rn(__fmul_rn(sd, nv[k].z), scale));
                            cv[k].w = __fadd_rn(cv[k].w, __fmul_rn(__fmul_rn(sd, nv[k].w), scale));
                            t4[k * HT] = cv[k];
                        }
                    // The half goes back only now: the stores above took every loaded value as an operand, so the loads
                    // are done.  (An arrive issued right after the loads can overtake them, see cost_tc_role's converters.)
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&nfree[h]);
                } else {
                    // rows that are not a multiple of 16 bytes (real collated batches: data_utils.py:151-214): the
                    // draw is packed with stride S, the private cost plane has stride ld; one element at a time
                    const uintptr_t addr = reinterpret_cast<uintptr_t>(p.noise + (size_t)b * plane) +
                                           ((size_t)c * R + (size_t)h * (R / 2)) * p.S * 4;
                    const float *nzf = reinterpret_cast<const float *>(nzh + (addr & 15));
                    float *tf = reinterpret_cast<float *>(t4 - gl);
                    for (int r = 0; r < rows_h; ++r)
                        for (int x = gl; x < p.S; x += HT) {
                            float *q = tf + r * ld + x;
                            *q = __fadd_rn(*q, __fmul_rn(__fmul_rn(sd, nzf[r * p.S + x]), scale));
                        }
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&nfree[h]);
                }
                if (MAS_TR(p)) {
                    const long long h3 = clock64();
                    hacc[1] += h2 - h1, hacc[2] += h3 - h2, h1 = h3;
                }
            }
            // generic stores into a stage the bulk-copy engine overwrites a few steps from now: order them for the
            // async proxy here, where there is slack
            fence_proxy_async();
        }
        bar_sync(bar, nthreads);
    }
    if (MAS_TR(p) && gl == 0) {
        p.trace[40960 + (size_t)b * 16 + 12] = (unsigned long long)(hacc[0] + hacc[1]);
        p.trace[40960 + (size_t)b * 16 + 13] = (unsigned long long)hacc[2];
    }
}

// Aligns utterance b.  Runs on the dp_threads(W) threads of one team; `slot` selects the team's region of
// the spill workspace; g_base is the running cost-tile counter of this CTA's stage ring (mbarrier phases
// continue across utterances).
// kHelp > 0: that many helper warps sit behind the producer warp and add VITS2's noise to every cost tile one chunk
// step ahead of the value warps (dp_noise_helper).
// (Inlined: as a separate function -- tried so that the 128-register limit of the fused kernels' 512-thread CTAs
// would not shape the value warps' loop -- the standalone kernel went from 43 to 59 us at config 2.
// Also tried and dropped: point-to-point progress counters in shared memory instead of the CTA-wide barrier after
// every chunk step (a warp polling only the neighbour it depends on).  All parity tests passed, but the two
// MEMBAR.CTA per chunk and warp cost more than the barrier's wait: maximum_path 43 -> 47 us at config 2 and
// 362 -> 474 us at config 4, where the chunks are 8-16 rows.)
// kFed: the cost tiles arrive from the partner CTA (see DpParams::fed).
template <int C, int R, int W, bool kVec, bool kNoise = false, bool kVK = false, int kHelp = 0, bool kFed = false>
__device__ __forceinline__ void dp_role(const DpParams &p, unsigned char *smem, int b, int slot, uint32_t &g_base, int tid,
                                        int bar)
{
    const int warp = tid >> 5;
    const int lane = tid & 31;
    const int T = p.T, S = p.S, ld = p.ld;
    const int t_y = p.t_ys[b], t_x = p.t_xs[b];
    constexpr int S_pad = W * 32 * C;
    constexpr int S_bits = S_pad + kBitsPad;
    static_assert(!kVK || R == kCheck, "the warp split replays whole decision words: one chunk = one word");
    static_assert(kHelp == 0 || (kVec && !kNoise), "helper warps work on 16-byte rows and replace the in-loop noise");
    static_assert(!kFed || (kVec && !kNoise && kHelp == 0), "the feeder delivers finished (noised) tiles with 16-byte rows");
    constexpr int kThreads = dp_threads(W, kVK) + 32 * kHelp;
    constexpr int kDpWarps = W;                       // warps that consume cost tiles
    constexpr int kProducerWarp = kVK ? 2 * W : W;
    constexpr int kHs = kHelp > 0 ? 1 : 0;            // the value warps trail the helpers by one chunk step
    const int esize = path_elem_size(p.path_dtype);
    const size_t plane = (size_t)T * S;
    unsigned char *path_b = p.path ? p.path + (size_t)b * plane * esize : nullptr;

    // ---- length contract (anything else is UB in the reference) ----
    if (!(t_x >= 1 && t_x <= t_y && t_y <= T && t_x <= S)) {
        if (path_b) {
            const size_t total = plane * esize;
            const size_t seg = align_up((total + kThreads / 32 - 1) / (kThreads / 32), 512);
            size_t lo = (size_t)warp * seg, hi = lo + seg;
            if (lo > total) lo = total;
            if (hi > total) hi = total;
            if (hi > lo) zero_bytes_warp(path_b + lo, hi - lo, lane);
        }
        if (p.dur)
            for (int x = tid; x < S; x += kThreads) p.dur[(size_t)b * S + x] = 0;
        if (p.idx)
            for (int y = tid; y < T; y += kThreads) p.idx[(size_t)b * T + y] = -1;
        if (p.status && tid == 0) p.status[b] = MAS_UTT_BAD_LENGTHS;
        return;
    }

    uint64_t *full = reinterpret_cast<uint64_t *>(smem + p.off_bar);
    // decision bits: word [(y >> 5) * S_bits + x], bit (y & 31); S_bits = S_pad + kBitsPad so that the backtrack's
    // walks (one thread per 32-row block, all near the same column) do not all hit one shared-memory bank
    uint32_t *bits = p.bits_in_smem ? reinterpret_cast<uint32_t *>(smem + p.off_bits)
                                    : p.bits_ws + (size_t)slot * p.bits_words_per_cta;
    unsigned char *hop = p.hop_in_smem ? (smem + p.off_hop) : p.hop_ws + (size_t)slot * p.hop_bytes_per_cta;
    float *bnd_v = reinterpret_cast<float *>(smem + p.off_bnd_v);  // [kDpWarps + 1][2R]
    int *bnd_o = reinterpret_cast<int *>(smem + p.off_bnd_o);
    uint16_t *idx_s = reinterpret_cast<uint16_t *>(smem + p.off_idx);
    // the column of mel row y sits at idx_s[ix(y)]: block-minor, so that the 32-row walks of the backtrack (one
    // thread per block, all at the same row offset) store to consecutive addresses instead of two banks
    // (short utterances only: at T = 4000 the linear layout measured 8 % faster for the whole kernel)
    const int idx_pitch = (T <= 2048) ? (T >> 5) + 2 : 0;
    auto ix = [idx_pitch](int y) { return idx_pitch ? (y & 31) * idx_pitch + (y >> 5) : y; };
    int *end_s = reinterpret_cast<int *>(smem + p.off_end);
    uint16_t *entry_s = reinterpret_cast<uint16_t *>(smem + p.off_entry);
    unsigned char *zero_s = smem + p.off_zero;

    constexpr int ring = 2 * R;
    const uint32_t n_stages = (uint32_t)p.stages;
    DpNoise nzp{0, 0.f, 0.f};
    if (kNoise) {
        // unbiased std over ALL cells, padding included (torch.std default, models.py:1243), from fp64 sums
        const double n = (double)p.B * (double)plane;
        const double mean = p.stats[0] / n;
        double var = (p.stats[1] - p.stats[0] * mean) / (n > 1.0 ? n - 1.0 : 1.0);
        if (var < 0) var = 0;
        nzp.sd = (float)sqrt(var);
        nzp.scale = p.noise_scale;
        nzp.nz_off = (int)(p.noise_off / 4);
    }
    int *nonfinite_s = reinterpret_cast<int *>(smem + p.off_misc);
    if (tid == 0) *nonfinite_s = 0;
    bar_sync(bar, kThreads);  // previous utterance's backtrack is done with the shared buffers

    if (MAS_TR(p) && tid == 0) p.trace[12288 + (size_t)b * 32 + 0] = globaltimer_ns();
    const int n_chunks = (t_y + R - 1) / R;
    const int n_steps = n_chunks + kDpWarps - 1 + (kVK ? 1 : 0) + kHs;  // the bookkeeping warps trail by one step
    const size_t utt_elem0 = (size_t)b * T * ld;  // first element of this utterance's cost plane
    const size_t total_bytes = (size_t)p.B * T * ld * 4;

    // Pass 0 runs the short-chain body; only if it met a non-finite cost does pass 1 redo the
    // forward DP with the reference's exact compare-select (NaN/Inf semantics of core.pyx).
    int passes = 0;
    for (int pass = 0; pass < 2; ++pass) {
        const uint32_t g0 = g_base + (uint32_t)pass * n_chunks;  // running tile number: stage = g % stages, parity = (g / stages) & 1
        bool saw_nonfinite = false;
        if (warp == kProducerWarp) {
            // =================== producer warp ===================
            long long iacc[3] = {0, 0, 0};  // diagnostics: cycles in flag wait + pad stores + fence, arrive, bulk issue
            auto issue_tile = [&](int c) {
                const long long i0 = MAS_TR(p) ? clock64() : 0;
                const uint32_t g = g0 + c;
                const uint32_t st = g % n_stages;
                const int row0 = c * R;
                const int rows = min(R, t_y - row0);
                if (p.flags && pass == 0 && (row0 & 127) == 0) {
                    // fused kernel: the cost tile holding these rows must have been published; consume the flag
                    uint32_t *f = p.flags + (size_t)b * p.flag_tiles + (row0 >> 7);
                    while (ld_acquire_gpu(f) < (uint32_t)p.flag_need) __nanosleep(64);
                    *f = 0u;
                    if (MAS_TR(p)) p.trace[12288 + (size_t)b * 32 + 2 + (row0 >> 7)] = globaltimer_ns();
                    fence_proxy_async_all();  // order the bulk (async-proxy) reads below after the acquire
                }
                const size_t start = (utt_elem0 + (size_t)row0 * ld) * 4;
                const uint32_t mis = (uint32_t)(start & 15);
                const size_t src0 = start - mis;
                const uint32_t want = mis + (uint32_t)rows * ld * 4;
                uint32_t bulk = (want + 15u) & ~15u;
                unsigned char *dst = smem + p.off_stage + (size_t)st * p.stage_bytes;
                const unsigned char *src = reinterpret_cast<const unsigned char *>(p.neg_cent) + src0;
                // finite pad behind the tile for the one thread whose columns straddle S
                *reinterpret_cast<uint4 *>(dst + ((want + 15u) & ~15u)) = make_uint4(0, 0, 0, 0);
                *reinterpret_cast<uint4 *>(dst + ((want + 15u) & ~15u) + 16) = make_uint4(0, 0, 0, 0);
                if (src0 + bulk > total_bytes) {
                    // the 16-byte round-up would run past the tensor: finish the tail by hand
                    bulk = want & ~15u;
                    for (uint32_t o = bulk; o < ((want + 15u) & ~15u); o += 4)
                        *reinterpret_cast<float *>(dst + o) =
                            (o < want) ? *reinterpret_cast<const float *>(src + o) : 0.0f;
                }
                if MAS_DBG(p, 4) bulk = 0;
                const long long i1 = MAS_TR(p) ? clock64() : 0;
                if (!kNoise && MAS_TR(p)) {
                    mbar_arrive_expect_tx(&full[st], bulk);
                    const long long i2 = clock64();
                    if (bulk) bulk_g2s(dst, src, bulk, &full[st]);
                    const long long i3 = clock64();
                    iacc[0] += i1 - i0, iacc[1] += i2 - i1, iacc[2] += i3 - i2;
                    return;
                }
                if (kNoise) {
                    // the noise tile of the same cells, nz_off behind the cost tile (same padding / tail rules)
                    unsigned char *ndst = dst + p.noise_off;
                    const unsigned char *nsrc = reinterpret_cast<const unsigned char *>(p.noise) + src0;
                    *reinterpret_cast<uint4 *>(ndst + ((want + 15u) & ~15u)) = make_uint4(0, 0, 0, 0);
                    *reinterpret_cast<uint4 *>(ndst + ((want + 15u) & ~15u) + 16) = make_uint4(0, 0, 0, 0);
                    for (uint32_t o = bulk; o < ((want + 15u) & ~15u); o += 4)
                        *reinterpret_cast<float *>(ndst + o) = (o < want) ? *reinterpret_cast<const float *>(nsrc + o) : 0.0f;
                    mbar_arrive_expect_tx(&full[st], 2 * bulk);
                    if (bulk) {
                        bulk_g2s(dst, src, bulk, &full[st]);
                        bulk_g2s(ndst, nsrc, bulk, &full[st]);
                    }
                } else {
                    mbar_arrive_expect_tx(&full[st], bulk);
                    if (bulk) bulk_g2s(dst, src, bulk, &full[st]);
                }
            };
            // kFed: the partner CTA writes the tiles; arm the stage's barrier with the bytes it is going to send
            auto arm_tile = [&](int c) {
                const uint32_t st_a = (g0 + (uint32_t)c) % n_stages;
                mbar_arrive_expect_tx(&full[st_a], (uint32_t)min(R, t_y - c * R) * ld * 4);
            };
            if (lane == 0) {
                const int pre = min((int)n_stages, n_chunks);
                for (int c = 0; c < pre; ++c) {
                    if (kFed)
                        arm_tile(c);
                    else
                        issue_tile(c);
                }
            }
            // zero-fill of the dense path, spread over the chunk steps: TMA bulk stores from a
            // zeroed shared buffer when the plane is 16-byte aligned, plain stores otherwise
            const size_t pbytes = (pass == 0 && !MAS_DBG(p, 1) && !p.zero_flags && path_b) ? plane * esize : 0;
            const bool bulk_ok = ((reinterpret_cast<uintptr_t>(path_b) | pbytes) & 15) == 0;
            const size_t quota = align_up((pbytes + n_steps - 1) / n_steps, 512);
            long long pacc[3] = {0, 0, 0};  // diagnostics: cycles in zero-fill issue, barrier, tile issue
            for (int step = 0; step < n_steps; ++step) {
                const long long q0 = MAS_TR(p) ? clock64() : 0;
                size_t lo = (size_t)step * quota, hi = lo + quota;
                if (lo > pbytes) lo = pbytes;
                if (hi > pbytes || step == n_steps - 1) hi = pbytes;
                if (hi > lo) {
                    if (bulk_ok) {
                        if (lane == 0) {
                            for (size_t o = lo; o < hi; o += kZeroBytes)
                                bulk_s2g(path_b + o, zero_s, (uint32_t)min((size_t)kZeroBytes, hi - o));
                            bulk_commit();
                        }
                    } else {
                        zero_bytes_warp(path_b + lo, hi - lo, lane);
                    }
                }
                const long long q1 = MAS_TR(p) ? clock64() : 0;
                bar_sync(bar, kThreads);
                const long long q2 = MAS_TR(p) ? clock64() : 0;
                const int freed = step - (kDpWarps - 1) - kHs;
                if (kFed) {
                    // the stage of chunk `freed` has been read by every value warp: hand it back to the feeder (one
                    // credit per chunk, so that its phase count stays in step with the running tile number)
                    if (lane == 0 && freed >= 0 && freed < n_chunks) {
                        const uint32_t st_f = (g0 + (uint32_t)freed) % n_stages;
                        if (freed + (int)n_stages < n_chunks) arm_tile(freed + (int)n_stages);   // before the credit
                        dsm_mbar_arrive_release(dsm_map(smem_u32(smem + p.fed_credit_off + st_f * 8), 1u));
                    }
                } else if (lane == 0 && freed >= 0 && freed + (int)n_stages < n_chunks) {
                    issue_tile(freed + (int)n_stages);
                }
                __syncwarp();
                if (MAS_TR(p)) {
                    const long long q3 = clock64();
                    pacc[0] += q1 - q0, pacc[1] += q2 - q1, pacc[2] += q3 - q2;
                }
            }
            if (MAS_TR(p) && lane == 0) {
                for (int j = 0; j < 3; ++j) p.trace[40960 + (size_t)b * 16 + 8 + j] = (unsigned long long)pacc[j];
                p.trace[40960 + (size_t)b * 16 + 11] = (unsigned long long)iacc[0];
                p.trace[40960 + (size_t)b * 16 + 14] = (unsigned long long)iacc[1];
                p.trace[40960 + (size_t)b * 16 + 15] = (unsigned long long)iacc[2];
            }
            if (pass == 0 && bulk_ok && lane == 0) bulk_wait_all();  // zeros land before the ones are scattered
        } else if (kHelp > 0 && warp > kProducerWarp) {
            // =================== noise helpers: chunk `step` while the value warps are on chunk step - 1 ===================
            dp_noise_helper<R, (kHelp > 0 ? kHelp : 2)>(p, smem, b, warp - kProducerWarp - 1, lane, g0, n_chunks, n_steps, t_y,
                                                      bar, kThreads);
        } else if (kVK && warp >= W) {
            // =================== origin warps (warp split), one step behind their value warp ===================
            // The value warps leave the decision word of every cell; these warps replay it into the
            // origin / hop / checkpoint bookkeeping, the other cross-lane chain of the forward pass.
            const int w = warp - W;
            const int x0 = (w * 32 + lane) * C;
            const bool lane0 = lane == 0, lane31 = lane == 31;
            int org[C];
#pragma unroll
            for (int k = 0; k < C; ++k) org[k] = x0 + k;
            int carry_o = 0;
            const int *bin_o = bnd_o + (size_t)w * ring;
            int *bout_o = bnd_o + (size_t)(w + 1) * ring;
            const int edge_rows = (w + 1) * 32 * C;
            long long kacc[2] = {0, 0};  // diagnostics: cycles in compute, barrier
            for (int step = 0; step < n_steps; ++step) {
                const int c = step - w - 1 - kHs;
                const long long f0 = MAS_TR(p) ? clock64() : 0;
                long long f1 = f0;
                if (c >= 0 && c < n_chunks && !MAS_DBG(p, 16)) {
                    const int row0 = c * R;
                    const int rows = min(R, t_y - row0);
                    const int slot0 = (c & 1) * R;
                    const uint32_t *wrow = bits + (size_t)(row0 >> 5) * S_bits + x0;
                    uint32_t wd[C];
#pragma unroll
                    for (int k = 0; k < C; ++k) wd[k] = wrow[k];
                    if (row0 < edge_rows)
                        org_chunk<C, true>(org, wd, rows, row0, carry_o, bin_o + slot0, bout_o + slot0, x0, lane0, lane31);
                    else
                        org_chunk<C, false>(org, wd, rows, row0, carry_o, bin_o + slot0, bout_o + slot0, x0, lane0, lane31);
                    const int end_row = row0 + rows;
                    if ((end_row & (kCheck - 1)) == 0) {
                        unsigned char *hrow = hop + (size_t)(end_row / kCheck) * S_pad;
#pragma unroll
                        for (int k = 0; k < C; ++k) {
                            hrow[x0 + k] = (unsigned char)(x0 + k - org[k]);
                            org[k] = x0 + k;
                        }
                        if (lane31) bout_o[slot0 + rows - 1] = x0 + C - 1;
                    }
                    if (MAS_TR(p)) f1 = clock64() + (long long)(org[0] & 0);
                }
                bar_sync(bar, kThreads);
                if (MAS_TR(p)) {
                    const long long f2 = clock64();
                    kacc[0] += f1 - f0, kacc[1] += f2 - f1;
                }
            }
            if (MAS_TR(p) && lane == 0 && w == 0)
                for (int j = 0; j < 2; ++j) p.trace[40960 + (size_t)b * 16 + 4 + j] = (unsigned long long)kacc[j];
#pragma unroll
            for (int k = 0; k < C; ++k) hop[x0 + k] = (unsigned char)(x0 + k - org[k]);
        } else {
            // =================== DP warps ===================
            const int w = warp;
            const int x0 = (w * 32 + lane) * C;
            const bool lane0 = lane == 0, lane31 = lane == 31;
            float v[C], fin = 0.0f;
            int org[C];
            uint32_t wl[C], wacc[C];
#pragma unroll
            for (int k = 0; k < C; ++k) {
                v[k] = kNeg;
                org[k] = x0 + k;
                wl[k] = 0u;
                wacc[k] = 0u;
            }
            // value left of column x0 for the first row: 0 for column 0 at y == 0 (core.pyx:22-23)
            float carry_v = (w == 0) ? 0.0f : kNeg;
            int carry_o = 0;
            const float *bin_v = bnd_v + (size_t)w * ring;
            const int *bin_o = bnd_o + (size_t)w * ring;
            float *bout_v = bnd_v + (size_t)(w + 1) * ring;
            int *bout_o = bnd_o + (size_t)(w + 1) * ring;
            const int edge_rows = (w + 1) * 32 * C;  // rows where a column of this warp is still above the diagonal
            long long dacc[4] = {0, 0, 0, 0};  // diagnostics: cycles in tile wait, compute, bits/hop, barrier
            uint32_t st = g0 % n_stages, st_par = (g0 / n_stages) & 1u;  // stage / mbarrier parity of chunk 0, then stepped
            for (int step = 0; step < n_steps; ++step) {
                const int c = step - w - kHs;
                long long d0 = MAS_TR(p) ? clock64() : 0, d1 = d0, d2 = d0, d3 = d0;
                if (c >= 0 && c < n_chunks) {
                    const int row0 = c * R;
                    const int rows = min(R, t_y - row0);
                    const int slot0 = (c & 1) * R;
                    const uint32_t mis = kVec ? 0u : (uint32_t)(((utt_elem0 + (size_t)row0 * ld) * 4) & 15);
                    const float *tile =
                        reinterpret_cast<const float *>(smem + p.off_stage + (size_t)st * p.stage_bytes + mis);
                    mbar_wait(&full[st], st_par);   // (kFed: the partner CTA's asynchronous stores complete the phase)
                    if (MAS_TR(p)) d1 = clock64();
                    const bool edge = row0 < edge_rows;
#define MAS_CHUNK(EDGE, EXACT)                                                                             \
    dp_chunk<C, R, EDGE, kVec, EXACT, kNoise, !kVK>(v, org, fin, wl, tile, S, ld, rows, row0, carry_v, carry_o,    \
                                                    bin_v + slot0, bin_o + slot0, bout_v + slot0, bout_o + slot0, \
                                                    x0, lane0, lane31, nzp)
                    if MAS_DBG(p, 2) {
                    } else if (pass == 0) {
                        if (edge)
                            MAS_CHUNK(true, false);
                        else
                            MAS_CHUNK(false, false);
                    } else {
                        if (edge)
                            MAS_CHUNK(true, true);
                        else
                            MAS_CHUNK(false, true);
                    }
#undef MAS_CHUNK
                    if (MAS_TR(p)) d2 = clock64() + (long long)(__float_as_int(v[0]) & 0);
                    // decision words: R < 32 accumulates 32 / R chunks per word
                    const int end_row = row0 + rows;
                    const bool word_done = ((end_row & (kCheck - 1)) == 0) || (c == n_chunks - 1);
#pragma unroll
                    for (int k = 0; k < C; ++k) {
                        if (R == kCheck)
                            wacc[k] = wl[k];
                        else
                            wacc[k] |= wl[k] << (row0 & (kCheck - 1));
                        wl[k] = 0u;
                    }
                    if (word_done) {
                        uint32_t *wrow = bits + (size_t)((end_row - 1) >> 5) * S_bits + x0;
                        if (C % 4 == 0) {
#pragma unroll
                            for (int k = 0; k < C; k += 4)
                                *reinterpret_cast<uint4 *>(wrow + k) =
                                    make_uint4(wacc[k], wacc[k + 1], wacc[k + 2], wacc[k + 3]);
                        } else if (C % 2 == 0) {
#pragma unroll
                            for (int k = 0; k < C; k += 2)
                                *reinterpret_cast<uint2 *>(wrow + k) = make_uint2(wacc[k], wacc[k + 1]);
                        } else {
#pragma unroll
                            for (int k = 0; k < C; ++k) wrow[k] = wacc[k];
                        }
#pragma unroll
                        for (int k = 0; k < C; ++k) wacc[k] = 0u;
                    }
                    // checkpoint after rows 31, 63, ...: remember where each column backtracks to, restart origins
                    // (the origin warps' job with the warp split)
                    if (!kVK && (end_row & (kCheck - 1)) == 0) {
                        unsigned char *hrow = hop + (size_t)(end_row / kCheck) * S_pad;
#pragma unroll
                        for (int k = 0; k < C; ++k) {
                            hrow[x0 + k] = (unsigned char)(x0 + k - org[k]);
                            org[k] = x0 + k;
                        }
                        // the right-hand warp must see the restarted origin of our last column
                        if (lane31) bout_o[slot0 + rows - 1] = x0 + C - 1;
                    }
                    if (MAS_TR(p)) d3 = clock64();
                    if (++st == n_stages) st = 0, st_par ^= 1u;
                }
                bar_sync(bar, kThreads);
                if (MAS_TR(p)) {
                    const long long d4 = clock64();
                    dacc[0] += d1 - d0, dacc[1] += d2 - d1, dacc[2] += d3 - d2, dacc[3] += d4 - d3;
                }
            }
            if (MAS_TR(p) && lane == 0 && (w == 0 || w == 3))
                for (int j = 0; j < 4; ++j) p.trace[40960 + (size_t)b * 16 + (w ? 4 : 0) + j] = (unsigned long long)dacc[j];
            // origin of the last row relative to its checkpoint -> hop row 0
            if (!kVK) {
#pragma unroll
                for (int k = 0; k < C; ++k) hop[x0 + k] = (unsigned char)(x0 + k - org[k]);
            }
            saw_nonfinite = !(fabsf(fin) <= 3.0e38f);  // NaN or Inf
        }
        if (saw_nonfinite) *nonfinite_s = 1;
        bar_sync(bar, kThreads);
        passes = pass + 1;
        const bool again = (pass == 0 && *nonfinite_s != 0);
        if (kFed && tid == 0) {
            // tell the feeder whether the tiles have to be streamed once more (exact pass)
            dsm_st_u32(dsm_map(smem_u32(smem + p.fed_verdict_off + 8), 1u), again ? 1u : 0u);
            dsm_mbar_arrive_release(dsm_map(smem_u32(smem + p.fed_verdict_off), 1u));
        }
        if (!again) break;
    }
    g_base += (uint32_t)passes * n_chunks;
    if (!p.bits_in_smem || !p.hop_in_smem) __threadfence_block();
    bar_sync(bar, kThreads);

    if (MAS_TR(p) && tid == 0) p.trace[12288 + (size_t)b * 32 + 1] = globaltimer_ns();
    // =================== backtrack ===================
    const int y_last = t_y - 1;
    const int J = t_y / kCheck;  // checkpoints stored: after rows 31, 63, ..., 32 J - 1
    if (!kFed && !p.hop_in_smem) {
        // Spilled hop bytes: level 1 below is J dependent reads, ~0.6 us each from L2 (75 us at T = 4000).  The tile
        // ring is idle now -- every tile this utterance needed has been consumed -- so bring the J + 1 rows back in
        // with coalesced loads first, when they fit.
        const size_t need = (size_t)(J + 1) * S_pad;
        if (need <= (size_t)p.stages * p.stage_bytes) {
            unsigned char *hop_s = smem + p.off_stage;
            const uint4 *src = reinterpret_cast<const uint4 *>(hop);   // region and row pitch are 16-byte multiples
            for (size_t i = tid; i < (need + 15) / 16; i += kThreads) reinterpret_cast<uint4 *>(hop_s)[i] = src[i];
            hop = hop_s;
            bar_sync(bar, kThreads);
        }
    }
    if (tid == 0) {
        // level 1: one dependent load per 32 rows
        int c = t_x - 1;
        c -= hop[c];  // column at row c_J
        entry_s[J] = (uint16_t)c;
        for (int j = J; j >= 1; --j) {
            c -= hop[(size_t)j * S_pad + c];  // column at row c_{j-1}
            entry_s[j - 1] = (uint16_t)c;
        }
        idx_s[ix(0)] = 0;
        if (MAS_TR(p)) p.trace[12288 + (size_t)b * 32 + 26] = globaltimer_ns();
    }
    bar_sync(bar, kThreads);
    // level 2: independent walks of <= 32 rows, one per thread
    for (int j = tid; j <= J; j += kThreads) {
        const int y_lo = check_row(j) + 1;
        int y_top, cur;
        if (j < J) {
            y_top = check_row(j + 1);
            cur = entry_s[j + 1];
        } else {
            y_top = y_last;
            cur = t_x - 1;
        }
        // (all rows of a walk lie in one 32-row block: one decision-word row, and the bit index is y - 32 j)
        const uint32_t *wrow = bits + (size_t)(y_top >> 5) * S_bits;
        for (int y = y_top; y >= y_lo; --y) {
            idx_s[ix(y)] = (uint16_t)cur;
            const uint32_t wd = wrow[cur];
            cur -= (cur != 0) ? (int)((wd >> (y & 31)) & 1u) : 0;  // core.pyx:32 "index != 0 and ..."
        }
    }
    for (int x = tid; x < S_pad; x += kThreads) end_s[x] = -1;
    if (MAS_TR(p) && tid == 0) p.trace[12288 + (size_t)b * 32 + 27] = globaltimer_ns();
    if (p.zero_flags && tid == 0) {
        // somebody else zero-fills this utterance's path plane: the ones may only follow the zeros
        uint32_t *f = p.zero_flags + b;
        while (ld_acquire_gpu(f) < (uint32_t)kZeroParts) __nanosleep(64);
        *f = 0u;
    }
    if (MAS_TR(p) && tid == 0) p.trace[12288 + (size_t)b * 32 + 28] = globaltimer_ns();
    bar_sync(bar, kThreads);

    // =================== outputs ===================
    for (int y = tid; y < t_y; y += kThreads) {
        const int c = idx_s[ix(y)];
        if (path_b) store_one(path_b, (size_t)y * S + c, p.path_dtype);
        if (y == y_last || idx_s[ix(y + 1)] != c) end_s[c] = y;
    }
    if (p.idx) {
        for (int y = tid; y < T; y += kThreads) p.idx[(size_t)b * T + y] = (y < t_y) ? (int)idx_s[ix(y)] : -1;
    }
    if (p.status && tid == 0) p.status[b] = MAS_UTT_OK;
    if (MAS_TR(p) && tid == 0) p.trace[12288 + (size_t)b * 32 + 30] = globaltimer_ns();
    if (p.dur) {
        bar_sync(bar, kThreads);
        for (int x = tid; x < S; x += kThreads) {
            int d = 0;
            if (x < t_x) d = end_s[x] - (x > 0 ? end_s[x - 1] : -1);
            p.dur[(size_t)b * S + x] = d;
        }
    }
}

// Zero-fills the dense path planes and raises their zero flags.  Runs on a whole CTA that has nothing else
// left to do (fused kernel); `zbuf` is kZeroFillBuf bytes of shared memory.  The DP role of the utterance's
// owner waits for the flag before it scatters the ones.
// Work item = one of kZeroParts slices of one plane, handed out through the queue counter `p.zero_queue` in
// plane order, so that a CTA that leaves the contraction early takes more of them than one that leaves late
// (a static split made every plane wait for the slowest CTA).  Two slices are in flight per CTA: the flag of
// a slice goes out once the stores of the next one have been issued.
__device__ __forceinline__ void zero_fill_role(const DpParams &p, unsigned char *zbuf)
{
    const int tid = threadIdx.x, nthr = blockDim.x;
    for (int i = tid; i < (int)(kZeroFillBuf / 16); i += nthr) reinterpret_cast<uint4 *>(zbuf)[i] = make_uint4(0, 0, 0, 0);
    fence_proxy_async();
    __syncthreads();
    const size_t pbytes = (size_t)p.T * p.S * path_elem_size(p.path_dtype);
    const size_t part_bytes = align_up((pbytes + kZeroParts - 1) / kZeroParts, 512);
    const int total = p.B * kZeroParts;
    const bool bulk_ok = ((reinterpret_cast<uintptr_t>(p.path) | pbytes) & 15) == 0;
    if (bulk_ok) {
        if (tid != 0) return;
        int prev = -1;
        for (;;) {
            const int wi = (int)atomicAdd(p.zero_queue, 1u);
            if (wi >= total) break;
            const int b = wi / kZeroParts, part = wi - b * kZeroParts;
            unsigned char *path_b = p.path + (size_t)b * pbytes;
            size_t lo = (size_t)part * part_bytes, hi = lo + part_bytes;
            if (lo > pbytes) lo = pbytes;
            if (hi > pbytes) hi = pbytes;
            for (size_t o = lo; o < hi; o += kZeroFillBuf)
                bulk_s2g(path_b + o, zbuf, (uint32_t)min((size_t)kZeroFillBuf, hi - o));
            bulk_commit();
            if (prev >= 0) {
                bulk_wait_group<1>();  // everything but the slice just issued has landed
                fence_proxy_async_all();
                __threadfence();
                atomicAdd(p.zero_flags + prev / kZeroParts, 1u);  // the DP waits for all kZeroParts slices
            }
            prev = wi;
        }
        if (prev >= 0) {
            bulk_wait_all();
            fence_proxy_async_all();
            __threadfence();
            atomicAdd(p.zero_flags + prev / kZeroParts, 1u);
        }
    } else {
        // planes that are not 16-byte aligned: plain stores by all warps, one slice at a time
        int *wi_s = reinterpret_cast<int *>(zbuf + kZeroFillBuf - 16);  // (the last bytes of zbuf, zero again afterwards)
        const int warp = tid >> 5, n_warps = nthr >> 5;
        for (;;) {
            __syncthreads();
            if (tid == 0) *wi_s = (int)atomicAdd(p.zero_queue, 1u);
            __syncthreads();
            const int wi = *wi_s;
            if (wi >= total) break;
            const int b = wi / kZeroParts, part = wi - b * kZeroParts;
            unsigned char *path_b = p.path + (size_t)b * pbytes;
            size_t lo = (size_t)part * part_bytes, hi = lo + part_bytes;
            if (lo > pbytes) lo = pbytes;
            if (hi > pbytes) hi = pbytes;
            const size_t seg = align_up((hi - lo + n_warps - 1) / n_warps, 512);
            size_t a = lo + (size_t)warp * seg, e = a + seg;
            if (a > hi) a = hi;
            if (e > hi) e = hi;
            if (e > a) zero_bytes_warp(path_b + a, e - a, tid & 31);
            __threadfence();
            __syncthreads();
            if (tid == 0) atomicAdd(p.zero_flags + b, 1u);
        }
    }
}

// host side (mas_dp.cu)
int dp_team_warps(int S);
size_t dp_workspace_bytes(int B, int T, int S);
int dp_prepare(DpPlan &pl, const float *neg_cent, const int32_t *t_ys, const int32_t *t_xs, void *path_out,
               int path_dtype, int32_t *dur_out, int32_t *idx_out, int32_t *status_out, void *workspace,
               size_t workspace_bytes, int B, int T, int S, int32_t **order_out, int R, size_t smem_budget,
               bool with_noise = false, int ld = 0, int help = 0);

}  // namespace mas
