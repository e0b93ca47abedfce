// mas_segsum.cu -- backward of the prior expansion (vits2/models.py:1270-1271, the transposed one-hot matmul):
//   g_in[b, d, s] = sum over the frames t aligned to text column s of g_out[b, d, t]
// The path is monotonic, so column s owns the contiguous frames [start_s, start_s + dur_s); which frame ends a
// column is the same for every channel of the utterance.  This kernel puts the CHANNELS on the lanes: a warp owns
// 32 channels and walks the frames in order with one running sum per lane, so "does a column end here" is a
// warp-uniform test, there are no shuffles, no atomics, and a long segment costs what a short one does (the
// column-per-thread kernel of mas_expand.cu waits for the thread that owns the longest segment).  The gradient rows
// arrive as [channels][32 frames] tiles through the tensor-map engine (128-byte swizzle: a lane reads its row 16
// bytes at a time without bank conflicts), a ring of kSegStages tiles per CTA and three CTAs per SM keep ~140 KB per SM in flight; the
// finished sums leave through a per-warp [columns][32 channels] transpose so the stores are 128-byte rows of
// g_in.  Frames past t_y are never read; empty columns (duration 0) are compacted away before the walk and zeroed
// after it.
#include <cuda.h>
#include <limits.h>

#include "mas_common.cuh"

namespace mas {

bool make_tmap_f32_3d(CUtensorMap *out, const void *base, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t stride1_bytes,
                      uint64_t stride2_bytes, uint32_t b0, uint32_t b1, uint32_t b2, bool swizzle128);

constexpr int kSegFrames = 32;     // frames per tile row: 128 bytes, one swizzle atom
constexpr int kSegStages = 8;     // most tiles in the ring (the launch picks how many are used)
constexpr int kSegRing = 32;      // parked column sums per warp (at most 4 close per step of 4 frames, 16 leave at a time)
constexpr int kSegOut = 16;       // columns per write-out: half a warp per column block, the halves take 16 channels each
constexpr int kSegMaxWarps = 8;    // consumer warps per CTA (32 channels each); one more warp issues the loads
constexpr int kSegMaxParts = 4;    // runs of frames per CTA, each with its own ring and consumer warps

struct SegParams {
    const int32_t *dur;
    float *g_m_p, *g_logs_p;
    int D, T, S, nw, np, stages;
};

// 16 parked sums x up to 32 channels of one warp -> g_in (a lane = one non-empty column and one half of the channels;
// neighbouring lanes are neighbouring columns unless an empty one lies between).  Out of line: it runs once per 16
// columns.
__device__ __noinline__ void seg_write_block(const float *tr_row, float *dst, int nch, int S, bool on)
{
    __syncwarp();
    if (on) {
#pragma unroll 4
        for (int c = 0; c < nch; ++c) dst[(size_t)c * S] = tr_row[c];
    }
    __syncwarp();
}

__global__ void __launch_bounds__(32 * (kSegMaxWarps + kSegMaxParts))
    mas_segsum_kernel(const __grid_constant__ CUtensorMap tm_m, const __grid_constant__ CUtensorMap tm_l, SegParams p)
{
    extern __shared__ __align__(1024) unsigned char seg_smem_raw[];
    const uint32_t raw = smem_u32(seg_smem_raw);
    unsigned char *smem = seg_smem_raw + (((raw + 1023u) & ~1023u) - raw);   // swizzled tiles need 1024-byte alignment
    const int nw = p.nw, np = p.np, n_st = p.stages, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n_cons = nw * np;                                // consumer warps: warp = part * nw + channel group
    const int b = blockIdx.y, d0 = blockIdx.x * 32 * nw;
    const bool logs = blockIdx.z != 0;
    const CUtensorMap *tm = logs ? &tm_l : &tm_m;
    float *out = logs ? p.g_logs_p : p.g_m_p;
    const uint32_t stage_bytes = (uint32_t)nw * 32u * 128u;
    unsigned char *rings = smem;                                                          // [np][n_st][stage]
    float *tr_all = reinterpret_cast<float *>(smem + (size_t)np * n_st * stage_bytes);    // [n_cons][64][33]
    float *fp_all = tr_all + n_cons * kSegRing * 33;                                      // [n_cons][32] first sums
    float *tail_all = fp_all + n_cons * 32;                                               // [n_cons][32] last sums
    int *closed_all = reinterpret_cast<int *>(tail_all + n_cons * 32);                    // [n_cons (+ pad)]
    int *nz_col = closed_all + 8;                                                         // [S]
    uint32_t *heads = reinterpret_cast<uint32_t *>(nz_col + ((p.S + 3) & ~3));            // [T / 32 + 2] bit masks
    uint32_t *emask = heads + ((p.T / 32 + 2 + 3) & ~3);                                  // [32]: empty columns
    uint64_t *full = reinterpret_cast<uint64_t *>(emask + 32);                            // [np][kSegStages]
    uint64_t *empty = full + kSegMaxParts * kSegStages;
    int *info = reinterpret_cast<int *>(empty + kSegMaxParts * kSegStages);               // {frames, non-empty columns}

    if (tid == 0) {
        for (int i = 0; i < kSegMaxParts * kSegStages; ++i) {
            mbar_init(&full[i], 1);
            mbar_init(&empty[i], (uint32_t)nw);
        }
        fence_mbar_init();
        tma_prefetch_desc(tm);
    }
    for (int i = tid; i < p.T / 32 + 2; i += blockDim.x) heads[i] = 0u;
    for (int s = tid; s < p.S; s += blockDim.x) nz_col[s] = max(p.dur[(size_t)b * p.S + s], 0);   // one round trip
    __syncthreads();
    if (warp == 0) {
        // The non-empty columns in order (nz_col[i]), and one bit per frame: "a column ends right before this frame"
        // (inclusive prefix sum of the durations).  Empty columns are skipped here and zeroed at the end, so a frame
        // closes at most one column, and whether it does is a bit test on a register -- nothing on the critical path
        // of the walk below depends on shared memory.
        int run = 0, n_nz = 0;
        for (int s0 = 0; s0 < p.S; s0 += 32) {
            const int s = s0 + lane;
            const int d = s < p.S ? nz_col[s] : 0;   // compacted in place below: index i <= s, this chunk is read first
            int incl = d;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int n = __shfl_up_sync(kFullMask, incl, o);
                if (lane >= o) incl += n;
            }
            const uint32_t m = __ballot_sync(kFullMask, d > 0);
            if (lane == 0) emask[s0 >> 5] = ~m;
            if (d > 0) {
                nz_col[n_nz + __popc(m & ((1u << lane) - 1u))] = s;
                const int e = run + incl;
                if (e <= p.T) atomicOr(&heads[e >> 5], 1u << (e & 31));
            }
            n_nz += __popc(m);
            run += __shfl_sync(kFullMask, incl, 31);
        }
        if (lane == 0) info[0] = run, info[1] = n_nz;
    }
    __syncthreads();
    const int t_end = min(info[0], p.T);                      // frames [0, t_end) belong to some column
    const int n_nz = info[1];
    const int n_tiles = (t_end + kSegFrames - 1) / kSegFrames;
    // the frames are cut into np runs of whole tiles, one per part; a column that straddles a cut is put together
    // from the parts' first / last sums once every part is through (below)
    const int tpp = (n_tiles + np - 1) / np;
    const int p_last = n_tiles > 0 ? (n_tiles - 1) / tpp : 0;

    if (warp >= n_cons) {
        // ---- producers: warp n_cons + r feeds the ring of part r.  One tile = [32 nw channels][32 frames]; rows past D
        // and frames past T arrive as zeros ----
        const int r = warp - n_cons;
        if (lane == 0) {
            const int k0 = r * tpp, n = max(0, min(tpp, n_tiles - k0));
            uint64_t *fl = full + r * kSegStages, *em = empty + r * kSegStages;
            unsigned char *ring = rings + (size_t)r * n_st * stage_bytes;
            for (int j = 0; j < n; ++j) {
                const int st = j % n_st;
                if (j >= n_st) mbar_wait(&em[st], (uint32_t)((j / n_st - 1) & 1));
                mbar_arrive_expect_tx(&fl[st], stage_bytes);
                tma_load_3d(ring + (size_t)st * stage_bytes, tm, (k0 + j) * kSegFrames, d0, b, &fl[st]);
            }
        }
        return;
    }

    // ---- consumers ----
    const int part = warp / nw, cg = warp - part * nw;
    const int k0 = part * tpp, n_mine = max(0, min(tpp, n_tiles - k0));
    uint64_t *fl = full + part * kSegStages, *em = empty + part * kSegStages;
    float *tr = tr_all + warp * kSegRing * 33;   // ring of parked sums: [non-empty column index & 31][channel]
    const int ch0 = d0 + cg * 32;
    const bool have_ch = ch0 < p.D;              // D not a multiple of 32 nw: a warp without channels only frees the stages
    const int nch = min(32, p.D - ch0);
    float *out_w = out + ((size_t)b * p.D + ch0) * p.S;
    // i_cur = index of the open column among the non-empty ones: the head bits below this part's first frame
    int i_open = 0;
    for (int w = lane; w < min(k0, n_tiles); w += 32) i_open += __popc(heads[w]);
    for (int o = 16; o > 0; o >>= 1) i_open += __shfl_xor_sync(kFullMask, i_open, o);
    int i_cur = i_open, i_out = part == 0 ? 0 : i_open + 1;
    // where the next closed column parks its sum: the ring -- except the first one of a later part, which began in
    // an earlier part and is finished there
    float *slot = part == 0 ? tr + (i_cur & (kSegRing - 1)) * 33 + lane : fp_all + warp * 32 + lane;
    float acc = 0.0f;
    auto write_blocks = [&](int upto) {   // 16 parked sums at a time; the last call takes what is left
        const int col = lane & (kSegOut - 1), c0 = (lane >> 4) * 16;
        while (i_out < upto) {
            const int i = i_out + col;
            const bool on = i < upto && c0 < nch;
            seg_write_block(tr + (i & (kSegRing - 1)) * 33 + c0, out_w + (size_t)c0 * p.S + (on ? nz_col[i] : 0),
                            min(16, nch - c0), p.S, on);
            i_out += kSegOut;
        }
    };
#define MAS_SEG_CLOSE()                                       \
    {                                                         \
        *slot = acc;                                          \
        acc = 0.0f;                                           \
        ++i_cur;                                              \
        slot = tr + (i_cur & (kSegRing - 1)) * 33 + lane;     \
    }
#define MAS_SEG_STEP(x, bit)       \
    if (g & (bit)) MAS_SEG_CLOSE() \
    acc += (x);
    if (have_ch) {
        const uint32_t row = smem_u32(rings) + (uint32_t)part * n_st * stage_bytes + (uint32_t)(cg * 32 + lane) * 128u;
        const uint32_t sw = (uint32_t)(lane & 7);
        uint32_t m = n_mine > 0 ? heads[k0] : 0u;
        for (int j = 0; j < n_mine; ++j) {
            const int st = j % n_st;
            mbar_wait(&fl[st], (uint32_t)((j / n_st) & 1));
            const uint32_t base = row + (uint32_t)st * stage_bytes;
            auto lds4 = [&](int c) {
                float4 r;
                asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                             : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                             : "r"(base + ((((uint32_t)c) ^ sw) << 4)));
                return r;
            };
            // four frames at a time, the next 16 bytes on their way while these are added.  (Kept rolled: with all 32
            // frames unrolled and the column-end handling inlined 32 times the loop was 5800 instructions, more than
            // the instruction cache holds -- 139 cycles per frame.)
            float4 cur = lds4(0);
            const uint32_t m_next = heads[k0 + j + 1];
#pragma unroll 1
            for (int c = 0; c < 8; ++c) {
                const float4 nxt = lds4((c + 1) & 7);
                const uint32_t g = (m >> (4 * c)) & 15u;
                if (g == 0u) {
                    acc = (((acc + cur.x) + cur.y) + cur.z) + cur.w;   // all four inside the open column
                } else {
                    MAS_SEG_STEP(cur.x, 1u)
                    MAS_SEG_STEP(cur.y, 2u)
                    MAS_SEG_STEP(cur.z, 4u)
                    MAS_SEG_STEP(cur.w, 8u)
                    if (i_cur - i_out >= kSegOut) write_blocks(i_out + ((i_cur - i_out) & ~(kSegOut - 1)));
                }
                cur = nxt;
            }
            m = m_next;
            // The tile goes back once its values have been consumed.  The store makes that explicit for the hardware
            // AND the compiler: it takes the running sum (hence every load of this tile) as its operand, and a memory
            // operation cannot move below the release-arrive, whereas the additions alone could be scheduled after it
            // (an arrive that overtakes pending shared-memory loads lets the refill land first: cost_tc_role).
            tail_all[warp * 32 + lane] = acc;
            __syncwarp();
            if (lane == 0) mbar_arrive(&em[st]);
        }
        if (part == p_last) {
            // the column that ends at t_end, and (durations adding up to more than T) the ones that never started
            while (i_cur < n_nz) {
                MAS_SEG_CLOSE()
                if (i_cur - i_out >= kSegOut) write_blocks(i_out + ((i_cur - i_out) & ~(kSegOut - 1)));
            }
        }
        tail_all[warp * 32 + lane] = acc;
        if (lane == 0) closed_all[warp] = i_cur != i_open;
    } else {
        for (int j = 0; j < n_mine; ++j) {
            const int st = j % n_st;
            mbar_wait(&fl[st], (uint32_t)((j / n_st) & 1));
            if (lane == 0) mbar_arrive(&em[st]);
        }
    }
    if (np > 1) asm volatile("bar.sync 1, %0;" ::"r"(n_cons * 32) : "memory");   // the consumer warps only
    if (!have_ch) return;
    if (part < p_last && (part == 0 || i_cur != i_open)) {
        // the column open at this part's last frame began here: its sum = this part's tail + every later part that
        // lies wholly inside it + the first sum of the part in which it ends -- a fixed order, whatever the timing
        float total = acc;
        int q = part + 1;
        while (q < p_last && !closed_all[q * nw + cg]) {
            total += tail_all[(q * nw + cg) * 32 + lane];
            ++q;
        }
        total += fp_all[(q * nw + cg) * 32 + lane];
        acc = total;
        MAS_SEG_CLOSE()
    }
#undef MAS_SEG_STEP
#undef MAS_SEG_CLOSE
    write_blocks(i_cur);
    if (part != 0) return;
    // empty columns receive nothing
    for (int s0 = 0; s0 < p.S; s0 += 32) {
        const int s = s0 + lane;
        if (s < p.S && ((emask[s0 >> 5] >> lane) & 1u))
            for (int c = 0; c < nch; ++c) out_w[(size_t)c * p.S + s] = 0.0f;
    }
}

static size_t segsum_smem(int nw, int np, int stages, int S, int T)
{
    const size_t n_cons = (size_t)nw * np;
    return 1024 + (size_t)np * stages * nw * 32 * 128 + n_cons * kSegRing * 33 * 4 + 2 * n_cons * 32 * 4 + 8 * 4 +
           (size_t)((S + 3) & ~3) * 4 + (size_t)((T / 32 + 2 + 3) & ~3) * 4 + 32 * 4 +
           2 * kSegMaxParts * kSegStages * 8 + 16;
}

// true when the launch was made; false = this shape / alignment needs the column-per-thread kernel
bool segsum_try_launch(const float *g_m, const float *g_logs, const int32_t *dur, float *g_m_p, float *g_logs_p, int B,
                       int D, int T, int S, cudaStream_t stream, int *rc)
{
    *rc = MAS_OK;
    if (T % 4 != 0 || B > 65535) return false;
    CUtensorMap tm_m, tm_l;
    const int n32 = (D + 31) / 32, ntens = g_logs ? 2 : 1;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    // The walk is a dependent chain per warp (~45 cycles per frame), so what counts is how many consumer warps an
    // SM holds (~9 fit, at ~20 KB of shared memory each) and that every SM has them:
    //   nw = channel groups per CTA: the most that divides the groups while the grid still covers the SMs;
    //   np = runs of frames per CTA (small batches): until there are ~8 consumer warps per SM;
    //   ring depth 3, or (np > 1) 2 when that lets the whole grid be resident at once.
    // tools/gpu_e2.sh sweeps the three (profiles/r2_segsum_sweep.txt).
    int nw = 1;
    for (int w = 4; w >= 2; --w)
        if (n32 % w == 0 && (n32 / w) * B * ntens >= sms) {
            nw = w;
            break;
        }
    int np = 1;
    const long warps1 = (long)n32 * B * ntens;
    while (np < kSegMaxParts && warps1 * np < 8L * sms && T / (np * 2) >= 128) np *= 2;
    if (config().seg_parts > 0) np = config().seg_parts > kSegMaxParts ? kSegMaxParts : config().seg_parts;
    while (nw > 1 && nw * np > 4) {   // CTAs of at most 4 consumer warps pack the SMs best
        int w = nw - 1;
        while (w > 1 && n32 % w != 0) --w;
        nw = w;
    }
    if (config().seg_nw > 0 && n32 % config().seg_nw == 0) nw = config().seg_nw;
    if (nw * np > kSegMaxWarps) return false;
    const long ctas = (long)((n32 + nw - 1) / nw) * B * ntens;
    int stages = config().seg_stages;
    if (stages < 2 || stages > kSegStages) {
        stages = 3;   // more CTAs per SM beat deeper rings (measured 3 / 4 / 6 / 8)
        const long slots3 = (long)sms * (long)((227 * 1024) / (segsum_smem(nw, np, 3, S, T) + 1024));
        const long slots2 = (long)sms * (long)((227 * 1024) / (segsum_smem(nw, np, 2, S, T) + 1024));
        if (np > 1 && ctas > slots3 && ctas <= slots2) stages = 2;   // (one run per CTA: 2 stages cost 10 %)
    }
    if (!make_tmap_f32_3d(&tm_m, g_m, (uint64_t)T, (uint64_t)D, (uint64_t)B, (uint64_t)T * 4, (uint64_t)D * T * 4,
                          kSegFrames, 32u * nw, 1, true))
        return false;
    if (g_logs) {
        if (!make_tmap_f32_3d(&tm_l, g_logs, (uint64_t)T, (uint64_t)D, (uint64_t)B, (uint64_t)T * 4, (uint64_t)D * T * 4,
                              kSegFrames, 32u * nw, 1, true))
            return false;
    } else {
        tm_l = tm_m;
    }
    const size_t smem = segsum_smem(nw, np, stages, S, T);
    static thread_local int configured_dev = -1;
    if (dev != configured_dev) {
        if (cudaFuncSetAttribute(mas_segsum_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)segsum_smem(4, 2, kSegStages / 2, MAS_MAX_TEXT, MAS_MAX_MEL)) != cudaSuccess) {
            cudaGetLastError();
            return false;
        }
        configured_dev = dev;
    }
    if (smem > segsum_smem(4, 2, kSegStages / 2, MAS_MAX_TEXT, MAS_MAX_MEL)) return false;
    SegParams p{dur, g_m_p, g_logs_p, D, T, S, nw, np, stages};
    const dim3 grid((unsigned)((n32 + nw - 1) / nw), (unsigned)B, (unsigned)ntens);
    mas_segsum_kernel<<<grid, 32 * (nw * np + np), smem, stream>>>(tm_m, tm_l, p);
    note_launch();
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) *rc = note_cuda_error(e, "mas_segsum_kernel");
    return true;
}

}  // namespace mas
