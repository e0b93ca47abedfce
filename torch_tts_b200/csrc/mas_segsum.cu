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
constexpr int kSegRing = 64;      // parked column sums per warp (at most 4 close per step of 4 frames, 32 leave at a time)
constexpr int kSegMaxWarps = 4;    // consumer warps per CTA (32 channels each); one more warp issues the loads

struct SegParams {
    const int32_t *dur;
    float *g_m_p, *g_logs_p;
    int D, T, S, nw, stages;
};

// 32 parked sums x up to 32 channels of one warp -> g_in (lane = one non-empty column, neighbouring lanes are
// neighbouring columns unless an empty one lies between).  Out of line: it runs once per 32 columns.
__device__ __noinline__ void seg_write_block(const float *tr_row, float *dst, int nch, int S, bool on)
{
    __syncwarp();
    if (on) {
#pragma unroll 4
        for (int c = 0; c < nch; ++c) dst[(size_t)c * S] = tr_row[c];
    }
    __syncwarp();
}

__global__ void __launch_bounds__(32 * (kSegMaxWarps + 1))
    mas_segsum_kernel(const __grid_constant__ CUtensorMap tm_m, const __grid_constant__ CUtensorMap tm_l, SegParams p)
{
    extern __shared__ __align__(1024) unsigned char seg_smem_raw[];
    const uint32_t raw = smem_u32(seg_smem_raw);
    unsigned char *smem = seg_smem_raw + (((raw + 1023u) & ~1023u) - raw);   // swizzled tiles need 1024-byte alignment
    const int nw = p.nw, n_st = p.stages, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = blockIdx.y, d0 = blockIdx.x * 32 * nw;
    const bool logs = blockIdx.z != 0;
    const CUtensorMap *tm = logs ? &tm_l : &tm_m;
    float *out = logs ? p.g_logs_p : p.g_m_p;
    const uint32_t stage_bytes = (uint32_t)nw * 32u * 128u;
    unsigned char *stages = smem;
    float *tr_all = reinterpret_cast<float *>(smem + n_st * stage_bytes);          // [nw][64][33]
    int *nz_col = reinterpret_cast<int *>(tr_all + nw * kSegRing * 33);                   // [S]
    uint32_t *heads = reinterpret_cast<uint32_t *>(nz_col + ((p.S + 3) & ~3));            // [T / 32 + 2] bit masks
    uint32_t *emask = heads + ((p.T / 32 + 2 + 3) & ~3);                                  // [32]: empty columns
    uint64_t *full = reinterpret_cast<uint64_t *>(emask + 32);
    uint64_t *empty = full + kSegStages;
    int *info = reinterpret_cast<int *>(empty + kSegStages);                              // {frames, non-empty columns}

    if (tid == 0) {
        for (int i = 0; i < kSegStages; ++i) {
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

    if (warp == nw) {
        // ---- producer: one tile = [32 nw channels][32 frames], rows past D / frames past T arrive as zeros ----
        if (lane == 0) {
            for (int k = 0; k < n_tiles; ++k) {
                const int st = k % n_st;
                if (k >= n_st) mbar_wait(&empty[st], (uint32_t)((k / n_st - 1) & 1));
                mbar_arrive_expect_tx(&full[st], stage_bytes);
                tma_load_3d(stages + (size_t)st * stage_bytes, tm, k * kSegFrames, d0, b, &full[st]);
            }
        }
        return;
    }

    // ---- consumers ----
    float *tr = tr_all + warp * kSegRing * 33;   // ring of parked sums: [non-empty column index & 63][channel]
    const int ch0 = d0 + warp * 32;
    if (ch0 >= p.D) {
        // a warp without channels (D not a multiple of 32 nw) still frees the stages
        for (int k = 0; k < n_tiles; ++k) {
            const int st = k % n_st;
            mbar_wait(&full[st], (uint32_t)((k / n_st) & 1));
            if (lane == 0) mbar_arrive(&empty[st]);
        }
        return;
    }
    const int nch = min(32, p.D - ch0);
    float *out_w = out + ((size_t)b * p.D + ch0) * p.S;
    // i_cur = index of the open column among the non-empty ones.  Sums park in the ring and leave 32 at a time.
    int i_cur = 0, i_out = 0;
    float acc = 0.0f;
    auto write_blocks = [&](int upto) {   // every complete block of 32 below `upto`, and the partial one if last
        while (i_out < upto) {
            const int i = i_out + lane;
            const bool on = i < upto;
            seg_write_block(tr + (i & (kSegRing - 1)) * 33, out_w + (on ? nz_col[i] : 0), nch, p.S, on);
            i_out += 32;
        }
    };
#define MAS_SEG_STEP(x, bit)                                  \
    if (g & (bit)) {                                          \
        tr[(i_cur & (kSegRing - 1)) * 33 + lane] = acc;       \
        acc = 0.0f;                                           \
        ++i_cur;                                              \
    }                                                         \
    acc += (x);
    const uint32_t row = smem_u32(stages) + (uint32_t)(warp * 32 + lane) * 128u;
    const uint32_t sw = (uint32_t)(lane & 7);
    uint32_t m = heads[0];
    for (int k = 0; k < n_tiles; ++k) {
        const int st = k % n_st;
        mbar_wait(&full[st], (uint32_t)((k / n_st) & 1));
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
        const uint32_t m_next = heads[k + 1];
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
                if (i_cur - i_out >= 32) write_blocks(i_cur & ~31);
            }
            cur = nxt;
        }
        m = m_next;
        __syncwarp();   // every lane has used its copy of the tile
        if (lane == 0) mbar_arrive(&empty[st]);
    }
#undef MAS_SEG_STEP
    // the column that ends at t_end, and (durations adding up to more than T) the ones that never started
    while (i_cur < n_nz) {
        tr[(i_cur & (kSegRing - 1)) * 33 + lane] = acc;
        acc = 0.0f;
        ++i_cur;
        if (i_cur - i_out >= 32) write_blocks(i_cur & ~31);
    }
    write_blocks(n_nz);
    // empty columns receive nothing
    for (int s0 = 0; s0 < p.S; s0 += 32) {
        const int s = s0 + lane;
        if (s < p.S && ((emask[s0 >> 5] >> lane) & 1u))
            for (int c = 0; c < nch; ++c) out_w[(size_t)c * p.S + s] = 0.0f;
    }
}

static size_t segsum_smem(int nw, int stages, int S, int T)
{
    return 1024 + (size_t)stages * nw * 32 * 128 + (size_t)nw * kSegRing * 33 * 4 + (size_t)((S + 3) & ~3) * 4 + (size_t)((T / 32 + 2 + 3) & ~3) * 4 + 32 * 4 +
           2 * kSegStages * 8 + 16;
}

// true when the launch was made; false = this shape / alignment needs the column-per-thread kernel
bool segsum_try_launch(const float *g_m, const float *g_logs, const int32_t *dur, float *g_m_p, float *g_logs_p, int B,
                       int D, int T, int S, cudaStream_t stream, int *rc)
{
    *rc = MAS_OK;
    if (T % 4 != 0 || B > 65535) return false;
    CUtensorMap tm_m, tm_l;
    const int n32 = (D + 31) / 32;
    // consumer warps per CTA: the most that divides the channel groups while the grid still covers the SMs
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int nw = 1;
    for (int w = kSegMaxWarps; w >= 2; --w)
        if (n32 % w == 0 && (n32 / w) * B * (g_logs ? 2 : 1) >= sms) {
            nw = w;
            break;
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
    int stages = config().seg_stages;
    if (stages < 2 || stages > kSegStages) stages = 3;   // more CTAs per SM beat deeper rings (measured 3 / 4 / 6 / 8)
    const size_t smem = segsum_smem(nw, stages, S, T);
    static thread_local int configured_dev = -1;
    if (dev != configured_dev) {
        if (cudaFuncSetAttribute(mas_segsum_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)segsum_smem(kSegMaxWarps, kSegStages, MAS_MAX_TEXT, MAS_MAX_MEL)) != cudaSuccess) {
            cudaGetLastError();
            return false;
        }
        configured_dev = dev;
    }
    SegParams p{dur, g_m_p, g_logs_p, D, T, S, nw, stages};
    const dim3 grid((unsigned)((n32 + nw - 1) / nw), (unsigned)B, g_logs ? 2u : 1u);
    mas_segsum_kernel<<<grid, 32 * (nw + 1), smem, stream>>>(tm_m, tm_l, p);
    note_launch();
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) *rc = note_cuda_error(e, "mas_segsum_kernel");
    return true;
}

}  // namespace mas
