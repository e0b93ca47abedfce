// mas_expand.cu -- what SynthesizerTrn does with the alignment right after MAS:
//   * prior expansion  m_p, logs_p [B,D,S] -> [B,D,T]:  the reference multiplies by the one-hot path
//     (`torch.matmul(attn.squeeze(1), m_p.transpose(1, 2)).transpose(1, 2)`, vits2/models.py:1270-1271), two
//     [T,S] x [S,D] GEMMs per utterance and a read of the dense plane each; with the compact idx [B,T] the
//     backtrack already emits it is a gather, and its backward (m_p, logs_p require grad) a segmented sum;
//   * logw_ = log(w + 1e-6) * x_mask with w = attn.sum(2) (models.py:1256, 1261): from the int32 durations.
// All three are HBM-bound streams; no tensor-core work.
#include "mas_common.cuh"

namespace mas {

// out[b, d, t] = idx[b, t] >= 0 ? in[b, d, idx[b, t]] : 0      (rows past t_y are all-zero in the path)
// One CTA = 4 * 128 consecutive mel frames x kDBlock channels of one utterance; a thread owns 4 consecutive
// frames (one 16-byte store per channel), looks its 4 columns up once and walks the channels.  Neighbouring
// frames map to the same or the next text column, so the gathered loads hit the same sectors.
constexpr int kGatherThreads = 128;
constexpr int kDBlock = 16;

template <bool kVec, bool kTwo>
__global__ void __launch_bounds__(kGatherThreads) mas_gather_prior_kernel(const float *__restrict__ m_p,
                                                                          const float *__restrict__ logs_p,
                                                                          const int32_t *__restrict__ idx,
                                                                          float *__restrict__ m_out,
                                                                          float *__restrict__ logs_out, int D, int T, int S)
{
    const int b = blockIdx.z, d0 = blockIdx.y * kDBlock;
    const int t0 = (blockIdx.x * kGatherThreads + threadIdx.x) * 4;
    if (t0 >= T) return;
    int c[4];
    if (kVec) {
        const int4 q = *reinterpret_cast<const int4 *>(idx + (size_t)b * T + t0);
        c[0] = q.x, c[1] = q.y, c[2] = q.z, c[3] = q.w;
    } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) c[j] = (t0 + j < T) ? idx[(size_t)b * T + t0 + j] : -1;
    }
    const int d1 = min(d0 + kDBlock, D);
    for (int d = d0; d < d1; ++d) {
        const size_t row_in = ((size_t)b * D + d) * S, row_out = ((size_t)b * D + d) * T + t0;
        float a[4], l[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const bool live = c[j] >= 0 && c[j] < S;
            a[j] = live ? __ldg(m_p + row_in + c[j]) : 0.0f;
            if (kTwo) l[j] = live ? __ldg(logs_p + row_in + c[j]) : 0.0f;
        }
        if (kVec) {
            *reinterpret_cast<float4 *>(m_out + row_out) = make_float4(a[0], a[1], a[2], a[3]);
            if (kTwo) *reinterpret_cast<float4 *>(logs_out + row_out) = make_float4(l[0], l[1], l[2], l[3]);
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (t0 + j < T) {
                    m_out[row_out + j] = a[j];
                    if (kTwo) logs_out[row_out + j] = l[j];
                }
        }
    }
}

// backward of the gather: g_in[b, d, s] = sum over the frames aligned to column s of g_out[b, d, t].
// The path is monotonic, so those frames are the contiguous range [start_s, start_s + dur_s) with start = the
// exclusive prefix sum of the durations: a segmented sum in a fixed (ascending t) order, no atomics.
// One CTA = one utterance x kScatterD channels.  The gradient rows stream through shared memory in chunks of
// kScatterChunk frames with coalesced 16-byte loads; a thread then owns a text column (up to 4 for S = 1024)
// and adds up the part of its segment that lies in the chunk.
// (Tried in round 2 and dropped: no shared tile at all -- a thread reading its own segment of 8 + 8 channel rows
// straight from global memory, neighbouring columns being neighbouring frames.  Every element comes from DRAM
// once, but the scalar loads of a warp straddle 4-5 lines each and the kernel ran 328 us against 93 us here.)
constexpr int kScatterThreads = 256;
constexpr int kScatterD = 4;
constexpr int kScatterChunk = 1024;   // frames per tile row = one 16-byte load per thread

template <bool kTwo, int kScatterCols>   // kScatterCols = ceil(S / 256) text columns per thread
__global__ void __launch_bounds__(kScatterThreads) mas_scatter_prior_kernel(const float *__restrict__ g_m,
                                                                            const float *__restrict__ g_logs,
                                                                            const int32_t *__restrict__ dur,
                                                                            float *__restrict__ g_m_p,
                                                                            float *__restrict__ g_logs_p, int D, int T, int S)
{
    constexpr int kRows = kTwo ? 2 * kScatterD : kScatterD;
    __shared__ __align__(16) float tile[kRows][kScatterChunk];
    __shared__ int start_s[MAS_MAX_TEXT + 1];
    __shared__ int warp_tot[kScatterThreads / 32];
    const int b = blockIdx.y, d0 = blockIdx.x * kScatterD, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // exclusive prefix sum of dur[b, :] -- each thread owns ceil(S / 256) consecutive columns
    const int per = (S + kScatterThreads - 1) / kScatterThreads;
    int local = 0;
    for (int j = 0; j < per; ++j) {
        const int s = tid * per + j;
        if (s < S) local += max(dur[(size_t)b * S + s], 0);
    }
    int incl = local;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int n = __shfl_up_sync(kFullMask, incl, o);
        if (lane >= o) incl += n;
    }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    int base = 0;
    for (int w = 0; w < warp; ++w) base += warp_tot[w];
    int run = base + incl - local;
    for (int j = 0; j < per; ++j) {
        const int s = tid * per + j;
        if (s < S) {
            start_s[s] = run;
            run += max(dur[(size_t)b * S + s], 0);
        }
    }
    if (tid == kScatterThreads - 1) start_s[S] = run;
    __syncthreads();

    // row r of the tile: channel d0 + (r % kScatterD) of g_m (r < kScatterD) or g_logs
    const float *rows[kRows];
#pragma unroll
    for (int r = 0; r < kRows; ++r) {
        const int d = min(d0 + (r % kScatterD), D - 1);
        rows[r] = ((kTwo && r >= kScatterD) ? g_logs : g_m) + ((size_t)b * D + d) * T;
    }
    const bool vec = (T % 4 == 0) && ((reinterpret_cast<uintptr_t>(g_m) & 15) == 0) &&
                     (!kTwo || (reinterpret_cast<uintptr_t>(g_logs) & 15) == 0);
    float acc[kScatterCols][kRows];
    int lo[kScatterCols], hi[kScatterCols];
#pragma unroll
    for (int c = 0; c < kScatterCols; ++c) {
        const int s = tid + c * kScatterThreads;
        lo[c] = s < S ? min(start_s[s], T) : 0;
        hi[c] = s < S ? min(start_s[s + 1], T) : 0;
#pragma unroll
        for (int r = 0; r < kRows; ++r) acc[c][r] = 0.0f;
    }
    for (int c0 = 0; c0 < T; c0 += kScatterChunk) {
        const int n = min(kScatterChunk, T - c0);
        if (vec) {
            for (int i = tid * 4; i < n; i += kScatterThreads * 4) {
#pragma unroll
                for (int r = 0; r < kRows; ++r)
                    *reinterpret_cast<float4 *>(&tile[r][i]) = __ldg(reinterpret_cast<const float4 *>(rows[r] + c0 + i));
            }
        } else {
            for (int i = tid; i < n; i += kScatterThreads) {
#pragma unroll
                for (int r = 0; r < kRows; ++r) tile[r][i] = __ldg(rows[r] + c0 + i);
            }
        }
        __syncthreads();
#pragma unroll
        for (int c = 0; c < kScatterCols; ++c) {
            const int a = max(lo[c], c0) - c0, e = min(hi[c], c0 + n) - c0;
            for (int t = a; t < e; ++t) {
#pragma unroll
                for (int r = 0; r < kRows; ++r) acc[c][r] += tile[r][t];
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int c = 0; c < kScatterCols; ++c) {
        const int s = tid + c * kScatterThreads;
        if (s >= S) continue;
#pragma unroll
        for (int r = 0; r < kRows; ++r) {
            const int d = d0 + (r % kScatterD);
            if (d >= D) continue;
            float *dst = (kTwo && r >= kScatterD) ? g_logs_p : g_m_p;
            dst[((size_t)b * D + d) * S + s] = acc[c][r];
        }
    }
}

// logw_[b, s] = log(w + 1e-6) * x_mask  (models.py:1256, 1261); x_mask[b, s] = s < t_x
__global__ void mas_logw_kernel(const int32_t *__restrict__ dur, const int32_t *__restrict__ t_xs, float *__restrict__ out,
                                int B, int S)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)B * S) return;
    const int b = (int)(i / S), s = (int)(i - (size_t)b * S);
    const float w = (float)dur[i];
    const float mask = (s < t_xs[b]) ? 1.0f : 0.0f;
    out[i] = __fmul_rn(logf(__fadd_rn(w, 1e-6f)), mask);
}

// inference side (commons.generate_path, commons.py:130-145): durations -> alignment.
// The reference builds the dense path as sequence_mask(cumsum(duration)) minus its copy shifted by one column,
// times the mask: cell (y, x) is 1 iff cum[x-1] <= y < cum[x], x < t_x, y < t_y, with cum the fp32 running sum and
// y compared as a float.  Here: the compact form idx[b, y] = that x (or -1), from an fp32 prefix sum and a binary
// search per frame; mas_expand_path / mas_expand_prior_f32 take it from there.
// Contract: durations finite and >= 0 (negative or NaN entries count as 0).  The running sum saturates at T + 1
// (every frame index is below it, so larger sums change nothing and Inf / huge values are harmless); for the
// integer-valued durations of models.py:1303 (ceil) every summation order gives the same exact sums, for
// fractional ones the scan's association order may differ from torch's in the last ulp (as torch's own CUDA
// cumsum does from its CPU one).
// One CTA per utterance.
__global__ void __launch_bounds__(kScatterThreads) mas_idx_from_durations_kernel(const float *__restrict__ dur_f,
                                                                                 const int32_t *__restrict__ t_xs,
                                                                                 const int32_t *__restrict__ t_ys,
                                                                                 int32_t *__restrict__ idx, int T, int S)
{
    __shared__ float cum_s[MAS_MAX_TEXT];
    __shared__ float warp_tot[kScatterThreads / 32];
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int t_x = min(max(t_xs[b], 0), S);
    const int per = (S + kScatterThreads - 1) / kScatterThreads;
    const float cap = (float)(T + 1);
    auto clean = [cap](float d) { return d > 0.0f ? fminf(d, cap) : 0.0f; };   // NaN and negatives -> 0
    float local = 0.0f;
    for (int j = 0; j < per; ++j) {
        const int s = tid * per + j;
        if (s < S) local = fminf(local + clean(dur_f[(size_t)b * S + s]), cap);
    }
    float incl = local;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const float n = __shfl_up_sync(kFullMask, incl, o);
        if (lane >= o) incl = fminf(incl + n, cap);
    }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    float run = 0.0f;
    for (int w = 0; w < warp; ++w) run = fminf(run + warp_tot[w], cap);
    run = fminf(run + __shfl_up_sync(kFullMask, incl, 1) * (lane > 0 ? 1.0f : 0.0f), cap);   // exclusive prefix of this thread
    for (int j = 0; j < per; ++j) {
        const int s = tid * per + j;
        if (s < S) {
            run = fminf(run + clean(dur_f[(size_t)b * S + s]), cap);
            cum_s[s] = run;  // inclusive: frames [cum[s-1], cum[s]) belong to column s
        }
    }
    __syncthreads();
    const int t_y = t_ys ? min(max(t_ys[b], 0), T) : T;
    for (int y = tid; y < T; y += kScatterThreads) {
        const float yf = (float)y;
        int lo = 0, hi = t_x;  // first column in [0, t_x) with cum > y, t_x if none
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (cum_s[mid] > yf)
                hi = mid;
            else
                lo = mid + 1;
        }
        idx[(size_t)b * T + y] = (y < t_y && lo < t_x) ? lo : -1;
    }
}

int idx_from_durations_launch(const float *dur_f, const int32_t *t_xs, const int32_t *t_ys, int32_t *idx, int B, int T,
                              int S, cudaStream_t stream)
{
    mas_idx_from_durations_kernel<<<(unsigned)B, kScatterThreads, 0, stream>>>(dur_f, t_xs, t_ys, idx, T, S);
    note_launch();
    MAS_CUDA_TRY(cudaGetLastError());
    return MAS_OK;
}

int expand_prior_launch(const float *m_p, const float *logs_p, const int32_t *idx, float *m_out, float *logs_out, int B,
                        int D, int T, int S, cudaStream_t stream)
{
    const dim3 grid((unsigned)((T + kGatherThreads * 4 - 1) / (kGatherThreads * 4)), (unsigned)((D + kDBlock - 1) / kDBlock),
                    (unsigned)B);
    const bool two = logs_p != nullptr;
    auto misaligned = [](const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15) != 0; };
    const bool vec = (T % 4 == 0) && !misaligned(idx) && !misaligned(m_out) && !(two && misaligned(logs_out));
    if (vec && two)
        mas_gather_prior_kernel<true, true><<<grid, kGatherThreads, 0, stream>>>(m_p, logs_p, idx, m_out, logs_out, D, T, S);
    else if (vec)
        mas_gather_prior_kernel<true, false><<<grid, kGatherThreads, 0, stream>>>(m_p, logs_p, idx, m_out, logs_out, D, T, S);
    else if (two)
        mas_gather_prior_kernel<false, true><<<grid, kGatherThreads, 0, stream>>>(m_p, logs_p, idx, m_out, logs_out, D, T, S);
    else
        mas_gather_prior_kernel<false, false><<<grid, kGatherThreads, 0, stream>>>(m_p, logs_p, idx, m_out, logs_out, D, T, S);
    note_launch();
    MAS_CUDA_TRY(cudaGetLastError());
    return MAS_OK;
}

// mas_segsum.cu: channels on the lanes, tiles through the tensor-map engine (T % 4 == 0, 16-byte aligned gradients)
bool segsum_try_launch(const float *g_m, const float *g_logs, const int32_t *dur, float *g_m_p, float *g_logs_p, int B,
                       int D, int T, int S, cudaStream_t stream, int *rc);

int expand_prior_backward_launch(const float *g_m, const float *g_logs, const int32_t *dur, float *g_m_p, float *g_logs_p,
                                 int B, int D, int T, int S, cudaStream_t stream)
{
    int rc = MAS_OK;
    if (config().segsum && segsum_try_launch(g_m, g_logs, dur, g_m_p, g_logs_p, B, D, T, S, stream, &rc)) return rc;
    const dim3 grid((unsigned)((D + kScatterD - 1) / kScatterD), (unsigned)B);
#define MAS_SCATTER(TWO, COLS) \
    mas_scatter_prior_kernel<TWO, COLS><<<grid, kScatterThreads, 0, stream>>>(g_m, g_logs, dur, g_m_p, g_logs_p, D, T, S)
    const int cols = (S + kScatterThreads - 1) / kScatterThreads;
    if (g_logs) {
        if (cols <= 1) MAS_SCATTER(true, 1);
        else if (cols == 2) MAS_SCATTER(true, 2);
        else MAS_SCATTER(true, 4);
    } else {
        if (cols <= 1) MAS_SCATTER(false, 1);
        else if (cols == 2) MAS_SCATTER(false, 2);
        else MAS_SCATTER(false, 4);
    }
#undef MAS_SCATTER
    note_launch();
    MAS_CUDA_TRY(cudaGetLastError());
    return MAS_OK;
}

int logw_launch(const int32_t *dur, const int32_t *t_xs, float *out, int B, int S, cudaStream_t stream)
{
    const size_t n = (size_t)B * S;
    mas_logw_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(dur, t_xs, out, B, S);
    note_launch();
    MAS_CUDA_TRY(cudaGetLastError());
    return MAS_OK;
}

}  // namespace mas
