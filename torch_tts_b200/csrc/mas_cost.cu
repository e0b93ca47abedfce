// mas_cost.cu -- the neg_cent cost block of SynthesizerTrn.forward
// (reference vits2/models.py:1226-1247): prior preparation, the contraction
// over the D-channel prior, the noise statistics and the VITS2 noise add.
//
//   neg_cent[b,t,s] = bias[b,s] + sum_d (-0.5 z^2)[d,t] * r[d,s] + z[d,t] * (m r)[d,s]
//   r = exp(-2 logs_p)                                   (models.py:1226)
//   bias = sum_d(-0.5 log 2pi - logs_p) + sum_d(-0.5 m^2 r)   (:1227-1229, :1236-1238)
//
// This file holds the SIMT fp32 contraction (exact-fp32 products, used as the
// on-device yardstick and for shapes the tensor-core kernel does not take) and
// the small prep / statistics / noise kernels.  The tcgen05 contraction lives
// in mas_cost_tc.cu.
#include "mas_common.cuh"

namespace mas {

// ---------------------------------------------------------------------------
// prior preparation: r, m*r, bias   (one thread per (b, s))
// ---------------------------------------------------------------------------
__global__ void mas_prior_prep_kernel(const float *__restrict__ m_p, const float *__restrict__ logs_p,
                                      float *__restrict__ r_out, float *__restrict__ mr_out,
                                      float *__restrict__ bias_out, int D, int S)
{
    const int b = blockIdx.y;
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= S) return;
    const size_t base = (size_t)b * D * S + s;
    const float c0 = -0.91893853320467274178f;  // -0.5*log(2*pi)
    float acc1 = 0.f, acc4 = 0.f;
    for (int d = 0; d < D; ++d) {
        const float l = logs_p[base + (size_t)d * S];
        const float m = m_p[base + (size_t)d * S];
        const float r = expf(-2.0f * l);
        r_out[base + (size_t)d * S] = r;
        mr_out[base + (size_t)d * S] = m * r;
        acc1 += c0 - l;
        acc4 += -0.5f * (m * m) * r;
    }
    bias_out[(size_t)b * S + s] = acc1 + acc4;
}

// ---------------------------------------------------------------------------
// SIMT fp32 contraction: 64(t) x 64(s) tile, 256 threads, 4x4 per thread
// ---------------------------------------------------------------------------
constexpr int kTM = 64, kTN = 64, kTK = 16;

__global__ void __launch_bounds__(256) mas_cost_simt_kernel(const float *__restrict__ z_p, const float *__restrict__ r_in,
                                                           const float *__restrict__ mr_in,
                                                           const float *__restrict__ bias, float *__restrict__ out,
                                                           double *stats, int D, int T, int S)
{
    __shared__ float zs[kTK][kTM];
    __shared__ float rs[kTK][kTN];
    __shared__ float ms[kTK][kTN];
    const int b = blockIdx.z;
    const int t0 = blockIdx.y * kTM, s0 = blockIdx.x * kTN;
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const float *zb = z_p + (size_t)b * D * T;
    const float *rb = r_in + (size_t)b * D * S;
    const float *mb = mr_in + (size_t)b * D * S;
    float acc[4][4] = {};
    for (int d0 = 0; d0 < D; d0 += kTK) {
        for (int i = tid; i < kTK * kTM; i += 256) {
            const int kk = i / kTM, c = i % kTM;
            const int d = d0 + kk;
            zs[kk][c] = (d < D && t0 + c < T) ? zb[(size_t)d * T + t0 + c] : 0.f;
            const bool ok = d < D && s0 + c < S;
            rs[kk][c] = ok ? rb[(size_t)d * S + s0 + c] : 0.f;
            ms[kk][c] = ok ? mb[(size_t)d * S + s0 + c] : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < kTK; ++kk) {
            float a[4], a2[4], br[4], bm[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                a[i] = zs[kk][ty * 4 + i];
                a2[i] = -0.5f * (a[i] * a[i]);
                br[i] = rs[kk][tx * 4 + i];
                bm[i] = ms[kk][tx * 4 + i];
            }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], bm[j], fmaf(a2[i], br[j], acc[i][j]));
        }
        __syncthreads();
    }
    double ssum = 0.0, ssq = 0.0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int t = t0 + ty * 4 + i;
        if (t >= T) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int s = s0 + tx * 4 + j;
            if (s >= S) continue;
            const float val = acc[i][j] + bias[(size_t)b * S + s];
            out[((size_t)b * T + t) * S + s] = val;
            ssum += (double)val;
            ssq += (double)val * (double)val;
        }
    }
    if (stats) {
        for (int o = 16; o > 0; o >>= 1) {
            ssum += __shfl_xor_sync(kFullMask, ssum, o);
            ssq += __shfl_xor_sync(kFullMask, ssq, o);
        }
        __shared__ double red[2][8];
        if ((tid & 31) == 0) {
            red[0][tid >> 5] = ssum;
            red[1][tid >> 5] = ssq;
        }
        __syncthreads();
        if (tid == 0) {
            double a = 0, q = 0;
            for (int i = 0; i < 8; ++i) {
                a += red[0][i];
                q += red[1][i];
            }
            atomicAdd(&stats[0], a);
            atomicAdd(&stats[1], q);
        }
    }
}

// ---------------------------------------------------------------------------
// VITS2 noise: out = nc + (std * noise) * scale     (models.py:1241-1247)
// std = unbiased standard deviation over all n cells from {sum, sumsq} in fp64
// ---------------------------------------------------------------------------
__global__ void mas_add_noise_kernel(const float *__restrict__ nc, const float *__restrict__ noise,
                                     const double *__restrict__ stats, float scale, float *__restrict__ out, size_t n)
{
    const double mean = stats[0] / (double)n;
    double var = (stats[1] - stats[0] * mean) / (double)(n > 1 ? n - 1 : 1);
    if (var < 0) var = 0;
    const float sd = (float)sqrt(var);
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        out[i] = __fadd_rn(nc[i], __fmul_rn(__fmul_rn(sd, noise[i]), scale));  // rounded after every operation
    }
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
// mas_cost_tc.cu
bool cost_tc_supported(int B, int D, int T, int S);
size_t cost_tc_workspace_bytes(int B, int D, int T, int S);
int cost_tc_launch(const float *z_p, const float *m_p, const float *logs_p, float *neg_cent_out, double *stats_out,
                   const int32_t *t_ys, void *workspace, size_t workspace_bytes, int B, int D, int T, int S,
                   cudaStream_t stream);

static bool use_tc(int B, int D, int T, int S)
{
    const char *e = getenv("MAS_COST_IMPL");  // "simt" forces the fp32 SIMT contraction (yardstick / A-B runs)
    if (e && e[0] == 's') return false;
    return cost_tc_supported(B, D, T, S);
}

size_t cost_workspace_bytes(int B, int D, int T, int S)
{
    const size_t prior = align_up((size_t)B * D * S * 4, 256);
    const size_t simt = 2 * prior + align_up((size_t)B * S * 4, 256);
    const size_t tc = cost_tc_supported(B, D, T, S) ? cost_tc_workspace_bytes(B, D, T, S) : 0;
    return simt > tc ? simt : tc;
}

int cost_launch(const float *z_p, const float *m_p, const float *logs_p, float *neg_cent_out, double *stats_out,
                const int32_t *t_ys, void *workspace, size_t workspace_bytes, int B, int D, int T, int S,
                cudaStream_t stream)
{
    if (workspace_bytes < cost_workspace_bytes(B, D, T, S) || !workspace) return MAS_ERR_WORKSPACE;
    if (use_tc(B, D, T, S))
        return cost_tc_launch(z_p, m_p, logs_p, neg_cent_out, stats_out, t_ys, workspace, workspace_bytes, B, D, T, S,
                              stream);
    unsigned char *ws = static_cast<unsigned char *>(workspace);
    const size_t prior = align_up((size_t)B * D * S * 4, 256);
    float *r = reinterpret_cast<float *>(ws);
    float *mr = reinterpret_cast<float *>(ws + prior);
    float *bias = reinterpret_cast<float *>(ws + 2 * prior);
    if (stats_out) MAS_CUDA_TRY(cudaMemsetAsync(stats_out, 0, 2 * sizeof(double), stream));
    mas_prior_prep_kernel<<<dim3((S + 127) / 128, B), 128, 0, stream>>>(m_p, logs_p, r, mr, bias, D, S);
    note_launch();
    MAS_CUDA_TRY(cudaGetLastError());
    mas_cost_simt_kernel<<<dim3((S + kTN - 1) / kTN, (T + kTM - 1) / kTM, B), 256, 0, stream>>>(
        z_p, r, mr, bias, neg_cent_out, stats_out, D, T, S);
    note_launch();
    MAS_CUDA_TRY(cudaGetLastError());
    return MAS_OK;
}

int add_noise_launch(const float *nc, const float *noise, const double *stats, float scale, float *out, size_t n,
                     cudaStream_t stream)
{
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    mas_add_noise_kernel<<<sms * 8, 256, 0, stream>>>(nc, noise, stats, scale, out, n);
    note_launch();
    MAS_CUDA_TRY(cudaGetLastError());
    return MAS_OK;
}

}  // namespace mas
