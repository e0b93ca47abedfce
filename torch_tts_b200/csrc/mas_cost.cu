// mas_cost.cu -- the neg_cent cost block of SynthesizerTrn.forward
// (reference vits2/models.py:1226-1247): prior preparation, the contraction
// over the D-channel prior, the noise statistics and the VITS2 noise add.
//
//   neg_cent[b,t,s] = bias[b,s] + sum_d (-0.5 z^2)[d,t] * r[d,s] + z[d,t] * (m r)[d,s]
//   r = exp(-2 logs_p)                                   (models.py:1226)
//   bias = sum_d(-0.5 log 2pi - logs_p) + sum_d(-0.5 m^2 r)   (:1227-1229, :1236-1238)
//
// The contraction itself is the tcgen05 kernel of mas_cost_tc.cu; this file holds the
// host dispatch and the explicit noise pass (used only when the caller asks for the
// noised cost plane itself).
#include "mas_common.cuh"

namespace mas {

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

// Rows of S cells between planes of different row pitch, with the noise folded in when there is any: the
// contraction writes planes whose rows are a multiple of 16 bytes (tensor-map stores), the caller's [B,T,S]
// tensor is packed.  src == dst is allowed when ld_src == ld_dst.  noise rows are packed (pitch S).
__global__ void mas_rows_kernel(const float *__restrict__ src, int ld_src, const float *__restrict__ noise,
                                const double *__restrict__ stats, float scale, float *__restrict__ dst, int ld_dst,
                                size_t rows, int S, size_t n_stat)
{
    float sd = 0.f;
    if (noise) {
        const double mean = stats[0] / (double)n_stat;
        double var = (stats[1] - stats[0] * mean) / (double)(n_stat > 1 ? n_stat - 1 : 1);
        if (var < 0) var = 0;
        sd = (float)sqrt(var);
    }
    const size_t n = rows * (size_t)S, stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const size_t r = i / (size_t)S;
        const int c = (int)(i - r * (size_t)S);
        float v = src[r * ld_src + c];
        if (noise) v = __fadd_rn(v, __fmul_rn(__fmul_rn(sd, noise[i]), scale));
        dst[r * ld_dst + c] = v;
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
                   cudaStream_t stream, int ld);

size_t cost_workspace_bytes(int B, int D, int T, int S) { return cost_tc_workspace_bytes(B, D, T, S); }

int cost_launch(const float *z_p, const float *m_p, const float *logs_p, float *neg_cent_out, double *stats_out,
                const int32_t *t_ys, void *workspace, size_t workspace_bytes, int B, int D, int T, int S,
                cudaStream_t stream, int ld)
{
    // rows of the output plane are `ld` floats apart; 16-byte multiples go out through the tensor-map engine, anything
    // else through the (much slower) per-cell stores of the fallback epilogue
    if (ld <= 0) ld = S;
    if (ld < S) return MAS_ERR_ALIGNMENT;
    if (!cost_tc_supported(B, D, T, S)) return MAS_ERR_UNSUPPORTED_SHAPE;
    if (workspace_bytes < cost_workspace_bytes(B, D, T, S) || !workspace) return MAS_ERR_WORKSPACE;
    return cost_tc_launch(z_p, m_p, logs_p, neg_cent_out, stats_out, t_ys, workspace, workspace_bytes, B, D, T, S, stream,
                          ld);
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

// dst[r][0..S) = src[r][0..S) (+ (std * noise[r]) * scale), rows `ld_src` / `ld_dst` floats apart
int rows_launch(const float *src, int ld_src, const float *noise, const double *stats, float scale, float *dst,
                int ld_dst, size_t rows, int S, cudaStream_t stream)
{
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    mas_rows_kernel<<<sms * 8, 256, 0, stream>>>(src, ld_src, noise, stats, scale, dst, ld_dst, rows, S,
                                                 rows * (size_t)S);
    note_launch();
    MAS_CUDA_TRY(cudaGetLastError());
    return MAS_OK;
}

}  // namespace mas
