// mas_api.cu -- the extern "C" surface declared in include/mas_b200.h.
// Argument validation on the host, then stream-ordered kernel launches; no
// host synchronisation, no CPU fallback.
#include <stdio.h>
#include <string.h>

#include <mutex>

#include "mas_common.cuh"
#include "mas_fused.cuh"

namespace mas {

static thread_local long g_launches = 0;
static thread_local char g_cuda_err[256] = "";

void note_launch(int n) { g_launches += n; }

static int env_int(const char *name, int dflt)
{
    const char *s = getenv(name);
    return (s && *s) ? atoi(s) : dflt;
}

static Config g_config;
static std::once_flag g_config_once;
static std::mutex g_config_mutex;

static void read_config()
{
    Config c{};
    c.no_fused = env_int("MAS_NO_FUSED", 0);
    c.dp_vk = env_int("MAS_DP_VK", 1);
    c.dp_warps = env_int("MAS_DP_WARPS", 0);
    c.dp_stages = env_int("MAS_DP_STAGES", 0);
    c.fused_dp_ctas = env_int("MAS_FUSED_DP_CTAS", 0);
    c.fused_rounds = env_int("MAS_FUSED_ROUNDS", -1);
    c.fused_zero_offload = env_int("MAS_FUSED_ZERO_OFFLOAD", 1);
    c.fused_pdl = env_int("MAS_FUSED_PDL", 1);
    c.noise_fused = env_int("MAS_NOISE_FUSED", 1);
    c.noise_feed = env_int("MAS_NOISE_FEED", 1);
    c.segsum = env_int("MAS_SEGSUM", 1);
    c.seg_stages = env_int("MAS_SEG_STAGES", 0);
    c.seg_parts = env_int("MAS_SEG_PARTS", 0);
    c.seg_nw = env_int("MAS_SEG_NW", 0);
    c.stage = env_int("MAS_STAGE", 0);
    c.tc_pair = 1;
    c.tc_no_tma = env_int("MAS_TC_NO_TMA", 0);   // host-side choice of the no-tensor-map fallbacks (tests force them)
    if (kTrace) {
        c.tc_debug = env_int("MAS_TC_DEBUG", 0);
        c.dp_debug = env_int("MAS_DP_DEBUG", 0);
        c.tc_grid = env_int("MAS_TC_GRID", 0);
        c.tc_pair = env_int("MAS_TC_PAIR", 1);
        c.trace = env_int("MAS_TRACE", 0);
    }
    std::lock_guard<std::mutex> lock(g_config_mutex);
    g_config = c;
}

const Config &config()
{
    std::call_once(g_config_once, read_config);
    return g_config;
}

// trace build + MAS_TRACE=1: one buffer per device, allocated on first use
unsigned long long *trace_buffer()
{
    if (!kTrace || !config().trace) return nullptr;
    static unsigned long long *bufs[64] = {};
    static std::mutex m;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
    std::lock_guard<std::mutex> lock(m);
    if (!bufs[dev]) {
        if (cudaMalloc(&bufs[dev], kTraceWords * sizeof(unsigned long long)) != cudaSuccess) bufs[dev] = nullptr;
        if (bufs[dev]) cudaMemset(bufs[dev], 0, kTraceWords * sizeof(unsigned long long));
    }
    return bufs[dev];
}

int note_cuda_error(cudaError_t e, const char *what)
{
    snprintf(g_cuda_err, sizeof(g_cuda_err), "%s: %s", what, cudaGetErrorString(e));
    return MAS_ERR_CUDA;
}

// mas_dp.cu
size_t dp_workspace_bytes(int B, int T, int S);
int dp_launch(const float *neg_cent, const int32_t *t_ys, const int32_t *t_xs, void *path_out, int path_dtype,
              int32_t *dur_out, int32_t *idx_out, int32_t *status_out, void *workspace, size_t workspace_bytes, int B,
              int T, int S, cudaStream_t stream, const float *noise = nullptr, const double *stats = nullptr,
              float noise_scale = 0.f, int ld = 0);
bool dp_noise_supported(const float *neg_cent, const float *noise, int S);
int lengths_launch(const float *mask, int32_t *t_ys, int32_t *t_xs, int B, int T, int S, cudaStream_t stream);
int expand_launch(const int32_t *idx, void *path_out, int path_dtype, int B, int T, int S, cudaStream_t stream);
// mas_expand.cu
int expand_prior_launch(const float *m_p, const float *logs_p, const int32_t *idx, float *m_out, float *logs_out, int B,
                        int D, int T, int S, cudaStream_t stream);
int expand_prior_backward_launch(const float *g_m, const float *g_logs, const int32_t *dur, float *g_m_p, float *g_logs_p,
                                 int B, int D, int T, int S, cudaStream_t stream);
int logw_launch(const int32_t *dur, const int32_t *t_xs, float *out, int B, int S, cudaStream_t stream);
int idx_from_durations_launch(const float *dur_f, const int32_t *t_xs, const int32_t *t_ys, int32_t *idx, int B, int T,
                              int S, cudaStream_t stream);
// mas_cost.cu
size_t cost_workspace_bytes(int B, int D, int T, int S);
int cost_launch(const float *z_p, const float *m_p, const float *logs_p, float *neg_cent_out, double *stats_out,
                const int32_t *t_ys, void *workspace, size_t workspace_bytes, int B, int D, int T, int S,
                cudaStream_t stream, int ld = 0);
int rows_launch(const float *src, int ld_src, const float *noise, const double *stats, float scale, float *dst,
                int ld_dst, size_t rows, int S, cudaStream_t stream);
int add_noise_launch(const float *nc, const float *noise, const double *stats, float scale, float *out, size_t n,
                     cudaStream_t stream);

static bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

static int check_shape(int B, int T, int S)
{
    if (B < 1 || T < 1 || S < 1) return MAS_ERR_BAD_SHAPE;
    if (S > MAS_MAX_TEXT || T > MAS_MAX_MEL) return MAS_ERR_UNSUPPORTED_SHAPE;
    return MAS_OK;
}

static int check_dtype(int path_dtype)
{
    return (path_dtype >= MAS_PATH_F32 && path_dtype <= MAS_PATH_I32) ? MAS_OK : MAS_ERR_BAD_DTYPE;
}

}  // namespace mas

using namespace mas;

extern "C" {

int mas_b200_abi_version(void) { return 2; }   // 2: path_out may be NULL (compact outputs only), mas_reload_config

void mas_reload_config(void)
{
    (void)config();
    read_config();
}

const char *mas_status_string(int code)
{
    switch (code) {
        case MAS_OK: return "ok";
        case MAS_ERR_NULL_POINTER: return "null pointer argument";
        case MAS_ERR_BAD_SHAPE: return "B, T, S and D must be >= 1";
        case MAS_ERR_UNSUPPORTED_SHAPE: return "shape outside the supported range (S <= 1024, T <= 65535)";
        case MAS_ERR_ALIGNMENT: return "base pointer must be 16-byte aligned";
        case MAS_ERR_WORKSPACE: return "workspace missing or too small";
        case MAS_ERR_BAD_DTYPE: return "unknown path dtype";
        case MAS_ERR_CUDA: return "CUDA runtime error (see mas_last_cuda_error)";
    }
    return "unknown status";
}

const char *mas_last_cuda_error(void) { return g_cuda_err; }

long mas_take_launch_count(void)
{
    long n = g_launches;
    g_launches = 0;
    return n;
}

int mas_lengths_from_mask_f32(const float *mask, int32_t *t_ys, int32_t *t_xs, int B, int T, int S, void *stream)
{
    if (!mask || !t_ys || !t_xs) return MAS_ERR_NULL_POINTER;
    if (B < 1 || T < 1 || S < 1) return MAS_ERR_BAD_SHAPE;
    return lengths_launch(mask, t_ys, t_xs, B, T, S, static_cast<cudaStream_t>(stream));
}

size_t mas_maximum_path_workspace_bytes(int B, int T, int S)
{
    if (check_shape(B, T, S) != MAS_OK) return 0;
    return dp_workspace_bytes(B, T, S);
}

int mas_maximum_path_f32(const float *neg_cent, const int32_t *t_ys, const int32_t *t_xs, void *path_out,
                         int path_dtype, int32_t *dur_out, int32_t *idx_out, int32_t *status_out, void *workspace,
                         size_t workspace_bytes, int B, int T, int S, void *stream)
{
    if (!neg_cent || !t_ys || !t_xs) return MAS_ERR_NULL_POINTER;
    if (!path_out && !dur_out && !idx_out) return MAS_ERR_NULL_POINTER;   // nothing to write
    int rc = check_shape(B, T, S);
    if (rc) return rc;
    rc = check_dtype(path_dtype);
    if (rc) return rc;
    if (!aligned16(neg_cent) || !aligned16(workspace)) return MAS_ERR_ALIGNMENT;
    return dp_launch(neg_cent, t_ys, t_xs, path_out, path_dtype, dur_out, idx_out, status_out, workspace,
                     workspace_bytes, B, T, S, static_cast<cudaStream_t>(stream));
}

size_t mas_neg_cent_workspace_bytes(int B, int D, int T, int S)
{
    if (check_shape(B, T, S) != MAS_OK || D < 1) return 0;
    // S % 4 != 0: the contraction stores 16-byte rows into a padded plane in the workspace, a second kernel packs them
    return align_up(cost_workspace_bytes(B, D, T, S), 256) + (S % 4 ? fused_plane_bytes(B, T, S) : 0);
}

int mas_neg_cent_f32(const float *z_p, const float *m_p, const float *logs_p, float *neg_cent_out, double *stats_out,
                     void *workspace, size_t workspace_bytes, int B, int D, int T, int S, void *stream)
{
    if (!z_p || !m_p || !logs_p || !neg_cent_out) return MAS_ERR_NULL_POINTER;
    int rc = check_shape(B, T, S);
    if (rc) return rc;
    if (D < 1) return MAS_ERR_BAD_SHAPE;
    if (!aligned16(z_p) || !aligned16(m_p) || !aligned16(logs_p) || !aligned16(neg_cent_out) || !aligned16(workspace))
        return MAS_ERR_ALIGNMENT;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (S % 4 == 0)
        return cost_launch(z_p, m_p, logs_p, neg_cent_out, stats_out, nullptr, workspace, workspace_bytes, B, D, T, S, st);
    if (!workspace || workspace_bytes < mas_neg_cent_workspace_bytes(B, D, T, S)) return MAS_ERR_WORKSPACE;
    const size_t cost_ws = align_up(cost_workspace_bytes(B, D, T, S), 256);
    float *plane = reinterpret_cast<float *>(static_cast<unsigned char *>(workspace) + cost_ws);
    const int ld = (S + 3) & ~3;
    rc = cost_launch(z_p, m_p, logs_p, plane, stats_out, nullptr, workspace, cost_ws, B, D, T, S, st, ld);
    if (rc) return rc;
    return rows_launch(plane, ld, nullptr, nullptr, 0.f, neg_cent_out, S, (size_t)B * T, S, st);
}

// fused workspace layout: [cost ws][stats 256 B][private cost plane, rows padded to 16 bytes][dp ws][flags]
size_t mas_fused_align_workspace_bytes(int B, int D, int T, int S, int with_noise)
{
    (void)with_noise;
    if (check_shape(B, T, S) != MAS_OK || D < 1) return 0;
    return align_up(cost_workspace_bytes(B, D, T, S), 256) + 256 + fused_plane_bytes(B, T, S) +
           align_up(dp_workspace_bytes(B, T, S), 256) + fused_flags_bytes(B, T);
}

int mas_fused_align_f32(const float *z_p, const float *m_p, const float *logs_p, const int32_t *t_ys,
                        const int32_t *t_xs, const float *noise, float noise_scale, void *path_out, int path_dtype,
                        int32_t *dur_out, int32_t *idx_out, int32_t *status_out, float *neg_cent_out, void *workspace,
                        size_t workspace_bytes, int B, int D, int T, int S, void *stream)
{
    if (!z_p || !m_p || !logs_p || !t_ys || !t_xs) return MAS_ERR_NULL_POINTER;
    if (!path_out && !dur_out && !idx_out) return MAS_ERR_NULL_POINTER;   // nothing to write
    int rc = check_shape(B, T, S);
    if (rc) return rc;
    if (D < 1) return MAS_ERR_BAD_SHAPE;
    rc = check_dtype(path_dtype);
    if (rc) return rc;
    if (!aligned16(z_p) || !aligned16(m_p) || !aligned16(logs_p) || !aligned16(workspace) ||
        (neg_cent_out && !aligned16(neg_cent_out)))
        return MAS_ERR_ALIGNMENT;
    if (!workspace || workspace_bytes < mas_fused_align_workspace_bytes(B, D, T, S, noise != nullptr))
        return MAS_ERR_WORKSPACE;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    unsigned char *ws = static_cast<unsigned char *>(workspace);
    const size_t cost_ws = align_up(cost_workspace_bytes(B, D, T, S), 256);
    double *stats = reinterpret_cast<double *>(ws + cost_ws);
    float *plane = reinterpret_cast<float *>(ws + cost_ws + 256);
    unsigned char *dp_ws = ws + cost_ws + 256 + fused_plane_bytes(B, T, S);
    const size_t dp_ws_bytes = align_up(dp_workspace_bytes(B, T, S), 256);
    // One kernel (plus the prior preparation) when the cost plane can stay private: without noise the DP trails
    // the contraction tile by tile; with noise the kernel has a grid barrier between them (mas_fused.cu).
    if (!neg_cent_out && (noise ? fused_noise_supported(B, D, T, S) : fused_supported(B, D, T, S))) {
        uint32_t *flags = reinterpret_cast<uint32_t *>(dp_ws + dp_ws_bytes);
        rc = fused_launch(z_p, m_p, logs_p, t_ys, t_xs, noise, noise_scale, stats, plane, path_out, path_dtype, dur_out,
                          idx_out, status_out, ws, cost_ws, dp_ws, dp_ws_bytes, flags, B, D, T, S, st);
        if (rc != kFusedFallback) return rc;
        // the cooperative grid does not fit this context: the same work as separate launches below
    }
    // mel tiles wholly past t_y are skipped only when the plane is private scratch: a caller who asked for
    // neg_cent_out gets every cell the reference would compute.  The contraction stores rows of 16-byte multiples:
    // S % 4 != 0 goes through the padded plane of the workspace.
    const int ld = (S + 3) & ~3;
    const bool packed = (ld == S);
    float *nc = (neg_cent_out && packed) ? neg_cent_out : plane;
    rc = cost_launch(z_p, m_p, logs_p, nc, noise ? stats : nullptr, neg_cent_out ? nullptr : t_ys, ws, cost_ws, B, D,
                     T, S, st, ld);
    if (rc) return rc;
    if (config().stage == 1) return MAS_OK;
    if (!packed) {
        // pack the rows into the caller's tensor, or (private plane) add the noise in place; the DP reads either pitch
        if (neg_cent_out || noise) {
            float *dst = neg_cent_out ? neg_cent_out : plane;
            rc = rows_launch(plane, ld, noise, stats, noise_scale, dst, neg_cent_out ? S : ld, (size_t)B * T, S, st);
            if (rc) return rc;
            noise = nullptr;
        }
        return dp_launch(neg_cent_out ? neg_cent_out : plane, t_ys, t_xs, path_out, path_dtype, dur_out, idx_out,
                         status_out, dp_ws, dp_ws_bytes, B, T, S, st, nullptr, nullptr, 0.f, neg_cent_out ? S : ld);
    }
    if (noise && (neg_cent_out || !dp_noise_supported(nc, noise, S))) {
        // the caller wants the noised cost plane itself (or the noise rows are not 16-byte aligned): one more pass
        rc = add_noise_launch(nc, noise, stats, noise_scale, nc, (size_t)B * T * S, st);
        if (rc) return rc;
        noise = nullptr;
    }
    // otherwise the DP adds (std * noise) * scale while the cost streams in
    return dp_launch(nc, t_ys, t_xs, path_out, path_dtype, dur_out, idx_out, status_out, dp_ws, dp_ws_bytes, B, T, S,
                     st, noise, stats, noise_scale);
}

int mas_debug_read_trace(unsigned long long *host_out, int n_words)
{
    unsigned long long *buf = trace_buffer();
    if (!buf || !host_out) return MAS_ERR_NULL_POINTER;
    if (n_words > kTraceWords) n_words = kTraceWords;
    MAS_CUDA_TRY(cudaMemcpy(host_out, buf, (size_t)n_words * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    return MAS_OK;
}

int mas_expand_path(const int32_t *idx, void *path_out, int path_dtype, int B, int T, int S, void *stream)
{
    if (!idx || !path_out) return MAS_ERR_NULL_POINTER;
    if (B < 1 || T < 1 || S < 1) return MAS_ERR_BAD_SHAPE;
    int rc = check_dtype(path_dtype);
    if (rc) return rc;
    return expand_launch(idx, path_out, path_dtype, B, T, S, static_cast<cudaStream_t>(stream));
}

int mas_expand_prior_f32(const float *m_p, const float *logs_p, const int32_t *idx, float *m_out, float *logs_out, int B,
                         int D, int T, int S, void *stream)
{
    if (!m_p || !idx || !m_out) return MAS_ERR_NULL_POINTER;
    if ((logs_p == nullptr) != (logs_out == nullptr)) return MAS_ERR_NULL_POINTER;
    if (B < 1 || D < 1 || T < 1 || S < 1) return MAS_ERR_BAD_SHAPE;
    if (S > MAS_MAX_TEXT || T > MAS_MAX_MEL) return MAS_ERR_UNSUPPORTED_SHAPE;
    return expand_prior_launch(m_p, logs_p, idx, m_out, logs_out, B, D, T, S, static_cast<cudaStream_t>(stream));
}

int mas_expand_prior_backward_f32(const float *g_m, const float *g_logs, const int32_t *dur, float *g_m_p, float *g_logs_p,
                                  int B, int D, int T, int S, void *stream)
{
    if (!g_m || !dur || !g_m_p) return MAS_ERR_NULL_POINTER;
    if ((g_logs == nullptr) != (g_logs_p == nullptr)) return MAS_ERR_NULL_POINTER;
    if (B < 1 || D < 1 || T < 1 || S < 1) return MAS_ERR_BAD_SHAPE;
    if (S > MAS_MAX_TEXT || T > MAS_MAX_MEL) return MAS_ERR_UNSUPPORTED_SHAPE;
    return expand_prior_backward_launch(g_m, g_logs, dur, g_m_p, g_logs_p, B, D, T, S, static_cast<cudaStream_t>(stream));
}

int mas_logw_f32(const int32_t *dur, const int32_t *t_xs, float *logw_out, int B, int S, void *stream)
{
    if (!dur || !t_xs || !logw_out) return MAS_ERR_NULL_POINTER;
    if (B < 1 || S < 1) return MAS_ERR_BAD_SHAPE;
    return logw_launch(dur, t_xs, logw_out, B, S, static_cast<cudaStream_t>(stream));
}

int mas_idx_from_durations_f32(const float *durations, const int32_t *t_xs, const int32_t *t_ys, int32_t *idx_out, int B,
                               int T, int S, void *stream)
{
    if (!durations || !t_xs || !idx_out) return MAS_ERR_NULL_POINTER;
    if (B < 1 || T < 1 || S < 1) return MAS_ERR_BAD_SHAPE;
    if (S > MAS_MAX_TEXT || T > MAS_MAX_MEL) return MAS_ERR_UNSUPPORTED_SHAPE;
    return idx_from_durations_launch(durations, t_xs, t_ys, idx_out, B, T, S, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
