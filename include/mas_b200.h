/*
 * mas_b200.h -- C ABI of the B200-native VITS2 alignment hot path
 * (neg_cent cost + monotonic alignment search).
 *
 * This is the drop-in boundary: what a maintainer of kgoba/torch-tts would bind
 * instead of the Cython extension `vits2/monotonic_align/core.pyx`.  Plain
 * pointers and sizes only; no torch types.  All data pointers are DEVICE
 * pointers on the current CUDA device unless stated otherwise; every call is
 * stream-ordered on `stream` (a cudaStream_t passed as void*), never
 * synchronises the host, and is re-entrant across streams and devices.
 *
 * Citations are relative to the reference tree (kgoba/torch-tts).
 *
 * Tensors (all C-contiguous):
 *   neg_cent [B, T, S] float32   cost plane, T = mel frames (reference t_t/t_y),
 *                                S = text tokens (reference t_s/t_x), S contiguous
 *   z_p      [B, D, T] float32   flow output, T contiguous      (models.py:1222)
 *   m_p      [B, D, S] float32   prior mean,   S contiguous     (models.py:1220)
 *   logs_p   [B, D, S] float32   prior log-std                  (models.py:1220)
 *   t_ys     [B] int32           mel lengths  (reference t_t_max, __init__.py:16)
 *   t_xs     [B] int32           text lengths (reference t_s_max, __init__.py:17)
 *   path     [B, T, S]           {0,1}, dtype selected by `path_dtype`
 *   dur      [B, S] int32        path.sum over T  (w of models.py:1256)
 *   idx      [B, T] int32        compact path: text column of mel row y, -1 for y >= t_y
 *   status   [B] int32           per-utterance MAS_UTT_* code (0 = aligned)
 *
 * Length contract: the reference is undefined behaviour unless
 * 1 <= t_x <= t_y <= T and t_x <= S (core.pyx:30-33 reads value[-1,..] /
 * writes path[y,-1] otherwise).  This library never emulates that: an
 * utterance that violates the contract gets an all-zero path, zero durations,
 * idx = -1 and status[b] = MAS_UTT_BAD_LENGTHS.
 */
#ifndef MAS_B200_H_
#define MAS_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* return codes (host-detectable argument errors; 0 = launched) */
enum {
    MAS_OK = 0,
    MAS_ERR_NULL_POINTER = 1,
    MAS_ERR_BAD_SHAPE = 2,        /* B, T, S or D < 1 */
    MAS_ERR_UNSUPPORTED_SHAPE = 3,/* S > MAS_MAX_TEXT, T > MAS_MAX_MEL */
    MAS_ERR_ALIGNMENT = 4,        /* a base pointer is not 16-byte aligned */
    MAS_ERR_WORKSPACE = 5,        /* workspace NULL or smaller than the *_workspace_bytes() answer */
    MAS_ERR_BAD_DTYPE = 6,
    MAS_ERR_CUDA = 7              /* a CUDA runtime call failed; see mas_last_cuda_error() */
};

/* per-utterance device status */
enum {
    MAS_UTT_OK = 0,
    MAS_UTT_BAD_LENGTHS = 1
};

/* dtype of the dense path written by the library (reference returns
 * neg_cent.dtype: __init__.py:19) */
enum {
    MAS_PATH_F32 = 0,
    MAS_PATH_F16 = 1,
    MAS_PATH_BF16 = 2,
    MAS_PATH_I32 = 3              /* the int32 plane core.pyx itself fills */
};

#define MAS_MAX_TEXT 1024         /* S  */
#define MAS_MAX_MEL 65535         /* T  */

int mas_b200_abi_version(void);
const char *mas_status_string(int code);
/* text of the last CUDA error seen by this thread inside the library */
const char *mas_last_cuda_error(void);

/*
 * Lengths from the reference's dense mask (replaces __init__.py:16-17:
 * t_y = mask.sum(1)[:,0], t_x = mask.sum(2)[:,0]).  Reads only column 0 and
 * row 0 of each mask plane (T + S elements, not T*S).
 */
int mas_lengths_from_mask_f32(const float *mask, int32_t *t_ys, int32_t *t_xs,
                              int B, int T, int S, void *stream);

/*
 * Replaces maximum_path_c (core.pyx:38-42) plus the zero-init and dtype cast
 * of the Python wrapper (__init__.py:14,19).
 *   - neg_cent is read-only (the reference's in-place DP works on a private copy,
 *     __init__.py:13).
 *   - path_out is fully written (zeros included); it need not be pre-zeroed.  It may be NULL (ABI >= 2): the
 *     alignment then comes back in compact form only (idx_out / dur_out), the dense plane -- half of the
 *     algorithmic bytes -- is neither zero-filled nor scattered.  mas_expand_path() rebuilds it on demand.
 *   - dur_out, idx_out, status_out may be NULL (not all three of path_out, dur_out, idx_out).
 *   - workspace: mas_maximum_path_workspace_bytes(B,T,S) bytes, 256-byte aligned.
 */
size_t mas_maximum_path_workspace_bytes(int B, int T, int S);
int mas_maximum_path_f32(const float *neg_cent, const int32_t *t_ys, const int32_t *t_xs,
                         void *path_out, int path_dtype,
                         int32_t *dur_out, int32_t *idx_out, int32_t *status_out,
                         void *workspace, size_t workspace_bytes,
                         int B, int T, int S, void *stream);

/*
 * The cost block of SynthesizerTrn.forward (models.py:1226-1239):
 *   neg_cent[b,t,s] = sum_d [ -0.5*log(2*pi) - logs_p - 0.5*(z_p - m_p)^2 * exp(-2*logs_p) ]
 * evaluated as the reference does (two contractions over D plus two column
 * biases).  If stats_out != NULL it receives {sum, sum of squares} of all
 * B*T*S cells as two float64 (device), for the noise std of models.py:1243.
 * S % 4 != 0: the workspace also holds a plane with 16-byte rows that the
 * contraction stores into; a second kernel packs it into neg_cent_out.
 */
size_t mas_neg_cent_workspace_bytes(int B, int D, int T, int S);
int mas_neg_cent_f32(const float *z_p, const float *m_p, const float *logs_p,
                     float *neg_cent_out, double *stats_out,
                     void *workspace, size_t workspace_bytes,
                     int B, int D, int T, int S, void *stream);

/*
 * The SynthesizerTrn alignment call as one unit (models.py:1224-1256):
 * cost -> optional VITS2 noise -> MAS -> path (+ durations, compact idx).
 *   - noise: NULL (mas_noise_scale is None) or the torch.randn_like draw
 *     [B,T,S] float32 supplied by the caller (models.py:1244); the library
 *     computes std over all B*T*S cells (unbiased, padding included, :1243)
 *     and adds (std * noise) * noise_scale (:1242-1247).
 *   - neg_cent_out: optional [B,T,S] float32 copy of the cost actually aligned (with noise: the noised
 *     cost; without this request the noised plane is never written -- the DP adds the noise on the fly).
 *     S % 4 != 0: computed into the padded plane of the workspace and packed by one more kernel.
 *   - path_out may be NULL (compact outputs only), as in mas_maximum_path_f32.
 *   - no neg_cent_out request, S <= 1024 without noise / S <= 256 with noise (any T, no multiple-of-4
 *     requirement): ONE kernel (after the prior
 *     preparation) runs contraction and DP -- concurrently without noise; with noise around a grid barrier (the
 *     statistics of models.py:1243 must exist before the first DP row), any batch size: B <= 74 with 16-byte rows
 *     of the draw runs {DP CTA, noise feeder CTA} pairs, larger batches noise-helper warps inside the DP CTAs.
 *     Cooperative launch: needs the whole GPU like any persistent kernel; where the context cannot hold the grid
 *     (MPS / MIG limits) the same work runs as separate launches.
 */
size_t mas_fused_align_workspace_bytes(int B, int D, int T, int S, int with_noise);
int mas_fused_align_f32(const float *z_p, const float *m_p, const float *logs_p,
                        const int32_t *t_ys, const int32_t *t_xs,
                        const float *noise, float noise_scale,
                        void *path_out, int path_dtype,
                        int32_t *dur_out, int32_t *idx_out, int32_t *status_out,
                        float *neg_cent_out,
                        void *workspace, size_t workspace_bytes,
                        int B, int D, int T, int S, void *stream);

/*
 * Compact <-> dense helpers for the multi-GPU path:
 * ranks all-gather idx [B,T] int32 and re-expand locally.
 */
int mas_expand_path(const int32_t *idx, void *path_out, int path_dtype,
                    int B, int T, int S, void *stream);

/*
 * The alignment's first consumers in SynthesizerTrn.forward, from the
 * compact outputs of the calls above instead of the dense path:
 *
 * mas_expand_prior_f32 -- replaces vits2/models.py:1270-1271
 *     m_p    = torch.matmul(attn.squeeze(1), m_p.transpose(1, 2)).transpose(1, 2)
 *     logs_p = torch.matmul(attn.squeeze(1), logs_p.transpose(1, 2)).transpose(1, 2)
 *   m_out[b,d,t] = m_p[b,d,idx[b,t]] (0 where idx[b,t] < 0, i.e. past t_y: the path row is all-zero there).
 *   m_p / logs_p [B,D,S], idx [B,T] int32, outputs [B,D,T]; logs_p and logs_out may both be NULL.
 *
 * mas_expand_prior_backward_f32 -- the gradient of that expansion with respect to m_p / logs_p
 *   (what autograd derives from the two matmuls): g_m_p[b,d,s] = sum of g_m[b,d,t] over the frames
 *   aligned to s, i.e. t in [start_s, start_s + dur[b,s]) with start = exclusive prefix sum of dur[b,:].
 *   Fixed summation order (ascending t; small batches sum a column that straddles a cut of the frame range from
 *   its pieces, the same way on every run), no atomics.  g_logs / g_logs_p may both be NULL.  Negative durations
 *   count as 0, frames past T are not read.  T % 4 == 0 with 16-byte aligned gradients takes the tensor-map kernel
 *   (csrc/mas_segsum.cu), anything else the column-per-thread kernel.
 *
 * mas_logw_f32 -- replaces models.py:1256 + 1261:  w = attn.sum(2);  logw_ = torch.log(w + 1e-6) * x_mask
 *   from the int32 durations [B,S] and the text lengths [B]; output [B,S] fp32.
 */
int mas_expand_prior_f32(const float *m_p, const float *logs_p, const int32_t *idx,
                         float *m_out, float *logs_out,
                         int B, int D, int T, int S, void *stream);
int mas_expand_prior_backward_f32(const float *g_m, const float *g_logs, const int32_t *dur,
                                  float *g_m_p, float *g_logs_p,
                                  int B, int D, int T, int S, void *stream);
int mas_logw_f32(const int32_t *dur, const int32_t *t_xs, float *logw_out,
                 int B, int S, void *stream);

/*
 * mas_idx_from_durations_f32 -- inference side, replaces commons.generate_path (vits2/commons.py:130-145, called at
 * models.py:1310) in compact form: idx_out[b,y] = the text column x with cum[x-1] <= y < cum[x], where cum is the
 * running sum of durations[b,:] (fp32 holding ceil()ed values, as models.py:1303 makes them), or -1 when
 * y >= t_ys[b] (t_ys may be NULL: T), x >= t_xs[b], or no column covers y.  mas_expand_path turns idx into the dense
 * path, mas_expand_prior_f32 does the expansion of models.py:1312-1317.
 */
int mas_idx_from_durations_f32(const float *durations, const int32_t *t_xs, const int32_t *t_ys,
                               int32_t *idx_out, int B, int T, int S, void *stream);

/* Tuning knobs (MAS_NO_FUSED, MAS_DP_VK, MAS_DP_WARPS, ... see csrc/mas_common.cuh) are read from the environment
 * once, on first use; this re-reads them (tests and benchmarks that switch variants inside one process). */
void mas_reload_config(void);

/* diagnostics (trace build of the library only, -DMAS_TRACE): with MAS_TRACE=1 in the environment the fused kernel records device timestamps
 * (ns, %globaltimer) of tile publications and DP milestones; this copies the first n_words of the
 * trace to the host (synchronises the device).  MAS_ERR_NULL_POINTER when tracing is off. */
int mas_debug_read_trace(unsigned long long *host_out, int n_words);

/* number of kernels this library launched on the calling thread since the
 * last call (bench.py's gpu_launches); resets the counter. */
long mas_take_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* MAS_B200_H_ */
