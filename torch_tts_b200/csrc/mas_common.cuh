// mas_common.cuh -- shared device helpers (sm_100a only) for the MAS hot path.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/mas_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "mas_b200 kernels are written for sm_100a only"
#endif

namespace mas {

constexpr float kNeg = -1e9f;  // core.pyx:7 max_neg_val (exactly representable)
constexpr unsigned kFullMask = 0xffffffffu;

// ---- diagnostics are a compile-time choice -----------------------------------
// The product library is built without MAS_TRACE: no timestamp stores, no debug-bit branches in the role
// loops.  `python -m torch_tts_b200.build --trace` builds libmas_b200_trace.so with -DMAS_TRACE for the
// timeline tools under tools/ (select it with MAS_LIB_PATH).
#ifdef MAS_TRACE
constexpr bool kTrace = true;
#else
constexpr bool kTrace = false;
#endif
// debug bit `bit` of a role's parameter block (always false in the product build)
#define MAS_DBG(p, bit) (::mas::kTrace && (((p).debug & (bit)) != 0))
// the role's trace buffer, or nullptr (always nullptr in the product build)
#define MAS_TR(p) (::mas::kTrace ? (p).trace : nullptr)

// ---- host-side bookkeeping (defined in mas_api.cu) -------------------------
void note_launch(int n = 1);
// trace build with MAS_TRACE=1 in the environment: device buffer of kTraceWords timestamps (per device), else nullptr
unsigned long long *trace_buffer();
constexpr int kTraceWords = 1 << 16;
int note_cuda_error(cudaError_t e, const char *what);

// Tuning knobs, read from the environment ONCE (first use) and again only on mas_reload_config():
// no getenv on the launch path.
struct Config {
    int no_fused;        // MAS_NO_FUSED=1: contraction and DP as separate launches
    int dp_vk;           // MAS_DP_VK=0: single-role DP warps instead of the value / origin split
    int dp_warps;        // MAS_DP_WARPS=4: four value warps with two columns per thread (128 < S <= 256)
    int dp_stages;       // MAS_DP_STAGES=n: force the cost-tile ring depth
    int fused_dp_ctas;   // MAS_FUSED_DP_CTAS=n: DP CTAs of the fused kernel (0 = cost model)
    int fused_rounds;    // MAS_FUSED_ROUNDS=k: contraction rounds before the DP CTAs leave (-1 = cost model)
    int fused_zero_offload;  // MAS_FUSED_ZERO_OFFLOAD=0: the DP CTAs zero-fill their own path planes
    int fused_pdl;       // MAS_FUSED_PDL=0: no programmatic dependent launch
    int noise_fused;     // MAS_NOISE_FUSED=0: noise-scaled alignment as separate launches
    int noise_feed;      // MAS_NOISE_FEED=0: no {DP CTA, noise feeder CTA} pairs (helper warps inside the DP CTAs instead)
    int segsum;          // MAS_SEGSUM=0: prior-expansion backward with the column-per-thread kernel (mas_expand.cu)
    int seg_stages;      // MAS_SEG_STAGES=n: tiles in that kernel's ring (2..8; default 3)
    int seg_parts;       // MAS_SEG_PARTS=n: runs of frames per CTA in that kernel (1, 2, 4; default: by batch size)
    int seg_nw;          // MAS_SEG_NW=n: channel groups (of 32) per CTA in that kernel
    int stage;           // MAS_STAGE: 1 = prior preparation only, 2 = skip it (reuse the images in the workspace);
                         // bench.py times the prior kernel alone with it
    int tc_no_tma;       // MAS_TC_NO_TMA: bit 1 plain z loads, bit 2 per-cell output stores -- the contraction's fallbacks for
                         // planes no tensor map describes; a host-side choice, forced by the parity tests
    int tc_debug, dp_debug, tc_grid, tc_pair, trace;   // trace build only (MAS_TC_DEBUG, MAS_DP_DEBUG, ...)
};
const Config &config();

#define MAS_CUDA_TRY(expr)                                        \
    do {                                                          \
        cudaError_t _e = (expr);                                  \
        if (_e != cudaSuccess) return ::mas::note_cuda_error(_e, #expr); \
    } while (0)

// ---- PTX wrappers -----------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}

// an mbarrier must be invalidated before its shared memory is used for anything else (PTX ISA, mbarrier.inval)
__device__ __forceinline__ void mbar_inval(uint64_t *bar)
{
    asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void fence_mbar_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void fence_proxy_async()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}

__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}

__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    while (!mbar_try_wait(bar, parity)) {
    }
}

// 1-D bulk copy global -> shared through the TMA engine (SASS: UBLKCP).
// dst/src 16-byte aligned, bytes a multiple of 16.
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar)
{
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(dst_smem)),
        "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

// ---- distributed shared memory (CTA pair of a cluster) ----------------------
// shared::cluster address of `local` (a shared::cta address) in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t dsm_map(uint32_t local, uint32_t rank)
{
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local), "r"(rank));
    return r;
}
__device__ __forceinline__ void dsm_st_v4(uint32_t cluster_addr, float4 v)
{
    asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(cluster_addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                 : "memory");
}
// asynchronous 16-byte store into another CTA's shared memory; its arrival is counted (16 bytes) on the mbarrier
// `cluster_mbar` of that CTA, so the reader needs nothing but its usual wait on the barrier and the writer no fence
__device__ __forceinline__ void dsm_st_async_v4(uint32_t cluster_addr, float4 v, uint32_t cluster_mbar)
{
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(
                     cluster_addr),
                 "r"(__float_as_uint(v.x)), "r"(__float_as_uint(v.y)), "r"(__float_as_uint(v.z)), "r"(__float_as_uint(v.w)),
                 "r"(cluster_mbar)
                 : "memory");
}
__device__ __forceinline__ void dsm_st_u32(uint32_t cluster_addr, uint32_t v)
{
    asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(cluster_addr), "r"(v) : "memory");
}
// arrive on an mbarrier of another CTA of the cluster; orders this thread's earlier (remote) stores before it
__device__ __forceinline__ void dsm_mbar_arrive_release(uint32_t cluster_addr)
{
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// wait on a local mbarrier that threads of another CTA of the cluster arrive on (acquire at cluster scope)
__device__ __forceinline__ void mbar_wait_acq_cluster(uint64_t *bar, uint32_t parity)
{
    uint32_t ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!ok);
}
__device__ __forceinline__ uint32_t cluster_rank()
{
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}

// prefetch the 128-byte line holding `p` into L2
__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// 1-D bulk copy shared -> global (SASS: UBLKCP); completion tracked by bulk groups.
__device__ __forceinline__ void bulk_s2g(void *dst_gmem, const void *src_smem, uint32_t bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(smem_u32(src_smem)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }


// ---- TMA tensor-map copies (SASS: UTMALDG / UTMASTG) --------------------------
// 3-D tile load global -> shared, completion on an mbarrier (coordinates innermost first).
__device__ __forceinline__ void tma_load_3d(void *dst_smem, const void *tmap, int c0, int c1, int c2, uint64_t *bar)
{
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
        ::"r"(smem_u32(dst_smem)), "l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
        : "memory");
}
// 3-D tile store shared -> global (bulk async-group; out-of-range elements are clipped).
__device__ __forceinline__ void tma_store_3d(const void *tmap, int c0, int c1, int c2, const void *src_smem)
{
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%1, %2, %3}], [%4];"
                 ::"l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(src_smem))
                 : "memory");
}
// pull a 3-D tile into L2 only (no shared-memory destination, no completion)
__device__ __forceinline__ void tma_prefetch_l2_3d(const void *tmap, int c0, int c1, int c2)
{
    asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];" ::"l"(tmap), "r"(c0), "r"(c1),
                 "r"(c2)
                 : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const void *tmap)
{
    asm volatile("prefetch.tensormap [%0];" ::"l"(tmap) : "memory");
}
// wait until at most N bulk groups are still pending at all (their global writes have landed)
template <int N>
__device__ __forceinline__ void bulk_wait_group() { asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory"); }
// wait until at most N bulk groups still READ their shared-memory source
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }

// ---- inter-CTA flags (fused kernel: cost tiles -> DP) ---------------------------
__device__ __forceinline__ void st_release_gpu(uint32_t *p, uint32_t v)
{
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_gpu(const uint32_t *p)
{
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

__device__ __forceinline__ unsigned long long globaltimer_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// named barrier over `count` threads (count a multiple of 32); id 0 is __syncthreads()
__device__ __forceinline__ void bar_sync(int id, int count)
{
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}

__device__ __forceinline__ void st_global_v4_zero(void *p)
{
    asm volatile("st.global.v4.u32 [%0], {%1, %1, %1, %1};" ::"l"(p), "r"(0u) : "memory");
}

__host__ __device__ __forceinline__ size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

__host__ __device__ __forceinline__ int path_elem_size(int path_dtype)
{
    return (path_dtype == MAS_PATH_F16 || path_dtype == MAS_PATH_BF16) ? 2 : 4;
}

// bit pattern of "1" in the path dtype
__host__ __device__ __forceinline__ uint32_t path_one_bits(int path_dtype)
{
    switch (path_dtype) {
        case MAS_PATH_F32: return 0x3f800000u;
        case MAS_PATH_F16: return 0x3c00u;
        case MAS_PATH_BF16: return 0x3f80u;
        default: return 1u;
    }
}

}  // namespace mas
