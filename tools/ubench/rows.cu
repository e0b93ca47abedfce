// micro-benchmark: the value warp's DP row (4 cells: FMNMX + FADD + decision bit, one shuffle) in isolation, then
// with the pieces of the real loop added one at a time -- which one takes the row from ~28 to ~57 cycles?
//   nvcc -arch=sm_100a -O3 -o /tmp/rows tools/ubench/rows.cu && /tmp/rows
#include <cstdio>
#include <cuda_runtime.h>
constexpr int C = 4, S = 256;
template <bool kSmemCost, bool kLane0, bool kFin, bool kStore, bool kTrips>
__device__ __forceinline__ long long run(const float *tile, float *ring, float (&v)[C], unsigned (&wl)[C], float &fin, int n)
{
    const int lane = threadIdx.x & 31;
    const bool lane0 = lane == 0, lane31 = lane == 31;
    const float *trow = tile + lane * C;
    float creg[C] = {-0.5f, -0.25f, -0.75f, -0.125f};
    const long long t0 = clock64();
    for (int it = 0; it < n; ++it) {        // one "chunk" of 32 rows
#pragma unroll 1
        for (int r = 0; r < 32; r += 8) {
            unsigned w8[C] = {0, 0, 0, 0};
#pragma unroll
            for (int blk = 0; blk < 2; ++blk) {
                float cost[4][C], lv[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    if (kSmemCost) {
                        const float4 t = *reinterpret_cast<const float4 *>(trow + (size_t)(r + blk * 4 + i) * S);
                        cost[i][0] = t.x, cost[i][1] = t.y, cost[i][2] = t.z, cost[i][3] = t.w;
                    } else {
#pragma unroll
                        for (int k = 0; k < C; ++k) cost[i][k] = creg[k];
                    }
                }
                if (kLane0) {
                    const float4 t = *reinterpret_cast<const float4 *>(ring + r + blk * 4);
                    lv[0] = t.x, lv[1] = t.y, lv[2] = t.z, lv[3] = t.w;
                }
                float ov[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    float up = __shfl_up_sync(0xffffffffu, v[C - 1], 1);
                    if (kLane0 && lane0) up = lv[i];
                    if (kFin) {
                        fin = fmaf(cost[i][0], cost[i][1], fin);
                        fin = fmaf(cost[i][2], cost[i][3], fin);
                    }
#pragma unroll
                    for (int k = C - 1; k >= 0; --k) {
                        const float p = k ? v[k - 1] : up, cur = v[k];
                        if (cur < p) w8[k] |= 1u << (blk * 4 + i);
                        v[k] = cost[i][k] + fmaxf(p, cur);
                    }
                    ov[i] = v[C - 1];
                }
                if (kStore && lane31) *reinterpret_cast<float4 *>(ring + 64 + r + blk * 4) = make_float4(ov[0], ov[1], ov[2], ov[3]);
            }
            if (kTrips) {
#pragma unroll
                for (int k = 0; k < C; ++k) wl[k] |= w8[k] << r;
            } else {
#pragma unroll
                for (int k = 0; k < C; ++k) wl[k] ^= w8[k];
            }
        }
    }
    return clock64() - t0;
}
__global__ void k(float *out, const float *in, long long *cyc, int n)
{
    __shared__ __align__(16) float tile[40 * S];
    __shared__ __align__(16) float ring[128];
    for (int i = threadIdx.x; i < 40 * S; i += blockDim.x) tile[i] = in[i % 12] - 3.0f;
    for (int i = threadIdx.x; i < 128; i += blockDim.x) ring[i] = -1e9f;
    __syncthreads();
    float v[C], fin = 0;
    unsigned wl[C] = {0, 0, 0, 0};
    for (int i = 0; i < C; ++i) v[i] = in[i] + threadIdx.x;
    long long t[6];
    t[0] = run<false, false, false, false, false>(tile, ring, v, wl, fin, n);
    t[1] = run<true, false, false, false, false>(tile, ring, v, wl, fin, n);
    t[2] = run<true, true, false, false, false>(tile, ring, v, wl, fin, n);
    t[3] = run<true, true, true, false, false>(tile, ring, v, wl, fin, n);
    t[4] = run<true, true, true, true, false>(tile, ring, v, wl, fin, n);
    t[5] = run<true, true, true, true, true>(tile, ring, v, wl, fin, n);
    float s = fin;
    for (int i = 0; i < C; ++i) s += v[i] + (float)wl[i];
    out[threadIdx.x] = s;
    if (threadIdx.x == 0)
        for (int i = 0; i < 6; ++i) cyc[i] = t[i];
}
int main()
{
    float *o, *in; long long *c, h[6];
    cudaMalloc(&o, 4096); cudaMalloc(&in, 64); cudaMalloc(&c, 48);
    float hin[12] = {1, 2, 3, 4, 5, 6, 7, 8, -0.5f, -0.25f, -0.75f, -0.125f};
    cudaMemcpy(in, hin, 48, cudaMemcpyHostToDevice);
    const int n = 512;
    k<<<1, 32>>>(o, in, c, n); k<<<1, 32>>>(o, in, c, n);
    cudaMemcpy(h, c, 48, cudaMemcpyDeviceToHost);
    const char *name[6] = {"registers only", "+ cost rows from shared memory", "+ lane 0 takes the left warp's value",
                           "+ non-finite tracking", "+ lane 31 stores the boundary", "+ decision word merge per trip"};
    for (int i = 0; i < 6; ++i) printf("%-40s %.1f cycles per row\n", name[i], h[i] / (32.0 * n));
    return 0;
}
