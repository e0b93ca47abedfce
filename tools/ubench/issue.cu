// micro-benchmark: what ONE warp per scheduler can issue on sm_100a (the DP warps run alone on their
// scheduler): independent FADD / FMNMX streams, and DP-like rows with and without the decision bits and the
// shuffle -- run under gpurun:  nvcc -arch=sm_100a -O3 -o /tmp/issue tools/ubench/issue.cu && /tmp/issue
#include <cstdio>
#include <cuda_runtime.h>
#define ROWS 8
__global__ void k(float *out, const float *in, long long *cyc, int n)
{
    float a[8], c[4];
    for (int i = 0; i < 8; ++i) a[i] = in[i] + threadIdx.x;
    for (int i = 0; i < 4; ++i) c[i] = in[8 + i];
    long long t[8];
    t[0] = clock64();
    for (int it = 0; it < n; ++it)                      // A: 8 independent FADD chains
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] += c[i & 3];
    t[1] = clock64();
    for (int it = 0; it < n; ++it)                      // B: 8 independent FMNMX chains
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = fmaxf(a[i], c[i & 3] + (float)0);
    t[2] = clock64();
    for (int it = 0; it < n; ++it)                      // C: 4 FMNMX + 4 FADD, independent
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            a[i] = fmaxf(a[i], a[i + 4]);
            a[i + 4] += c[i];
        }
    t[3] = clock64();
    float v[4] = {a[0], a[1], a[2], a[3]};
    for (int it = 0; it < n; ++it)                      // D: DP rows, no bits, no shuffle
#pragma unroll
        for (int r = 0; r < ROWS; ++r) {
            const float up = v[3];
#pragma unroll
            for (int kk = 3; kk >= 0; --kk) v[kk] = c[kk] + fmaxf(kk ? v[kk - 1] : up, v[kk]);
        }
    t[4] = clock64();
    unsigned w[4] = {0, 0, 0, 0};
    for (int it = 0; it < n; ++it)                      // E: D + decision bits
#pragma unroll
        for (int r = 0; r < ROWS; ++r) {
            const float up = v[3];
#pragma unroll
            for (int kk = 3; kk >= 0; --kk) {
                const float p = kk ? v[kk - 1] : up;
                if (v[kk] < p) w[kk] |= 1u << r;
                v[kk] = c[kk] + fmaxf(p, v[kk]);
            }
        }
    t[5] = clock64();
    for (int it = 0; it < n; ++it)                      // F: D + shuffle
#pragma unroll
        for (int r = 0; r < ROWS; ++r) {
            const float up = __shfl_up_sync(0xffffffffu, v[3], 1);
#pragma unroll
            for (int kk = 3; kk >= 0; --kk) v[kk] = c[kk] + fmaxf(kk ? v[kk - 1] : up, v[kk]);
        }
    t[6] = clock64();
    for (int it = 0; it < n; ++it)                      // G: D + bits + shuffle
#pragma unroll
        for (int r = 0; r < ROWS; ++r) {
            const float up = __shfl_up_sync(0xffffffffu, v[3], 1);
#pragma unroll
            for (int kk = 3; kk >= 0; --kk) {
                const float p = kk ? v[kk - 1] : up;
                if (v[kk] < p) w[kk] |= 1u << r;
                v[kk] = c[kk] + fmaxf(p, v[kk]);
            }
        }
    t[7] = clock64();
    float s = 0;
    for (int i = 0; i < 8; ++i) s += a[i];
    for (int i = 0; i < 4; ++i) s += v[i] + (float)w[i];
    out[threadIdx.x + blockIdx.x * blockDim.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0)
        for (int i = 0; i < 7; ++i) cyc[i] = t[i + 1] - t[i];
}
int main()
{
    float *o, *in; long long *c, h[7];
    cudaMalloc(&o, 4096); cudaMalloc(&in, 64); cudaMalloc(&c, 56);
    float hin[12] = {1, 2, 3, 4, 5, 6, 7, 8, -0.5f, -0.25f, -0.75f, -0.125f};
    cudaMemcpy(in, hin, 48, cudaMemcpyHostToDevice);
    const int n = 2048;
    for (int threads = 32; threads <= 256; threads *= 2) {   // 1, 2, 4, 8 warps: 1 or 2 per scheduler
        k<<<1, threads>>>(o, in, c, n); k<<<1, threads>>>(o, in, c, n);
        cudaMemcpy(h, c, 56, cudaMemcpyDeviceToHost);
        printf("%d warps | per instr: 8xFADD %.2f  8xFMNMX %.2f  4+4 mixed %.2f | per DP row (4 cells): plain %.1f  +bits %.1f  +shfl %.1f  +bits+shfl %.1f cycles\n",
               threads / 32, h[0] / (8.0 * n), h[1] / (8.0 * n), h[2] / (8.0 * n), h[3] / (double)(ROWS * n), h[4] / (double)(ROWS * n),
               h[5] / (double)(ROWS * n), h[6] / (double)(ROWS * n));
    }
    return 0;
}
