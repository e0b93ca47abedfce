// micro-benchmark: dependent-chain latencies on sm_100a (SHFL.UP, FMNMX+FADD, FSEL) -- run under gpurun
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(float *out, long long *cyc, int n)
{
    float v = threadIdx.x * 0.5f, w = 1.0f + threadIdx.x;
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) v = __shfl_up_sync(0xffffffffu, v, 1) + 1.0f;   // SHFL + FADD chain
    long long t1 = clock64();
    for (int i = 0; i < n; ++i) w = fmaxf(w, v) + 1.0f;                          // FMNMX + FADD chain
    long long t2 = clock64();
    float u = v;
    for (int i = 0; i < n; ++i) {                                                // SHFL + FSEL + FMNMX + FADD chain
        float s = __shfl_up_sync(0xffffffffu, u, 1);
        s = (threadIdx.x == 0) ? w : s;
        u = fmaxf(s, u) + 1.0f;
    }
    long long t3 = clock64();
    out[threadIdx.x] = v + w + u;
    if (threadIdx.x == 0) { cyc[0] = t1 - t0; cyc[1] = t2 - t1; cyc[2] = t3 - t2; }
}
int main()
{
    float *o; long long *c, h[3];
    cudaMalloc(&o, 128); cudaMalloc(&c, 24);
    const int n = 4096;
    k<<<1, 32>>>(o, c, n); k<<<1, 32>>>(o, c, n);
    cudaMemcpy(h, c, 24, cudaMemcpyDeviceToHost);
    printf("SHFL+FADD %.1f cyc/iter, FMNMX+FADD %.1f, SHFL+FSEL+FMNMX+FADD %.1f\n", h[0] / (double)n, h[1] / (double)n, h[2] / (double)n);
    return 0;
}
