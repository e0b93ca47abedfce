// micro-benchmark: ways to record the MAS decision bit of a cell next to FMNMX + FADD, one warp per scheduler,
// sm_100a.  The alu pipe (FMNMX, FSETP, LOP3, SHF, SEL) issues one warp instruction per 2 cycles, the fma pipe
// (FADD, FFMA, IMAD) one per cycle, so the variants move the bit bookkeeping between them.
//   nvcc -arch=sm_100a -O3 -o /tmp/bits tools/ubench/bits.cu && /tmp/bits
#include <cstdio>
#include <cuda_runtime.h>
#define ROWS 8
template <int VAR>
__device__ __forceinline__ void cell(float &v, float p, float c, unsigned &w, float &fin, float &facc, int r)
{
    const float cur = v;
    v = c + fmaxf(p, cur);
    if (VAR == 0) {                      // compare + predicated or
        if (cur < p) w |= 1u << r;
    } else if (VAR == 1) {               // sign through a funnel shift
        const float d = cur - p;
        fin = fmaf(d, 0.0f, fin);
        w = __funnelshift_l(__float_as_uint(d), w, 1);
    } else if (VAR == 2) {               // sign through mul.hi + mad (fma pipe)
        const float d = cur - p;
        fin = fmaf(d, 0.0f, fin);
        unsigned t;
        asm("mul.hi.u32 %0, %1, 2;" : "=r"(t) : "r"(__float_as_uint(d)));
        asm("mad.lo.u32 %0, %0, 2, %1;" : "+r"(w) : "r"(t));
    } else if (VAR == 3) {               // sign as a float through two saturating multiplies, fp accumulator
        const float d = cur - p;
        fin = fmaf(d, 0.0f, fin);
        float s;
        asm("mul.sat.f32 %0, %1, 0fFE800000;" : "=f"(s) : "f"(d));   // -2^126
        asm("mul.sat.f32 %0, %0, 0f7E800000;" : "+f"(s));            //  2^126
        facc = fmaf(facc, 2.0f, s);
    } else if (VAR == 4) {               // sign via shift right + mad
        const float d = cur - p;
        fin = fmaf(d, 0.0f, fin);
        w = w * 2 + (__float_as_uint(d) >> 31);
    }
}
template <int VAR>
__device__ __forceinline__ long long run(float (&v)[4], const float (&c)[4], unsigned (&w)[4], float &fin, float (&fa)[4], int n)
{
    const long long t0 = clock64();
    for (int it = 0; it < n; ++it)
#pragma unroll
        for (int r = 0; r < ROWS; ++r) {
            const float up = __shfl_up_sync(0xffffffffu, v[3], 1);
#pragma unroll
            for (int kk = 3; kk >= 0; --kk) cell<VAR>(v[kk], kk ? v[kk - 1] : up, c[kk], w[kk], fin, fa[kk], r);
        }
    return clock64() - t0;
}
__global__ void k(float *out, const float *in, long long *cyc, int n)
{
    float v[4], c[4], fa[4] = {0, 0, 0, 0}, fin = 0;
    unsigned w[4] = {0, 0, 0, 0};
    for (int i = 0; i < 4; ++i) v[i] = in[i] + threadIdx.x, c[i] = in[8 + i];
    long long t[5];
    t[0] = run<0>(v, c, w, fin, fa, n);
    t[1] = run<1>(v, c, w, fin, fa, n);
    t[2] = run<2>(v, c, w, fin, fa, n);
    t[3] = run<3>(v, c, w, fin, fa, n);
    t[4] = run<4>(v, c, w, fin, fa, n);
    float s = fin;
    for (int i = 0; i < 4; ++i) s += v[i] + (float)w[i] + fa[i];
    out[threadIdx.x] = s;
    if (threadIdx.x == 0)
        for (int i = 0; i < 5; ++i) cyc[i] = t[i];
}
int main()
{
    float *o, *in; long long *c, h[5];
    cudaMalloc(&o, 4096); cudaMalloc(&in, 64); cudaMalloc(&c, 40);
    float hin[12] = {1, 2, 3, 4, 5, 6, 7, 8, -0.5f, -0.25f, -0.75f, -0.125f};
    cudaMemcpy(in, hin, 48, cudaMemcpyHostToDevice);
    const int n = 2048;
    k<<<1, 32>>>(o, in, c, n); k<<<1, 32>>>(o, in, c, n);
    cudaMemcpy(h, c, 40, cudaMemcpyDeviceToHost);
    const char *name[5] = {"FSETP + predicated or", "FADD + funnel shift", "FADD + mul.hi + mad", "FADD + 2 mul.sat + fma", "FADD + shr + mad"};
    for (int i = 0; i < 5; ++i) printf("%-26s %.1f cycles per row (4 cells, with shuffle)\n", name[i], h[i] / (double)(ROWS * n));
    return 0;
}
