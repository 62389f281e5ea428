// Micro-benchmark: issue rate of packed fp32 (FFMA2/FADD2, sm_100) vs scalar FFMA/FADD.  nvcc -arch=sm_100a ffma2.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(float* out, int iters)
{
    float2 a0 = make_float2(threadIdx.x, 1.f), a1 = make_float2(2.f, 3.f), a2 = make_float2(4.f, 5.f), a3 = make_float2(6.f, 7.f);
    float2 a4 = a0, a5 = a1, a6 = a2, a7 = a3;
    const float2 m = make_float2(1.0001f, 0.9999f), c = make_float2(0.5f, 0.25f);
    for (int i = 0; i < iters; i++) {
        if (MODE == 0) {   // scalar: 16 FFMA
            a0.x = fmaf(a0.x, m.x, c.x); a0.y = fmaf(a0.y, m.y, c.y); a1.x = fmaf(a1.x, m.x, c.x); a1.y = fmaf(a1.y, m.y, c.y);
            a2.x = fmaf(a2.x, m.x, c.x); a2.y = fmaf(a2.y, m.y, c.y); a3.x = fmaf(a3.x, m.x, c.x); a3.y = fmaf(a3.y, m.y, c.y);
            a4.x = fmaf(a4.x, m.x, c.x); a4.y = fmaf(a4.y, m.y, c.y); a5.x = fmaf(a5.x, m.x, c.x); a5.y = fmaf(a5.y, m.y, c.y);
            a6.x = fmaf(a6.x, m.x, c.x); a6.y = fmaf(a6.y, m.y, c.y); a7.x = fmaf(a7.x, m.x, c.x); a7.y = fmaf(a7.y, m.y, c.y);
        } else {           // packed: 8 FFMA2 = same 16 FMAs
            a0 = __ffma2_rn(a0, m, c); a1 = __ffma2_rn(a1, m, c); a2 = __ffma2_rn(a2, m, c); a3 = __ffma2_rn(a3, m, c);
            a4 = __ffma2_rn(a4, m, c); a5 = __ffma2_rn(a5, m, c); a6 = __ffma2_rn(a6, m, c); a7 = __ffma2_rn(a7, m, c);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0.x + a0.y + a1.x + a1.y + a2.x + a2.y + a3.x + a3.y + a4.x + a4.y + a5.x + a5.y + a6.x + a6.y + a7.x + a7.y;
}
int main()
{
    float* d; cudaMalloc(&d, 148 * 8 * 1024 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000;
    for (int mode = 0; mode < 2; mode++) {
        for (int rep = 0; rep < 2; rep++) {
            cudaEventRecord(e0);
            if (mode == 0) k<0><<<148 * 8, 1024>>>(d, iters); else k<1><<<148 * 8, 1024>>>(d, iters);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            double fma = 148.0 * 8 * 1024 * iters * 16;
            if (rep) printf("%s: %.3f ms  %.1f TFMA/s  (%.1f TFLOP/s)\n", mode ? "FFMA2 " : "FFMA  ", ms, fma / ms / 1e9, 2 * fma / ms / 1e9);
        }
    }
    return 0;
}
