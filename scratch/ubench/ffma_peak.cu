// Peak issue rate of 3-register FFMA vs packed FFMA2 (fma.rn.f32x2) on sm_100a: 32 independent accumulators per thread,
// operands from registers, `warps` warps per SM.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 ffma_peak.cu -o ffma_peak
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned long long ffma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long d;
    asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
template <int MODE>
__global__ void k(float* out, int iters, float a0, float b0) {
    float acc[32];
    unsigned long long acc2[16];
    for (int i = 0; i < 32; ++i) acc[i] = i;
    for (int i = 0; i < 16; ++i) acc2[i] = i;
    float a[4] = {a0, a0 + 1, a0 + 2, a0 + 3}, b[8];
    for (int i = 0; i < 8; ++i) b[i] = b0 + i + threadIdx.x;
    unsigned long long b2[4], a2[4];
    for (int i = 0; i < 4; ++i) {
        asm("mov.b64 %0, {%1, %2};" : "=l"(b2[i]) : "f"(b[2 * i]), "f"(b[2 * i + 1]));
        asm("mov.b64 %0, {%1, %2};" : "=l"(a2[i]) : "f"(a[i]), "f"(a[i]));
    }
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0) {
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i * 8 + j] = __fmaf_rn(a[i], b[j], acc[i * 8 + j]);
        } else {
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc2[i * 4 + j] = ffma2(a2[i], b2[j], acc2[i * 4 + j]);
        }
    }
    float s = 0;
    for (int i = 0; i < 32; ++i) s += acc[i];
    for (int i = 0; i < 16; ++i) s += (float)(acc2[i] & 0xffff);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
    float* out;
    cudaMalloc(&out, 148 * 1024 * 4 * sizeof(float));
    const int iters = 20000;
    for (int warps : {4, 8, 16, 32}) {
        for (int mode = 0; mode < 2; ++mode) {
            cudaEvent_t e0, e1;
            cudaEventCreate(&e0); cudaEventCreate(&e1);
            dim3 grid(148), block(warps * 32);
            for (int rep = 0; rep < 2; ++rep) {
                cudaEventRecord(e0);
                if (mode == 0) k<0><<<grid, block>>>(out, iters, 1.0f, 2.0f); else k<1><<<grid, block>>>(out, iters, 1.0f, 2.0f);
                cudaEventRecord(e1);
                cudaEventSynchronize(e1);
            }
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            double fma = (double)148 * warps * 32 * iters * 32;
            printf("warps/SM %2d  %s  %8.3f ms  %7.2f TFLOP/s\n", warps, mode ? "FFMA2" : "FFMA ", ms, 2 * fma / ms / 1e9);
        }
    }
    return 0;
}
