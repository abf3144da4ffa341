// Dependent-chain latencies of the warp collectives / barriers the FPS kernels are made of (cycles per op).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o scratch/ubench/warp_lat scratch/ubench/warp_lat.cu
#include <cstdio>
#include <cstdint>
#define REP 256
__global__ void k(unsigned* out, long long* cyc, int nwarps_active) {
    __shared__ unsigned long long bar;
    __shared__ unsigned sm[1024];
    unsigned v = threadIdx.x * 2654435761u;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    long long t0, t1;
    int n = 0;
    sm[threadIdx.x] = v;
    if (threadIdx.x == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(&bar)), "r"(blockDim.x / 32));
    __syncthreads();
#define MEASURE(body) \
    __syncthreads(); t0 = clock64(); \
    _Pragma("unroll 1") for (int i = 0; i < REP; ++i) { body; } \
    t1 = clock64(); if (threadIdx.x == 0) cyc[n] = (t1 - t0); n++;
    MEASURE(v = __reduce_max_sync(0xffffffffu, v) + lane)                                    // 0 redux
    MEASURE(v = __ballot_sync(0xffffffffu, (v & 1) != 0) + lane)                              // 1 vote
    MEASURE(v = __shfl_sync(0xffffffffu, v, (v + 1) & 31))                                    // 2 shfl
    MEASURE(v = __ffs(v | 0x80000000u) + v)                                                  // 3 ffs
    MEASURE(v = sm[(v + lane) & 1023])                                                       // 4 lds
    MEASURE(v = v * 3 + 1)                                                                   // 5 imad
    MEASURE(float f = __uint_as_float(v & 0x3fffffff); f = fmaxf(fmaxf(f - 1.f, 1.f - f), 0.f); v = __float_as_uint(f) + 1)  // 6 fadd+fmnmx3
    MEASURE(sm[threadIdx.x] = v; __syncthreads(); v += sm[(threadIdx.x + 32) & (blockDim.x - 1)])                               // 7 sts+bar+lds
    {
        unsigned b = (unsigned)__cvta_generic_to_shared(&bar);
        MEASURE(
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(b) : "memory");
            { unsigned done = 0; while (!done) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(b), "r"((unsigned)(i & 1)) : "memory"); }
            v += 1)                                                                          // 8 mbarrier round (all warps)
    }
    MEASURE(unsigned hi = __reduce_max_sync(0xffffffffu, v); unsigned m = __ballot_sync(0xffffffffu, v == hi); int src = __ffs(m) - 1; v = __shfl_sync(0xffffffffu, v ^ 0x5555u, src) + lane)  // 9 argmax combo
    out[threadIdx.x] = v;
}
int main() {
    unsigned* out; long long* cyc;
    cudaMalloc(&out, 4096); cudaMalloc(&cyc, 256);
    const char* names[] = {"redux.max", "ballot", "shfl.idx", "ffs", "lds", "imad", "fadd+fmnmx3", "sts+bar.sync+lds", "mbarrier arrive+try_wait", "argmax (redux+ballot+ffs+shfl)"};
    for (int threads : {32, 256, 512, 1024}) {
        k<<<1, threads>>>(out, cyc, 0);
        cudaDeviceSynchronize();
        long long h[16]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
        printf("threads %d:", threads);
        for (int i = 0; i < 10; ++i) printf(" %s %.0f;", names[i], (double)h[i] / REP);
        printf("\n");
    }
    return 0;
}
