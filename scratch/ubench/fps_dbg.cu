// Phase-level cycle breakdown of the bucket-pruned FPS kernel (csrc/fps_pruned.cu compiled with -DFP_DEBUG).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -DFP_DEBUG -o scratch/ubench/fps_dbg scratch/ubench/fps_dbg.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
namespace mpc { unsigned long long* g_fp_dbg = nullptr; }
#include "../../markov-process-analysis-on-point-cloud_b200/csrc/fps_pruned.cu"
namespace mpc { int64_t g_knob[8] = {0}; }

int main(int argc, char** argv) {
    const int B = argc > 1 ? atoi(argv[1]) : 8, N = argc > 2 ? atoi(argv[2]) : 24000, np = argc > 3 ? atoi(argv[3]) : N / 2;
    std::vector<float> h((size_t)B * N * 3);
    srand(1);
    for (auto& v : h) v = 2.f * rand() / RAND_MAX - 1.f;
    float* d; int64_t *st, *out; unsigned long long* dbg;
    cudaMalloc(&d, h.size() * 4); cudaMalloc(&st, B * 8); cudaMalloc(&out, (size_t)B * np * 8); cudaMalloc(&dbg, 64 * 8 + 16 * 32 * 8 * 8);
    cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    cudaMemset(st, 0, B * 8); cudaMemset(dbg, 0, 64);
    mpc::g_fp_dbg = dbg;
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    for (int rep = 0; rep < 2; ++rep) {
        cudaMemset(dbg, 0, 64);
        cudaEventRecord(a);
        const int cfg = argc > 4 ? atoi(argv[4]) : 0;
        int rc;
        switch (cfg) {
            case 1: rc = mpc::launch_pruned<1024, 12, 1>(d, st, out, B, N, np, 0); break;
            case 2: rc = mpc::launch_pruned<1024, 12, 2>(d, st, out, B, N, np, 0); break;
            case 3: rc = mpc::launch_pruned<256, 24, 2>(d, st, out, B, N, np, 0); break;
            case 4: rc = mpc::launch_pruned<128, 24, 1>(d, st, out, B, N, np, 0); break;
            case 5: rc = mpc::launch_pruned<128, 24, 2>(d, st, out, B, N, np, 0); break;
            case 6: rc = mpc::launch_pruned<256, 24, 1>(d, st, out, B, N, np, 0); break;
            case 7: rc = mpc::launch_pruned<512, 12, 2>(d, st, out, B, N, np, 0); break;
            case 8: rc = mpc::launch_pruned<512, 12, 1>(d, st, out, B, N, np, 0); break;
            default: rc = mpc::fps_pruned_dispatch(d, st, out, B, N, np, 0);
        }
        cudaEventRecord(b);
        cudaError_t e = cudaDeviceSynchronize();
        if (rc || e) { printf("rc %d err %s\n", rc, cudaGetErrorString(e)); return 1; }
    }
    float ms; cudaEventElapsedTime(&ms, a, b);
    unsigned long long r[8]; cudaMemcpy(r, dbg, 64, cudaMemcpyDeviceToHost);
    if (argc > 5) {
        const int NW = atoi(argv[5]);
        std::vector<unsigned long long> tr(16 * NW * 8);
        cudaMemcpy(tr.data(), dbg + 64, tr.size() * 8, cudaMemcpyDeviceToHost);
        for (int r = 0; r < 16; ++r) {
            unsigned long long base = ~0ull;
            for (int w = 0; w < NW; ++w) if (tr[(r * NW + w) * 8 + 1] && tr[(r * NW + w) * 8 + 1] < base) base = tr[(r * NW + w) * 8 + 1];
            printf("round %d (test | update | publish+refresh | wait | reduce end, cycles after the round's first test end):\n", 1000 + r);
            for (int w = 0; w < NW; ++w) {
                printf("  w%02d", w);
                for (int i = 1; i <= 5; ++i) printf(" %6lld", (long long)(tr[(r * NW + w) * 8 + i] - base));
                printf("\n");
            }
        }
    }
    printf("B %d N %d np %d: %.3f ms = %.3f us/round; touched buckets/round/cloud %.2f of %d; warp0 cycles/round: test %.0f update %.0f publish %.0f wait %.0f reduce %.0f\n",
           B, N, np, ms, 1e3 * ms / np, (double)r[0] / np / B, (N + 31) / 32, (double)r[1] / np, (double)r[2] / np,
           (double)r[3] / np, (double)r[4] / np, (double)r[5] / np);
    return 0;
}
