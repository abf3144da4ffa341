// Shared helpers for the sm_100a point-set kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/mpc_b200.h"

#define MPC_API extern "C" __attribute__((visibility("default")))

#define MPC_LAUNCH_CHECK()                        \
    do {                                          \
        cudaError_t e__ = cudaGetLastError();     \
        if (e__ != cudaSuccess) return (int)e__;  \
    } while (0)

#define MPC_CUDA(call)                            \
    do {                                          \
        cudaError_t e__ = (call);                 \
        if (e__ != cudaSuccess) return (int)e__;  \
    } while (0)

namespace mpc {

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs

// debug / tuning knobs (mpc_debug_set_knob, defined in bnact.cu; 0 = built-in default)
extern int64_t g_knob[8];

__host__ __device__ inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// Clamp an API-level int64 index into [0, n): out-of-range indices are never dereferenced.
__device__ __forceinline__ int clamp_index(int64_t i, int n) {
    return i < 0 ? 0 : (i >= n ? n - 1 : (int)i);
}

// Programmatic dependent launch.  Every kernel of the short streaming family starts with pdl_prologue(): it lets the
// NEXT kernel in the stream be scheduled right away (its CTAs park in their own griddepcontrol.wait until this grid
// has completed and flushed), and waits for the PREVIOUS kernel before touching global memory.  Launched through
// pdl_launch() the ~2-3 us of launch/drain latency between two dependent kernels overlaps; launched normally both
// instructions are no-ops.  (Long-running kernels -- FPS, kNN -- do not trigger early: parked dependents would hold
// SM slots that other streams need.)
__device__ __forceinline__ void pdl_prologue() {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
}
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

template <typename... KArgs, typename... Args>
inline void pdl_launch(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    (void)cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);  // errors surface in MPC_LAUNCH_CHECK()
}

// 128-bit reduction into global memory (sm_90+: red.global.add.v4.f32).
__device__ __forceinline__ void red_add_f32x4(float* addr, float4 v) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(v.x), "f"(v.y), "f"(v.z),
                 "f"(v.w)
                 : "memory");
}
__device__ __forceinline__ void red_add_f32(float* addr, float v) {
    asm volatile("red.global.add.f32 [%0], %1;" ::"l"(addr), "f"(v) : "memory");
}

// Expanded-form squared distance of the reference (pointnet2_utils.py:204-208); see mpc_b200.h.
// -2*dot is exact, so (-2*dot)+qn may be contracted freely; written with explicit roundings anyway.
__device__ __forceinline__ float sqdist_from_dot(float dot, float qn, float rn) {
    return __fadd_rn(__fadd_rn(__fmul_rn(-2.0f, dot), qn), rn);
}

// Sequential, non-fused squared norm ((x0*x0 + x1*x1) + x2*x2) + ...
__device__ __forceinline__ float sqnorm3(float x, float y, float z) {
    return __fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z));
}

}  // namespace mpc
