// bf16-I/O shared-MLP block for inference on the 5th-generation tensor cores (tcgen05 kind::f16, bf16 operands, fp32
// accumulation in TMEM, TMA-fed): the reference's whole `Linear` block in eval() (R/modules/pointnet2_utils.py:413-425:
// nn.Linear -> BatchNorm1d with running statistics -> LeakyReLU(0.2)) [+ the residual LocalTrans / Fuse add to it,
// :515,574,640-709] as ONE kernel:
//     out[m, n] = act( acc[m, n] * scale[n] + shift[n] ) (+ residual[m, n]),   acc = x[m, :] . w[n, :]
// With running statistics BatchNorm is a per-channel affine map, so scale = gamma / sqrt(var + eps) and
// shift = beta + (bias - mean) * scale are applied in fp32 to the fp32 accumulator in the epilogue: normalise +
// activation never make their own pass over HBM, activations travel as bf16 (half the bytes of the fp32 path), and
// one MMA pass replaces the three of the 3xTF32 split (no splitter warps: the tensor core reads the TMA tiles
// directly).  Weights are rounded to bf16 once per layer by the host side; nothing else is rounded before the final
// store.
//
// Structure (persistent, one CTA per SM, 192 threads): warps 0-3 epilogue (thread = output row: tcgen05.ld of its TMEM
// lane, affine + activation + residual in registers, 16-byte stores), warp 4 TMA producer, warp 5 MMA issuer + TMEM
// owner.  Two TMEM accumulator stages: the epilogue of tile i overlaps the main loop of tile i + 1.
#include <cuda_bf16.h>

#include "tc_common.cuh"

namespace mpc {
namespace tcb {
using namespace tc;

constexpr int BM = 128;
constexpr int BK = 64;          // bf16 elements = one 128-byte swizzle row
constexpr int UK = 16;          // kind::f16: 32 bytes per instruction
constexpr int THREADS = 192;
constexpr int MAX_STAGES = 6;
constexpr int A_BYTES = BM * BK * 2;  // 16 KB

struct Params {
    int M, N, K, block_n, n_tiles, m_tiles, stages, k_chunks;
    const float* scale;  // [N] or null (= 1)
    const float* shift;  // [N] or null (= 0)
    float slope;         // LeakyReLU slope (1 = no activation)
    const __nv_bfloat16* residual;  // [M, ldr] or null
    int ldr;
    void* out;           // bf16 [M, ldo] or f32 [M, ldo]
    int ldo;
    int out_f32;
};

__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}

__global__ void __launch_bounds__(THREADS, 1)
linear_bf16_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, const Params p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ uint64_t bar_full[MAX_STAGES], bar_empty[MAX_STAGES], bar_tmem_full[2], bar_tmem_empty[2];
    __shared__ uint32_t tmem_base_slot;
    __shared__ float scale_s[256], shift_s[256];

    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const int b_bytes = p.block_n * BK * 2;
    const int stage_bytes = A_BYTES + b_bytes;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int total_tiles = p.m_tiles * p.n_tiles;

    if (threadIdx.x == 0) {
        for (int s = 0; s < p.stages; ++s) {
            mbar_init(&bar_full[s], 1);
            mbar_init(&bar_empty[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&bar_tmem_full[a], 1);
            mbar_init(&bar_tmem_empty[a], 128);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x == 32) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
    }
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (warp == 5) {  // TMEM: 2 accumulator stages x 256 fp32 columns
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)),
                     "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_slot;

    if (warp == 4) {
        // ===== TMA producer =====
        if (lane == 0) {
            int it = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                const int mt = tile / p.n_tiles, nt = tile - mt * p.n_tiles;
                for (int kc = 0; kc < p.k_chunks; ++kc, ++it) {
                    const int s = it % p.stages;
                    const uint32_t ph = (it / p.stages) & 1;
                    mbar_wait(&bar_empty[s], ph ^ 1);
                    uint8_t* st = smem + (size_t)s * stage_bytes;
                    mbar_arrive_expect_tx(&bar_full[s], A_BYTES + b_bytes);
                    tma_load_2d(&map_a, &bar_full[s], st, kc * BK, mt * BM);
                    tma_load_2d(&map_b, &bar_full[s], st + A_BYTES, kc * BK, nt * p.block_n);
                }
            }
        }
    } else if (warp == 5) {
        // ===== MMA issuer =====
        if (lane == 0) {
            // instruction descriptor: D = F32 (1 << 4), A = B = BF16 (1 << 7, 1 << 10), both K-major, N >> 3 at bit 17,
            // M >> 4 at bit 24
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.block_n >> 3) << 17) |
                                   ((uint32_t)(BM >> 4) << 24);
            int it = 0, tile_it = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++tile_it) {
                const int as = tile_it & 1;
                const uint32_t aph = (tile_it >> 1) & 1;
                mbar_wait(&bar_tmem_empty[as], aph ^ 1);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + (uint32_t)(as * 256);
                for (int kc = 0; kc < p.k_chunks; ++kc, ++it) {
                    const int s = it % p.stages;
                    const uint32_t ph = (it / p.stages) & 1;
                    mbar_wait(&bar_full[s], ph);
                    tc_fence_after();
                    const uint32_t st = smem_u32(smem + (size_t)s * stage_bytes);
                    const uint64_t a_desc = make_desc(st), b_desc = make_desc(st + A_BYTES);
#pragma unroll
                    for (int kk = 0; kk < BK / UK; ++kk)  // +32 B inside the swizzle row per K step
                        umma_bf16(tmem_d, a_desc + (uint64_t)(kk * 2), b_desc + (uint64_t)(kk * 2), idesc,
                                  (kc != 0 || kk != 0) ? 1u : 0u);
                    umma_commit(&bar_empty[s]);
                }
                umma_commit(&bar_tmem_full[as]);
            }
        }
    } else {
        // ===== epilogue: warp w owns TMEM lanes 32w..32w+31 = output rows 32w..32w+31 of the tile =====
        int tile_it = 0, cur_nt = -1;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++tile_it) {
            const int mt = tile / p.n_tiles, nt = tile - mt * p.n_tiles;
            const int as = tile_it & 1;
            const uint32_t aph = (tile_it >> 1) & 1;
            const int row = mt * BM + warp * 32 + lane;
            const int n0 = nt * p.block_n;
            if (nt != cur_nt) {  // (uniform over the four warps: they walk the same tile sequence)
                epi_barrier();
                for (int i = threadIdx.x; i < p.block_n; i += 128) {
                    const bool in = n0 + i < p.N;
                    scale_s[i] = (in && p.scale) ? __ldg(p.scale + n0 + i) : 1.0f;
                    shift_s[i] = (in && p.shift) ? __ldg(p.shift + n0 + i) : 0.0f;
                }
                epi_barrier();
                cur_nt = nt;
            }
            mbar_wait(&bar_tmem_full[as], aph);
            tc_fence_after();
            const uint32_t taddr = tmem_base + (uint32_t)(as * 256) + ((uint32_t)(warp * 32) << 16);
            for (int c = 0; c < p.block_n; c += 16) {
                uint32_t r[16];
                tmem_ld16(taddr + c, r);
                tmem_ld_wait();
                if (row < p.M) {
                    float v[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        float t = fmaf(__uint_as_float(r[i]), scale_s[c + i], shift_s[c + i]);
                        v[i] = t > 0.f ? t : t * p.slope;
                    }
                    const bool full = n0 + c + 16 <= p.N;
                    if (p.residual) {
                        const __nv_bfloat16* rr = p.residual + (size_t)row * p.ldr + n0 + c;
                        if (full && (p.ldr & 7) == 0) {
                            const uint4 q0 = __ldg(reinterpret_cast<const uint4*>(rr));
                            const uint4 q1 = __ldg(reinterpret_cast<const uint4*>(rr) + 1);
                            const uint32_t w8[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                v[2 * i] += __uint_as_float(w8[i] << 16);
                                v[2 * i + 1] += __uint_as_float(w8[i] & 0xffff0000u);
                            }
                        } else {
#pragma unroll
                            for (int i = 0; i < 16; ++i)
                                if (n0 + c + i < p.N) v[i] += __bfloat162float(rr[i]);
                        }
                    }
                    if (p.out_f32) {
                        float* o = static_cast<float*>(p.out) + (size_t)row * p.ldo + n0 + c;
                        if (full && (p.ldo & 3) == 0) {
#pragma unroll
                            for (int i = 0; i < 16; i += 4)
                                *reinterpret_cast<float4*>(o + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
                        } else {
#pragma unroll
                            for (int i = 0; i < 16; ++i)
                                if (n0 + c + i < p.N) o[i] = v[i];
                        }
                    } else {
                        __nv_bfloat16* o = static_cast<__nv_bfloat16*>(p.out) + (size_t)row * p.ldo + n0 + c;
                        if (full && (p.ldo & 7) == 0) {
                            uint32_t w8[8];
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                const __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
                                w8[i] = *reinterpret_cast<const uint32_t*>(&h);
                            }
                            reinterpret_cast<uint4*>(o)[0] = make_uint4(w8[0], w8[1], w8[2], w8[3]);
                            reinterpret_cast<uint4*>(o)[1] = make_uint4(w8[4], w8[5], w8[6], w8[7]);
                        } else {
#pragma unroll
                            for (int i = 0; i < 16; ++i)
                                if (n0 + c + i < p.N) o[i] = __float2bfloat16_rn(v[i]);
                        }
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(&bar_tmem_empty[as]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 5) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
    }
}

// 2D bf16 row-major [rows, cols] (row stride ld elements), box = 64 columns x box_rows rows, 128B swizzle.
static int make_map_bf16(CUtensorMap* map, const void* base, int64_t rows, int64_t cols, int64_t ld, int box_rows) {
    if (((uintptr_t)base & 15u) || (ld & 7)) return MPC_ERR_UNSUPPORTED;
    EncodeTiledFn fn = encode_fn();
    if (!fn) return MPC_ERR_UNSUPPORTED;
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
    cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? MPC_OK : MPC_ERR_INVALID;
}

__global__ void f32_to_bf16_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y, int64_t rows, int64_t cols,
                                   int64_t ldx, int64_t ldy) {
    pdl_prologue();
    const int64_t total = rows * ldy;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = t / ldy, c = t - r * ldy;
        y[t] = __float2bfloat16_rn(c < cols ? x[r * ldx + c] : 0.0f);  // pad columns (ldy > cols) are zero
    }
}

}  // namespace tcb
}  // namespace mpc

MPC_API int mpc_linear_bf16(const void* x, int64_t ldx, const void* w, int64_t ldw, const float* scale,
                            const float* shift, float slope, const void* residual, int64_t ldr, void* out, int64_t ldo,
                            int64_t out_is_f32, int64_t M, int64_t K, int64_t N, mpc_stream_t stream) {
    using namespace mpc;
    using namespace mpc::tcb;
    if (!x || !w || !out || M <= 0 || K <= 0 || N <= 0) return MPC_ERR_INVALID;
    if (K % BK || ldx < K || ldw < K || ldo < N || (residual && ldr < N)) return MPC_ERR_UNSUPPORTED;
    if (M > INT32_MAX || N > 65536 || K > 65536) return MPC_ERR_UNSUPPORTED;
    Params p;
    p.M = (int)M;
    p.N = (int)N;
    p.K = (int)K;
    int bn = (int)(N < 256 ? N : 256);
    bn = (bn + 15) & ~15;
    while (bn > 64 && ceil_div(M, BM) * ceil_div(N, bn) < kNumSMs) bn = (bn / 2 + 15) & ~15;
    p.block_n = bn;
    p.n_tiles = (int)ceil_div(N, bn);
    p.m_tiles = (int)ceil_div(M, BM);
    p.k_chunks = (int)(K / BK);
    const int stage_bytes = A_BYTES + bn * BK * 2;
    int stages = (200 * 1024) / stage_bytes;
    if (stages > MAX_STAGES) stages = MAX_STAGES;
    if (stages > p.k_chunks * 2 && stages > 2) stages = p.k_chunks * 2 > 2 ? p.k_chunks * 2 : 2;
    p.stages = stages;
    p.scale = scale;
    p.shift = shift;
    p.slope = slope;
    p.residual = static_cast<const __nv_bfloat16*>(residual);
    p.ldr = (int)ldr;
    p.out = out;
    p.ldo = (int)ldo;
    p.out_f32 = out_is_f32 ? 1 : 0;
    CUtensorMap map_a, map_b;
    int rc = make_map_bf16(&map_a, x, M, K, ldx, BM);
    if (rc) return rc;
    rc = make_map_bf16(&map_b, w, N, K, ldw, bn);
    if (rc) return rc;
    const size_t smem = (size_t)stages * stage_bytes + 1024;
    static bool optin[64] = {};
    int dev = 0;
    MPC_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64 || !optin[dev]) {
        MPC_CUDA(cudaFuncSetAttribute(linear_bf16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 216 * 1024));
        if (dev >= 0 && dev < 64) optin[dev] = true;
    }
    const int total_tiles = p.m_tiles * p.n_tiles;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(total_tiles < kNumSMs ? total_tiles : kNumSMs));
    cfg.blockDim = dim3(THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    MPC_CUDA(cudaLaunchKernelEx(&cfg, linear_bf16_kernel, map_a, map_b, p));
    return MPC_OK;
}

MPC_API int mpc_f32_to_bf16(const float* x, int64_t ldx, void* y, int64_t ldy, int64_t rows, int64_t cols,
                            mpc_stream_t stream) {
    using namespace mpc;
    if (rows < 0 || cols < 0 || ldx < cols || ldy < cols) return MPC_ERR_INVALID;
    if (rows == 0 || cols == 0) return MPC_OK;
    if (!x || !y) return MPC_ERR_INVALID;
    const int64_t total = rows * ldy;
    const unsigned grid = (unsigned)(ceil_div(total, 256) < (int64_t)kNumSMs * 8 ? ceil_div(total, 256) : (int64_t)kNumSMs * 8);
    pdl_launch(tcb::f32_to_bf16_kernel, dim3(grid), dim3(256), 0, (cudaStream_t)stream, x,
               static_cast<__nv_bfloat16*>(y), rows, cols, ldx, ldy);
    MPC_LAUNCH_CHECK();
    return MPC_OK;
}
