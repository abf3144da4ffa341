// tcgen05 / TMEM / TMA / mbarrier plumbing shared by the tensor-core kernels of this library (linear_tc.cu: the
// shared-MLP GEMMs; knn_tc.cu: the distance GEMM of the feature-space neighbour search): PTX wrappers, shared-memory
// matrix descriptors for fp32 (kind::tf32) operands, and the host-side tensor-map encoding with its small cache.
#pragma once
#include <cuda.h>

#include <mutex>

#include "common.cuh"

namespace mpc {
namespace tc {

constexpr int BLOCK_K = 32;               // fp32 elements = one 128-byte swizzle row
constexpr int UMMA_K = 8;                 // tf32: 32 bytes per instruction
constexpr uint32_t TF32_MASK = 0xffffe000u;
extern long long* g_trace;  // debug timeline buffer (mpc_debug_trace_buffer, linear_tc.cu)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!done);
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
            smem_u32(dst)),
        "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, "
        "[%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_proxy() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map),
                 "r"(smem_u32(src)), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void tma_reduce_add_2d(const CUtensorMap* map, const void* src, int c0, int c1) {
    asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map),
                 "r"(smem_u32(src)), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void epi_barrier() { asm volatile("bar.sync 1, 128;" ::: "memory"); }


// K-major, 128-byte-swizzled shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start address >> 4,
// LBO = 1 (unused for swizzled K-major), SBO = 1024 B (8 rows x 128 B) >> 4, version 1, layout SWIZZLE_128B (2).
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3fffu) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
// MN-major descriptor for the weight-gradient GEMM, whose operands are contiguous along the OUTPUT dimension:
// 32-element (128 B) rows run along M/N, successive rows are successive reduction indices.  For 32-bit (tf32)
// MN-major operands the only legal shared-memory layout is SWIZZLE_128B_BASE32B (layout type 1, Swizzle<2,5,2>:
// 32-byte pieces of a 128-byte row XORed with row % 4), which is what TMA's SWIZZLE_128B_ATOM_32B writes.
// One TMA box = 32 (M/N) x 32 (reduction) = 4 KB; LBO = distance between 32-wide M/N atoms = 4096 B,
// SBO = distance between 4-row reduction groups = 512 B.
__device__ __forceinline__ uint64_t make_desc_mn(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3fffu) | (256ull << 16) | (32ull << 32) | (1ull << 46) | (1ull << 61);
}


// ---- host side ----------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// 2D fp32 row-major [rows, cols] (row stride ld floats), box = 32 columns x box_rows rows, 128B swizzle, zero fill.
// Encoding a tensor map costs a few microseconds of host time; layers are called with the same buffers step after
// step (the caching allocator hands the same addresses back), so encoded maps are kept in a small direct-mapped
// cache keyed by everything that goes into the encoding.
struct MapKey {
    const void* base;
    int64_t rows, cols, ld;
    int box_rows, mn;
    bool operator==(const MapKey& o) const {
        return base == o.base && rows == o.rows && cols == o.cols && ld == o.ld && box_rows == o.box_rows && mn == o.mn;
    }
};
struct MapSlot {
    MapKey key;
    CUtensorMap map;
    bool valid;
};
constexpr int MAP_CACHE_SLOTS = 1024;
static MapSlot g_map_cache[MAP_CACHE_SLOTS];
static std::mutex g_map_mutex;

static int make_map(CUtensorMap* map, const float* base, int64_t rows, int64_t cols, int64_t ld, int box_rows,
                    bool mn_major = false) {
    if (((uintptr_t)base & 15u) || (ld & 3)) return MPC_ERR_UNSUPPORTED;
    const MapKey key{base, rows, cols, ld, box_rows, mn_major ? 1 : 0};
    uint64_t h = (uint64_t)(uintptr_t)base * 0x9E3779B97F4A7C15ull;
    h ^= (uint64_t)rows * 0xC2B2AE3D27D4EB4Full + (uint64_t)cols * 0x165667B19E3779F9ull + (uint64_t)ld * 31 +
         (uint64_t)box_rows * 7 + (uint64_t)key.mn;
    MapSlot& slot = g_map_cache[(h >> 20) % MAP_CACHE_SLOTS];
    {
        std::lock_guard<std::mutex> lock(g_map_mutex);
        if (slot.valid && slot.key == key) {
            *map = slot.map;
            return MPC_OK;
        }
    }
    EncodeTiledFn fn = encode_fn();
    if (!fn) return MPC_ERR_UNSUPPORTED;
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
    cuuint32_t box[2] = {(cuuint32_t)BLOCK_K, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE,
                    mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return MPC_ERR_INVALID;
    std::lock_guard<std::mutex> lock(g_map_mutex);
    slot.key = key;
    slot.map = *map;
    slot.valid = true;
    return MPC_OK;
}


}  // namespace tc
}  // namespace mpc
