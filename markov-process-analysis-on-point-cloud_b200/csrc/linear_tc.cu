// Shared-MLP projection Y = X W^T + b on the 5th-generation tensor cores (tcgen05 + TMEM + TMA) for sm_100a.
// Replaces the nn.Linear inside the reference's `Linear` / q,k,v projections (pointnet2_utils.py:408,415,
// 489-491) -- the one family of ops on this path that is a real dense contraction.
//
// Precision: the reference is fp32 and the parity bar for features is rtol 1e-5..1e-4, which a single TF32
// pass (10-bit mantissa) cannot meet.  The kernel therefore runs the 3xTF32 split: every fp32 operand tile is
// split in shared memory into hi = top 19 bits and lo = (x - hi) truncated to TF32, and each K step issues
//      D += A_lo B_hi ; D += A_hi B_lo ; D += A_hi B_hi        (fp32 accumulation in TMEM)
// which recovers ~fp32 accuracy at one third of the TF32 rate -- still ~5x the FP32 SIMT pipe, enough to put
// every layer width of both models back under the HBM roofline.
//
// Structure (one persistent CTA per SM, 448 threads, warp-specialised):
//   warps 0-3   epilogue  : each warp owns 32 accumulator rows (TMEM lanes) end to end: tcgen05.ld, + bias, its own
//                           4 KB swizzled staging slab, its own 32 x 32 TMA store -- no CTA-level barrier per slab
//   warps 4-11  splitters : wait for a TMA stage, write lo = x - trunc_tf32(x) next to the raw tile (which doubles as
//                           the hi term), fence to the async proxy, hand the stage to the MMA warp
//   warp  12    TMA producer (one elected lane): cp.async.bulk.tensor 2D, 128B swizzle, fp32 boxes 32 x 128 (A)
//                           and 32 x BLOCK_N (B) per stage
//   warp  13    MMA issuer (one elected lane) + TMEM allocation: 12 tcgen05.mma.kind::tf32 per stage
//                           (4 K-steps of 8 x 3 split terms), tcgen05.commit frees the stage / publishes the tile
// Pipelines: smem stages (full -> split -> empty), 2 TMEM accumulator stages (tmem_full / tmem_empty), so the
// epilogue of tile i overlaps the main loop of tile i+1.
#include "tc_common.cuh"

namespace mpc {
namespace tc {

constexpr int BLOCK_M = 128;
constexpr int A_TILE_BYTES = BLOCK_M * BLOCK_K * 4;  // 16 KB
constexpr int THREADS = 448;  // warps 0-3 epilogue, 4-11 splitters, 12 TMA producer, 13 MMA issuer
constexpr int SPLIT_THREADS = 256;
constexpr int MAX_STAGES = 6;

constexpr int STAGING_BYTES = BLOCK_M * 128;  // one 32-column slab of a tile: 128 rows x 128 B
constexpr int EPI_BIAS_BYTES = 2048;  // bias (shift) and scale for up to 256 columns

struct Params {
    int M, N, K;
    int block_n;      // accumulator columns per tile (multiple of 16, <= 256)
    int n_tiles;      // ceil(N / block_n)
    int m_tiles;      // ceil(M / 128)
    int stages;
    const float* bias;  // may be null
    const float* bias2; // optional [groups][N]: a second bias shared by `bias2_rows` consecutive rows (a per-cloud
    int bias2_rows;     //   vector: the projection of channels that are constant over a cloud's points); any multiple of 128 is looked up per tile, anything else per row
    float* y;
    int ldy;
    int a_mn, b_mn;   // operand layouts: 0 = K-major (reduction index contiguous), 1 = MN-major (output index
                      // contiguous).  fwd: 0/0 (y = x w^T); dgrad: 0/1 (gx = gy w); wgrad: 1/1 (gw = gy^T x)
    int splits;       // reduction split across CTAs (wgrad; partial results are combined with TMA reduce-add)
    int k_chunks;     // ceil(reduction length / 32)
    long long* trace; // debug: per-role clock64() timeline of CTA 0 (null in production)
    int tma_out;      // 1: results leave through a swizzled staging slab + TMA store / TMA reduce-add
    int epi_bufs;     // staging slabs (1 or 2): with 2 the TMA store of slab i overlaps the fill of slab i+1
    double* stat_sum; // optional [2][N]: per-column sum and sum of squares of y (BatchNorm batch statistics),
                      // accumulated from the staged tile while it is still in shared memory (must be zero on entry)
    float4* zero_ptr; // optional: buffer this launch clears on behalf of a LATER launch in the same stream (the
    long long zero_n4;//   split-reduction target of the weight-gradient GEMM that follows a dgrad), in float4 units
    // inference epilogue (mpc_linear_affine_act_f32): y = LeakyReLU_slope(acc * scale[n] + bias[n]) (+ residual[m,n]) --
    // BatchNorm with running statistics folded to a per-channel affine map, activation and residual add in registers
    const float* scale;     // [N] or null (null: the plain y = acc + bias epilogue)
    float slope;
    const float* residual;  // [M, ldr] or null
    int ldr;
};

// debug timeline: role r in [0,4) records up to 255 timestamps
#define MPC_TRACE(role, n)                                                              \
    do {                                                                                \
        if (p.trace && blockIdx.x == 0 && (n) < 255) p.trace[(role) * 256 + (n)++] = clock64(); \
    } while (0)

__global__ void __launch_bounds__(THREADS, 1)
linear_3xtf32_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                     const __grid_constant__ CUtensorMap map_y, const Params p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ uint64_t bar_full[MAX_STAGES], bar_split[MAX_STAGES], bar_empty[MAX_STAGES];
    __shared__ uint64_t bar_tmem_full[2], bar_tmem_empty[2];
    __shared__ uint32_t tmem_base_slot;
    __shared__ float stat_accw[4][2][256];  // per-epilogue-warp column sums over all of this CTA's tiles; combined in
                                            // fp64 at the end: one atomic per column per CTA

    // programmatic dependent launch: let the next kernel in the stream start its own prologue now ...
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    // dynamic smem is requested with 1024 B of slack and aligned here (swizzle-128B atoms need 1024 B alignment)
    // (pointer arithmetic on the __shared__ array, not an integer round-trip, so accesses stay LDS/STS)
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const int b_tile_bytes = p.block_n * BLOCK_K * 4;
    const int stage_bytes = 2 * A_TILE_BYTES + 2 * b_tile_bytes;
    uint8_t* staging0 = smem + (size_t)p.stages * stage_bytes;                // epi_bufs x 16 KB, 1024-aligned
    float* bias_s = reinterpret_cast<float*>(staging0 + p.epi_bufs * STAGING_BYTES);  // 256 floats
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int total_tiles = p.m_tiles * p.n_tiles * p.splits;
    const int chunks_per_split = (p.k_chunks + p.splits - 1) / p.splits;
    // work item -> (row tile, column tile, reduction range)
    auto decode = [&](int item, int& mt, int& nt, int& kc0, int& kc1) {
        const int sp = item % p.splits;
        const int t = item / p.splits;
        mt = t / p.n_tiles;
        nt = t - mt * p.n_tiles;
        kc0 = sp * chunks_per_split;
        kc1 = min(p.k_chunks, kc0 + chunks_per_split);
    };

    if (threadIdx.x == 0) {
        for (int s = 0; s < p.stages; ++s) {
            mbar_init(&bar_full[s], 1);
            mbar_init(&bar_split[s], SPLIT_THREADS);
            mbar_init(&bar_empty[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&bar_tmem_full[a], 1);
            mbar_init(&bar_tmem_empty[a], 128);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x == 32) {  // descriptor fetches (~1 DRAM round trip each) off the critical path
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
        if (p.tma_out) asm volatile("prefetch.tensormap [%0];" ::"l"(&map_y) : "memory");
    }
    // ... and wait here, after the part of our prologue that touches no global data (barrier init, descriptor
    // prefetch), until the kernel before us has completed and its writes are visible
    asm volatile("griddepcontrol.wait;" ::: "memory");
    __syncthreads();  // barriers initialised: the TMA producer and the splitters start right away
    uint32_t tmem_base = 0;
    if (p.zero_ptr && warp < 8) {  // epilogue + splitter warps are idle until the first TMA round trip completes
        const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
        for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < p.zero_n4; i += (long long)gridDim.x * 256)
            p.zero_ptr[i] = z;
    }
    if (warp == 13) {  // TMEM: 512 columns = 2 accumulator stages x 256 fp32 columns
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)),
                     "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    if (warp == 13 || warp < 4) {  // only the MMA issuer and the epilogue warps need the TMEM address
        tc_fence_before();
        asm volatile("bar.sync 2, 160;" ::: "memory");
        tc_fence_after();
        tmem_base = tmem_base_slot;
    }

    if (warp == 12) {
        // ===== TMA producer =====
        if (lane == 0) {
            int it = 0, tn_ = 0;
            MPC_TRACE(0, tn_);
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                int mt, nt, kc0, kc1;
                decode(tile, mt, nt, kc0, kc1);
                for (int kc = kc0; kc < kc1; ++kc, ++it) {
                    const int s = it % p.stages;
                    const uint32_t ph = (it / p.stages) & 1;
                    mbar_wait(&bar_empty[s], ph ^ 1);
                    MPC_TRACE(0, tn_);
                    uint8_t* st = smem + (size_t)s * stage_bytes;
                    mbar_arrive_expect_tx(&bar_full[s], A_TILE_BYTES + b_tile_bytes);
                    // K-major: one box of 32 reduction columns x rows.  MN-major: 4 KB boxes of 32 output
                    // indices x 32 reduction rows.
                    if (!p.a_mn) {
                        tma_load_2d(&map_a, &bar_full[s], st, kc * BLOCK_K, mt * BLOCK_M);
                    } else {
                        for (int a = 0; a < BLOCK_M / 32; ++a)
                            tma_load_2d(&map_a, &bar_full[s], st + a * 4096, mt * BLOCK_M + a * 32, kc * BLOCK_K);
                    }
                    if (!p.b_mn) {
                        tma_load_2d(&map_b, &bar_full[s], st + 2 * A_TILE_BYTES, kc * BLOCK_K, nt * p.block_n);
                    } else {
                        for (int b = 0; b < p.block_n / 32; ++b)
                            tma_load_2d(&map_b, &bar_full[s], st + 2 * A_TILE_BYTES + b * 4096, nt * p.block_n + b * 32,
                                        kc * BLOCK_K);
                    }
                }
            }
        }
    } else if (warp >= 4 && warp < 12) {
        // ===== splitters: lo = x - trunc_tf32(x) next to the raw tile, 16 bytes per thread per step =====
        const int t = threadIdx.x - 128;
        int it = 0, tn_ = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            int mt, nt, kc0, kc1;
            decode(tile, mt, nt, kc0, kc1);
            for (int kc = kc0; kc < kc1; ++kc, ++it) {
                const int s = it % p.stages;
                const uint32_t ph = (it / p.stages) & 1;
                mbar_wait(&bar_full[s], ph);
                if (t == 0) MPC_TRACE(1, tn_);
                uint8_t* st = smem + (size_t)s * stage_bytes;
                uint4* a_hi = reinterpret_cast<uint4*>(st);
                uint4* a_lo = reinterpret_cast<uint4*>(st + A_TILE_BYTES);
                uint4* b_hi = reinterpret_cast<uint4*>(st + 2 * A_TILE_BYTES);
                uint4* b_lo = reinterpret_cast<uint4*>(st + 2 * A_TILE_BYTES + b_tile_bytes);
                // kind::tf32 reads the top 19 bits of each 32-bit container and ignores the rest, so the tile as
                // TMA delivered it already IS the hi term; only lo = x - trunc(x) has to be materialised
                auto split = [](uint4 v, uint4& lo) {
                    lo.x = __float_as_uint(__uint_as_float(v.x) - __uint_as_float(v.x & TF32_MASK));
                    lo.y = __float_as_uint(__uint_as_float(v.y) - __uint_as_float(v.y & TF32_MASK));
                    lo.z = __float_as_uint(__uint_as_float(v.z) - __uint_as_float(v.z & TF32_MASK));
                    lo.w = __float_as_uint(__uint_as_float(v.w) - __uint_as_float(v.w & TF32_MASK));
                };
                {   // all loads first (independent), then split + store: one shared-memory latency, not eight
                    uint4 v[A_TILE_BYTES / 16 / SPLIT_THREADS];
#pragma unroll
                    for (int i = 0; i < A_TILE_BYTES / 16 / SPLIT_THREADS; ++i) v[i] = a_hi[t + i * SPLIT_THREADS];
#pragma unroll
                    for (int i = 0; i < A_TILE_BYTES / 16 / SPLIT_THREADS; ++i) {
                        uint4 lo;
                        split(v[i], lo);
                        a_lo[t + i * SPLIT_THREADS] = lo;
                    }
                }
                for (int i0 = t; i0 < b_tile_bytes / 16; i0 += 4 * SPLIT_THREADS) {
                    uint4 v[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u)
                        if (i0 + u * SPLIT_THREADS < b_tile_bytes / 16) v[u] = b_hi[i0 + u * SPLIT_THREADS];
#pragma unroll
                    for (int u = 0; u < 4; ++u)
                        if (i0 + u * SPLIT_THREADS < b_tile_bytes / 16) {
                            uint4 lo;
                            split(v[u], lo);
                            b_lo[i0 + u * SPLIT_THREADS] = lo;
                        }
                }
                fence_async_proxy();  // generic-proxy writes -> visible to the tensor core's async-proxy reads
                mbar_arrive(&bar_split[s]);
                if (t == 0) MPC_TRACE(1, tn_);
            }
        }
    } else if (warp == 13) {
        // ===== MMA issuer =====
        if (lane == 0) {
            // instruction descriptor (cute::UMMA::InstrDescriptor): D = F32 (1 << 4), A = B = TF32 (2 << 7, 2 << 10),
            // both K-major, N >> 3 at bit 17, M >> 4 at bit 24
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(p.block_n >> 3) << 17) |
                                   ((uint32_t)(BLOCK_M >> 4) << 24) |
                                   (p.a_mn ? (1u << 15) : 0u) | (p.b_mn ? (1u << 16) : 0u);  // a_major / b_major
            int it = 0, tile_it = 0, tn_ = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++tile_it) {
                const int as = tile_it & 1;
                const uint32_t aph = (tile_it >> 1) & 1;
                mbar_wait(&bar_tmem_empty[as], aph ^ 1);
                MPC_TRACE(2, tn_);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + (uint32_t)(as * 256);
                int mt, nt, kc0, kc1;
                decode(tile, mt, nt, kc0, kc1);
                for (int kc = kc0; kc < kc1; ++kc, ++it) {
                    const int s = it % p.stages;
                    const uint32_t ph = (it / p.stages) & 1;
                    mbar_wait(&bar_split[s], ph);
                    MPC_TRACE(2, tn_);
                    tc_fence_after();
                    const uint32_t st = smem_u32(smem + (size_t)s * stage_bytes);
                    const uint32_t o_alo = A_TILE_BYTES, o_bhi = 2 * A_TILE_BYTES, o_blo = 2 * A_TILE_BYTES + b_tile_bytes;
                    const uint64_t a_hi = p.a_mn ? make_desc_mn(st) : make_desc(st);
                    const uint64_t a_lo = p.a_mn ? make_desc_mn(st + o_alo) : make_desc(st + o_alo);
                    const uint64_t b_hi = p.b_mn ? make_desc_mn(st + o_bhi) : make_desc(st + o_bhi);
                    const uint64_t b_lo = p.b_mn ? make_desc_mn(st + o_blo) : make_desc(st + o_blo);
                    // K-major: +32 B inside the swizzle row per K step; MN-major: +1024 B (next 8 reduction rows)
                    const uint64_t step_a = p.a_mn ? (1024u >> 4) : ((UMMA_K * 4) >> 4);
                    const uint64_t step_b = p.b_mn ? (1024u >> 4) : ((UMMA_K * 4) >> 4);
#pragma unroll
                    for (int kk = 0; kk < BLOCK_K / UMMA_K; ++kk) {
                        const uint64_t ada = step_a * kk, adb = step_b * kk;
                        umma_tf32(tmem_d, a_lo + ada, b_hi + adb, idesc, (kc != kc0 || kk != 0) ? 1u : 0u);
                        umma_tf32(tmem_d, a_hi + ada, b_lo + adb, idesc, 1u);
                        umma_tf32(tmem_d, a_hi + ada, b_hi + adb, idesc, 1u);
                    }
                    umma_commit(&bar_empty[s]);  // the stage may be refilled once these MMAs have read it
                }
                umma_commit(&bar_tmem_full[as]);  // accumulator complete
            }
        }
    } else {
        // ===== epilogue: warp w owns TMEM lanes 32w..32w+31 = output rows 32w..32w+31 of the tile =====
        int tile_it = 0, tn_ = 0, slab_it = 0;
        int stat_n0 = -1;  // column offset the statistics accumulators currently belong to
        int bias_n0 = -1;  // column offset bias_s currently holds
        int bias_grp = -1; // row group (per-cloud bias) bias_s currently holds
        float* my_acc0 = &stat_accw[warp][0][0];
        float* my_acc1 = &stat_accw[warp][1][0];
        auto stat_flush = [&]() {  // all four epilogue warps: combine the per-warp sums in fp64, warp 0 publishes
            if (stat_n0 < 0) return;
            epi_barrier();
            if (warp == 0)
                for (int i = lane; i < p.block_n; i += 32)
                    if (stat_n0 + i < p.N) {
                        const double a = (double)stat_accw[0][0][i] + (double)stat_accw[1][0][i] +
                                         (double)stat_accw[2][0][i] + (double)stat_accw[3][0][i];
                        const double b = (double)stat_accw[0][1][i] + (double)stat_accw[1][1][i] +
                                         (double)stat_accw[2][1][i] + (double)stat_accw[3][1][i];
                        atomicAdd(p.stat_sum + stat_n0 + i, a);
                        atomicAdd(p.stat_sum + p.N + stat_n0 + i, b);
                    }
            epi_barrier();
        };
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++tile_it) {
            int mt, nt, kc0, kc1;
            decode(tile, mt, nt, kc0, kc1);
            const int as = tile_it & 1;
            const uint32_t aph = (tile_it >> 1) & 1;
            const int row = mt * BLOCK_M + warp * 32 + lane;
            const int n0 = nt * p.block_n;
            if (p.stat_sum && n0 != stat_n0) {  // (the four warps walk the same tile sequence: uniform decision)
                stat_flush();
                stat_n0 = n0;
                for (int i = lane; i < p.block_n; i += 32) my_acc0[i] = my_acc1[i] = 0.f;
            }
            const bool b2_row = p.bias2 && (p.bias2_rows % BLOCK_M) != 0;  // groups not tile aligned: per-row lookup
            const int grp = (p.bias2 && !b2_row) ? (mt * BLOCK_M) / p.bias2_rows : 0;
            if (p.tma_out && (n0 != bias_n0 || grp != bias_grp)) {
                // bias of this column tile -> shared (broadcast reads below); reloaded only when the column tile (or
                // the row group of a per-cloud bias) changes -- once per kernel in every plain layer of both models
                epi_barrier();
                for (int i = threadIdx.x; i < p.block_n; i += 128) {
                    float v = (p.bias && n0 + i < p.N) ? __ldg(p.bias + n0 + i) : 0.f;
                    if (p.bias2 && !b2_row && n0 + i < p.N) v += __ldg(p.bias2 + (size_t)grp * p.N + n0 + i);
                    bias_s[i] = v;
                    if (p.scale) bias_s[256 + i] = n0 + i < p.N ? __ldg(p.scale + n0 + i) : 1.0f;
                }
                epi_barrier();
                bias_n0 = n0;
                bias_grp = grp;
            }
            mbar_wait(&bar_tmem_full[as], aph);
            if (threadIdx.x == 0) MPC_TRACE(3, tn_);
            tc_fence_after();
            const uint32_t taddr = tmem_base + (uint32_t)(as * 256) + ((uint32_t)(warp * 32) << 16);
            if (p.tma_out) {
                const int rows_valid = min(32, p.M - (mt * BLOCK_M + warp * 32));  // rows of this warp inside M
                for (int c = 0; c < p.block_n; c += 32, ++slab_it) {
                    // this warp's 4 KB slab (32 rows x 128 B, 1024-aligned: the 128B swizzle pattern repeats every 8 rows)
                    uint8_t* staging = staging0 + (p.epi_bufs == 2 ? (slab_it & 1) * STAGING_BYTES : 0) + warp * 4096;
                    uint4* srow = reinterpret_cast<uint4*>(staging + lane * 128);
                    uint32_t r0[16], r1[16];
                    tmem_ld16(taddr + c, r0);
                    tmem_ld16(taddr + c + 16, r1);
                    tmem_ld_wait();
                    // the TMA store of this warp that last used this slab must have finished READING it
                    if (lane == 0) {
                        if (p.epi_bufs == 2)
                            bulk_wait_read1();
                        else
                            bulk_wait_read0();
                    }
                    __syncwarp();
                    if (p.scale) {
                        // inference epilogue: affine (BatchNorm with running statistics) + LeakyReLU + residual
                        const bool row_ok = lane < rows_valid;
                        const float* rrow = p.residual ? p.residual + (size_t)row * p.ldr + n0 + c : nullptr;
#pragma unroll
                        for (int q4 = 0; q4 < 8; ++q4) {
                            const uint32_t* src = q4 < 4 ? &r0[q4 * 4] : &r1[(q4 - 4) * 4];
                            float v[4];
#pragma unroll
                            for (int u = 0; u < 4; ++u) {
                                const float t = fmaf(__uint_as_float(src[u]), bias_s[256 + c + q4 * 4 + u],
                                                     bias_s[c + q4 * 4 + u]);
                                v[u] = t > 0.f ? t : t * p.slope;
                            }
                            if (rrow && row_ok) {
                                if (n0 + c + q4 * 4 + 4 <= p.N && (p.ldr & 3) == 0) {
                                    const float4 rv = __ldg(reinterpret_cast<const float4*>(rrow) + q4);
                                    v[0] += rv.x; v[1] += rv.y; v[2] += rv.z; v[3] += rv.w;
                                } else {
#pragma unroll
                                    for (int u = 0; u < 4; ++u)
                                        if (n0 + c + q4 * 4 + u < p.N) v[u] += __ldg(rrow + q4 * 4 + u);
                                }
                            }
                            srow[q4 ^ (lane & 7)] = make_uint4(__float_as_uint(v[0]), __float_as_uint(v[1]),
                                                               __float_as_uint(v[2]), __float_as_uint(v[3]));
                        }
                    } else if (b2_row) {
                        // per-cloud bias whose row groups are not tile aligned (24 000-point blocks): this thread's row
                        // looks its cloud's vector up itself (a [clouds, N] table, L2 resident)
                        const int rr = row < p.M ? row : p.M - 1;
                        const float* b2 = p.bias2 + (size_t)(rr / p.bias2_rows) * p.N + n0 + c;
#pragma unroll
                        for (int q4 = 0; q4 < 8; ++q4) {
                            const uint32_t* src = q4 < 4 ? &r0[q4 * 4] : &r1[(q4 - 4) * 4];
                            float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
                            if (n0 + c + q4 * 4 + 4 <= p.N) {
                                g = __ldg(reinterpret_cast<const float4*>(b2) + q4);
                            } else {
                                if (n0 + c + q4 * 4 + 0 < p.N) g.x = __ldg(b2 + q4 * 4 + 0);
                                if (n0 + c + q4 * 4 + 1 < p.N) g.y = __ldg(b2 + q4 * 4 + 1);
                                if (n0 + c + q4 * 4 + 2 < p.N) g.z = __ldg(b2 + q4 * 4 + 2);
                            }
                            uint4 o;
                            o.x = __float_as_uint(__uint_as_float(src[0]) + (bias_s[c + q4 * 4 + 0] + g.x));
                            o.y = __float_as_uint(__uint_as_float(src[1]) + (bias_s[c + q4 * 4 + 1] + g.y));
                            o.z = __float_as_uint(__uint_as_float(src[2]) + (bias_s[c + q4 * 4 + 2] + g.z));
                            o.w = __float_as_uint(__uint_as_float(src[3]) + (bias_s[c + q4 * 4 + 3] + g.w));
                            srow[q4 ^ (lane & 7)] = o;
                        }
                    } else {
#pragma unroll
                    for (int q4 = 0; q4 < 8; ++q4) {  // 8 x 16 B of this row, 128B-swizzled like the TMA box expects
                        const uint32_t* src = q4 < 4 ? &r0[q4 * 4] : &r1[(q4 - 4) * 4];
                        uint4 o;
                        o.x = __float_as_uint(__uint_as_float(src[0]) + bias_s[c + q4 * 4 + 0]);
                        o.y = __float_as_uint(__uint_as_float(src[1]) + bias_s[c + q4 * 4 + 1]);
                        o.z = __float_as_uint(__uint_as_float(src[2]) + bias_s[c + q4 * 4 + 2]);
                        o.w = __float_as_uint(__uint_as_float(src[3]) + bias_s[c + q4 * 4 + 3]);
                        srow[q4 ^ (lane & 7)] = o;
                    }
                    }
                    fence_async_proxy();
                    __syncwarp();
                    if (lane == 0 && rows_valid > 0) {
                        if (p.splits > 1)
                            tma_reduce_add_2d(&map_y, staging, n0 + c, mt * BLOCK_M + warp * 32);
                        else
                            tma_store_2d(&map_y, staging, n0 + c, mt * BLOCK_M + warp * 32);
                    }
                    if (lane == 0) bulk_commit();
                    if (p.stat_sum) {
                        // BatchNorm statistics of this warp's 32 x 32 slab straight from shared memory: lane = column
                        // (conflict-free: one 128-byte row per step), rows beyond M excluded
                        const float* sf = reinterpret_cast<const float*>(staging);
                        float su = 0.f, sq2 = 0.f;
#pragma unroll 8
                        for (int r = 0; r < 32; ++r) {
                            if (r < rows_valid) {
                                const float v = sf[r * 32 + ((((lane >> 2) ^ (r & 7)) << 2) | (lane & 3))];
                                su += v;
                                sq2 = fmaf(v, v, sq2);
                            }
                        }
                        my_acc0[c + lane] += su;
                        my_acc1[c + lane] += sq2;
                    }
                }
            } else {
            float* yrow = p.y + (size_t)row * p.ldy + n0;
            for (int c = 0; c < p.block_n; c += 16) {
                uint32_t r[16];
                tmem_ld16(taddr + c, r);
                tmem_ld_wait();
                if (row < p.M) {
                    float v[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        v[i] = __uint_as_float(r[i]);
                        if (p.bias && n0 + c + i < p.N) v[i] += __ldg(p.bias + n0 + c + i);
                    }
                    if (p.splits > 1) {  // partial sums of a split reduction
                        if (n0 + c + 16 <= p.N && (p.ldy & 3) == 0) {
#pragma unroll
                            for (int i = 0; i < 16; i += 4)
                                red_add_f32x4(yrow + c + i, make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]));
                        } else {
#pragma unroll
                            for (int i = 0; i < 16; ++i)
                                if (n0 + c + i < p.N) red_add_f32(yrow + c + i, v[i]);
                        }
                    } else if (n0 + c + 16 <= p.N && (p.ldy & 3) == 0) {
#pragma unroll
                        for (int i = 0; i < 16; i += 4)
                            *reinterpret_cast<float4*>(yrow + c + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
                    } else {
#pragma unroll
                        for (int i = 0; i < 16; ++i)
                            if (n0 + c + i < p.N) yrow[c + i] = v[i];
                    }
                }
            }
            }
            tc_fence_before();
            mbar_arrive(&bar_tmem_empty[as]);
            if (threadIdx.x == 0) MPC_TRACE(3, tn_);
        }
        if (p.stat_sum) stat_flush();
        if (lane == 0 && p.tma_out) bulk_wait0();  // this warp's staged results have left shared memory
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 13) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
    }
}

long long* g_trace = nullptr;  // debug only, see mpc_debug_trace_buffer (also read by knn_tc.cu)

// Launch with programmatic stream serialization: the kernel may be scheduled while its predecessor in the stream
// is still draining (it blocks in griddepcontrol.wait before reading anything the predecessor wrote).
static cudaError_t launch_pdl(int grid, size_t smem, cudaStream_t st, const CUtensorMap& a, const CUtensorMap& b,
                              const CUtensorMap& y, const Params& p);

// Shared-memory budget (224 KB opt-in, 1 KB alignment slack): as many operand stages as fit; a second epilogue
// staging slab only if at least 3 operand stages remain (TMA latency ~1.3 us needs >= 3 stages in flight).
static void pick_pipeline(int stage_bytes, int* stages, int* epi_bufs) {
    const int budget = 214 * 1024 - EPI_BIAS_BYTES;  // 227 KB per CTA minus static shared memory and alignment slack
    int s2 = (budget - 2 * STAGING_BYTES) / stage_bytes;
    int s1 = (budget - STAGING_BYTES) / stage_bytes;
    if (s2 >= 3) {
        *stages = s2 > MAX_STAGES ? MAX_STAGES : s2;
        *epi_bufs = 2;
    } else {
        *stages = s1 > MAX_STAGES ? MAX_STAGES : s1;
        *epi_bufs = 1;
    }
}

static cudaError_t ensure_smem_optin() {
    // the attribute is per device (and idempotent: a benign race at worst sets it twice)
    static bool done[64] = {};
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev >= 0 && dev < 64 && done[dev]) return cudaSuccess;
    e = cudaFuncSetAttribute(linear_3xtf32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 216 * 1024);  // + ~10 KB static <= 227 KB
    if (dev >= 0 && dev < 64) done[dev] = e == cudaSuccess;
    return e;
}

static cudaError_t launch_pdl(int grid, size_t smem, cudaStream_t st, const CUtensorMap& a, const CUtensorMap& b,
                              const CUtensorMap& y, const Params& p) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, linear_3xtf32_kernel, a, b, y, p);
}

}  // namespace tc
}  // namespace mpc

static int linear_fwd_impl(const float* x, int64_t ldx, const float* w, int64_t ldw, const float* bias, float* y,
                           int64_t ldy, double* stat_scratch, const float* group_bias, int64_t rows_per_group,
                           int64_t M, int64_t K, int64_t N, const float* scale, float slope, const float* residual,
                           int64_t ldr, mpc_stream_t stream) {
    using namespace mpc;
    using namespace mpc::tc;
    if (!x || !w || !y || M <= 0 || K <= 0 || N <= 0) return MPC_ERR_INVALID;
    if (K % BLOCK_K || ldx < K || ldw < K || ldy < N || (ldx & 3) || (ldw & 3)) return MPC_ERR_UNSUPPORTED;
    if (((uintptr_t)x | (uintptr_t)w) & 15u) return MPC_ERR_UNSUPPORTED;
    if (M > INT32_MAX || N > 65536 || K > 65536) return MPC_ERR_UNSUPPORTED;
    Params p;
    p.M = (int)M;
    p.N = (int)N;
    p.K = (int)K;
    // results leave through TMA when y rows are 16-byte aligned; the accumulator tile is then a multiple of 32 columns
    const int tma_out = ((ldy & 3) == 0 && ((uintptr_t)y & 15u) == 0) ? 1 : 0;
    int bn = (int)(N < 256 ? N : 256);
    bn = tma_out ? ((bn + 31) & ~31) : ((bn + 15) & ~15);
    // few rows: narrower column tiles until the tile count reaches the SM count (each CTA then loads and splits a
    // proportionally smaller slice of the weight matrix instead of 32 CTAs each chewing through all of it)
    while (bn > 64 && ceil_div(M, BLOCK_M) * ceil_div(N, bn) < kNumSMs) bn = (bn / 2 + 31) & ~31;
    p.block_n = bn;
    p.n_tiles = (int)ceil_div(N, bn);
    p.m_tiles = (int)ceil_div(M, BLOCK_M);
    p.tma_out = tma_out;
    p.stat_sum = stat_scratch;
    if (stat_scratch && !tma_out) return MPC_ERR_UNSUPPORTED;  // zero on entry is the caller's contract
    if (group_bias && (!tma_out || rows_per_group <= 0 || (N & 3))) return MPC_ERR_UNSUPPORTED;
    p.bias2 = group_bias;
    p.bias2_rows = (int)(group_bias ? rows_per_group : 1);
    p.zero_ptr = nullptr;
    p.zero_n4 = 0;
    const int stage_bytes = 2 * A_TILE_BYTES + 2 * bn * BLOCK_K * 4;
    int stages;
    pick_pipeline(stage_bytes, &stages, &p.epi_bufs);
    if (stages < 2) return MPC_ERR_UNSUPPORTED;
    p.stages = stages;
    p.bias = bias;
    p.y = y;
    p.ldy = (int)ldy;
    p.a_mn = 0;
    p.b_mn = 0;
    p.splits = 1;
    p.k_chunks = (int)(K / BLOCK_K);
    p.trace = g_trace;
    p.scale = scale;
    p.slope = slope;
    p.residual = residual;
    p.ldr = (int)ldr;
    if (scale && !tma_out) return MPC_ERR_UNSUPPORTED;
    CUtensorMap map_a, map_b;
    int rc = make_map(&map_a, x, M, K, ldx, BLOCK_M);
    if (rc) return rc;
    rc = make_map(&map_b, w, N, K, ldw, bn);
    if (rc) return rc;
    CUtensorMap map_y = map_a;  // placeholder when results are stored directly
    if (tma_out) {
        rc = make_map(&map_y, y, M, N, ldy, 32);  // box 32 columns x 32 rows (one epilogue warp), 128B swizzle
        if (rc) return rc;
    }
    const size_t smem = (size_t)stages * stage_bytes + p.epi_bufs * STAGING_BYTES + EPI_BIAS_BYTES + 1024;
    MPC_CUDA(ensure_smem_optin());
    const int total_tiles = p.m_tiles * p.n_tiles;
    const int grid = total_tiles < kNumSMs ? total_tiles : kNumSMs;
    MPC_CUDA(launch_pdl(grid, smem, (cudaStream_t)stream, map_a, map_b, map_y, p));
    return MPC_OK;
}

MPC_API int mpc_linear_fwd_f32(const float* x, int64_t ldx, const float* w, int64_t ldw, const float* bias, float* y,
                               int64_t ldy, double* stat_scratch, const float* group_bias, int64_t rows_per_group,
                               int64_t M, int64_t K, int64_t N, mpc_stream_t stream) {
    return linear_fwd_impl(x, ldx, w, ldw, bias, y, ldy, stat_scratch, group_bias, rows_per_group, M, K, N, nullptr,
                           1.0f, nullptr, 0, stream);
}

MPC_API int mpc_linear_affine_act_f32(const float* x, int64_t ldx, const float* w, int64_t ldw, const float* scale,
                                      const float* shift, float slope, const float* residual, int64_t ldr, float* y,
                                      int64_t ldy, int64_t M, int64_t K, int64_t N, mpc_stream_t stream) {
    if (!scale || !shift || (residual && ldr < N)) return MPC_ERR_INVALID;
    return linear_fwd_impl(x, ldx, w, ldw, shift, y, ldy, nullptr, nullptr, 0, M, K, N, scale, slope, residual, ldr,
                           stream);
}

// grad_w[N,K] = gy[M,N]^T x[M,K]: both operands are contiguous along the OUTPUT dimensions (MN-major for the
// tensor core), the reduction runs over all M points and is split across the SMs.
MPC_API int mpc_linear_wgrad_f32(const float* gy, int64_t ldg, const float* x, int64_t ldx, float* gw, int64_t ldw,
                                 int64_t M, int64_t K, int64_t N, int64_t gw_is_zero, mpc_stream_t stream) {
    using namespace mpc;
    using namespace mpc::tc;
    if (!gy || !x || !gw || M <= 0 || K <= 0 || N <= 0) return MPC_ERR_INVALID;
    if (K % 32 || ldg < N || ldx < K || ldw < K || (ldg & 3) || (ldx & 3)) return MPC_ERR_UNSUPPORTED;
    if (((uintptr_t)gy | (uintptr_t)x) & 15u) return MPC_ERR_UNSUPPORTED;
    if (M > INT32_MAX || N > 65536 || K > 65536) return MPC_ERR_UNSUPPORTED;
    cudaStream_t st = (cudaStream_t)stream;
    Params p;
    p.M = (int)N;  // output rows  = N (tile of 128, TMEM lanes)
    p.N = (int)K;  // output cols  = K (accumulator columns)
    p.K = (int)M;  // reduction    = M points
    int bn = (int)(K < 256 ? K : 256);  // multiple of 32 because K % 32 == 0
    p.block_n = bn;
    p.n_tiles = (int)ceil_div(K, bn);
    p.m_tiles = (int)ceil_div(N, BLOCK_M);
    p.k_chunks = (int)ceil_div(M, BLOCK_K);
    const int tiles = p.m_tiles * p.n_tiles;
    int splits = kNumSMs / tiles;
    if (splits < 1) splits = 1;
    if (splits > p.k_chunks) splits = p.k_chunks;
    // every split must own at least one chunk: shrink until ceil-division leaves none empty
    while (splits > 1 && (int)ceil_div(p.k_chunks, splits) * (splits - 1) >= p.k_chunks) --splits;
    p.splits = splits;
    const int stage_bytes = 2 * A_TILE_BYTES + 2 * bn * BLOCK_K * 4;
    int stages;
    pick_pipeline(stage_bytes, &stages, &p.epi_bufs);
    if (stages < 2) return MPC_ERR_UNSUPPORTED;
    p.stages = stages;
    p.bias = nullptr;
    p.bias2 = nullptr;
    p.bias2_rows = 1;
    p.y = gw;
    p.ldy = (int)ldw;
    p.a_mn = 1;
    p.b_mn = 1;
    p.stat_sum = nullptr;
    p.zero_ptr = nullptr;
    p.zero_n4 = 0;
    p.tma_out = ((ldw & 3) == 0 && ((uintptr_t)gw & 15u) == 0) ? 1 : 0;
    p.trace = g_trace;
    p.scale = nullptr;
    p.slope = 1.0f;
    p.residual = nullptr;
    p.ldr = 0;
    CUtensorMap map_a, map_b;
    int rc = make_map(&map_a, gy, M, N, ldg, 32, true);  // dims {N, M}: inner = output index n, box 32 x 32
    if (rc) return rc;
    rc = make_map(&map_b, x, M, K, ldx, 32, true);
    if (rc) return rc;
    if (splits > 1 && !gw_is_zero) {
        if (ldw == K) {
            MPC_CUDA(cudaMemsetAsync(gw, 0, (size_t)N * K * sizeof(float), st));
        } else {
            MPC_CUDA(cudaMemset2DAsync(gw, (size_t)ldw * 4, 0, (size_t)K * 4, (size_t)N, st));
        }
    }
    CUtensorMap map_y = map_a;
    if (p.tma_out) {
        rc = make_map(&map_y, gw, N, K, ldw, 32);  // TMA reduce-add (or store) of 32 x 32 slabs
        if (rc) return rc;
    }
    const size_t smem = (size_t)stages * stage_bytes + p.epi_bufs * STAGING_BYTES + EPI_BIAS_BYTES + 1024;
    MPC_CUDA(ensure_smem_optin());
    const int items = tiles * splits;
    const int grid = items < kNumSMs ? items : kNumSMs;
    MPC_CUDA(launch_pdl(grid, smem, st, map_a, map_b, map_y, p));
    return MPC_OK;
}

MPC_API int mpc_debug_trace_buffer(void* device_buffer) {
    mpc::tc::g_trace = static_cast<long long*>(device_buffer);
    return MPC_OK;
}

// grad_x[M,K] = gy[M,N] w[N,K]: A = gy is K-major (the reduction index n is contiguous), B = w is MN-major (the
// output index k is contiguous), so the weight matrix is consumed as stored -- no transposed copy.
MPC_API int mpc_linear_dgrad_f32(const float* gy, int64_t ldg, const float* w, int64_t ldw, float* gx, int64_t ldx,
                                 int64_t M, int64_t K, int64_t N, float* zero_buf, int64_t zero_count,
                                 mpc_stream_t stream) {
    using namespace mpc;
    using namespace mpc::tc;
    if (!gy || !w || !gx || M <= 0 || K <= 0 || N <= 0) return MPC_ERR_INVALID;
    if (K % 32 || ldg < N || ldw < K || ldx < K || (ldg & 3) || (ldw & 3)) return MPC_ERR_UNSUPPORTED;
    if (((uintptr_t)gy | (uintptr_t)w) & 15u) return MPC_ERR_UNSUPPORTED;
    if (M > INT32_MAX || N > 65536 || K > 65536) return MPC_ERR_UNSUPPORTED;
    if (zero_buf && (((uintptr_t)zero_buf & 15u) || (zero_count & 3) || zero_count < 0)) return MPC_ERR_UNSUPPORTED;
    Params p;
    p.zero_ptr = reinterpret_cast<float4*>(zero_buf);
    p.zero_n4 = zero_buf ? zero_count / 4 : 0;
    p.M = (int)M;  // output rows
    p.N = (int)K;  // output cols = layer input width
    p.K = (int)N;  // reduction   = layer output width
    int bn = (int)(K < 256 ? K : 256);  // multiple of 32
    while (bn > 64 && ceil_div(M, BLOCK_M) * ceil_div(K, bn) < kNumSMs) bn = (bn / 2 + 31) & ~31;  // see fwd
    p.block_n = bn;
    p.n_tiles = (int)ceil_div(K, bn);
    p.m_tiles = (int)ceil_div(M, BLOCK_M);
    p.k_chunks = (int)ceil_div(N, BLOCK_K);
    p.splits = 1;
    p.tma_out = ((ldx & 3) == 0 && ((uintptr_t)gx & 15u) == 0) ? 1 : 0;
    const int stage_bytes = 2 * A_TILE_BYTES + 2 * bn * BLOCK_K * 4;
    int stages;
    pick_pipeline(stage_bytes, &stages, &p.epi_bufs);
    if (stages < 2) return MPC_ERR_UNSUPPORTED;
    p.stages = stages;
    p.bias = nullptr;
    p.bias2 = nullptr;
    p.bias2_rows = 1;
    p.y = gx;
    p.ldy = (int)ldx;
    p.a_mn = 0;
    p.b_mn = 1;
    p.stat_sum = nullptr;
    p.trace = g_trace;
    p.scale = nullptr;
    p.slope = 1.0f;
    p.residual = nullptr;
    p.ldr = 0;
    CUtensorMap map_a, map_b;
    int rc = make_map(&map_a, gy, M, N, ldg, BLOCK_M);        // K-major: box 32 reduction columns x 128 rows
    if (rc) return rc;
    rc = make_map(&map_b, w, N, K, ldw, 32, true);            // MN-major: box 32 output columns x 32 reduction rows
    if (rc) return rc;
    CUtensorMap map_y = map_a;
    if (p.tma_out) {
        rc = make_map(&map_y, gx, M, K, ldx, 32);
        if (rc) return rc;
    }
    const size_t smem = (size_t)stages * stage_bytes + p.epi_bufs * STAGING_BYTES + EPI_BIAS_BYTES + 1024;
    MPC_CUDA(ensure_smem_optin());
    const int total_tiles = p.m_tiles * p.n_tiles;
    const int grid = total_tiles < kNumSMs ? total_tiles : kNumSMs;
    MPC_CUDA(launch_pdl(grid, smem, (cudaStream_t)stream, map_a, map_b, map_y, p));
    return MPC_OK;
}
