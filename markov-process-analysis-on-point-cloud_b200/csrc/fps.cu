// Farthest point sampling for sm_100a.  See include/mpc_b200.h (mpc_fps_f32) for the contract and the
// reference lines it replaces (pointnet2_utils.py:84-109).
//
// FPS is a latency chain: `npoint` strictly sequential (distance update, global argmax) rounds with
// essentially no HBM traffic.  Design:
//   * one persistent CTA per cloud; every thread keeps PPT points (x, y, z, running min distance) in
//     REGISTERS for the whole kernel, so a round touches no memory except one broadcast shared-memory read of
//     the winner's coordinates;
//   * argmax = packed 64-bit key (distance bits << 32 | ~index): two redux.sync per warp, one
//     __syncthreads per round (double-buffered warp slots), lowest index wins ties like torch.max;
//   * clouds too large for one CTA's registers use one thread-block CLUSTER per cloud: each CTA owns a
//     slice, publishes its local winner (key + coordinates) into every peer's shared memory through DSMEM and
//     one cluster barrier per round decides the global winner.
#include <cooperative_groups.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace mpc {

__device__ __forceinline__ unsigned long long pack_key(float d, int idx) {
    return ((unsigned long long)__float_as_uint(d) << 32) | (unsigned)(~idx);
}

// max over the warp of a 64-bit key whose high word is a non-negative float's bits.
__device__ __forceinline__ unsigned long long warp_max_key(unsigned hi, unsigned lo) {
    unsigned mhi = __reduce_max_sync(0xffffffffu, hi);
    unsigned mlo = __reduce_max_sync(0xffffffffu, hi == mhi ? lo : 0u);
    return ((unsigned long long)mhi << 32) | mlo;
}

// ---- per-thread slice update: returns the local best (distance bits, ~index) ---------------------------
template <int PPT>
__device__ __forceinline__ void update_slice(const float (&px)[PPT], const float (&py)[PPT],
                                             const float (&pz)[PPT], float (&md)[PPT], float cx, float cy,
                                             float cz, int base, int stride, unsigned& bhi, unsigned& blo) {
    float best = -1.0f;
    int besti = 0;
#pragma unroll
    for (int i = 0; i < PPT; ++i) {
        float dx = __fsub_rn(px[i], cx), dy = __fsub_rn(py[i], cy), dz = __fsub_rn(pz[i], cz);
        float d = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
        float m = md[i];
        m = d < m ? d : m;
        md[i] = m;
        if (m > best) {  // ascending index within the thread: strict > keeps the lowest
            best = m;
            besti = base + i * stride;
        }
    }
    // a thread that owns only padding must lose against every real candidate: key 0 (real keys have a
    // non-zero low word); note -1.0f reinterpreted as unsigned would compare ABOVE every real distance
    const bool any = best >= 0.0f;
    bhi = any ? __float_as_uint(best) : 0u;
    blo = any ? (unsigned)(~besti) : 0u;
}

// ---- variant A: one CTA per cloud, C == 3 ----------------------------------------------------------------
template <int PPT, int THREADS>
__global__ void __launch_bounds__(THREADS, 1)
fps_cta_kernel(const float* __restrict__ xyz, const int64_t* __restrict__ start, int64_t* __restrict__ out,
               int N, int npoint) {
    extern __shared__ float smem[];  // sx[N] sy[N] sz[N]
    float* sx = smem;
    float* sy = sx + N;
    float* sz = sy + N;
    __shared__ unsigned long long slots[2][32];
    constexpr int NW = THREADS / 32;

    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float* p = xyz + (size_t)b * N * 3;
    for (int i = tid; i < N * 3; i += THREADS) {  // coalesced AoS read, SoA in shared memory
        int n = i / 3, c = i - n * 3;
        smem[c * N + n] = p[i];
    }
    __syncthreads();
    float px[PPT], py[PPT], pz[PPT], md[PPT];
#pragma unroll
    for (int i = 0; i < PPT; ++i) {
        int n = tid + i * THREADS;
        bool ok = n < N;
        px[i] = ok ? sx[n] : 0.f;
        py[i] = ok ? sy[n] : 0.f;
        pz[i] = ok ? sz[n] : 0.f;
        md[i] = ok ? 1e10f : -1.0f;  // padding can never win (real distances are >= 0)
    }
    int far = clamp_index(start[b], N);
    int64_t* o = out + (size_t)b * npoint;
    for (int it = 0; it < npoint; ++it) {
        if (tid == 0) o[it] = far;
        float cx = sx[far], cy = sy[far], cz = sz[far];
        unsigned bhi, blo;
        update_slice<PPT>(px, py, pz, md, cx, cy, cz, tid, THREADS, bhi, blo);
        unsigned long long k = warp_max_key(bhi, blo);
        if (lane == 0) slots[it & 1][warp] = k;
        __syncthreads();
        unsigned long long v = lane < NW ? slots[it & 1][lane] : 0ull;
        k = warp_max_key((unsigned)(v >> 32), (unsigned)v);
        far = (int)(~(unsigned)k);
    }
}

// ---- variant C: one cluster per cloud, C == 3 --------------------------------------------------------------
struct __align__(16) Cand {
    unsigned long long key;
    float x, y, z;
    float pad;
};

template <int PPT, int THREADS>
__global__ void __launch_bounds__(THREADS, 1)
fps_cluster_kernel(const float* __restrict__ xyz, const int64_t* __restrict__ start,
                   int64_t* __restrict__ out, int N, int npoint) {
    cg::cluster_group cluster = cg::this_cluster();
    const int CS = cluster.num_blocks();
    const int rank = cluster.block_rank();
    extern __shared__ float smem[];  // this CTA's slice, SoA: sx[SL] sy[SL] sz[SL]
    __shared__ unsigned long long slots[2][32];
    __shared__ Cand cand[2][16];
    constexpr int NW = THREADS / 32;
    constexpr int SL = PPT * THREADS;  // slice capacity

    const int b = blockIdx.x / CS, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int base = rank * SL;  // first global point index of this slice
    float* sx = smem;
    float* sy = sx + SL;
    float* sz = sy + SL;
    const float* p = xyz + (size_t)b * N * 3;
    const int cntp = max(0, min(SL, N - base));
    for (int i = tid; i < cntp * 3; i += THREADS) {
        int n = i / 3, c = i - n * 3;
        smem[c * SL + n] = p[(size_t)base * 3 + i];
    }
    __syncthreads();
    float px[PPT], py[PPT], pz[PPT], md[PPT];
#pragma unroll
    for (int i = 0; i < PPT; ++i) {
        int n = tid + i * THREADS;
        bool ok = n < cntp;
        px[i] = ok ? sx[n] : 0.f;
        py[i] = ok ? sy[n] : 0.f;
        pz[i] = ok ? sz[n] : 0.f;
        md[i] = ok ? 1e10f : -1.0f;
    }
    int far = clamp_index(start[b], N);
    float cx = p[(size_t)far * 3 + 0], cy = p[(size_t)far * 3 + 1], cz = p[(size_t)far * 3 + 2];
    int64_t* o = out + (size_t)b * npoint;
    cluster.sync();  // every CTA of the cluster is resident before any DSMEM write
    for (int it = 0; it < npoint; ++it) {
        if (rank == 0 && tid == 0) o[it] = far;
        unsigned bhi, blo;
        update_slice<PPT>(px, py, pz, md, cx, cy, cz, base + tid, THREADS, bhi, blo);
        unsigned long long k = warp_max_key(bhi, blo);
        if (lane == 0) slots[it & 1][warp] = k;
        __syncthreads();
        if (warp == 0) {
            unsigned long long v = lane < NW ? slots[it & 1][lane] : 0ull;
            k = warp_max_key((unsigned)(v >> 32), (unsigned)v);
            if (lane < CS) {  // lane r publishes this CTA's winner into CTA r's shared memory
                int li = (int)(~(unsigned)k) - base;
                li = li < 0 ? 0 : (li >= SL ? SL - 1 : li);  // a slice of pure padding publishes key < any real one
                Cand c;
                c.key = k;
                c.x = sx[li];
                c.y = sy[li];
                c.z = sz[li];
                c.pad = 0.f;
                Cand* remote = cluster.map_shared_rank(&cand[it & 1][rank], lane);
                *remote = c;
            }
        }
        cluster.sync();  // release/acquire: all candidates of this round are visible
        Cand c = cand[it & 1][lane < CS ? lane : 0];
        unsigned long long kk = lane < CS ? c.key : 0ull;
        unsigned long long kmax = warp_max_key((unsigned)(kk >> 32), (unsigned)kk);
        unsigned src = __ffs(__ballot_sync(0xffffffffu, kk == kmax && lane < CS)) - 1;
        far = (int)(~(unsigned)kmax);
        cx = __shfl_sync(0xffffffffu, c.x, src);
        cy = __shfl_sync(0xffffffffu, c.y, src);
        cz = __shfl_sync(0xffffffffu, c.z, src);
    }
    cluster.sync();  // no CTA exits while a peer may still write into its shared memory
}

// ---- variant D: one cloud across the whole GPU (N > 262144, up to 148 x 8192 points), C == 3 ---------------
// The running minimum distances of a 1M-point cloud (16 MB with the coordinates) only fit in the register files of
// ~128 SMs together.  One cooperative launch per cloud: CTA g keeps points [g*8192, (g+1)*8192) in registers, publishes
// its local winner (key + coordinates) into a double-buffered global slot array and a grid-wide barrier (one
// monotonically increasing counter, release/acquire) decides the round: ~3 us per round, HBM traffic still ~0.
// Barrier words and candidate slots are library-global: one such launch at a time per device.
constexpr int GRID_MAX = kNumSMs;
__device__ unsigned g_fps_barrier;
__device__ Cand g_fps_cand[2][GRID_MAX];

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

template <int PPT, int THREADS>
__global__ void __launch_bounds__(THREADS, 1)
fps_grid_kernel(const float* __restrict__ p, const int64_t* __restrict__ start, int64_t* __restrict__ o, int N,
                int npoint) {
    extern __shared__ float smem[];  // this CTA's slice, SoA
    __shared__ unsigned long long slots[2][32];
    __shared__ float win[4];
    constexpr int NW = THREADS / 32;
    constexpr int SL = PPT * THREADS;
    const int G = gridDim.x, rank = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int base = rank * SL;
    float* sx = smem;
    float* sy = sx + SL;
    float* sz = sy + SL;
    const int cntp = max(0, min(SL, N - base));
    for (int i = tid; i < cntp * 3; i += THREADS) {
        int n = i / 3, c = i - n * 3;
        smem[c * SL + n] = p[(size_t)base * 3 + i];
    }
    __syncthreads();
    float px[PPT], py[PPT], pz[PPT], md[PPT];
#pragma unroll
    for (int i = 0; i < PPT; ++i) {
        int n = tid + i * THREADS;
        bool ok = n < cntp;
        px[i] = ok ? sx[n] : 0.f;
        py[i] = ok ? sy[n] : 0.f;
        pz[i] = ok ? sz[n] : 0.f;
        md[i] = ok ? 1e10f : -1.0f;
    }
    int far = clamp_index(start[0], N);
    float cx = p[(size_t)far * 3 + 0], cy = p[(size_t)far * 3 + 1], cz = p[(size_t)far * 3 + 2];
    for (int it = 0; it < npoint; ++it) {
        if (rank == 0 && tid == 0) o[it] = far;
        unsigned bhi, blo;
        update_slice<PPT>(px, py, pz, md, cx, cy, cz, base + tid, THREADS, bhi, blo);
        unsigned long long k = warp_max_key(bhi, blo);
        if (lane == 0) slots[it & 1][warp] = k;
        __syncthreads();
        if (warp == 0) {
            unsigned long long v = lane < NW ? slots[it & 1][lane] : 0ull;
            k = warp_max_key((unsigned)(v >> 32), (unsigned)v);
            if (lane == 0) {  // publish this CTA's winner, then arrive at the grid barrier of this round
                int li = (int)(~(unsigned)k) - base;
                li = li < 0 ? 0 : (li >= SL ? SL - 1 : li);
                Cand c;
                c.key = k;
                c.x = sx[li];
                c.y = sy[li];
                c.z = sz[li];
                c.pad = 0.f;
                g_fps_cand[it & 1][rank] = c;
                __threadfence();
                atomicAdd(&g_fps_barrier, 1u);
                const unsigned target = (unsigned)(it + 1) * (unsigned)G;
                while (ld_acquire_u32(&g_fps_barrier) < target) {
                }
            }
            __syncwarp();
            // every candidate of this round is visible: lane l reduces candidates l, l + 32, ...
            unsigned long long best = 0ull;
            int bsrc = 0;
            for (int r = lane; r < G; r += 32) {
                const unsigned long long kk = __ldcg(&g_fps_cand[it & 1][r].key);
                if (kk > best) {
                    best = kk;
                    bsrc = r;
                }
            }
            const unsigned long long kmax = warp_max_key((unsigned)(best >> 32), (unsigned)best);
            const unsigned src_lane = __ffs(__ballot_sync(0xffffffffu, best == kmax)) - 1;
            if (lane == src_lane) {
                const Cand* c = &g_fps_cand[it & 1][bsrc];
                win[0] = __ldcg(&c->x);
                win[1] = __ldcg(&c->y);
                win[2] = __ldcg(&c->z);
                win[3] = __int_as_float((int)(~(unsigned)kmax));
            }
        }
        __syncthreads();
        cx = win[0];
        cy = win[1];
        cz = win[2];
        far = __float_as_int(win[3]);
    }
    // leave the barrier word zero for the next launch: the last CTA through the final barrier resets it
    __syncthreads();
    if (tid == 0) {
        const unsigned done = atomicAdd(&g_fps_barrier, 1u) + 1u;
        if (done == (unsigned)(npoint + 1) * (unsigned)G) g_fps_barrier = 0u;
    }
}

// ---- variant B: generic channel count (feature-space FPS), one CTA per cloud -----------------------------
__global__ void __launch_bounds__(512, 1)
fps_generic_kernel(const float* __restrict__ xyz, const int64_t* __restrict__ start,
                   int64_t* __restrict__ out, int N, int C, int npoint) {
    extern __shared__ float smem[];  // md[N] then centroid[C]
    float* md = smem;
    float* cen = smem + N;
    __shared__ unsigned long long slots[2][32];
    constexpr int THREADS = 512, NW = THREADS / 32;
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float* p = xyz + (size_t)b * N * C;
    for (int n = tid; n < N; n += THREADS) md[n] = 1e10f;
    int far = clamp_index(start[b], N);
    int64_t* o = out + (size_t)b * npoint;
    for (int it = 0; it < npoint; ++it) {
        if (tid == 0) o[it] = far;
        __syncthreads();  // previous round's readers of cen[] are done
        for (int c = tid; c < C; c += THREADS) cen[c] = p[(size_t)far * C + c];
        __syncthreads();
        float best = -1.0f;
        int besti = 0;
        for (int n = tid; n < N; n += THREADS) {
            const float* q = p + (size_t)n * C;
            float d0 = __fsub_rn(q[0], cen[0]);
            float acc = __fmul_rn(d0, d0);
            for (int c = 1; c < C; ++c) {
                float dc = __fsub_rn(q[c], cen[c]);
                acc = __fadd_rn(acc, __fmul_rn(dc, dc));
            }
            float m = md[n];
            m = acc < m ? acc : m;
            md[n] = m;
            if (m > best) {
                best = m;
                besti = n;
            }
        }
        const bool any = best >= 0.0f;  // threads beyond N own no point
        unsigned long long k = warp_max_key(any ? __float_as_uint(best) : 0u, any ? (unsigned)(~besti) : 0u);
        if (lane == 0) slots[it & 1][warp] = k;
        __syncthreads();
        unsigned long long v = lane < NW ? slots[it & 1][lane] : 0ull;
        k = warp_max_key((unsigned)(v >> 32), (unsigned)v);
        far = (int)(~(unsigned)k);
    }
}

// FPS is a latency chain: every CTA wants an SM's issue slots to itself.  Small CTAs (128-512 threads) launched
// while other kernels hold most SMs get STACKED several to an SM by the block scheduler (measured: 3.3x slower
// rounds inside the multi-stream training step).  Asking for enough dynamic shared memory that only
// ceil(ctas / SMs) of them fit on one SM spreads the clouds over distinct SMs.
static size_t spread_smem(size_t needed, int ctas) {
    const int per_sm = (ctas + kNumSMs - 1) / kNumSMs;
    if (per_sm >= 8) return needed;
    const size_t want = (size_t)(226 * 1024) / (per_sm + 1) + 1024;
    return needed > want ? needed : want;
}

template <int PPT, int THREADS>
static int launch_cta(const float* xyz, const int64_t* start, int64_t* out, int B, int N, int npoint,
                      cudaStream_t st) {
    size_t smem = spread_smem((size_t)N * 3 * sizeof(float), B);
    auto kern = fps_cta_kernel<PPT, THREADS>;
    if (smem > 40 * 1024)  // static shared memory (warp slots) counts against the default 48 KB too
        MPC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<B, THREADS, smem, st>>>(xyz, start, out, N, npoint);
    MPC_LAUNCH_CHECK();
    return MPC_OK;
}

template <int PPT, int THREADS>
static int launch_cluster(const float* xyz, const int64_t* start, int64_t* out, int B, int N, int npoint,
                          int CS, cudaStream_t st) {
    size_t smem = spread_smem((size_t)PPT * THREADS * 3 * sizeof(float), B * CS);
    auto kern = fps_cluster_kernel<PPT, THREADS>;
    MPC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (CS > 8) MPC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(B * CS);
    cfg.blockDim = dim3(THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CS;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    MPC_CUDA(cudaLaunchKernelEx(&cfg, kern, xyz, start, out, N, npoint));
    return MPC_OK;
}

template <int PPT, int THREADS>
static int launch_grid(const float* xyz, const int64_t* start, int64_t* out, int B, int N, int npoint, cudaStream_t st) {
    const int SL = PPT * THREADS;
    const int G = (N + SL - 1) / SL;
    if (G > GRID_MAX) return MPC_ERR_UNSUPPORTED;
    if ((long long)(npoint + 1) * G >= 0xffffffffll) return MPC_ERR_UNSUPPORTED;  // barrier counter range
    size_t smem = spread_smem((size_t)SL * 3 * sizeof(float), G);
    auto kern = fps_grid_kernel<PPT, THREADS>;
    MPC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    for (int b = 0; b < B; ++b) {  // one cloud at a time: the grid barrier words are shared
        const float* p = xyz + (size_t)b * N * 3;
        const int64_t* s0 = start + b;
        int64_t* o = out + (size_t)b * npoint;
        void* args[] = {(void*)&p, (void*)&s0, (void*)&o, (void*)&N, (void*)&npoint};
        MPC_CUDA(cudaLaunchCooperativeKernel((const void*)kern, dim3(G), dim3(THREADS), args, smem, st));
    }
    return MPC_OK;
}

// fps_pruned.cu: exact bucket-pruned variant for mid-sized clouds
int fps_pruned_dispatch(const float* xyz, const int64_t* start, int64_t* out, int B, int N, int npoint,
                        cudaStream_t st);

}  // namespace mpc

MPC_API int mpc_fps_f32(const float* xyz, const int64_t* start, int64_t* out, int64_t B, int64_t N, int64_t C,
                        int64_t npoint, mpc_stream_t stream) {
    using namespace mpc;
    if (B < 0 || N <= 0 || C <= 0 || npoint < 0) return MPC_ERR_INVALID;
    if (B == 0 || npoint == 0) return MPC_OK;  // nothing to do: empty tensors carry null pointers
    if (!xyz || !start || !out) return MPC_ERR_INVALID;
    cudaStream_t st = (cudaStream_t)stream;
    if (C != 3) {
        size_t smem = (size_t)(N + C) * sizeof(float);
        if (smem > 200 * 1024) return MPC_ERR_UNSUPPORTED;
        if (smem > 40 * 1024)
            MPC_CUDA(cudaFuncSetAttribute(fps_generic_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        fps_generic_kernel<<<(int)B, 512, smem, st>>>(xyz, start, out, (int)N, (int)C, (int)npoint);
        MPC_LAUNCH_CHECK();
        return MPC_OK;
    }
    const int b = (int)B, n = (int)N, np = (int)npoint;
    if (n > 8192 && g_knob[3] > 0) return launch_grid<8, 1024>(xyz, start, out, b, n, np, st);  // debug: force variant D
    // mid-sized clouds: exact bucket pruning (fps_pruned.cu).  knob 6: 0 = default threshold, < 0 = plain kernels
    // only (A/B runs and the tests that compare the two), > 0 = smallest N that takes the pruned variant
    {
        const int64_t min_n = g_knob[6] == 0 ? 8193 : g_knob[6];
        if (min_n > 0 && n >= min_n && n <= 24576 && np > 1) return fps_pruned_dispatch(xyz, start, out, b, n, np, st);
    }
    // single CTA: THREADS * PPT >= N, xyz copy N*12 bytes of shared memory (<= 192 KB at N = 16384)
    if (n <= 128) return launch_cta<1, 128>(xyz, start, out, b, n, np, st);
    if (n <= 256) return launch_cta<2, 128>(xyz, start, out, b, n, np, st);
    if (n <= 512) return launch_cta<4, 128>(xyz, start, out, b, n, np, st);
    if (n <= 1024) return launch_cta<4, 256>(xyz, start, out, b, n, np, st);
    if (n <= 2048) return launch_cta<8, 256>(xyz, start, out, b, n, np, st);
    if (n <= 4096) return launch_cta<8, 512>(xyz, start, out, b, n, np, st);
    if (n <= 8192) return launch_cta<8, 1024>(xyz, start, out, b, n, np, st);
    // cluster: a round costs (points per thread) x ~12 instructions per warp plus one cluster barrier, so mid-sized
    // clouds are spread thinly (512 threads x 8 points per CTA) over up to 16 CTAs; larger ones use fatter CTAs
    if (n <= 2 * 4096) return launch_cluster<8, 512>(xyz, start, out, b, n, np, 2, st);
    if (n <= 4 * 4096) return launch_cluster<8, 512>(xyz, start, out, b, n, np, 4, st);
    if (n <= 8 * 4096) return launch_cluster<8, 512>(xyz, start, out, b, n, np, 8, st);
    if (n <= 16 * 4096) return launch_cluster<8, 512>(xyz, start, out, b, n, np, 16, st);
    if (n <= 16 * 8192) return launch_cluster<8, 1024>(xyz, start, out, b, n, np, 16, st);
    if (n <= 16 * 16384) return launch_cluster<16, 1024>(xyz, start, out, b, n, np, 16, st);
    if (n <= GRID_MAX * 8192) return launch_grid<8, 1024>(xyz, start, out, b, n, np, st);  // up to 1 212 416 points
    return MPC_ERR_UNSUPPORTED;
}
