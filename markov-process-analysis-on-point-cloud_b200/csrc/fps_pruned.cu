// Farthest point sampling with exact bucket pruning for sm_100a (mid-sized clouds: 1K .. 24K points, the states of
// the 24 000-point blocks).  Same contract as every other variant behind mpc_fps_f32 (include/mpc_b200.h; replaces
// R/modules/pointnet2_utils.py:84-109): bit-identical indices for the same start index.
//
// FPS is `npoint` strictly sequential rounds (update every point's running minimum distance against the newest
// sample, then a global arg-max).  The plain kernels (fps.cu) touch all N points every round; at 24 000 points that
// costs 8 points x ~14 instructions per thread per round on 8 CTAs plus a cluster barrier: ~1.0 us per round (0.93 us
// at 12 000 points on 4 CTAs).  This variant measures 0.77 us per round at both sizes on 2 / 1 CTAs (the round is a
// chain of ~25 dependent warp collectives and shared-memory round trips, ~0.3 us even when nothing is touched), so it
// is the default above 8192 points, where it is 20 % faster and leaves 3/4 of the SMs the cluster variant occupies to
// the rest of the forward pass.  But
// after a few hundred samples almost no running minimum changes any more: a point's minimum only drops if the new
// sample is closer than its current minimum.  So:
//   * the CTA first sorts its cloud into Morton order of a 16^3 grid (counting sort in shared memory) and cuts the
//     sorted sequence into buckets of 32 spatially adjacent points; lane j of a warp keeps the bounding box, the
//     largest running minimum (as the packed arg-max key) and that point's slot of the warp's j-th bucket;
//   * a round first tests every bucket: if the (downward-rounded) squared distance from the new sample to the
//     bucket's box exceeds the bucket's largest running minimum, NO point of the bucket can change -- the bucket is
//     skipped and its cached key stays valid.  Only touched buckets run the exact update (the contract's expression
//     ((dx*dx + dy*dy) + dz*dz), strict <) and refresh their key;
//   * the arg-max runs over cached bucket keys: two redux.sync per warp, one 32-byte slot per warp, one mbarrier
//     phase per round, every warp reduces the slots redundantly (no second barrier).  Keys are (distance bits,
//     ~original index), so ties resolve to the lowest ORIGINAL index exactly like torch.max / the plain kernels.
//   * clouds above 12 288 points run on a 2-CTA cluster: each CTA owns one half of the sorted sequence and every
//     warp writes its slot into both CTAs' shared memory (st.async + mbarrier complete_tx over DSMEM), so a round
//     costs one DSMEM hop instead of a cluster barrier.
// Skipping is conservative (it only ever skips provable no-ops), so the selected indices are those of the plain
// algorithm bit for bit; the order of points inside a grid cell (atomic arrival) only changes which points share a
// bucket, never the result.
#include <cooperative_groups.h>
#include <math.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace mpc {

constexpr int FP_G = 16;                       // Morton grid: cells per axis
constexpr int FP_CELLS = FP_G * FP_G * FP_G;   // 4096

struct __align__(16) FpSlot {
    unsigned long long key;
    float x, y, z;
    unsigned pad[3];
};
static_assert(sizeof(FpSlot) == 32, "slot layout");

__device__ __forceinline__ uint32_t fp_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ unsigned fp_spread4(unsigned v) {  // b3 b2 b1 b0 -> b3 0 0 b2 0 0 b1 0 0 b0
    return (v & 1u) | ((v & 2u) << 2) | ((v & 4u) << 4) | ((v & 8u) << 6);
}

__device__ __forceinline__ unsigned long long fp_warp_max_key(unsigned hi, unsigned lo, unsigned& khi, unsigned& klo) {
    khi = __reduce_max_sync(0xffffffffu, hi);
    klo = __reduce_max_sync(0xffffffffu, hi == khi ? lo : 0u);
    return ((unsigned long long)khi << 32) | klo;
}

// Lane holding the largest (hi, lo) key of the warp, and that key.  One redux when the largest high word is unique (the
// common case: distinct distances); the low words only break ties.
__device__ __forceinline__ int fp_warp_argmax(unsigned hi, unsigned lo, unsigned& khi, unsigned& klo) {
    khi = __reduce_max_sync(0xffffffffu, hi);
    unsigned m = __ballot_sync(0xffffffffu, hi == khi);
    if (m & (m - 1)) {  // several lanes share the largest distance: the larger low word (= lower index) wins
        klo = __reduce_max_sync(0xffffffffu, hi == khi ? lo : 0u);
        m = __ballot_sync(0xffffffffu, hi == khi && lo == klo);
    }
    const int src = __ffs(m) - 1;
    klo = __shfl_sync(0xffffffffu, lo, src);
    return src;
}

template <int THREADS, int BPW, int CS>
__global__ void __launch_bounds__(THREADS, 1)
fps_pruned_kernel(const float* __restrict__ xyz, const int64_t* __restrict__ start, int64_t* __restrict__ out, int N,
                  int npoint
#ifdef FP_DEBUG
                  , unsigned long long* dbg  // [0] touched buckets, [1..5] cycles of warp 0 per phase, [6] sort cycles
#endif
) {
    constexpr int NW = THREADS / 32;
    constexpr int SL = THREADS * BPW;  // slots per CTA
    static_assert(BPW <= 32, "one lane per bucket of a warp");
    extern __shared__ __align__(16) unsigned char raw[];
    float* sx = reinterpret_cast<float*>(raw);
    float* sy = sx + SL;
    float* sz = sy + SL;
    constexpr int MDR = SL > FP_CELLS + 4 ? SL : FP_CELLS + 4;  // running minima; the sort's histogram lives here first
    float* smd = sz + SL;
    int* hist = reinterpret_cast<int*>(smd);
    unsigned short* sid = reinterpret_cast<unsigned short*>(smd + MDR);
    FpSlot* slots = reinterpret_cast<FpSlot*>(sid + SL);
    uint64_t* bars = reinterpret_cast<uint64_t*>(slots + 2 * NW * CS);
    __shared__ float red[32][8];
    __shared__ int wtot[32];
    __shared__ int s_cb, s_sb;

    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    unsigned rank = 0;
    if (CS > 1) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    const int b = blockIdx.x / CS;
    const float* p = xyz + (size_t)b * N * 3;

    // ---- bounding box of the cloud (every CTA of the cluster computes the same values) ----
    float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int i = tid; i < N; i += THREADS) {
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            const float v = p[3 * i + d];
            mn[d] = fminf(mn[d], v);
            mx[d] = fmaxf(mx[d], v);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            mn[d] = fminf(mn[d], __shfl_xor_sync(0xffffffffu, mn[d], o));
            mx[d] = fmaxf(mx[d], __shfl_xor_sync(0xffffffffu, mx[d], o));
        }
    }
    if (lane == 0) {
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            red[w][d] = mn[d];
            red[w][3 + d] = mx[d];
        }
    }
    for (int c = tid; c <= FP_CELLS; c += THREADS) hist[c] = 0;
    if (tid == 0) {
        s_cb = -1;
        s_sb = 0;
    }
    __syncthreads();
    float inv[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        float lo = red[0][d], hi = red[0][3 + d];
        for (int k = 1; k < NW; ++k) {
            lo = fminf(lo, red[k][d]);
            hi = fmaxf(hi, red[k][3 + d]);
        }
        mn[d] = lo;
        const float ext = hi - lo;
        inv[d] = ext > 0.f ? (float)FP_G / ext : 0.f;
    }
    auto cell_of = [&](float x, float y, float z) -> int {
        const unsigned cx = (unsigned)fminf((x - mn[0]) * inv[0], (float)(FP_G - 1));
        const unsigned cy = (unsigned)fminf((y - mn[1]) * inv[1], (float)(FP_G - 1));
        const unsigned cz = (unsigned)fminf((z - mn[2]) * inv[2], (float)(FP_G - 1));
        return (int)(fp_spread4(cx) | (fp_spread4(cy) << 1) | (fp_spread4(cz) << 2));
    };

    // ---- counting sort by Morton cell: histogram, exclusive scan, fill ----
    for (int i = tid; i < N; i += THREADS) atomicAdd(&hist[cell_of(p[3 * i], p[3 * i + 1], p[3 * i + 2])], 1);
    __syncthreads();
    {
        constexpr int PER = FP_CELLS / THREADS;  // cells per thread (4096 / 512 = 8)
        int v[PER], tsum = 0;
#pragma unroll
        for (int u = 0; u < PER; ++u) {
            v[u] = hist[tid * PER + u];
            tsum += v[u];
        }
        int inc = tsum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        if (lane == 31) wtot[w] = inc;
        __syncthreads();
        if (w == 0) {
            int x = lane < NW ? wtot[lane] : 0;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, x, o);
                if (lane >= o) x += t;
            }
            wtot[lane] = x;
        }
        __syncthreads();
        int run = (w > 0 ? wtot[w - 1] : 0) + inc - tsum;
#pragma unroll
        for (int u = 0; u < PER; ++u) {
            hist[tid * PER + u] = run;
            run += v[u];
        }
        if (tid == THREADS - 1) hist[FP_CELLS] = run;  // = N
    }
    __syncthreads();
    // CTA 0 owns sorted positions [0, H), CTA 1 owns [H, N); H is a multiple of 32 (bucket aligned)
    const int H = CS == 1 ? N : min(SL, (((N + 1) / 2) + 31) & ~31);
    const int base = (int)rank * H;
    const int cnt = CS == 1 ? N : (rank == 0 ? min(H, N) : max(N - H, 0));
    if (CS > 1) {
        // the one cell that straddles H (if any) must be split identically by both CTAs: its points are ranked by
        // original index instead of atomic arrival
        constexpr int PER = FP_CELLS / THREADS;
#pragma unroll
        for (int u = 0; u < PER; ++u) {
            const int c = tid * PER + u;
            if (hist[c] < H && hist[c + 1] > H) {
                s_cb = c;
                s_sb = hist[c];
            }
        }
        __syncthreads();
    }
    const int cb = s_cb, sb = s_sb;
    int carry = 0;
    for (int i0 = 0; i0 < N; i0 += THREADS) {
        const int i = i0 + tid;
        const bool valid = i < N;
        float x = 0.f, y = 0.f, z = 0.f;
        int c = 0;
        if (valid) {
            x = p[3 * i];
            y = p[3 * i + 1];
            z = p[3 * i + 2];
            c = cell_of(x, y, z);
        }
        int pos = -1;
        if (CS > 1 && cb >= 0) {  // (uniform branch)
            const bool flag = valid && c == cb;
            const unsigned bal = __ballot_sync(0xffffffffu, flag);
            if (lane == 0) wtot[w] = __popc(bal);
            __syncthreads();
            int x2 = lane < NW ? wtot[lane] : 0;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, x2, o);
                if (lane >= o) x2 += t;
            }
            const int before = w > 0 ? __shfl_sync(0xffffffffu, x2, w - 1) : 0;
            const int total = __shfl_sync(0xffffffffu, x2, 31);
            if (flag) pos = sb + carry + before + __popc(bal & ((1u << lane) - 1u));
            carry += total;
            __syncthreads();
            if (valid && !flag) pos = atomicAdd(&hist[c], 1);
        } else if (valid) {
            pos = atomicAdd(&hist[c], 1);
        }
        const int local = pos - base;
        if (pos >= 0 && local >= 0 && local < cnt) {
            sx[local] = x;
            sy[local] = y;
            sz[local] = z;
            sid[local] = (unsigned short)i;
        }
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(fp_smem_u32(&bars[0])), "r"(NW));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(fp_smem_u32(&bars[1])), "r"(NW));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    // ---- running minima (shared memory, the histogram is dead now) and per-bucket state (lane j <-> bucket j*NW+w) ----
    for (int sl = tid; sl < SL; sl += THREADS) smd[sl] = sl < cnt ? 1e10f : -1.0f;  // padding can never win
    for (int e = tid; e < 2 * NW * CS; e += THREADS) {
        slots[e].key = 0ull;
        slots[e].x = slots[e].y = slots[e].z = 0.f;
    }
    float blo[3] = {INFINITY, INFINITY, INFINITY}, bhi[3] = {-INFINITY, -INFINITY, -INFINITY};
    float bmax = -1.0f;            // largest running minimum of my bucket (-1: bucket holds no point)
    unsigned bkhi = 0u, bklo = 0u;  // its packed key
    int bslot = 0;                  // and the slot of the point that holds it
#pragma unroll 1
    for (int j = 0; j < BPW; ++j) {
        const int slot = (j * NW + w) * 32 + lane;
        const bool ok = slot < cnt;
        float lo3[3], hi3[3];
        lo3[0] = ok ? sx[slot] : INFINITY;
        lo3[1] = ok ? sy[slot] : INFINITY;
        lo3[2] = ok ? sz[slot] : INFINITY;
#pragma unroll
        for (int d = 0; d < 3; ++d) hi3[d] = ok ? lo3[d] : -INFINITY;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                lo3[d] = fminf(lo3[d], __shfl_xor_sync(0xffffffffu, lo3[d], o));
                hi3[d] = fmaxf(hi3[d], __shfl_xor_sync(0xffffffffu, hi3[d], o));
            }
        }
        const bool any = (j * NW + w) * 32 < cnt;
        if (lane == j) {
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                blo[d] = lo3[d];
                bhi[d] = hi3[d];
            }
            bmax = any ? 1e10f : -1.0f;
        }
    }
    __syncthreads();
    uint32_t peer_slots = 0, peer_bars = 0;
    if (CS > 1) {
        const unsigned peer = rank ^ 1u;
        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(peer_slots) : "r"(fp_smem_u32(slots)), "r"(peer));
        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(peer_bars) : "r"(fp_smem_u32(bars)), "r"(peer));
        cg::this_cluster().sync();  // both CTAs resident, barriers initialised, before any remote write
    }

    int far = clamp_index(start[b], N);
    float cx = p[(size_t)far * 3], cy = p[(size_t)far * 3 + 1], cz = p[(size_t)far * 3 + 2];
    int64_t* o = out + (size_t)b * npoint;
#ifdef FP_DEBUG
    long long t0 = clock64();
    unsigned long long acc[6] = {0, 0, 0, 0, 0, 0};
#define FP_MARK(i) { long long t1 = clock64(); acc[i] += (unsigned long long)(t1 - t0); t0 = t1; \
    if (dbg && it >= 1000 && it < 1016 && lane == 0 && blockIdx.x == 0) dbg[64 + ((it - 1000) * NW + w) * 8 + i] = (unsigned long long)t1; }
#else
#define FP_MARK(i)
#endif
    for (int it = 0; it < npoint; ++it) {
        // (1) which of this warp's buckets can change at all?
        bool touched = false;
        if (lane < BPW) {
            const float dx = fmaxf(fmaxf(blo[0] - cx, cx - bhi[0]), 0.f);
            const float dy = fmaxf(fmaxf(blo[1] - cy, cy - bhi[1]), 0.f);
            const float dz = fmaxf(fmaxf(blo[2] - cz, cz - bhi[2]), 0.f);
            const float lb = dx * dx + dy * dy + dz * dz;
            // the exact per-point distance is >= lb * (1 - 10 * 2^-24); skip only when even that exceeds the
            // bucket's largest running minimum (an empty bucket has lb = inf or bmax = -1: never touched)
            touched = !(lb * 0.999999f > bmax);
        }
#if defined(FP_EXPERIMENT)
        const unsigned mask = FP_EXPERIMENT == 1 ? 0u : (__ballot_sync(0xffffffffu, touched) & (w == (it & (NW - 1)) ? 1u : 0u));
#else
        const unsigned mask = __ballot_sync(0xffffffffu, touched);
#endif
#ifdef FP_DEBUG
        acc[0] += __popc(mask);
#endif
        FP_MARK(1)
        const int buf = it & 1;
        unsigned long long wkey;
        float wx, wy, wz;
        int pub;  // the lane that publishes this warp's slot
        // candidate of this lane so far: the cached key of its own bucket if that bucket cannot change this round
        unsigned chi = 0u, clo = 0u;
        int cslot = 0;
        if (lane < BPW && !((mask >> lane) & 1u)) {
            chi = bkhi;
            clo = bklo;
            cslot = bslot;
        }
        if (mask) {
            // (2) exact update of the touched buckets.  No warp collective in here: every lane folds its own points
            // into its candidate, two buckets per iteration so that their loads overlap; the buckets' cached keys are
            // refreshed after the warp has published (off the round's critical path)
            unsigned m2 = mask;
            do {
                const int j0 = __ffs(m2) - 1;
                m2 &= m2 - 1;
                const bool two = m2 != 0;
                const int j1 = two ? __ffs(m2) - 1 : j0;
                m2 &= m2 - 1;
                const int s0 = (j0 * NW + w) * 32 + lane, s1 = (j1 * NW + w) * 32 + lane;
                const float x0 = sx[s0], y0 = sy[s0], z0 = sz[s0], x1 = sx[s1], y1 = sy[s1], z1 = sz[s1];
                float m0 = smd[s0], m1 = smd[s1];
                const unsigned i0 = sid[s0], i1 = sid[s1];
                {
                    const float dx = __fsub_rn(x0, cx), dy = __fsub_rn(y0, cy), dz = __fsub_rn(z0, cz);
                    const float d = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
                    if (d < m0) {
                        m0 = d;
                        smd[s0] = d;
                    }
                    const bool real = m0 >= 0.0f;  // padding slots hold -1 (and arbitrary coordinates)
                    const unsigned hi = real ? __float_as_uint(m0) : 0u, lo = real ? ~i0 : 0u;
                    if (hi > chi || (hi == chi && lo > clo)) {
                        chi = hi;
                        clo = lo;
                        cslot = s0;
                    }
                }
                if (two) {  // (warp-uniform)
                    const float dx = __fsub_rn(x1, cx), dy = __fsub_rn(y1, cy), dz = __fsub_rn(z1, cz);
                    const float d = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
                    if (d < m1) {
                        m1 = d;
                        smd[s1] = d;
                    }
                    const bool real = m1 >= 0.0f;
                    const unsigned hi = real ? __float_as_uint(m1) : 0u, lo = real ? ~i1 : 0u;
                    if (hi > chi || (hi == chi && lo > clo)) {
                        chi = hi;
                        clo = lo;
                        cslot = s1;
                    }
                }
            } while (m2);
            // (3) the warp's winner over the lanes' candidates; the lane that owns it publishes
            unsigned khi, klo;
            pub = fp_warp_argmax(chi, clo, khi, klo);
            wkey = ((unsigned long long)khi << 32) | klo;
            wx = sx[cslot];
            wy = sy[cslot];
            wz = sz[cslot];
        } else {
            // nothing changed in this warp: its slot of the previous round is still right (round 0: the zero key)
            pub = 0;
            const FpSlot prev = slots[((buf ^ 1) * CS + (int)rank) * NW + w];
            wkey = prev.key;
            wx = prev.x;
            wy = prev.y;
            wz = prev.z;
        }
        FP_MARK(2)
        // (4) publish the warp's slot (into both CTAs of a cluster), one mbarrier phase per round
        if (lane == pub) {
            FpSlot* mine = &slots[(buf * CS + (int)rank) * NW + w];
            mine->key = wkey;
            mine->x = wx;
            mine->y = wy;
            mine->z = wz;
            const uint32_t bar = fp_smem_u32(&bars[buf]);
            if (CS > 1) {
                const uint32_t raddr = peer_slots + (uint32_t)(((buf * CS + (int)rank) * NW + w) * sizeof(FpSlot));
                const uint32_t rbar = peer_bars + (uint32_t)(buf * sizeof(uint64_t));
                asm volatile(
                    "st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];" ::
                        "r"(raddr), "r"((unsigned)wkey), "r"((unsigned)(wkey >> 32)), "r"(__float_as_uint(wx)),
                    "r"(__float_as_uint(wy)), "r"(rbar)
                    : "memory");
                asm volatile(
                    "st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.b32 [%0], {%1, %2}, [%3];" ::"r"(
                        raddr + 16u),
                    "r"(__float_as_uint(wz)), "r"(0u), "r"(rbar)
                    : "memory");
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(24u) : "memory");
            } else {
#ifndef FP_BARSYNC
                asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
#endif
            }
        }
        if (rank == 0 && tid == THREADS - 1) o[it] = far;  // (after publishing: the store's address arithmetic is off the chain)
        // (4b) refresh the cached keys of the touched buckets (needed by the NEXT round's test and candidates): the
        // latency of these collectives overlaps with waiting for the other warps
        if (mask) {
            unsigned m2 = mask;
            do {
                const int j = __ffs(m2) - 1;
                m2 &= m2 - 1;
                const int slot = (j * NW + w) * 32 + lane;
                const float m = smd[slot];
                const bool real = m >= 0.0f;
                const unsigned hi = real ? __float_as_uint(m) : 0u;
                const unsigned lo = real ? ~(unsigned)sid[slot] : 0u;
                unsigned khi, klo;
                const int src = fp_warp_argmax(hi, lo, khi, klo);
                if (lane == j) {
                    bkhi = khi;
                    bklo = klo;
                    bslot = (j * NW + w) * 32 + src;
                    bmax = (khi | klo) ? __uint_as_float(khi) : -1.0f;
                }
            } while (m2);
        }
        FP_MARK(3)
#ifdef FP_BARSYNC
        if (CS == 1) __syncthreads(); else
#endif
        {
            const uint32_t bar = fp_smem_u32(&bars[buf]);
            const uint32_t parity = (uint32_t)(it >> 1) & 1u;
            uint32_t done = 0;
            while (!done) {
                asm volatile(
                    "{\n\t.reg .pred p;\n\t"
                    "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                    "selp.u32 %0, 1, 0, p;\n\t}"
                    : "=r"(done)
                    : "r"(bar), "r"(parity)
                    : "memory");
            }
        }
        FP_MARK(4)
        // (5) every warp reduces all slots: the round's winner and its coordinates
        const FpSlot* sl = &slots[buf * CS * NW];
        unsigned long long k = 0ull;
        float ex = 0.f, ey = 0.f, ez = 0.f;
#pragma unroll
        for (int r = 0; r < (CS * NW + 31) / 32; ++r) {
            const int e = r * 32 + lane;
            if (e < CS * NW) {
                const FpSlot s = sl[e];
                if (s.key > k) {
                    k = s.key;
                    ex = s.x;
                    ey = s.y;
                    ez = s.z;
                }
            }
        }
        unsigned khi, klo;
        const int owner = fp_warp_argmax((unsigned)(k >> 32), (unsigned)k, khi, klo);
        far = (int)(~klo);
        cx = __shfl_sync(0xffffffffu, ex, owner);
        cy = __shfl_sync(0xffffffffu, ey, owner);
        cz = __shfl_sync(0xffffffffu, ez, owner);
        FP_MARK(5)
    }
#ifdef FP_DEBUG
    if (dbg && lane == 0) {
        atomicAdd(&dbg[0], acc[0]);
        if (w == 0 && blockIdx.x == 0)
            for (int i = 1; i < 6; ++i) dbg[i] = acc[i];
    }
#endif
    if (CS > 1) cg::this_cluster().sync();  // no CTA exits while its peer may still write into its shared memory
}

template <int THREADS, int BPW, int CS>
static int launch_pruned(const float* xyz, const int64_t* start, int64_t* out, int B, int N, int npoint,
                         cudaStream_t st) {
    constexpr int SL = THREADS * BPW;
    constexpr size_t smem = (size_t)SL * 14 + (size_t)(SL > FP_CELLS + 4 ? SL : FP_CELLS + 4) * 4 +
                            2 * (THREADS / 32) * CS * sizeof(FpSlot) + 64;
    auto kern = fps_pruned_kernel<THREADS, BPW, CS>;
    MPC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
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
#ifdef FP_DEBUG
    extern unsigned long long* g_fp_dbg;
    MPC_CUDA(cudaLaunchKernelEx(&cfg, kern, xyz, start, out, N, npoint, g_fp_dbg));
#else
    MPC_CUDA(cudaLaunchKernelEx(&cfg, kern, xyz, start, out, N, npoint));
#endif
    return MPC_OK;
}

// Entry for mpc_fps_f32's dispatcher (fps.cu): returns MPC_ERR_UNSUPPORTED when N is outside this variant's reach.
int fps_pruned_dispatch(const float* xyz, const int64_t* start, int64_t* out, int B, int N, int npoint,
                        cudaStream_t st) {
    if (N <= 3072) return launch_pruned<256, 12, 1>(xyz, start, out, B, N, npoint, st);
    if (N <= 6144) return launch_pruned<256, 24, 1>(xyz, start, out, B, N, npoint, st);  // (plain kernels are as fast here)
    if (N <= 12288) return launch_pruned<512, 24, 1>(xyz, start, out, B, N, npoint, st);
    if (N <= 24576) return launch_pruned<512, 24, 2>(xyz, start, out, B, N, npoint, st);
    return MPC_ERR_UNSUPPORTED;
}

}  // namespace mpc
