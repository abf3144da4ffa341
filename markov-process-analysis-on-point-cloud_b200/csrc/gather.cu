// Row gather / scatter family for sm_100a: index_points forward/backward, the sparse Markov transition
// (`upsample`) forward/backward and three_interpolate forward/backward.  See include/mpc_b200.h for the
// contracts and the reference lines replaced (pointnet2_utils.py:13-50, 64-81, 903-906).
//
// All of these are HBM/L2-bound row movers: rows are moved with 128-bit accesses (one float4 lane per 4
// channels, consecutive lanes on consecutive 16-byte pieces of a row), reductions into rows use
// red.global.add.v4.f32, and grids are sized from the element count so every SM has several CTAs in flight.
#include "common.cuh"

namespace mpc {

constexpr int GT = 256;  // threads per CTA for the row movers

static inline unsigned grid_for(int64_t work_items) {
    int64_t g = ceil_div(work_items, GT);
    const int64_t cap = (int64_t)kNumSMs * 32;  // grid-stride beyond 32 CTAs per SM
    return (unsigned)(g < 1 ? 1 : (g > cap ? cap : g));
}

// ---- index_points forward ---------------------------------------------------------------------------------
template <typename VEC>
__global__ void __launch_bounds__(GT)
gather_kernel(const VEC* __restrict__ points, const int64_t* __restrict__ idx, VEC* __restrict__ out,
              int N, int64_t M, int CV, int64_t total) {
    pdl_prologue();
    // total = B*M*CV vector elements; CV = vectors per row
    for (int64_t t = (int64_t)blockIdx.x * GT + threadIdx.x; t < total; t += (int64_t)gridDim.x * GT) {
        const int64_t row = t / CV;
        const int v = (int)(t - row * CV);
        const int64_t b = row / M;
        const int n = clamp_index(__ldg(idx + row), N);
        out[t] = __ldg(points + ((size_t)b * N + n) * CV + v);
    }
}

// ---- index_points backward: grad_points[b, idx[b,m], :] += grad_out[b,m,:] ----------------------------------
__global__ void __launch_bounds__(GT)
scatter_add_v4_kernel(const float4* __restrict__ grad_out, const int64_t* __restrict__ idx,
                      float* __restrict__ grad_points, int N, int64_t M, int CV, int64_t total) {
    pdl_prologue();
    for (int64_t t = (int64_t)blockIdx.x * GT + threadIdx.x; t < total; t += (int64_t)gridDim.x * GT) {
        const int64_t row = t / CV;
        const int v = (int)(t - row * CV);
        const int64_t b = row / M;
        const int n = clamp_index(__ldg(idx + row), N);
        red_add_f32x4(grad_points + (((size_t)b * N + n) * CV + v) * 4, __ldg(grad_out + t));
    }
}
// bf16 payloads (2-byte features, half the HBM bytes of the fp32 path): the forward gather is a byte mover and shares
// gather_kernel; the backward converts on the fly and accumulates in an fp32 buffer (8 neighbours summed in bf16
// would lose ~1 % of the gradient), two channels per thread.
__global__ void __launch_bounds__(GT)
scatter_add_bf16_kernel(const unsigned* __restrict__ grad_out, const int64_t* __restrict__ idx,
                        float* __restrict__ grad_points, int N, int64_t M, int C2, int64_t total) {
    pdl_prologue();
    for (int64_t t = (int64_t)blockIdx.x * GT + threadIdx.x; t < total; t += (int64_t)gridDim.x * GT) {
        const int64_t row = t / C2;
        const int c2 = (int)(t - row * C2);
        const int64_t b = row / M;
        const int n = clamp_index(__ldg(idx + row), N);
        const unsigned pr = __ldg(grad_out + t);  // two bf16: low half = even channel
        float* dst = grad_points + ((size_t)b * N + n) * (2 * C2) + 2 * c2;
        red_add_f32(dst, __uint_as_float(pr << 16));
        red_add_f32(dst + 1, __uint_as_float(pr & 0xffff0000u));
    }
}

__global__ void __launch_bounds__(GT)
scatter_add_s_kernel(const float* __restrict__ grad_out, const int64_t* __restrict__ idx,
                     float* __restrict__ grad_points, int N, int64_t M, int C, int64_t total) {
    pdl_prologue();
    for (int64_t t = (int64_t)blockIdx.x * GT + threadIdx.x; t < total; t += (int64_t)gridDim.x * GT) {
        const int64_t row = t / C;
        const int c = (int)(t - row * C);
        const int64_t b = row / M;
        const int n = clamp_index(__ldg(idx + row), N);
        red_add_f32(grad_points + ((size_t)b * N + n) * C + c, __ldg(grad_out + t));
    }
}

// ---- Markov transition ---------------------------------------------------------------------------------------
// scatter phase: one work item = (row (b,s), k, vector v).  A neighbour index repeated at an earlier k of the
// same row is skipped (the reference's scatter_ overwrites, it does not accumulate).
template <bool VEC4>
__global__ void __launch_bounds__(GT)
transition_scatter_kernel(const float* __restrict__ points, const int64_t* __restrict__ idx,
                          float* __restrict__ out, float* __restrict__ cnt, int S, int K, int C, int N,
                          int64_t total) {
    pdl_prologue();
    const int CV = VEC4 ? C / 4 : C;
    for (int64_t t = (int64_t)blockIdx.x * GT + threadIdx.x; t < total; t += (int64_t)gridDim.x * GT) {
        const int64_t rk = t / CV;  // (row, k)
        const int v = (int)(t - rk * CV);
        const int64_t row = rk / K;
        const int k = (int)(rk - row * K);
        const int64_t b = row / S;
        const int64_t* irow = idx + row * K;
        const int64_t raw = __ldg(irow + k);
        if (raw < 0 || raw >= N) continue;  // validated on the host; never touch memory out of range
        bool dup = false;
        for (int j = 0; j < k; ++j) dup |= (__ldg(irow + j) == raw);
        if (dup) continue;
        const size_t dst = ((size_t)b * N + (size_t)raw);
        if (VEC4) {
            const float4 p = __ldg(reinterpret_cast<const float4*>(points) + row * CV + v);
            red_add_f32x4(out + (dst * CV + v) * 4, p);
            if (v == 0 && p.x != 0.0f) red_add_f32(cnt + dst, 1.0f);
        } else {
            const float p = __ldg(points + row * C + v);
            red_add_f32(out + dst * C + v, p);
            if (v == 0 && p != 0.0f) red_add_f32(cnt + dst, 1.0f);
        }
    }
}

// normalise phase: out[b,n,:] /= (cnt == 0 ? 1 : cnt); cnt is rewritten with the divisor actually used.
template <bool VEC4>
__global__ void __launch_bounds__(GT)
transition_normalise_kernel(float* __restrict__ out, float* __restrict__ cnt, int C, int64_t total) {
    pdl_prologue();
    const int CV = VEC4 ? C / 4 : C;
    for (int64_t t = (int64_t)blockIdx.x * GT + threadIdx.x; t < total; t += (int64_t)gridDim.x * GT) {
        const int64_t row = t / CV;
        const int v = (int)(t - row * CV);
        float d = cnt[row];
        d = d == 0.0f ? 1.0f : d;
        if (VEC4) {
            float4* o = reinterpret_cast<float4*>(out) + t;
            float4 x = *o;
            x.x = __fdiv_rn(x.x, d);
            x.y = __fdiv_rn(x.y, d);
            x.z = __fdiv_rn(x.z, d);
            x.w = __fdiv_rn(x.w, d);
            *o = x;
        } else {
            out[t] = __fdiv_rn(out[t], d);
        }
        (void)v;
    }
}
__global__ void __launch_bounds__(GT) transition_fix_cnt_kernel(float* __restrict__ cnt, int64_t total) {
    pdl_prologue();
    for (int64_t t = (int64_t)blockIdx.x * GT + threadIdx.x; t < total; t += (int64_t)gridDim.x * GT)
        if (cnt[t] == 0.0f) cnt[t] = 1.0f;
}

// ---- Markov transition, gather form -----------------------------------------------------------------------------
// out = D^-1 A^T X is a gather once A^T (for every fine point n, the coarse points s whose neighbour list contains
// n) is available as per-cloud CSR lists.  Building them costs S*K integer atomics and a scan -- 16x fewer atomics
// than reducing S*K*C floats into out -- and the product itself then streams every source row exactly once per
// incidence and writes every output row exactly once (no zero-fill, no normalise pass).  Lists are sorted, so the
// fp32 summation order is fixed (ascending s, as in the oracle): the op is deterministic.
__device__ __forceinline__ bool dup_before(const int64_t* irow, int k, int64_t raw) {
    bool dup = false;
    for (int j = 0; j < k; ++j) dup |= (__ldg(irow + j) == raw);
    return dup;
}

__global__ void __launch_bounds__(GT)
csr_count_kernel(const int64_t* __restrict__ idx, int* __restrict__ deg, int S, int K, int N, int64_t total) {
    pdl_prologue();
    for (int64_t t = (int64_t)blockIdx.x * GT + threadIdx.x; t < total; t += (int64_t)gridDim.x * GT) {
        const int64_t row = t / K;
        const int k = (int)(t - row * K);
        const int64_t b = row / S;
        const int64_t raw = __ldg(idx + t);
        if (raw < 0 || raw >= N || dup_before(idx + row * K, k, raw)) continue;
        atomicAdd(deg + b * (int64_t)(N + 1) + raw, 1);
    }
}

// exclusive scan of deg[b, 0..N] in place (one CTA per cloud); cursor[b, n] = start of list n.  Every thread owns
// one contiguous chunk: local sum, one block-wide scan of the 1024 chunk sums, local rewrite (two barriers in all).
__global__ void __launch_bounds__(1024)
csr_scan_kernel(int* __restrict__ deg, int* __restrict__ cursor, int N) {
    pdl_prologue();
    __shared__ int warp_sums[32];
    int* d = deg + (int64_t)blockIdx.x * (N + 1);
    int* cur = cursor + (int64_t)blockIdx.x * N;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int chunk = (N + 1 + 1023) / 1024;
    const int lo = min(N + 1, (int)threadIdx.x * chunk), hi = min(N + 1, lo + chunk);
    int sum = 0;
    for (int i = lo; i < hi; ++i) sum += d[i];
    int x = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
    }
    if (lane == 31) warp_sums[warp] = x;
    __syncthreads();
    if (warp == 0) {
        int w = warp_sums[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int y = __shfl_up_sync(0xffffffffu, w, o);
            if (lane >= o) w += y;
        }
        warp_sums[lane] = w;
    }
    __syncthreads();
    int run = (warp > 0 ? warp_sums[warp - 1] : 0) + x - sum;
    for (int i = lo; i < hi; ++i) {
        const int v = d[i];
        d[i] = run;
        if (i < N) cur[i] = run;
        run += v;
    }
}

// Large clouds: three-phase scan over 1024-element blocks (grid = blocks x clouds) so that no thread walks a long
// dependent chain: (1) block sums, (2) scan of the block sums (one CTA per cloud), (3) block rescans + offset.
__device__ __forceinline__ int block_inclusive_scan_1024(int v, int* warp_sums) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
    }
    if (lane == 31) warp_sums[warp] = x;
    __syncthreads();
    if (warp == 0) {
        int w = warp_sums[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int y = __shfl_up_sync(0xffffffffu, w, o);
            if (lane >= o) w += y;
        }
        warp_sums[lane] = w;
    }
    __syncthreads();
    return x + (warp > 0 ? warp_sums[warp - 1] : 0);
}

__global__ void __launch_bounds__(1024)
csr_block_sums_kernel(const int* __restrict__ deg, int* __restrict__ bsum, int N, int nblk) {
    pdl_prologue();
    __shared__ int warp_sums[32];
    const int* d = deg + (int64_t)blockIdx.y * (N + 1);
    const int i = blockIdx.x * 1024 + threadIdx.x;
    const int incl = block_inclusive_scan_1024(i <= N ? d[i] : 0, warp_sums);
    if (threadIdx.x == 1023) bsum[(int64_t)blockIdx.y * nblk + blockIdx.x] = incl;
}

__global__ void __launch_bounds__(1024)
csr_scan_block_sums_kernel(int* __restrict__ bsum, int nblk) {
    pdl_prologue();  // nblk <= 1024*... handled by chunks, carry in smem
    __shared__ int warp_sums[32];
    __shared__ int carry;
    int* bs = bsum + (int64_t)blockIdx.x * nblk;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < nblk; base += 1024) {
        const int i = base + threadIdx.x;
        const int v = i < nblk ? bs[i] : 0;
        const int incl = block_inclusive_scan_1024(v, warp_sums);
        const int c = carry;
        __syncthreads();
        if (i < nblk) bs[i] = c + incl - v;  // exclusive
        if (threadIdx.x == 1023) carry = c + incl;
        __syncthreads();
    }
}

__global__ void __launch_bounds__(1024)
csr_block_rescan_kernel(int* __restrict__ deg, int* __restrict__ cursor, const int* __restrict__ bsum, int N,
                        int nblk) {
    pdl_prologue();
    __shared__ int warp_sums[32];
    int* d = deg + (int64_t)blockIdx.y * (N + 1);
    int* cur = cursor + (int64_t)blockIdx.y * N;
    const int i = blockIdx.x * 1024 + threadIdx.x;
    const int v = i <= N ? d[i] : 0;
    const int incl = block_inclusive_scan_1024(v, warp_sums);
    const int excl = bsum[(int64_t)blockIdx.y * nblk + blockIdx.x] + incl - v;
    if (i <= N) d[i] = excl;
    if (i < N) cur[i] = excl;
}

__global__ void __launch_bounds__(GT)
csr_fill_kernel(const int64_t* __restrict__ idx, int* __restrict__ cursor, int* __restrict__ list, int S, int K, int N,
                int64_t total) {
    pdl_prologue();
    for (int64_t t = (int64_t)blockIdx.x * GT + threadIdx.x; t < total; t += (int64_t)gridDim.x * GT) {
        const int64_t row = t / K;
        const int k = (int)(t - row * K);
        const int64_t b = row / S;
        const int64_t raw = __ldg(idx + t);
        if (raw < 0 || raw >= N || dup_before(idx + row * K, k, raw)) continue;
        const int pos = atomicAdd(cursor + b * (int64_t)N + raw, 1);
        list[b * (int64_t)S * K + pos] = (int)(row - b * S);
    }
}

// one thread per (fine point, 4 channels): sorted walk over the reverse-neighbour list
template <bool VEC4>
__global__ void __launch_bounds__(GT)
transition_gather_kernel(const float* __restrict__ points, const int* __restrict__ offs, const int* __restrict__ list,
                         float* __restrict__ out, float* __restrict__ cnt, int S, int K, int C, int N, int64_t total) {
    pdl_prologue();
    const int CV = VEC4 ? C / 4 : C;
    for (int64_t t = (int64_t)blockIdx.x * GT + threadIdx.x; t < total; t += (int64_t)gridDim.x * GT) {
        const int64_t row = t / CV;  // (b, n)
        const int v = (int)(t - row * CV);
        const int64_t b = row / N;
        const int n = (int)(row - b * N);
        const int* o = offs + b * (int64_t)(N + 1) + n;
        const int lo = __ldg(o), hi = __ldg(o + 1);
        const int* lst = list + b * (int64_t)S * K;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        float nz = 0.f;
        int prev = -1;
        for (int e = lo; e < hi; ++e) {
            // next source in ascending order (lists are short: K*S/N entries on average): selection walk
            int s = 0x7fffffff;
            for (int f = lo; f < hi; ++f) {
                const int c = __ldg(lst + f);
                if (c > prev && c < s) s = c;
            }
            prev = s;
            const float* src = points + ((size_t)b * S + s) * C;
            if (__ldg(src) != 0.0f) nz += 1.0f;  // count_nonzero of channel 0 (reference :44)
            if (VEC4) {
                const float4 p = __ldg(reinterpret_cast<const float4*>(src) + v);
                acc.x += p.x; acc.y += p.y; acc.z += p.z; acc.w += p.w;
            } else {
                acc.x += __ldg(src + v);
            }
        }
        const float dv = nz == 0.0f ? 1.0f : nz;
        if (VEC4) {
            reinterpret_cast<float4*>(out)[t] = make_float4(__fdiv_rn(acc.x, dv), __fdiv_rn(acc.y, dv),
                                                            __fdiv_rn(acc.z, dv), __fdiv_rn(acc.w, dv));
        } else {
            out[t] = __fdiv_rn(acc.x, dv);
        }
        if (v == 0) cnt[row] = dv;
    }
}

// Same product, G = min(C/4, 32) lanes per output row (C/4 a power of two): the lanes load the row's reverse-neighbour
// list together (one entry per lane), rank-sort it with shuffles (ascending s: the oracle's summation order) and then
// stream the source rows in that order, one float4 per lane and row -- each list entry is loaded once per row instead
// of L times per thread, and the channel-0 count is taken by the lane that holds channel 0.  Rows whose list is
// longer than G fall back to the selection walk.
template <int G>
__global__ void __launch_bounds__(GT)
transition_gather_group_kernel(const float* __restrict__ points, const int* __restrict__ offs,
                               const int* __restrict__ list, float* __restrict__ out, float* __restrict__ cnt, int S,
                               int K, int C, int N, int64_t rows_total) {
    pdl_prologue();
    const int CV = C / 4;
    const int slices = CV / G;  // 32-lane slices of one row when C/4 > 32
    const int lane = threadIdx.x & 31, gl = lane & (G - 1);
    const unsigned gmask = G == 32 ? 0xffffffffu : (((1u << G) - 1u) << (lane & ~(G - 1)));
    const int64_t groups_total = rows_total * slices;
    const int64_t gstride = (int64_t)gridDim.x * (GT / G);
    // all lanes of a warp run the same number of iterations (the shuffles are group-wide, the loop bound warp-wide)
    const int64_t g0 = (int64_t)blockIdx.x * (GT / G) + threadIdx.x / G;
    const int64_t w0 = g0 - (lane / G);  // first group of this warp
    for (int64_t wg = w0; wg < groups_total; wg += gstride) {
        const int64_t g = wg + lane / G;
        const bool live = g < groups_total;
        const int64_t row = live ? g / slices : 0;  // (b, n)
        const int v = (int)(live ? g - row * slices : 0) * G + gl;
        const int64_t b = row / N;
        const int n = (int)(row - b * N);
        const int* o = offs + b * (int64_t)(N + 1) + n;
        const int lo = live ? __ldg(o) : 0, hi = live ? __ldg(o + 1) : 0;
        const int L = hi - lo;
        const int* lst = list + b * (int64_t)S * K;
        const float* pb = points + (size_t)b * S * C;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        float nz = 0.f;
        const int Lmax = __reduce_max_sync(0xffffffffu, L <= G ? L : 0);
        // sorted list in registers: lane p of the group ends up with the p-th smallest source
        const int mine = (gl < L && L <= G) ? __ldg(lst + lo + gl) : 0x7fffffff;
        int rank = 0;
        for (int e = 0; e < Lmax; ++e) {
            const int other = __shfl_sync(gmask, mine, (lane & ~(G - 1)) + e);
            rank += (other < mine) ? 1 : 0;  // sources of one list are distinct
        }
        int sorted = 0x7fffffff;
        for (int e = 0; e < Lmax; ++e) {
            const int val = __shfl_sync(gmask, mine, (lane & ~(G - 1)) + e);
            const int r = __shfl_sync(gmask, rank, (lane & ~(G - 1)) + e);
            if (r == gl) sorted = val;
        }
        for (int e0 = 0; e0 < Lmax; e0 += 4) {  // four source rows in flight per lane, summed in list order
            float4 pr[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int s = __shfl_sync(gmask, sorted, (lane & ~(G - 1)) + ((e0 + u) & (G - 1)));
                pr[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (e0 + u < L && L <= G) pr[u] = __ldg(reinterpret_cast<const float4*>(pb + (size_t)s * C) + v);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (e0 + u < L && L <= G) {
                    if (v == 0 && pr[u].x != 0.0f) nz += 1.0f;  // count_nonzero of channel 0 (reference :44)
                    acc.x += pr[u].x; acc.y += pr[u].y; acc.z += pr[u].z; acc.w += pr[u].w;
                }
            }
        }
        if (L > G) {  // long list: selection walk (rare)
            int prev = -1;
            for (int e = lo; e < hi; ++e) {
                int s = 0x7fffffff;
                for (int f = lo; f < hi; ++f) {
                    const int c = __ldg(lst + f);
                    if (c > prev && c < s) s = c;
                }
                prev = s;
                const float* src = pb + (size_t)s * C;
                if (v == 0 && __ldg(src) != 0.0f) nz += 1.0f;
                const float4 p = __ldg(reinterpret_cast<const float4*>(src) + v);
                acc.x += p.x; acc.y += p.y; acc.z += p.z; acc.w += p.w;
            }
        }
        // the count lives in the lane that holds channel 0 (slice 0, group lane 0); other slices of the row recount
        float dv;
        if (slices == 1) {
            nz = __shfl_sync(gmask, nz, lane & ~(G - 1));
            dv = nz == 0.0f ? 1.0f : nz;
        } else {
            float c0 = 0.f;  // (C/4 > 32: the row spans several warps; every slice counts channel 0 itself)
            if (gl == 0 && live) {
                for (int e = lo; e < hi; ++e)
                    if (__ldg(pb + (size_t)__ldg(lst + e) * C) != 0.0f) c0 += 1.0f;
            }
            c0 = __shfl_sync(gmask, c0, lane & ~(G - 1));
            dv = c0 == 0.0f ? 1.0f : c0;
        }
        if (live) {
            reinterpret_cast<float4*>(out)[row * CV + v] = make_float4(__fdiv_rn(acc.x, dv), __fdiv_rn(acc.y, dv),
                                                                       __fdiv_rn(acc.z, dv), __fdiv_rn(acc.w, dv));
            if (v == 0) cnt[row] = dv;
        }
    }
}

// backward: a pure gather (no atomics): grad_points[row,:] = sum_k grad_out[b, idx[row,k], :] / cnt[b, idx[row,k]]
template <bool VEC4>
__global__ void __launch_bounds__(GT)
transition_bwd_kernel(const float* __restrict__ grad_out, const int64_t* __restrict__ idx,
                      const float* __restrict__ cnt, float* __restrict__ grad_points, int S, int K, int C,
                      int N, int64_t total) {
    pdl_prologue();
    const int CV = VEC4 ? C / 4 : C;
    for (int64_t t = (int64_t)blockIdx.x * GT + threadIdx.x; t < total; t += (int64_t)gridDim.x * GT) {
        const int64_t row = t / CV;
        const int v = (int)(t - row * CV);
        const int64_t b = row / S;
        const int64_t* irow = idx + row * K;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int k = 0; k < K; ++k) {
            const int64_t raw = __ldg(irow + k);
            if (raw < 0 || raw >= N) continue;
            bool dup = false;
            for (int j = 0; j < k; ++j) dup |= (__ldg(irow + j) == raw);
            if (dup) continue;
            const size_t src = (size_t)b * N + (size_t)raw;
            const float d = __ldg(cnt + src);
            if (VEC4) {
                const float4 g = __ldg(reinterpret_cast<const float4*>(grad_out) + src * CV + v);
                acc.x += __fdiv_rn(g.x, d);
                acc.y += __fdiv_rn(g.y, d);
                acc.z += __fdiv_rn(g.z, d);
                acc.w += __fdiv_rn(g.w, d);
            } else {
                acc.x += __fdiv_rn(__ldg(grad_out + src * C + v), d);
            }
        }
        if (VEC4)
            reinterpret_cast<float4*>(grad_points)[t] = acc;
        else
            grad_points[t] = acc.x;
    }
}

// ---- three_interpolate ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(GT)
three_weights_kernel(const float* __restrict__ dist, float* __restrict__ weight, int64_t rows) {
    pdl_prologue();
    for (int64_t r = (int64_t)blockIdx.x * GT + threadIdx.x; r < rows; r += (int64_t)gridDim.x * GT) {
        const float r0 = __fdiv_rn(1.0f, __fadd_rn(dist[r * 3 + 0], 1e-8f));
        const float r1 = __fdiv_rn(1.0f, __fadd_rn(dist[r * 3 + 1], 1e-8f));
        const float r2 = __fdiv_rn(1.0f, __fadd_rn(dist[r * 3 + 2], 1e-8f));
        const float norm = __fadd_rn(__fadd_rn(r0, r1), r2);
        weight[r * 3 + 0] = __fdiv_rn(r0, norm);
        weight[r * 3 + 1] = __fdiv_rn(r1, norm);
        weight[r * 3 + 2] = __fdiv_rn(r2, norm);
    }
}

template <bool VEC4>
__global__ void __launch_bounds__(GT)
three_interp_fwd_kernel(const float* __restrict__ points2, const float* __restrict__ weight,
                        const int64_t* __restrict__ idx, float* __restrict__ out, int N, int S, int C,
                        int64_t total) {
    pdl_prologue();
    const int CV = VEC4 ? C / 4 : C;
    for (int64_t t = (int64_t)blockIdx.x * GT + threadIdx.x; t < total; t += (int64_t)gridDim.x * GT) {
        const int64_t row = t / CV;
        const int v = (int)(t - row * CV);
        const int64_t b = row / N;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const int s = clamp_index(__ldg(idx + row * 3 + i), S);
            const float w = __ldg(weight + row * 3 + i);
            if (VEC4) {
                const float4 p = __ldg(reinterpret_cast<const float4*>(points2) + ((size_t)b * S + s) * CV + v);
                // products rounded separately, summed left to right (torch.sum over the 3-axis)
                acc.x = i == 0 ? __fmul_rn(w, p.x) : __fadd_rn(acc.x, __fmul_rn(w, p.x));
                acc.y = i == 0 ? __fmul_rn(w, p.y) : __fadd_rn(acc.y, __fmul_rn(w, p.y));
                acc.z = i == 0 ? __fmul_rn(w, p.z) : __fadd_rn(acc.z, __fmul_rn(w, p.z));
                acc.w = i == 0 ? __fmul_rn(w, p.w) : __fadd_rn(acc.w, __fmul_rn(w, p.w));
            } else {
                const float p = __ldg(points2 + ((size_t)b * S + s) * C + v);
                acc.x = i == 0 ? __fmul_rn(w, p) : __fadd_rn(acc.x, __fmul_rn(w, p));
            }
        }
        if (VEC4)
            reinterpret_cast<float4*>(out)[t] = acc;
        else
            out[t] = acc.x;
    }
}

template <bool VEC4>
__global__ void __launch_bounds__(GT)
three_interp_bwd_kernel(const float* __restrict__ grad_out, const float* __restrict__ weight,
                        const int64_t* __restrict__ idx, float* __restrict__ grad_points2, int N, int S, int C,
                        int64_t total) {
    pdl_prologue();
    const int CV = VEC4 ? C / 4 : C;
    for (int64_t t = (int64_t)blockIdx.x * GT + threadIdx.x; t < total; t += (int64_t)gridDim.x * GT) {
        const int64_t row = t / CV;
        const int v = (int)(t - row * CV);
        const int64_t b = row / N;
        float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
        if (VEC4)
            g = __ldg(reinterpret_cast<const float4*>(grad_out) + t);
        else
            g.x = __ldg(grad_out + t);
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const int s = clamp_index(__ldg(idx + row * 3 + i), S);
            const float w = __ldg(weight + row * 3 + i);
            if (VEC4)
                red_add_f32x4(grad_points2 + (((size_t)b * S + s) * CV + v) * 4,
                              make_float4(w * g.x, w * g.y, w * g.z, w * g.w));
            else
                red_add_f32(grad_points2 + ((size_t)b * S + s) * C + v, w * g.x);
        }
    }
}

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace mpc

using namespace mpc;

MPC_API int mpc_gather_f32(const float* points, const int64_t* idx, float* out, int64_t B, int64_t N, int64_t M,
                           int64_t C, mpc_stream_t stream) {
    if (B < 0 || N <= 0 || M < 0 || C <= 0 || N > INT32_MAX) return MPC_ERR_INVALID;
    if (B == 0 || M == 0) return MPC_OK;
    if (!points || !idx || !out) return MPC_ERR_INVALID;
    cudaStream_t st = (cudaStream_t)stream;
    if (C % 4 == 0 && aligned16(points) && aligned16(out)) {
        const int CV = (int)(C / 4);
        const int64_t total = B * M * CV;
        pdl_launch(gather_kernel<float4>, dim3(grid_for(total)), dim3(GT), 0, st, reinterpret_cast<const float4*>(points), idx,
                                                             reinterpret_cast<float4*>(out), (int)N, M, CV, total);
    } else {
        const int64_t total = B * M * C;
        pdl_launch(gather_kernel<float>, dim3(grid_for(total)), dim3(GT), 0, st, points, idx, out, (int)N, M, (int)C, total);
    }
    MPC_LAUNCH_CHECK();
    return MPC_OK;
}

MPC_API int mpc_gather_i64(const int64_t* values, const int64_t* idx, int64_t* out, int64_t B, int64_t N,
                           int64_t M, mpc_stream_t stream) {
    if (B < 0 || N <= 0 || M < 0 || N > INT32_MAX) return MPC_ERR_INVALID;
    if (B == 0 || M == 0) return MPC_OK;
    if (!values || !idx || !out) return MPC_ERR_INVALID;
    const int64_t total = B * M;
    pdl_launch(gather_kernel<long long>, dim3(grid_for(total)), dim3(GT), 0, (cudaStream_t)stream, 
        reinterpret_cast<const long long*>(values), idx, reinterpret_cast<long long*>(out), (int)N, M, 1, total);
    MPC_LAUNCH_CHECK();
    return MPC_OK;
}

/* bf16 rows: [B,N,C] bf16 -> [B,M,C] bf16 (a byte mover: 16-byte, 4-byte or 2-byte vectors by alignment). */
MPC_API int mpc_gather_bf16(const void* points, const int64_t* idx, void* out, int64_t B, int64_t N, int64_t M,
                            int64_t C, mpc_stream_t stream) {
    if (B < 0 || N <= 0 || M < 0 || C <= 0 || N > INT32_MAX) return MPC_ERR_INVALID;
    if (B == 0 || M == 0) return MPC_OK;
    if (!points || !idx || !out) return MPC_ERR_INVALID;
    cudaStream_t st = (cudaStream_t)stream;
    if (C % 8 == 0 && aligned16(points) && aligned16(out)) {
        const int CV = (int)(C / 8);
        const int64_t total = B * M * CV;
        pdl_launch(gather_kernel<float4>, dim3(grid_for(total)), dim3(GT), 0, st, reinterpret_cast<const float4*>(points),
                   idx, reinterpret_cast<float4*>(out), (int)N, M, CV, total);
    } else if (C % 2 == 0) {
        const int CV = (int)(C / 2);
        const int64_t total = B * M * CV;
        pdl_launch(gather_kernel<float>, dim3(grid_for(total)), dim3(GT), 0, st, reinterpret_cast<const float*>(points),
                   idx, reinterpret_cast<float*>(out), (int)N, M, CV, total);
    } else {
        const int64_t total = B * M * C;
        pdl_launch(gather_kernel<unsigned short>, dim3(grid_for(total)), dim3(GT), 0, st,
                   reinterpret_cast<const unsigned short*>(points), idx, reinterpret_cast<unsigned short*>(out), (int)N, M,
                   (int)C, total);
    }
    MPC_LAUNCH_CHECK();
    return MPC_OK;
}

/* grad_points [B,N,C] F32 (cleared by the call) += grad_out [B,M,C] bf16 scattered by idx; C % 2 == 0. */
MPC_API int mpc_gather_bwd_bf16(const void* grad_out, const int64_t* idx, float* grad_points, int64_t B, int64_t N,
                                int64_t M, int64_t C, mpc_stream_t stream) {
    if (B < 0 || N <= 0 || M < 0 || C <= 0 || N > INT32_MAX) return MPC_ERR_INVALID;
    if (C % 2) return MPC_ERR_UNSUPPORTED;
    if (B == 0) return MPC_OK;
    if (!grad_points) return MPC_ERR_INVALID;
    cudaStream_t st = (cudaStream_t)stream;
    MPC_CUDA(cudaMemsetAsync(grad_points, 0, (size_t)B * N * C * sizeof(float), st));
    if (M == 0) return MPC_OK;
    if (!grad_out || !idx) return MPC_ERR_INVALID;
    const int C2 = (int)(C / 2);
    const int64_t total = B * M * C2;
    pdl_launch(scatter_add_bf16_kernel, dim3(grid_for(total)), dim3(GT), 0, st, reinterpret_cast<const unsigned*>(grad_out),
               idx, grad_points, (int)N, M, C2, total);
    MPC_LAUNCH_CHECK();
    return MPC_OK;
}

/* Bytes of the reduction scratch a layer of C channels needs under the scratch contract (2C sums + 2 tickets). */
MPC_API int mpc_reduction_scratch_bytes(int64_t C) {
    return (C < 0 || C > (INT32_MAX / 16) - 2) ? MPC_ERR_INVALID : (int)((2 * C + 2) * sizeof(double));
}

MPC_API int mpc_gather_bwd_f32(const float* grad_out, const int64_t* idx, float* grad_points, int64_t B,
                               int64_t N, int64_t M, int64_t C, mpc_stream_t stream) {
    if (B < 0 || N <= 0 || M < 0 || C <= 0 || N > INT32_MAX) return MPC_ERR_INVALID;
    if (B == 0) return MPC_OK;
    if (!grad_points) return MPC_ERR_INVALID;
    cudaStream_t st = (cudaStream_t)stream;
    MPC_CUDA(cudaMemsetAsync(grad_points, 0, (size_t)B * N * C * sizeof(float), st));
    if (M == 0) return MPC_OK;
    if (!grad_out || !idx) return MPC_ERR_INVALID;
    if (C % 4 == 0 && aligned16(grad_out) && aligned16(grad_points)) {
        const int CV = (int)(C / 4);
        const int64_t total = B * M * CV;
        pdl_launch(scatter_add_v4_kernel, dim3(grid_for(total)), dim3(GT), 0, st, reinterpret_cast<const float4*>(grad_out), idx,
                                                              grad_points, (int)N, M, CV, total);
    } else {
        const int64_t total = B * M * C;
        pdl_launch(scatter_add_s_kernel, dim3(grid_for(total)), dim3(GT), 0, st, grad_out, idx, grad_points, (int)N, M, (int)C, total);
    }
    MPC_LAUNCH_CHECK();
    return MPC_OK;
}

MPC_API int mpc_transition_fwd_f32(const float* points, const int64_t* idx, float* out, float* cnt, int64_t B,
                                   int64_t S, int64_t K, int64_t C, int64_t N, mpc_stream_t stream) {
    if (B < 0 || S < 0 || K <= 0 || C <= 0 || N <= 0 || N > INT32_MAX) return MPC_ERR_INVALID;
    if (K > 32) return MPC_ERR_UNSUPPORTED;
    if (B == 0) return MPC_OK;
    if (!out || !cnt || (S > 0 && (!points || !idx))) return MPC_ERR_INVALID;
    cudaStream_t st = (cudaStream_t)stream;
    MPC_CUDA(cudaMemsetAsync(out, 0, (size_t)B * N * C * sizeof(float), st));
    MPC_CUDA(cudaMemsetAsync(cnt, 0, (size_t)B * N * sizeof(float), st));
    const bool v4 = C % 4 == 0 && aligned16(points) && aligned16(out);
    const int CV = (int)(v4 ? C / 4 : C);
    if (S > 0) {
        const int64_t total = B * S * K * CV;
        if (v4)
            pdl_launch(transition_scatter_kernel<true>, dim3(grid_for(total)), dim3(GT), 0, st, points, idx, out, cnt, (int)S, (int)K,
                                                                           (int)C, (int)N, total);
        else
            pdl_launch(transition_scatter_kernel<false>, dim3(grid_for(total)), dim3(GT), 0, st, points, idx, out, cnt, (int)S, (int)K,
                                                                            (int)C, (int)N, total);
        MPC_LAUNCH_CHECK();
    }
    const int64_t total = B * N * CV;
    if (v4)
        pdl_launch(transition_normalise_kernel<true>, dim3(grid_for(total)), dim3(GT), 0, st, out, cnt, (int)C, total);
    else
        pdl_launch(transition_normalise_kernel<false>, dim3(grid_for(total)), dim3(GT), 0, st, out, cnt, (int)C, total);
    MPC_LAUNCH_CHECK();
    pdl_launch(transition_fix_cnt_kernel, dim3(grid_for(B * N)), dim3(GT), 0, st, cnt, B * N);
    MPC_LAUNCH_CHECK();
    return MPC_OK;
}

MPC_API int mpc_transition_csr_build(const int64_t* idx, int32_t* workspace, int64_t B, int64_t S, int64_t K,
                                     int64_t N, mpc_stream_t stream) {
    if (B < 0 || S < 0 || K <= 0 || N <= 0 || N > INT32_MAX) return MPC_ERR_INVALID;
    if (K > 32 || S * K > INT32_MAX) return MPC_ERR_UNSUPPORTED;
    if (B == 0) return MPC_OK;
    if (!workspace || (S > 0 && !idx)) return MPC_ERR_INVALID;
    cudaStream_t st = (cudaStream_t)stream;
    // workspace: offsets [B, N+1] | cursor [B, N] | lists [B, S*K]
    int* offs = workspace;
    int* cursor = offs + B * (N + 1);
    int* list = cursor + B * N;
    MPC_CUDA(cudaMemsetAsync(offs, 0, sizeof(int) * (size_t)B * (N + 1), st));
    if (S > 0) {
        const int64_t total = B * S * K;
        pdl_launch(csr_count_kernel, dim3(grid_for(total)), dim3(GT), 0, st, idx, offs, (int)S, (int)K, (int)N, total);
        MPC_LAUNCH_CHECK();
    }
    if (N + 1 <= 8192) {
        pdl_launch(csr_scan_kernel, dim3((unsigned)B), dim3(1024), 0, st, offs, cursor, (int)N);
        MPC_LAUNCH_CHECK();
    } else {  // the block sums live at the head of the (not yet filled) list area
        const int nblk = (int)ceil_div(N + 1, 1024);
        if ((int64_t)nblk > S * K) return MPC_ERR_UNSUPPORTED;
        int* bsum = list;
        pdl_launch(csr_block_sums_kernel, dim3(dim3((unsigned)nblk, (unsigned)B)), dim3(1024), 0, st, offs, bsum, (int)N, nblk);
        MPC_LAUNCH_CHECK();
        pdl_launch(csr_scan_block_sums_kernel, dim3((unsigned)B), dim3(1024), 0, st, bsum, nblk);
        MPC_LAUNCH_CHECK();
        pdl_launch(csr_block_rescan_kernel, dim3(dim3((unsigned)nblk, (unsigned)B)), dim3(1024), 0, st, offs, cursor, bsum, (int)N, nblk);
        MPC_LAUNCH_CHECK();
    }
    if (S > 0) {
        const int64_t total = B * S * K;
        pdl_launch(csr_fill_kernel, dim3(grid_for(total)), dim3(GT), 0, st, idx, cursor, list, (int)S, (int)K, (int)N, total);
        MPC_LAUNCH_CHECK();
    }
    return MPC_OK;
}

MPC_API int mpc_transition_csr_apply_f32(const float* points, const int32_t* workspace, float* out, float* cnt,
                                         int64_t B, int64_t S, int64_t K, int64_t C, int64_t N, mpc_stream_t stream) {
    if (B < 0 || S < 0 || K <= 0 || C <= 0 || N <= 0 || N > INT32_MAX) return MPC_ERR_INVALID;
    if (K > 32 || S * K > INT32_MAX) return MPC_ERR_UNSUPPORTED;
    if (B == 0) return MPC_OK;
    if (!out || !cnt || !workspace || (S > 0 && !points)) return MPC_ERR_INVALID;
    cudaStream_t st = (cudaStream_t)stream;
    const int* offs = workspace;
    const int* list = offs + B * (N + 1) + B * N;
    const bool v4 = C % 4 == 0 && aligned16(points) && aligned16(out);
    const int64_t CV = v4 ? C / 4 : C;
    if (v4 && (CV & (CV - 1)) == 0 && g_knob[7] == 0) {
        // G lanes per output row; grid: enough warps in flight to cover the gather latency, whole waves of SMs
        const int G = (int)(CV < 32 ? CV : 32);
        const int64_t groups = B * N * (CV / G);
        const int64_t want = ceil_div(groups * G, GT);
        const int64_t cap = (int64_t)kNumSMs * (g_knob[5] > 0 ? g_knob[5] : 16);
        const unsigned grid = (unsigned)(want < cap ? (want > 0 ? want : 1) : cap);
        switch (G) {
#define MPC_TG(GG) case GG: pdl_launch(transition_gather_group_kernel<GG>, dim3(grid), dim3(GT), 0, st, points, offs, list, out, cnt, (int)S, (int)K, (int)C, (int)N, B * N); break;
            MPC_TG(1) MPC_TG(2) MPC_TG(4) MPC_TG(8) MPC_TG(16) MPC_TG(32)
#undef MPC_TG
        }
        MPC_LAUNCH_CHECK();
        return MPC_OK;
    }
    const int64_t total = B * N * CV;
    if (v4)
        pdl_launch(transition_gather_kernel<true>, dim3(grid_for(total)), dim3(GT), 0, st, points, offs, list, out, cnt, (int)S, (int)K, (int)C,
                                                                      (int)N, total);
    else
        pdl_launch(transition_gather_kernel<false>, dim3(grid_for(total)), dim3(GT), 0, st, points, offs, list, out, cnt, (int)S, (int)K,
                                                                       (int)C, (int)N, total);
    MPC_LAUNCH_CHECK();
    return MPC_OK;
}

MPC_API int mpc_transition_fwd_csr_f32(const float* points, const int64_t* idx, float* out, float* cnt,
                                       int32_t* workspace, int64_t B, int64_t S, int64_t K, int64_t C, int64_t N,
                                       mpc_stream_t stream) {
    if (!out || !cnt) return B == 0 ? MPC_OK : MPC_ERR_INVALID;
    const int rc = mpc_transition_csr_build(idx, workspace, B, S, K, N, stream);
    if (rc != MPC_OK) return rc;
    return mpc_transition_csr_apply_f32(points, workspace, out, cnt, B, S, K, C, N, stream);
}

MPC_API int mpc_transition_bwd_f32(const float* grad_out, const int64_t* idx, const float* cnt,
                                   float* grad_points, int64_t B, int64_t S, int64_t K, int64_t C, int64_t N,
                                   mpc_stream_t stream) {
    if (B < 0 || S < 0 || K <= 0 || C <= 0 || N <= 0 || N > INT32_MAX) return MPC_ERR_INVALID;
    if (K > 32) return MPC_ERR_UNSUPPORTED;
    if (B == 0 || S == 0) return MPC_OK;
    if (!grad_out || !idx || !cnt || !grad_points) return MPC_ERR_INVALID;
    cudaStream_t st = (cudaStream_t)stream;
    const bool v4 = C % 4 == 0 && aligned16(grad_out) && aligned16(grad_points);
    const int CV = (int)(v4 ? C / 4 : C);
    const int64_t total = B * S * CV;
    if (v4)
        pdl_launch(transition_bwd_kernel<true>, dim3(grid_for(total)), dim3(GT), 0, st, grad_out, idx, cnt, grad_points, (int)S, (int)K,
                                                                   (int)C, (int)N, total);
    else
        pdl_launch(transition_bwd_kernel<false>, dim3(grid_for(total)), dim3(GT), 0, st, grad_out, idx, cnt, grad_points, (int)S, (int)K,
                                                                    (int)C, (int)N, total);
    MPC_LAUNCH_CHECK();
    return MPC_OK;
}

MPC_API int mpc_three_interpolate_fwd_f32(const float* points2, const float* dist, const int64_t* idx,
                                          float* weight_out, float* out, int64_t B, int64_t N, int64_t S,
                                          int64_t C, mpc_stream_t stream) {
    if (B < 0 || N < 0 || S <= 0 || C <= 0 || S > INT32_MAX) return MPC_ERR_INVALID;
    if (B == 0 || N == 0) return MPC_OK;
    if (!points2 || !dist || !idx || !weight_out || !out) return MPC_ERR_INVALID;
    cudaStream_t st = (cudaStream_t)stream;
    pdl_launch(three_weights_kernel, dim3(grid_for(B * N)), dim3(GT), 0, st, dist, weight_out, B * N);
    MPC_LAUNCH_CHECK();
    const bool v4 = C % 4 == 0 && aligned16(points2) && aligned16(out);
    const int64_t total = B * N * (v4 ? C / 4 : C);
    if (v4)
        pdl_launch(three_interp_fwd_kernel<true>, dim3(grid_for(total)), dim3(GT), 0, st, points2, weight_out, idx, out, (int)N, (int)S,
                                                                     (int)C, total);
    else
        pdl_launch(three_interp_fwd_kernel<false>, dim3(grid_for(total)), dim3(GT), 0, st, points2, weight_out, idx, out, (int)N, (int)S,
                                                                      (int)C, total);
    MPC_LAUNCH_CHECK();
    return MPC_OK;
}

MPC_API int mpc_three_interpolate_bwd_f32(const float* grad_out, const float* weight, const int64_t* idx,
                                          float* grad_points2, int64_t B, int64_t N, int64_t S, int64_t C,
                                          mpc_stream_t stream) {
    if (B < 0 || N < 0 || S <= 0 || C <= 0 || S > INT32_MAX) return MPC_ERR_INVALID;
    if (B == 0) return MPC_OK;
    if (!grad_points2) return MPC_ERR_INVALID;
    cudaStream_t st = (cudaStream_t)stream;
    MPC_CUDA(cudaMemsetAsync(grad_points2, 0, (size_t)B * S * C * sizeof(float), st));
    if (N == 0) return MPC_OK;
    if (!grad_out || !weight || !idx) return MPC_ERR_INVALID;
    const bool v4 = C % 4 == 0 && aligned16(grad_out) && aligned16(grad_points2);
    const int64_t total = B * N * (v4 ? C / 4 : C);
    if (v4)
        pdl_launch(three_interp_bwd_kernel<true>, dim3(grid_for(total)), dim3(GT), 0, st, grad_out, weight, idx, grad_points2, (int)N,
                                                                     (int)S, (int)C, total);
    else
        pdl_launch(three_interp_bwd_kernel<false>, dim3(grid_for(total)), dim3(GT), 0, st, grad_out, weight, idx, grad_points2, (int)N,
                                                                      (int)S, (int)C, total);
    MPC_LAUNCH_CHECK();
    return MPC_OK;
}
