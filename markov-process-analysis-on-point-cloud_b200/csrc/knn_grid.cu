// Coordinate-space (C = 3) k nearest neighbours through a uniform grid, bit-identical to the brute-force contract of
// mpc_knn_f32 (include/mpc_b200.h; replaces R/modules/pointnet2_utils.py:190-222 for the coordinate searches of
// LocalMerge / Fuse / three_nn).
//
// Brute force evaluates B*S*N distances (4.6 G at 8 x 24000 x 24000) to keep 8 per query.  Here the reference set of
// every cloud is counting-sorted into the cells of a uniform grid (about K/4 points per cell) and a query only
// evaluates the points of the cells its neighbourhood can reach.  What makes the result IDENTICAL rather than
// approximately equal:
//   * every candidate's distance is the contract's expression, evaluated by the same instruction sequence as
//     knn3_kernel (fma chain over x, y, z; ((-2 dot) + |q|^2) + |r|^2; norms sequential non-fused);
//   * candidates are ranked by (distance, index) lexicographically, so the order in which cells are visited (and the
//     arbitrary order of points inside a cell) does not matter;
//   * the search stops only when every cell that intersects the ball of radius sqrt(d_K + eps) around the query has
//     been visited, where d_K is the current K-th best EXPANDED distance and eps = 64 * 2^-24 * (|q|^2 + max|r|^2)
//     bounds (with a 5x margin) how far the expanded form can lie below the true squared distance: a point outside
//     that ball cannot have an expanded distance <= d_K.  Cell coordinates are a monotone function of the coordinate
//     (subtract, multiply, floor, clamp), and the ball's cell range is computed with the same function on outward-
//     rounded bounds, so no point of the ball can sit in an unvisited cell.
#include <math.h>

#include "common.cuh"

namespace mpc {

constexpr int GRID_PARTS = 32;       // bounding-box partials per cloud
constexpr int GRID_MAX_DIM = 256;    // cells per axis
constexpr int GRID_THREADS = 256;
constexpr int GRID_QT = 128;         // queries per CTA of the search kernel (64 was tried for single-cloud launches: no gain
                                     // at 24 000 queries, 14 % slower at 262 144)

struct GridParams {
    float minx, miny, minz, inv_h;
    int gx, gy, gz;
    float max_rn;  // max |r|^2 over the cloud (bounds the rounding error of the expanded form)
};

__device__ __forceinline__ int cell_coord(float v, float mn, float inv_h, int g) {
    const float t = floorf(__fmul_rn(__fsub_rn(v, mn), inv_h));
    return (int)fminf(fmaxf(t, 0.0f), (float)(g - 1));
}

// Grid geometry of one cloud from its bounding-box partials; identical code (hence identical results) wherever it is
// evaluated.  The cell edge starts at the value that gives `target` cells in the box and grows until the cell count
// fits the workspace (flat or degenerate clouds).
__device__ GridParams grid_params(const float* __restrict__ part, int cap, int target) {
    float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY}, rn = 0.f;
    for (int p = 0; p < GRID_PARTS; ++p) {
        const float* q = part + p * 8;
        for (int d = 0; d < 3; ++d) {
            mn[d] = fminf(mn[d], q[d]);
            mx[d] = fmaxf(mx[d], q[3 + d]);
        }
        rn = fmaxf(rn, q[6]);
    }
    float ext[3], maxext = 0.f, vol = 1.f;
    for (int d = 0; d < 3; ++d) {
        ext[d] = fmaxf(mx[d] - mn[d], 0.f);
        maxext = fmaxf(maxext, ext[d]);
        vol *= ext[d];
    }
    float h = fmaxf(cbrtf(vol / (float)target), maxext / (float)GRID_MAX_DIM);
    h = fmaxf(h, 1e-30f);
    int g[3];
    for (int it = 0; it < 200; ++it) {
        for (int d = 0; d < 3; ++d) g[d] = min((int)(ext[d] / h) + 1, GRID_MAX_DIM);
        if ((int64_t)g[0] * g[1] * g[2] <= cap) break;
        h *= 1.125f;
    }
    GridParams gp;
    gp.minx = mn[0];
    gp.miny = mn[1];
    gp.minz = mn[2];
    gp.inv_h = 1.0f / h;
    gp.gx = g[0];
    gp.gy = g[1];
    gp.gz = g[2];
    gp.max_rn = rn;
    return gp;
}

// (1) bounding-box partials of every cloud; the same launch clears the cell counters.
__global__ void __launch_bounds__(GRID_THREADS)
grid_bbox_kernel(const float* __restrict__ ref, float* __restrict__ part, int* __restrict__ start, int N, int cap) {
    const int b = blockIdx.y, p = blockIdx.x;
    const float* rb = ref + (size_t)b * N * 3;
    const int chunk = (N + GRID_PARTS - 1) / GRID_PARTS;
    const int i0 = p * chunk, i1 = min(N, i0 + chunk);
    float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY}, rn = 0.f;
    for (int i = i0 + threadIdx.x; i < i1; i += GRID_THREADS) {
        const float x = rb[3 * i], y = rb[3 * i + 1], z = rb[3 * i + 2];
        mn[0] = fminf(mn[0], x); mx[0] = fmaxf(mx[0], x);
        mn[1] = fminf(mn[1], y); mx[1] = fmaxf(mx[1], y);
        mn[2] = fminf(mn[2], z); mx[2] = fmaxf(mx[2], z);
        rn = fmaxf(rn, sqnorm3(x, y, z));
    }
    __shared__ float red[GRID_THREADS / 32][8];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            mn[d] = fminf(mn[d], __shfl_xor_sync(0xffffffffu, mn[d], o));
            mx[d] = fmaxf(mx[d], __shfl_xor_sync(0xffffffffu, mx[d], o));
        }
        rn = fmaxf(rn, __shfl_xor_sync(0xffffffffu, rn, o));
    }
    if (lane == 0) {
        for (int d = 0; d < 3; ++d) {
            red[w][d] = mn[d];
            red[w][3 + d] = mx[d];
        }
        red[w][6] = rn;
    }
    __syncthreads();
    if (threadIdx.x < 7) {
        const int d = threadIdx.x;
        float v = red[0][d];
        for (int k = 1; k < GRID_THREADS / 32; ++k) v = d < 3 ? fminf(v, red[k][d]) : fmaxf(v, red[k][d]);
        part[((size_t)b * GRID_PARTS + p) * 8 + d] = v;
    }
    int* sb = start + (size_t)b * (cap + 1);
    for (int c = p * GRID_THREADS + threadIdx.x; c <= cap; c += GRID_PARTS * GRID_THREADS) sb[c] = 0;
}

// (2) cell of every point + cell histogram.
__global__ void __launch_bounds__(GRID_THREADS)
grid_count_kernel(const float* __restrict__ ref, const float* __restrict__ part, int* __restrict__ cell_of,
                  int* __restrict__ start, int N, int cap, int target) {
    __shared__ GridParams gp;
    const int b = blockIdx.y;
    if (threadIdx.x == 0) gp = grid_params(part + (size_t)b * GRID_PARTS * 8, cap, target);
    __syncthreads();
    const int i = blockIdx.x * GRID_THREADS + threadIdx.x;
    if (i >= N) return;
    const float* r = ref + ((size_t)b * N + i) * 3;
    const int cx = cell_coord(r[0], gp.minx, gp.inv_h, gp.gx);
    const int cy = cell_coord(r[1], gp.miny, gp.inv_h, gp.gy);
    const int cz = cell_coord(r[2], gp.minz, gp.inv_h, gp.gz);
    const int c = (cz * gp.gy + cy) * gp.gx + cx;
    cell_of[(size_t)b * N + i] = c;
    atomicAdd(start + (size_t)b * (cap + 1) + c, 1);
}

// (3) exclusive scan of the histogram (one CTA per cloud), cursor copy for the fill, grid geometry for the queries.
__global__ void __launch_bounds__(1024)
grid_scan_kernel(const float* __restrict__ part, int* __restrict__ start, int* __restrict__ cursor,
                 GridParams* __restrict__ params, int N, int cap, int target) {
    __shared__ GridParams gp;
    __shared__ int wsum[32];
    __shared__ int carry_s;
    const int b = blockIdx.x;
    if (threadIdx.x == 0) {
        gp = grid_params(part + (size_t)b * GRID_PARTS * 8, cap, target);
        params[b] = gp;
        carry_s = 0;
    }
    __syncthreads();
    const int ncell = gp.gx * gp.gy * gp.gz;
    int* sb = start + (size_t)b * (cap + 1);
    int* cb = cursor + (size_t)b * cap;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int base = 0; base < ncell; base += 4096) {
        const int c0 = base + threadIdx.x * 4;
        int v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = c0 + u < ncell ? sb[c0 + u] : 0;
        const int tsum = v[0] + v[1] + v[2] + v[3];
        int inc = tsum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        if (lane == 31) wsum[w] = inc;
        __syncthreads();
        if (w == 0) {
            int x = wsum[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, x, o);
                if (lane >= o) x += t;
            }
            wsum[lane] = x;
        }
        __syncthreads();
        const int carry = carry_s;
        int run = carry + (w > 0 ? wsum[w - 1] : 0) + inc - tsum;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (c0 + u < ncell) {
                sb[c0 + u] = run;
                cb[c0 + u] = run;
            }
            run += v[u];
        }
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = carry + wsum[31];
        __syncthreads();
    }
    if (threadIdx.x == 0) sb[ncell] = N;
}

// (4) counting-sort fill: (x, y, z, original index) of every point into its cell's slot range.  The order inside a
// cell depends on atomic arrival and is irrelevant (candidates are ranked by (distance, index)).
__global__ void __launch_bounds__(GRID_THREADS)
grid_fill_kernel(const float* __restrict__ ref, const int* __restrict__ cell_of, int* __restrict__ cursor,
                 float4* __restrict__ sorted, int N, int cap) {
    const int b = blockIdx.y;
    const int i = blockIdx.x * GRID_THREADS + threadIdx.x;
    if (i >= N) return;
    const float* r = ref + ((size_t)b * N + i) * 3;
    const int pos = atomicAdd(cursor + (size_t)b * cap + cell_of[(size_t)b * N + i], 1);
    sorted[(size_t)b * N + pos] = make_float4(r[0], r[1], r[2], __int_as_float(i));
}

template <int K>
__device__ __forceinline__ void lex_insert(float (&bd)[K], int (&bi)[K], float d, int n) {
    if (!(d < bd[K - 1] || (d == bd[K - 1] && n < bi[K - 1]))) return;
    bd[K - 1] = d;
    bi[K - 1] = n;
#pragma unroll
    for (int p = K - 1; p > 0; --p) {
        const bool lt = bd[p] < bd[p - 1] || (bd[p] == bd[p - 1] && bi[p] < bi[p - 1]);
        if (lt) {
            const float td = bd[p];
            bd[p] = bd[p - 1];
            bd[p - 1] = td;
            const int ti = bi[p];
            bi[p] = bi[p - 1];
            bi[p - 1] = ti;
        }
    }
}

// (5) thread = query.
template <int K>
__global__ void __launch_bounds__(GRID_QT)
knn3_grid_kernel(const float4* __restrict__ sorted, const int* __restrict__ start,
                 const GridParams* __restrict__ params, const float* __restrict__ qry, float* __restrict__ dist_out,
                 int64_t* __restrict__ idx_out, int N, int S, int cap) {
    const int b = blockIdx.y;
    const int s = blockIdx.x * GRID_QT + threadIdx.x;
    if (s >= S) return;
    const GridParams gp = params[b];
    const float4* pts = sorted + (size_t)b * N;
    const int* st = start + (size_t)b * (cap + 1);
    const float* q = qry + ((size_t)b * S + s) * 3;
    const float qx = q[0], qy = q[1], qz = q[2];
    const float qn = sqnorm3(qx, qy, qz);
    float bd[K];
    int bi[K];
#pragma unroll
    for (int k = 0; k < K; ++k) {
        bd[k] = INFINITY;
        bi[k] = 0x7fffffff;
    }
    auto scan = [&](int c0, int c1) {  // every point of cells c0..c1 (one x-run of a grid row)
        const int p1 = st[c1 + 1];
        for (int p = st[c0]; p < p1; ++p) {
            const float4 r = pts[p];
            float dot = __fmul_rn(qx, r.x);
            dot = __fmaf_rn(qy, r.y, dot);
            dot = __fmaf_rn(qz, r.z, dot);
            const float d = sqdist_from_dot(dot, qn, sqnorm3(r.x, r.y, r.z));
            lex_insert<K>(bd, bi, d, __float_as_int(r.w));
        }
    };
    const int g[3] = {gp.gx, gp.gy, gp.gz};
    const float mn[3] = {gp.minx, gp.miny, gp.minz};
    const float qv[3] = {qx, qy, qz};
    int lo[3], hi[3], plo[3] = {0, 0, 0}, phi[3] = {-1, -1, -1};
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        const int c = cell_coord(qv[d], mn[d], gp.inv_h, g[d]);
        lo[d] = max(c - 1, 0);
        hi[d] = min(c + 1, g[d] - 1);
    }
    const float eps = 64.0f * 5.9604645e-8f * (qn + gp.max_rn) + 1e-30f;
    for (;;) {
        const bool have_prev = phi[0] >= plo[0];
        for (int z = lo[2]; z <= hi[2]; ++z) {
            for (int y = lo[1]; y <= hi[1]; ++y) {
                const int row = (z * g[1] + y) * g[0];
                const bool inside = have_prev && z >= plo[2] && z <= phi[2] && y >= plo[1] && y <= phi[1];
                if (!inside) {
                    scan(row + lo[0], row + hi[0]);
                } else {
                    if (lo[0] < plo[0]) scan(row + lo[0], row + plo[0] - 1);
                    if (hi[0] > phi[0]) scan(row + phi[0] + 1, row + hi[0]);
                }
            }
        }
        int nlo[3], nhi[3];
        if (bd[K - 1] < INFINITY) {
            // every point whose expanded distance can be <= bd[K-1] lies within R of the query
            const float R = sqrtf(fmaxf(bd[K - 1], 0.0f) + eps) * 1.000001f;
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                const float m = R + 9.5367432e-7f * (fabsf(qv[d]) + R) + 1e-37f;  // outward rounding margin (2^-20)
                nlo[d] = min(lo[d], cell_coord(qv[d] - m, mn[d], gp.inv_h, g[d]));
                nhi[d] = max(hi[d], cell_coord(qv[d] + m, mn[d], gp.inv_h, g[d]));
            }
        } else {
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                nlo[d] = max(lo[d] - 1, 0);
                nhi[d] = min(hi[d] + 1, g[d] - 1);
            }
        }
        bool same = true;
#pragma unroll
        for (int d = 0; d < 3; ++d) same = same && nlo[d] == lo[d] && nhi[d] == hi[d];
        if (same) break;
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            plo[d] = lo[d];
            phi[d] = hi[d];
            lo[d] = nlo[d];
            hi[d] = nhi[d];
        }
    }
    const size_t o = ((size_t)b * S + s) * K;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        if (dist_out) dist_out[o + k] = bd[k];
        idx_out[o + k] = bi[k];
    }
}

struct GridLayout {
    size_t part, cell_of, start, cursor, params, sorted, total;
    int cap, target;
};

static GridLayout grid_layout(int64_t B, int64_t N, int64_t K) {
    GridLayout L;
    const int64_t ppc = K >= 8 ? K / 4 : 2;
    int64_t target = N / ppc;
    if (target < 8) target = 8;
    if (target > (1 << 21)) target = 1 << 21;
    L.target = (int)target;
    L.cap = (int)(2 * target);
    auto align = [](size_t x) { return (x + 255) & ~(size_t)255; };
    size_t off = 0;
    L.part = off;    off = align(off + (size_t)B * GRID_PARTS * 8 * sizeof(float));
    L.cell_of = off; off = align(off + (size_t)B * N * sizeof(int));
    L.start = off;   off = align(off + (size_t)B * (L.cap + 1) * sizeof(int));
    L.cursor = off;  off = align(off + (size_t)B * L.cap * sizeof(int));
    L.params = off;  off = align(off + (size_t)B * sizeof(GridParams));
    L.sorted = off;  off = align(off + (size_t)B * N * sizeof(float4));
    L.total = off;
    return L;
}

template <int K>
static int launch_knn3_grid(const float* ref, const float* qry, float* dist_out, int64_t* idx_out, char* ws,
                            const GridLayout& L, int B, int N, int S, cudaStream_t st) {
    float* part = reinterpret_cast<float*>(ws + L.part);
    int* cell_of = reinterpret_cast<int*>(ws + L.cell_of);
    int* start = reinterpret_cast<int*>(ws + L.start);
    int* cursor = reinterpret_cast<int*>(ws + L.cursor);
    GridParams* params = reinterpret_cast<GridParams*>(ws + L.params);
    float4* sorted = reinterpret_cast<float4*>(ws + L.sorted);
    const dim3 per_point((unsigned)ceil_div(N, GRID_THREADS), (unsigned)B);
    grid_bbox_kernel<<<dim3(GRID_PARTS, B), GRID_THREADS, 0, st>>>(ref, part, start, N, L.cap);
    grid_count_kernel<<<per_point, GRID_THREADS, 0, st>>>(ref, part, cell_of, start, N, L.cap, L.target);
    grid_scan_kernel<<<B, 1024, 0, st>>>(part, start, cursor, params, N, L.cap, L.target);
    grid_fill_kernel<<<per_point, GRID_THREADS, 0, st>>>(ref, cell_of, cursor, sorted, N, L.cap);
    knn3_grid_kernel<K><<<dim3((unsigned)ceil_div(S, GRID_QT), (unsigned)B), GRID_QT, 0, st>>>(
        sorted, start, params, qry, dist_out, idx_out, N, S, L.cap);
    MPC_LAUNCH_CHECK();
    return MPC_OK;
}

}  // namespace mpc

MPC_API int mpc_knn3_grid_workspace_bytes(int64_t B, int64_t N, int64_t S, int64_t K, int64_t* bytes_out) {
    if (B < 0 || N <= 0 || S < 0 || K <= 0 || !bytes_out) return MPC_ERR_INVALID;
    *bytes_out = (int64_t)mpc::grid_layout(B, N, K).total;
    return MPC_OK;
}

MPC_API int mpc_knn3_grid_f32(const float* ref, const float* qry, float* dist_out, int64_t* idx_out, void* workspace,
                              int64_t workspace_bytes, int64_t B, int64_t N, int64_t S, int64_t K,
                              mpc_stream_t stream) {
    using namespace mpc;
    if (B < 0 || N <= 0 || S < 0 || K <= 0 || K > N) return MPC_ERR_INVALID;
    if (B == 0 || S == 0) return MPC_OK;
    if (!ref || !qry || !idx_out || !workspace) return MPC_ERR_INVALID;
    if (K > 32 || N > (1 << 30) || S > INT32_MAX || B > 65535) return MPC_ERR_UNSUPPORTED;
    const GridLayout L = grid_layout(B, N, K);
    if ((size_t)workspace_bytes < L.total || (reinterpret_cast<uintptr_t>(workspace) & 255)) return MPC_ERR_INVALID;
    cudaStream_t st = (cudaStream_t)stream;
    char* ws = static_cast<char*>(workspace);
    const int b = (int)B, n = (int)N, s = (int)S;
    switch (K) {
        case 1: return launch_knn3_grid<1>(ref, qry, dist_out, idx_out, ws, L, b, n, s, st);
        case 3: return launch_knn3_grid<3>(ref, qry, dist_out, idx_out, ws, L, b, n, s, st);
        case 8: return launch_knn3_grid<8>(ref, qry, dist_out, idx_out, ws, L, b, n, s, st);
        case 9: return launch_knn3_grid<9>(ref, qry, dist_out, idx_out, ws, L, b, n, s, st);
        case 16: return launch_knn3_grid<16>(ref, qry, dist_out, idx_out, ws, L, b, n, s, st);
        case 32: return launch_knn3_grid<32>(ref, qry, dist_out, idx_out, ws, L, b, n, s, st);
        default: return MPC_ERR_UNSUPPORTED;  // the host wrapper rounds K up to a supported list length and slices
    }
}
