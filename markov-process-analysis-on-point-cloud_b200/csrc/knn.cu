// Brute-force k nearest neighbours and ball query for sm_100a.  See include/mpc_b200.h (mpc_knn_f32,
// mpc_ball_query_f32) for the contract and the reference lines replaced (pointnet2_utils.py:112-134,190-222).
//
// The reference materialises the [B,S,N] distance matrix (537 MB per call at 32x2048x2048) and runs a full
// topk over it.  Here a thread owns one query and a sorted K-list in registers; reference points stream
// through shared memory in tiles, so HBM sees each coordinate once per CTA and the distance matrix never
// exists.  Arithmetic must stay fp32 FFMA in the oracle's order (indices have to be bit-exact), so the
// tensor cores are deliberately not used.
#include "common.cuh"

namespace mpc {

// Sorted insertion of (d, n) into an ascending K-list held in registers.  Strict < keeps equal distances
// in ascending index order (candidates arrive in ascending n).
template <int K>
__device__ __forceinline__ void topk_insert(float (&bd)[K], int (&bi)[K], float d, int n) {
    bd[K - 1] = d;
    bi[K - 1] = n;
#pragma unroll
    for (int p = K - 1; p > 0; --p) {
        if (bd[p] < bd[p - 1]) {
            float td = bd[p];
            bd[p] = bd[p - 1];
            bd[p - 1] = td;
            int ti = bi[p];
            bi[p] = bi[p - 1];
            bi[p - 1] = ti;
        }
    }
}

// Candidates that beat a thread's current K-th distance are not inserted immediately: in a warp of 32
// independent queries some lane passes the threshold on ~70% of the reference points, so an immediate
// insertion network would run (mostly masked off) on almost every iteration.  Instead each thread appends the
// candidate to a small per-thread queue in shared memory, and the warp drains all queues together when any of
// them is close to full: the insertion network then runs ~20x less often and with most lanes busy.  Arrival
// order (ascending index) is preserved, so equal distances still resolve to the lower index.
constexpr int KNN_FEW_MAX = 48;  // listed queries per cloud up to which knn_few_kernel (one CTA per query) takes them
constexpr int KNN_THREADS = 128;
constexpr int KNN_Q = 8;      // queue slots per thread
constexpr int KNN_QFLUSH = 4; // drain when any lane holds more than this many (<= KNN_Q - unroll)

struct CandQueue {
    float d[KNN_Q][KNN_THREADS];
    int i[KNN_Q][KNN_THREADS];
};

template <int K>
__device__ __forceinline__ void drain_queue(CandQueue& q, int& qn, float (&bd)[K], int (&bi)[K], float& thr) {
    const int m = __reduce_max_sync(0xffffffffu, qn);
    for (int e = 0; e < m; ++e) {
        if (e < qn) {
            const float d = q.d[e][threadIdx.x];
            if (d < bd[K - 1]) topk_insert<K>(bd, bi, d, q.i[e][threadIdx.x]);
        }
    }
    qn = 0;
    thr = bd[K - 1];
}

// ---- C == 3 ------------------------------------------------------------------------------------------------
// A CTA serves 32 queries with SPLIT warps: lane = query, warp w scans the w-th part of every reference tile
// (so the shared-memory reads stay warp-wide broadcasts) and keeps its own sorted K-list; the lists are merged
// at the end with a lexicographic (distance, index) comparison.  SPLIT is picked on the host so that even the
// small query sets of the coarse states (128 queries per cloud) put ~150k threads in flight: with one thread per
// query those calls were pure latency chains (one warp per scheduler, 2048 dependent iterations).
template <int K>
__device__ __forceinline__ void topk_insert_lex(float (&bd)[K], int (&bi)[K], float d, int n) {
    bd[K - 1] = d;
    bi[K - 1] = n;
#pragma unroll
    for (int p = K - 1; p > 0; --p) {
        const bool lt = bd[p] < bd[p - 1] || (bd[p] == bd[p - 1] && bi[p] < bi[p - 1]);
        if (lt) {
            float td = bd[p];
            bd[p] = bd[p - 1];
            bd[p - 1] = td;
            int ti = bi[p];
            bi[p] = bi[p - 1];
            bi[p - 1] = ti;
        }
    }
}

template <int K, int SPLIT>
__global__ void __launch_bounds__(32 * SPLIT)
knn3_kernel(const float* __restrict__ ref, const float* __restrict__ qry, float* __restrict__ dist_out,
            int64_t* __restrict__ idx_out, int N, int S) {
    constexpr int THREADS = 32 * SPLIT;
    constexpr int KNN3_TILE = SPLIT >= 16 ? 1024 : 2048;  // reference points per tile (float4 x,y,z,|r|^2)
    // one static block <= 48 KB: [tile | candidate queues]; reused as the merge buffer at the end
    // (SPLIT * K * 32 * 8 B, the host only launches combinations with SPLIT * K <= 128)
    __shared__ __align__(16) unsigned char raw[KNN3_TILE * 16 + KNN_Q * THREADS * 8];
    float4* tile = reinterpret_cast<float4*>(raw);
    float (*qd)[THREADS] = reinterpret_cast<float (*)[THREADS]>(raw + KNN3_TILE * 16);
    int (*qi)[THREADS] = reinterpret_cast<int (*)[THREADS]>(raw + KNN3_TILE * 16 + KNN_Q * THREADS * 4);
    const int b = blockIdx.y;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int s = blockIdx.x * 32 + lane;
    const bool active = s < S;
    const float* rb = ref + (size_t)b * N * 3;
    float qx = 0.f, qy = 0.f, qz = 0.f;
    if (active) {
        const float* q = qry + ((size_t)b * S + s) * 3;
        qx = q[0];
        qy = q[1];
        qz = q[2];
    }
    const float qn = sqnorm3(qx, qy, qz);
    float bd[K];
    int bi[K];
#pragma unroll
    for (int k = 0; k < K; ++k) {
        bd[k] = __int_as_float(0x7f800000);  // +inf
        bi[k] = 0x7fffffff;
    }
    float thr = active ? __int_as_float(0x7f800000) : -__int_as_float(0x7f800000);
    int cnt = 0;
    auto drain = [&]() {
        const int m = __reduce_max_sync(0xffffffffu, cnt);
        for (int e = 0; e < m; ++e) {
            if (e < cnt) {
                const float d = qd[e][threadIdx.x];
                if (d < bd[K - 1]) topk_insert<K>(bd, bi, d, qi[e][threadIdx.x]);
            }
        }
        cnt = 0;
        thr = bd[K - 1];
    };
    for (int t0 = 0; t0 < N; t0 += KNN3_TILE) {
        const int tn = min(KNN3_TILE, N - t0);
        __syncthreads();
        for (int j = threadIdx.x; j < tn; j += THREADS) {
            const float* r = rb + (size_t)(t0 + j) * 3;
            float x = r[0], y = r[1], z = r[2];
            tile[j] = make_float4(x, y, z, sqnorm3(x, y, z));
        }
        __syncthreads();
        const int part = (((tn + SPLIT - 1) / SPLIT) + 3) & ~3;
        const int jb = w * part, je = min(tn, jb + part);
        for (int j0 = jb; j0 < je; j0 += 4) {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int j = j0 + u;
                const float4 r = tile[j < je ? j : je - 1];  // warp-wide broadcast
                float dot = __fmul_rn(qx, r.x);
                dot = __fmaf_rn(qy, r.y, dot);
                dot = __fmaf_rn(qz, r.z, dot);
                const float d = sqdist_from_dot(dot, qn, r.w);
                if (j < je && d < thr) {
                    qd[cnt][threadIdx.x] = d;
                    qi[cnt][threadIdx.x] = t0 + j;
                    ++cnt;
                }
            }
            if (__any_sync(0xffffffffu, cnt > KNN_QFLUSH)) drain();
        }
    }
    drain();
    if (SPLIT > 1) {  // merge the SPLIT partial lists of each query (warp 0 does the merge)
        __syncthreads();
        float* md = reinterpret_cast<float*>(raw);             // [SPLIT][K][32]
        int* mi = reinterpret_cast<int*>(raw) + SPLIT * K * 32;
        if (w > 0) {
#pragma unroll
            for (int k = 0; k < K; ++k) {
                md[(w * K + k) * 32 + lane] = bd[k];
                mi[(w * K + k) * 32 + lane] = bi[k];
            }
        }
        __syncthreads();
        if (w == 0) {
            for (int w2 = 1; w2 < SPLIT; ++w2) {
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    const float d = md[(w2 * K + k) * 32 + lane];
                    const int n = mi[(w2 * K + k) * 32 + lane];
                    const bool lt = d < bd[K - 1] || (d == bd[K - 1] && n < bi[K - 1]);
                    if (lt) topk_insert_lex<K>(bd, bi, d, n);
                }
            }
        }
    }
    if (active && w == 0) {
        const size_t o = ((size_t)b * S + s) * K;
#pragma unroll
        for (int k = 0; k < K; ++k) {
            if (dist_out) dist_out[o + k] = bd[k];
            idx_out[o + k] = bi[k];
        }
    }
}

// ---- C == 64 (the feature width of every large feature-space search in both models) ------------------------
// The query vector lives in 64 registers; reference tiles of 64 points are stored transposed [c][64+4] so one
// broadcast LDS.128 feeds 4 reference points; a thread advances 16 reference points at once (16 independent
// fma chains, each sequential in c exactly as the oracle prescribes).
constexpr int KNN64_TR = 64;
constexpr int KNN64_LD = KNN64_TR + 4;

template <int K>
__global__ void __launch_bounds__(KNN_THREADS)
knn64_kernel(const float* __restrict__ ref, const float* __restrict__ qry, float* __restrict__ dist_out,
             int64_t* __restrict__ idx_out, int N, int S) {
    constexpr int C = 64;
    __shared__ __align__(16) float rt[C * KNN64_LD];
    __shared__ float rn[KNN64_TR];
    __shared__ CandQueue queue;
    const int b = blockIdx.y;
    const int tid = threadIdx.x;
    const int s = blockIdx.x * KNN_THREADS + tid;
    const bool active = s < S;
    const float* rb = ref + (size_t)b * N * C;
    float q[C];
    {
        const float4* qp = reinterpret_cast<const float4*>(qry + ((size_t)b * S + (active ? s : 0)) * C);
#pragma unroll
        for (int c = 0; c < C / 4; ++c) {
            const float4 v = __ldg(qp + c);
            q[4 * c + 0] = v.x;
            q[4 * c + 1] = v.y;
            q[4 * c + 2] = v.z;
            q[4 * c + 3] = v.w;
        }
    }
    float qn = __fmul_rn(q[0], q[0]);
#pragma unroll
    for (int c = 1; c < C; ++c) qn = __fadd_rn(qn, __fmul_rn(q[c], q[c]));
    float bd[K];
    int bi[K];
#pragma unroll
    for (int k = 0; k < K; ++k) {
        bd[k] = __int_as_float(0x7f800000);
        bi[k] = 0;
    }
    float thr = active ? __int_as_float(0x7f800000) : -__int_as_float(0x7f800000);
    int cnt = 0;
    for (int t0 = 0; t0 < N; t0 += KNN64_TR) {
        const int tn = min(KNN64_TR, N - t0);
        __syncthreads();
        // coalesced float4 reads of 64 rows x 64 channels, transposed into [c][j]
        for (int i = tid; i < KNN64_TR * (C / 4); i += KNN_THREADS) {
            const int j = i / (C / 4), c4 = (i - j * (C / 4)) * 4;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (j < tn) v = __ldg(reinterpret_cast<const float4*>(rb + (size_t)(t0 + j) * C + c4));
            rt[(c4 + 0) * KNN64_LD + j] = v.x;
            rt[(c4 + 1) * KNN64_LD + j] = v.y;
            rt[(c4 + 2) * KNN64_LD + j] = v.z;
            rt[(c4 + 3) * KNN64_LD + j] = v.w;
        }
        __syncthreads();
        if (tid < KNN64_TR) {
            float a = __fmul_rn(rt[tid], rt[tid]);
#pragma unroll 8
            for (int c = 1; c < C; ++c) {
                const float v = rt[c * KNN64_LD + tid];
                a = __fadd_rn(a, __fmul_rn(v, v));
            }
            rn[tid] = a;
        }
        __syncthreads();
#pragma unroll 1
        for (int j0 = 0; j0 < KNN64_TR; j0 += 16) {
            float acc[16];
#pragma unroll
            for (int u = 0; u < 16; ++u) acc[u] = 0.f;
#pragma unroll
            for (int c = 0; c < C; ++c) {
                const float qv = q[c];
                const float4* rp = reinterpret_cast<const float4*>(rt + c * KNN64_LD + j0);
#pragma unroll
                for (int v4 = 0; v4 < 4; ++v4) {
                    const float4 r = rp[v4];
                    acc[4 * v4 + 0] = __fmaf_rn(qv, r.x, acc[4 * v4 + 0]);
                    acc[4 * v4 + 1] = __fmaf_rn(qv, r.y, acc[4 * v4 + 1]);
                    acc[4 * v4 + 2] = __fmaf_rn(qv, r.z, acc[4 * v4 + 2]);
                    acc[4 * v4 + 3] = __fmaf_rn(qv, r.w, acc[4 * v4 + 3]);
                }
            }
#pragma unroll
            for (int u0 = 0; u0 < 16; u0 += 4) {
#pragma unroll
                for (int u = u0; u < u0 + 4; ++u) {
                    const int j = j0 + u;
                    const float d = sqdist_from_dot(acc[u], qn, rn[j]);
                    if (j < tn && d < thr) {
                        queue.d[cnt][tid] = d;
                        queue.i[cnt][tid] = t0 + j;
                        ++cnt;
                    }
                }
                if (__any_sync(0xffffffffu, cnt > KNN_QFLUSH)) drain_queue<K>(queue, cnt, bd, bi, thr);
            }
        }
    }
    drain_queue<K>(queue, cnt, bd, bi, thr);
    if (active) {
        const size_t o = ((size_t)b * S + s) * K;
#pragma unroll
        for (int k = 0; k < K; ++k) {
            if (dist_out) dist_out[o + k] = bd[k];
            idx_out[o + k] = bi[k];
        }
    }
}

// ---- C == 64, two queries per thread ----------------------------------------------------------------------------
// knn64_kernel is shared-memory-issue bound (one broadcast LDS.128 per 4 FFMA; ncu: 74% of the LSU shared
// wavefront peak).  Holding TWO query vectors in registers halves the loads per FMA: per channel step a thread
// issues 2 LDS.128 (8 reference points) and 16 FFMA.  CTAs are 64 threads (128 queries) so that the mid-sized
// query sets still spread over every SM; used when that gives at least one CTA per SM.
constexpr int KNN64X2_THREADS = 64;

template <int K>
__global__ void __launch_bounds__(KNN64X2_THREADS, 4)
knn64x2_kernel(const float* __restrict__ ref, const float* __restrict__ qry, float* __restrict__ dist_out,
               int64_t* __restrict__ idx_out, int N, int S) {
    constexpr int C = 64, T = KNN64X2_THREADS;
    __shared__ __align__(16) float rt[C * KNN64_LD];
    __shared__ float rn[KNN64_TR];
    __shared__ float qd[2][KNN_Q][T];
    __shared__ int qi[2][KNN_Q][T];
    const int b = blockIdx.y;
    const int tid = threadIdx.x;
    const int s0 = blockIdx.x * (2 * T) + tid, s1 = s0 + T;
    const bool act0 = s0 < S, act1 = s1 < S;
    const float* rb = ref + (size_t)b * N * C;
    float q0[C], q1[C];
    {
        const float4* p0 = reinterpret_cast<const float4*>(qry + ((size_t)b * S + (act0 ? s0 : 0)) * C);
        const float4* p1 = reinterpret_cast<const float4*>(qry + ((size_t)b * S + (act1 ? s1 : 0)) * C);
#pragma unroll
        for (int c = 0; c < C / 4; ++c) {
            const float4 u = __ldg(p0 + c), v = __ldg(p1 + c);
            q0[4 * c + 0] = u.x; q0[4 * c + 1] = u.y; q0[4 * c + 2] = u.z; q0[4 * c + 3] = u.w;
            q1[4 * c + 0] = v.x; q1[4 * c + 1] = v.y; q1[4 * c + 2] = v.z; q1[4 * c + 3] = v.w;
        }
    }
    float qn0 = __fmul_rn(q0[0], q0[0]), qn1 = __fmul_rn(q1[0], q1[0]);
#pragma unroll
    for (int c = 1; c < C; ++c) {
        qn0 = __fadd_rn(qn0, __fmul_rn(q0[c], q0[c]));
        qn1 = __fadd_rn(qn1, __fmul_rn(q1[c], q1[c]));
    }
    float bd0[K], bd1[K];
    int bi0[K], bi1[K];
#pragma unroll
    for (int k = 0; k < K; ++k) {
        bd0[k] = bd1[k] = __int_as_float(0x7f800000);
        bi0[k] = bi1[k] = 0;
    }
    const float ninf = -__int_as_float(0x7f800000);
    float thr0 = act0 ? -ninf : ninf, thr1 = act1 ? -ninf : ninf;
    int cnt0 = 0, cnt1 = 0;
    auto drain = [&](int which, int& cnt, float (&bd)[K], int (&bi)[K], float& thr) {
        const int m = __reduce_max_sync(0xffffffffu, cnt);
        for (int e = 0; e < m; ++e) {
            if (e < cnt) {
                const float d = qd[which][e][tid];
                if (d < bd[K - 1]) topk_insert<K>(bd, bi, d, qi[which][e][tid]);
            }
        }
        cnt = 0;
        thr = bd[K - 1];
    };
    for (int t0 = 0; t0 < N; t0 += KNN64_TR) {
        const int tn = min(KNN64_TR, N - t0);
        __syncthreads();
        for (int i = tid; i < KNN64_TR * (C / 4); i += T) {
            const int j = i / (C / 4), c4 = (i - j * (C / 4)) * 4;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (j < tn) v = __ldg(reinterpret_cast<const float4*>(rb + (size_t)(t0 + j) * C + c4));
            rt[(c4 + 0) * KNN64_LD + j] = v.x;
            rt[(c4 + 1) * KNN64_LD + j] = v.y;
            rt[(c4 + 2) * KNN64_LD + j] = v.z;
            rt[(c4 + 3) * KNN64_LD + j] = v.w;
        }
        __syncthreads();
        {
            const int j = tid;  // T == KNN64_TR == 64: one reference norm per thread
            float a = __fmul_rn(rt[j], rt[j]);
#pragma unroll 8
            for (int c = 1; c < C; ++c) {
                const float v = rt[c * KNN64_LD + j];
                a = __fadd_rn(a, __fmul_rn(v, v));
            }
            rn[j] = a;
        }
        __syncthreads();
#pragma unroll 1
        for (int j0 = 0; j0 < KNN64_TR; j0 += 8) {
            float a0[8], a1[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) a0[u] = a1[u] = 0.f;
#pragma unroll
            for (int c = 0; c < C; ++c) {
                const float4 ra = *reinterpret_cast<const float4*>(rt + c * KNN64_LD + j0);
                const float4 rb4 = *reinterpret_cast<const float4*>(rt + c * KNN64_LD + j0 + 4);
                const float u = q0[c], v = q1[c];
                a0[0] = __fmaf_rn(u, ra.x, a0[0]); a0[1] = __fmaf_rn(u, ra.y, a0[1]);
                a0[2] = __fmaf_rn(u, ra.z, a0[2]); a0[3] = __fmaf_rn(u, ra.w, a0[3]);
                a0[4] = __fmaf_rn(u, rb4.x, a0[4]); a0[5] = __fmaf_rn(u, rb4.y, a0[5]);
                a0[6] = __fmaf_rn(u, rb4.z, a0[6]); a0[7] = __fmaf_rn(u, rb4.w, a0[7]);
                a1[0] = __fmaf_rn(v, ra.x, a1[0]); a1[1] = __fmaf_rn(v, ra.y, a1[1]);
                a1[2] = __fmaf_rn(v, ra.z, a1[2]); a1[3] = __fmaf_rn(v, ra.w, a1[3]);
                a1[4] = __fmaf_rn(v, rb4.x, a1[4]); a1[5] = __fmaf_rn(v, rb4.y, a1[5]);
                a1[6] = __fmaf_rn(v, rb4.z, a1[6]); a1[7] = __fmaf_rn(v, rb4.w, a1[7]);
            }
#pragma unroll
            for (int u0 = 0; u0 < 8; u0 += 4) {
#pragma unroll
                for (int u = u0; u < u0 + 4; ++u) {
                    const int j = j0 + u;
                    const float d0 = sqdist_from_dot(a0[u], qn0, rn[j]);
                    const float d1 = sqdist_from_dot(a1[u], qn1, rn[j]);
                    if (j < tn && d0 < thr0) {
                        qd[0][cnt0][tid] = d0;
                        qi[0][cnt0][tid] = t0 + j;
                        ++cnt0;
                    }
                    if (j < tn && d1 < thr1) {
                        qd[1][cnt1][tid] = d1;
                        qi[1][cnt1][tid] = t0 + j;
                        ++cnt1;
                    }
                }
                if (__any_sync(0xffffffffu, cnt0 > KNN_QFLUSH)) drain(0, cnt0, bd0, bi0, thr0);
                if (__any_sync(0xffffffffu, cnt1 > KNN_QFLUSH)) drain(1, cnt1, bd1, bi1, thr1);
            }
        }
    }
    drain(0, cnt0, bd0, bi0, thr0);
    drain(1, cnt1, bd1, bi1, thr1);
    if (act0) {
        const size_t o = ((size_t)b * S + s0) * K;
#pragma unroll
        for (int k = 0; k < K; ++k) {
            if (dist_out) dist_out[o + k] = bd0[k];
            idx_out[o + k] = bi0[k];
        }
    }
    if (act1) {
        const size_t o = ((size_t)b * S + s1) * K;
#pragma unroll
        for (int k = 0; k < K; ++k) {
            if (dist_out) dist_out[o + k] = bd1[k];
            idx_out[o + k] = bi1[k];
        }
    }
}

// ---- feature-space search, register-tiled (C % 64 == 0, C <= 256) ---------------------------------------------
// knn64 / knn64x2 keep whole query vectors in registers and stream reference points through broadcast shared-memory
// loads: 4-8 FFMA per LDS.128, shared-memory-issue bound (ncu) at 30-46 % of the FP32 peak.  This variant tiles the
// distance computation like an SGEMM: a CTA of 128 threads owns 128 queries (transposed in shared memory for the whole
// kernel) and walks the reference set in tiles of 64 points; a thread accumulates an 8 x 8 block of dot products
// (64 FFMA per 4 LDS.128 = 16 per load), every dot product still ONE fma chain over c = 0..C-1 from 0 -- the
// oracle's order, so indices stay bit-exact.  The finished 128 x 64 distance tile goes through shared memory to the
// selection phase, where thread = query scans its row against its current K-th distance (the same queued insertion
// as the other kernels).
__device__ __forceinline__ unsigned long long pack2(float lo, float hi) {
    unsigned long long d;
    asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(lo), "f"(hi));
    return d;
}
__device__ __forceinline__ void unpack2(unsigned long long v, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ unsigned long long ffma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}

constexpr int KT_THREADS = 128;
constexpr int KT_Q = 128;   // queries per CTA
constexpr int KT_R = 64;    // reference points per tile
constexpr int KT_CK = 32;   // channels per shared-memory chunk of the reference tile
constexpr int KT_DH = 32;   // reference points per selection pass: the distance tile is handed over in two halves
                            // [128 queries][32], XOR-swizzled, so that a CTA needs 58 KB at C = 64 and THREE fit on an SM

template <int K>
__global__ void __launch_bounds__(KT_THREADS, 3)
knn_tiled_kernel(const float* __restrict__ ref, const float* __restrict__ qry, float* __restrict__ dist_out,
                 int64_t* __restrict__ idx_out, int N, int S, int C, const int* __restrict__ qlist,
                 const int* __restrict__ qcount) {
    // Indirect mode (qlist != null): the CTA grid covers all S query slots of a cloud, but only the first qcount[b]
    // slots are live and slot i stands for query row qlist[b*S + i] (the rows mpc_knn_tc_f32's filter could not decide);
    // CTAs beyond the live range leave at once.
    const int S_live = qcount ? qcount[blockIdx.y] : S;
    if ((int)(blockIdx.x * KT_Q) >= S_live) return;
    if (qcount && S_live <= KNN_FEW_MAX) return;  // a handful of listed queries: knn_few_kernel serves them
    extern __shared__ __align__(16) float sm[];
    float* qt = sm;                              // [C][KT_Q]   queries, transposed
    float* rt = qt + (size_t)C * KT_Q;           // [KT_CK][KT_R] reference chunk, transposed
    float* dt = rt + KT_CK * KT_R;               // [KT_Q][KT_DH] half distance tile, column j of row i at j ^ (i & 31)
    float* qn_s = dt + KT_Q * KT_DH;             // [KT_Q]
    float* rn_s = qn_s + KT_Q;                   // [KT_R]
    __shared__ CandQueue queue;
    const int b = blockIdx.y, tid = threadIdx.x;
    const int tx = tid & 7, ty = tid >> 3;       // 8 x 16 thread grid: columns tx*4+{0..3}, 32+tx*4+{0..3}; rows alike
    const int q0 = blockIdx.x * KT_Q;
    const float* rb = ref + (size_t)b * N * C;
    const float* qb = qry + (size_t)b * S * C;
    // Transposed tiles are stored with the 4-column groups of row c XOR-swizzled by (c / 4): lanes that read 16
    // consecutive float4 of one point (coalesced 256-byte global rows) then hit 16 different bank groups, and the
    // compute loop's float4 reads of one row stay conflict-free (the XOR term is uniform per row).
    auto swz = [](int c, int col, int groups) { return ((((col >> 2) ^ (c >> 2)) & (groups - 1)) << 2) | (col & 3); };
    // queries -> shared (transposed), once
    for (int i = tid; i < KT_Q * (C / 4); i += KT_THREADS) {
        const int c4 = (i % (C / 4)) * 4, q = i / (C / 4);
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (q0 + q < S_live) {
            const int qrow = qlist ? qlist[(size_t)b * S + q0 + q] : q0 + q;
            v = __ldg(reinterpret_cast<const float4*>(qb + (size_t)qrow * C + c4));
        }
        const int col = swz(c4, q, KT_Q / 4);
        qt[(c4 + 0) * KT_Q + col] = v.x;
        qt[(c4 + 1) * KT_Q + col] = v.y;
        qt[(c4 + 2) * KT_Q + col] = v.z;
        qt[(c4 + 3) * KT_Q + col] = v.w;
    }
    __syncthreads();
    {   // |q|^2, sequential non-fused like the oracle; thread = query
        float a = __fmul_rn(qt[swz(0, tid, KT_Q / 4)], qt[swz(0, tid, KT_Q / 4)]);
        for (int c = 1; c < C; ++c) {
            const float v = qt[c * KT_Q + swz(c, tid, KT_Q / 4)];
            a = __fadd_rn(a, __fmul_rn(v, v));
        }
        qn_s[tid] = a;
    }
    const bool active = q0 + tid < S_live;
    const int s = !active ? 0 : (qlist ? qlist[(size_t)b * S + q0 + tid] : q0 + tid);
    float bd[K];
    int bi[K];
#pragma unroll
    for (int k = 0; k < K; ++k) {
        bd[k] = __int_as_float(0x7f800000);
        bi[k] = 0;
    }
    float thr = active ? __int_as_float(0x7f800000) : -__int_as_float(0x7f800000);
    int qcnt = 0;
    for (int t0 = 0; t0 < N; t0 += KT_R) {
        const int tn = min(KT_R, N - t0);
        // accumulators as packed pairs (acc[i][2m], acc[i][2m+1]): Blackwell's FFMA2 (fma.rn.f32x2) retires two IEEE
        // fp32 FMAs per lane per issue slot -- bit-identical to two scalar fmaf, twice the SIMT FP32 rate
        unsigned long long acc2[8][4];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int m = 0; m < 4; ++m) acc2[i][m] = 0ull;
        float rnorm = 0.f;  // thread j < 64: running |r_j|^2 across the channel chunks
        for (int c0 = 0; c0 < C; c0 += KT_CK) {
            __syncthreads();  // previous chunk / previous tile's selection are done with rt, dt
            for (int i = tid; i < KT_R * (KT_CK / 4); i += KT_THREADS) {
                const int c4 = (i % (KT_CK / 4)) * 4, j = i / (KT_CK / 4);
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (j < tn) v = __ldg(reinterpret_cast<const float4*>(rb + (size_t)(t0 + j) * C + c0 + c4));
                const int col = swz(c4, j, KT_R / 4);
                rt[(c4 + 0) * KT_R + col] = v.x;
                rt[(c4 + 1) * KT_R + col] = v.y;
                rt[(c4 + 2) * KT_R + col] = v.z;
                rt[(c4 + 3) * KT_R + col] = v.w;
            }
            __syncthreads();
            if (tid < KT_R) {
                int c = 0;
                if (c0 == 0) {
                    const float v0 = rt[swz(0, tid, KT_R / 4)];
                    rnorm = __fmul_rn(v0, v0);
                    c = 1;
                }
                for (; c < KT_CK; ++c) {
                    const float v = rt[c * KT_R + swz(c, tid, KT_R / 4)];
                    rnorm = __fadd_rn(rnorm, __fmul_rn(v, v));
                }
            }
            const float* qc = qt + (size_t)c0 * KT_Q;
#pragma unroll 4
            for (int c = 0; c < KT_CK; ++c) {
                const int g = c >> 2, gq = (c0 + c) >> 2;  // swizzle terms: chunk-local row (rt), global channel (qt)
                const float4 qa = *reinterpret_cast<const float4*>(qc + c * KT_Q + (((ty) ^ gq) & 31) * 4);
                const float4 qb4 = *reinterpret_cast<const float4*>(qc + c * KT_Q + (((16 + ty) ^ gq) & 31) * 4);
                const float4 ra = *reinterpret_cast<const float4*>(rt + c * KT_R + (((tx) ^ g) & 15) * 4);
                const float4 rb4 = *reinterpret_cast<const float4*>(rt + c * KT_R + (((8 + tx) ^ g) & 15) * 4);
                const float qv[8] = {qa.x, qa.y, qa.z, qa.w, qb4.x, qb4.y, qb4.z, qb4.w};
                const unsigned long long rp[4] = {pack2(ra.x, ra.y), pack2(ra.z, ra.w), pack2(rb4.x, rb4.y),
                                                  pack2(rb4.z, rb4.w)};
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const unsigned long long qq = pack2(qv[i], qv[i]);
#pragma unroll
                    for (int m = 0; m < 4; ++m) acc2[i][m] = ffma2(qq, rp[m], acc2[i][m]);
                }
            }
        }
        if (tid < KT_R) rn_s[tid] = rnorm;
        __syncthreads();
        // distance tile: d = ((-2 dot) + |q|^2) + |r|^2, the reference's order; handed to the selection phase in two
        // halves of 32 reference points (ascending index order is preserved: points 0..31, then 32..63)
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            if (half) __syncthreads();  // the first half has been scanned
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int qi = (i < 4 ? 0 : 60) + ty * 4 + i;  // rows ty*4+{0..3}, 64+ty*4+{0..3}
                const float qn = qn_s[qi];
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                    const int j = half * 4 + jj;      // accumulator column; reference point half*32 + tx*4 + jj
                    const int col = tx * 4 + jj;      // column inside the half tile
                    float lo, hi;
                    unpack2(acc2[i][j >> 1], lo, hi);
                    dt[qi * KT_DH + (col ^ (qi & 31))] =
                        sqdist_from_dot((j & 1) ? hi : lo, qn, rn_s[half * KT_DH + col]);
                }
            }
            __syncthreads();
            // selection: thread = query, candidates in ascending index order
            const float* row = dt + tid * KT_DH;
            const int sw = tid & 31;
#pragma unroll 1
            for (int j0 = 0; j0 < KT_DH; j0 += 4) {
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int j = j0 + u;
                    const float d = row[j ^ sw];
                    if (half * KT_DH + j < tn && d < thr) {
                        queue.d[qcnt][tid] = d;
                        queue.i[qcnt][tid] = t0 + half * KT_DH + j;
                        ++qcnt;
                    }
                }
                if (__any_sync(0xffffffffu, qcnt > KNN_QFLUSH)) drain_queue<K>(queue, qcnt, bd, bi, thr);
            }
        }
    }
    drain_queue<K>(queue, qcnt, bd, bi, thr);
    if (active) {
        const size_t o = ((size_t)b * S + s) * K;
#pragma unroll
        for (int k = 0; k < K; ++k) {
            if (dist_out) dist_out[o + k] = bd[k];
            idx_out[o + k] = bi[k];
        }
    }
}

// ---- generic C ------------------------------------------------------------------------------------------------
// CTA = 128 queries.  Queries live in shared memory [q][C+1]; reference tiles of 32 points are stored
// transposed [c][32+4]; a thread advances 8 reference points at once.
constexpr int KNNG_THREADS = KNN_THREADS;
constexpr int KNNG_TR = 32;           // reference points per tile
constexpr int KNNG_LD = KNNG_TR + 4;  // padded row (floats), keeps 16-byte alignment

template <int K>
__global__ void __launch_bounds__(KNNG_THREADS)
knng_kernel(const float* __restrict__ ref, const float* __restrict__ qry, float* __restrict__ dist_out,
            int64_t* __restrict__ idx_out, int N, int S, int C) {
    extern __shared__ __align__(16) float smem[];
    float* qs = smem;                                // [128][C+1]
    float* rt = qs + KNNG_THREADS * (C + 1);         // [C][KNNG_LD]   (offset is a multiple of 4 floats)
    float* rn = rt + (size_t)C * KNNG_LD;            // [KNNG_TR]
    __shared__ CandQueue queue;
    const int b = blockIdx.y;
    const int s0 = blockIdx.x * KNNG_THREADS;
    const int tid = threadIdx.x;
    const int s = s0 + tid;
    const bool active = s < S;
    const float* rb = ref + (size_t)b * N * C;
    const float* qb = qry + ((size_t)b * S + s0) * C;
    const int nq = min(KNNG_THREADS, S - s0);
    for (int i = tid; i < nq * C; i += KNNG_THREADS) {
        int q = i / C, c = i - q * C;
        qs[q * (C + 1) + c] = qb[i];
    }
    __syncthreads();
    const float* myq = qs + tid * (C + 1);
    float qn = 0.f;
    if (active) {
        qn = __fmul_rn(myq[0], myq[0]);
        for (int c = 1; c < C; ++c) qn = __fadd_rn(qn, __fmul_rn(myq[c], myq[c]));
    }
    float bd[K];
    int bi[K];
#pragma unroll
    for (int k = 0; k < K; ++k) {
        bd[k] = __int_as_float(0x7f800000);
        bi[k] = 0;
    }
    float thr = active ? __int_as_float(0x7f800000) : -__int_as_float(0x7f800000);
    int cnt = 0;
    for (int t0 = 0; t0 < N; t0 += KNNG_TR) {
        const int tn = min(KNNG_TR, N - t0);
        __syncthreads();
        for (int i = tid; i < KNNG_TR * C; i += KNNG_THREADS) {
            int j = i / C, c = i - j * C;
            rt[c * KNNG_LD + j] = j < tn ? rb[(size_t)(t0 + j) * C + c] : 0.f;
        }
        __syncthreads();
        if (tid < KNNG_TR) {
            float a = __fmul_rn(rt[tid], rt[tid]);
            for (int c = 1; c < C; ++c) {
                float v = rt[c * KNNG_LD + tid];
                a = __fadd_rn(a, __fmul_rn(v, v));
            }
            rn[tid] = a;
        }
        __syncthreads();
#pragma unroll 1
        for (int j0 = 0; j0 < KNNG_TR; j0 += 8) {
            float acc[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) acc[u] = 0.f;
#pragma unroll 4
            for (int c = 0; c < C; ++c) {
                const float qv = myq[c];
                const float4 r0 = *reinterpret_cast<const float4*>(rt + c * KNNG_LD + j0);
                const float4 r1 = *reinterpret_cast<const float4*>(rt + c * KNNG_LD + j0 + 4);
                acc[0] = __fmaf_rn(qv, r0.x, acc[0]);
                acc[1] = __fmaf_rn(qv, r0.y, acc[1]);
                acc[2] = __fmaf_rn(qv, r0.z, acc[2]);
                acc[3] = __fmaf_rn(qv, r0.w, acc[3]);
                acc[4] = __fmaf_rn(qv, r1.x, acc[4]);
                acc[5] = __fmaf_rn(qv, r1.y, acc[5]);
                acc[6] = __fmaf_rn(qv, r1.z, acc[6]);
                acc[7] = __fmaf_rn(qv, r1.w, acc[7]);
            }
#pragma unroll
            for (int u0 = 0; u0 < 8; u0 += 4) {
#pragma unroll
                for (int u = u0; u < u0 + 4; ++u) {
                    const int j = j0 + u;
                    const float d = sqdist_from_dot(acc[u], qn, rn[j]);
                    if (j < tn && d < thr) {
                        queue.d[cnt][tid] = d;
                        queue.i[cnt][tid] = t0 + j;
                        ++cnt;
                    }
                }
                if (__any_sync(0xffffffffu, cnt > KNN_QFLUSH)) drain_queue<K>(queue, cnt, bd, bi, thr);
            }
        }
    }
    drain_queue<K>(queue, cnt, bd, bi, thr);
    if (active) {
        const size_t o = ((size_t)b * S + s) * K;
#pragma unroll
        for (int k = 0; k < K; ++k) {
            if (dist_out) dist_out[o + k] = bd[k];
            idx_out[o + k] = bi[k];
        }
    }
}

template <int K>
static int launch_knn(const float* ref, const float* qry, float* dist_out, int64_t* idx_out, int B, int N,
                      int S, int C, cudaStream_t st) {
    if (C == 3) {
        // threads per query: enough to put ~150k threads in flight, bounded by the merge buffer (SPLIT*K <= 128)
        // and by the work available per warp (>= 64 reference points)
        const int64_t queries = (int64_t)B * S;
        int split = 1;
        while (split < 16 && queries * split < 150000 && split * 4 * K <= 128 && N / (split * 4) >= 64) split *= 4;
        dim3 grid((unsigned)ceil_div(S, 32), (unsigned)B);
        if (split == 1)
            knn3_kernel<K, 1><<<grid, 32, 0, st>>>(ref, qry, dist_out, idx_out, N, S);
        else if (split == 4)
            knn3_kernel<K, 4><<<grid, 128, 0, st>>>(ref, qry, dist_out, idx_out, N, S);
        else
            knn3_kernel<K, 16><<<grid, 512, 0, st>>>(ref, qry, dist_out, idx_out, N, S);
        MPC_LAUNCH_CHECK();
        return MPC_OK;
    }
    const bool al = (reinterpret_cast<uintptr_t>(ref) & 15u) == 0 && (reinterpret_cast<uintptr_t>(qry) & 15u) == 0;
    if (al && C % 64 == 0 && C <= 256 && K <= 16 && g_knob[4] != 1) {
        // register-tiled variant (knob 4 = 1 forces the query-in-registers / generic kernels below).  Measured, k = 8:
        // C = 128 / 256: 2x the generic kernel; C = 64, 24 000 x 24 000 x 8 clouds: 20.8 vs 21.8 ms, 12 000 x 24 000:
        // 11.6 vs 14.4 ms, 2048 x 2048 x 32 clouds: 0.90 vs 0.79 ms alone but 5 % better for the training step, where
        // the search shares the GPU with the other two branches of its LocalMerge (12.97 vs 13.62 ms, A/B on one box).
        // Both designs are latency-bound (ncu: no eligible warp on ~50 % of the cycles) far below the 72 TFLOP/s FFMA
        // peak of scratch/ubench/ffma_peak.cu; three resident CTAs (58 KB each at C = 64) are what lifted the large
        // clouds from 19-21 to 24-29 TFLOP/s.
        // Occupancy: three CTAs per SM pay off when the launch has many waves of CTAs (24 000-point blocks: +6 % on the
        // fwd+bwd step); with ~1-2 waves (32 clouds x 2048 queries = 512 CTAs) two fatter-share CTAs per SM are as fast
        // or slightly faster inside the training step (11.63 / 11.69 vs 11.70 / 11.74 ms, A/B), so small launches pad
        // their shared-memory request.  knob 5 (KB of padding) overrides.
        const int64_t ctas = (int64_t)B * ceil_div(S, KT_Q);
        const int64_t pad_kb = g_knob[5] > 0 ? g_knob[5] : ((C == 64 && ctas <= 4 * kNumSMs) ? 22 : 0);
        const size_t smem = ((size_t)C * KT_Q + KT_CK * KT_R + KT_Q * KT_DH + KT_Q + KT_R) * sizeof(float) +
                            (size_t)pad_kb * 1024;
        auto kern = knn_tiled_kernel<(K <= 16 ? K : 16)>;
        MPC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        dim3 grid((unsigned)ceil_div(S, KT_Q), (unsigned)B);
        kern<<<grid, KT_THREADS, smem, st>>>(ref, qry, dist_out, idx_out, N, S, C, nullptr, nullptr);
        MPC_LAUNCH_CHECK();
        return MPC_OK;
    }
    if (C == 64 && (reinterpret_cast<uintptr_t>(ref) & 15u) == 0 && (reinterpret_cast<uintptr_t>(qry) & 15u) == 0) {
        if (K <= 16 && (int64_t)B * ceil_div(S, 2 * KNN64X2_THREADS) >= kNumSMs) {
            dim3 grid2((unsigned)ceil_div(S, 2 * KNN64X2_THREADS), (unsigned)B);
            knn64x2_kernel<(K <= 16 ? K : 16)><<<grid2, KNN64X2_THREADS, 0, st>>>(ref, qry, dist_out, idx_out, N, S);
            MPC_LAUNCH_CHECK();
            return MPC_OK;
        }
        dim3 grid((unsigned)ceil_div(S, KNN_THREADS), (unsigned)B);
        knn64_kernel<K><<<grid, KNN_THREADS, 0, st>>>(ref, qry, dist_out, idx_out, N, S);
        MPC_LAUNCH_CHECK();
        return MPC_OK;
    }
    // qs occupies 128*(C+1) floats; round the rt offset up to a multiple of 4 floats by padding C+1 -> handled
    // here by requiring 128*(C+1) % 4 == 0, which holds for every C (128 is a multiple of 4).
    size_t smem = ((size_t)KNNG_THREADS * (C + 1) + (size_t)C * KNNG_LD + KNNG_TR) * sizeof(float);
    if (smem > 210 * 1024) return MPC_ERR_UNSUPPORTED;  // + 8 KB static candidate queue
    auto kern = knng_kernel<K>;
    if (smem > 36 * 1024) MPC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((unsigned)ceil_div(S, KNNG_THREADS), (unsigned)B);
    kern<<<grid, KNNG_THREADS, smem, st>>>(ref, qry, dist_out, idx_out, N, S, C);
    MPC_LAUNCH_CHECK();
    return MPC_OK;
}

// A handful of listed queries (<= KNN_FEW_MAX per cloud): one CTA per query.  The tiled kernel works on 64-query tiles,
// so three undecided queries cost it a full tile's sweep over the reference set on a single SM (4.4 ms at 24 000
// points -- on the critical path of the step); here the reference rows are split over the CTA's threads, every thread
// keeps its own sorted K-list of exact distances (the contract's expression: fma chain over c from 0, sequential
// non-fused norms, ((-2 dot) + |q|^2) + |r|^2) and the lists are merged by K rounds of a block-wide lexicographic
// (distance, index) arg-min: ~20 us.
constexpr int KNN_FEW_THREADS = 1024;  // (the CTA is alone on its query: reference rows over as many threads as fit)

__device__ __forceinline__ unsigned long long few_key(float d, int n) {
    const unsigned u = __float_as_uint(d);
    const unsigned ord = (u & 0x80000000u) ? ~u : (u | 0x80000000u);  // order-preserving map of float to unsigned
    return ((unsigned long long)ord << 32) | (unsigned)n;
}

template <int K>
__global__ void __launch_bounds__(KNN_FEW_THREADS)
knn_few_kernel(const float* __restrict__ ref, const float* __restrict__ qry, float* __restrict__ dist_out,
               int64_t* __restrict__ idx_out, int N, int S, int C, const int* __restrict__ qlist,
               const int* __restrict__ qcount) {
    const int b = blockIdx.y;
    const int live = qcount[b];
    if (live > KNN_FEW_MAX || (int)blockIdx.x >= live) return;
    extern __shared__ __align__(16) float qs[];  // [C]
    __shared__ unsigned long long wbest[KNN_FEW_THREADS / 32];
    __shared__ unsigned long long round_best;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int s = qlist[(size_t)b * S + blockIdx.x];
    const float* q = qry + ((size_t)b * S + s) * C;
    for (int c = tid; c < C; c += KNN_FEW_THREADS) qs[c] = q[c];
    __syncthreads();
    float qn = __fmul_rn(qs[0], qs[0]);
    for (int c = 1; c < C; ++c) qn = __fadd_rn(qn, __fmul_rn(qs[c], qs[c]));
    float bd[K];
    int bi[K];
#pragma unroll
    for (int k = 0; k < K; ++k) {
        bd[k] = __int_as_float(0x7f800000);
        bi[k] = 0x7fffffff;
    }
    const float* rb = ref + (size_t)b * N * C;
    for (int n = tid; n < N; n += KNN_FEW_THREADS) {
        const float4* r4 = reinterpret_cast<const float4*>(rb + (size_t)n * C);
        float dot = 0.f, rn = 0.f;
        for (int c4 = 0; c4 < C / 4; ++c4) {
            const float4 r = __ldg(r4 + c4);
            dot = __fmaf_rn(qs[4 * c4 + 0], r.x, dot);
            dot = __fmaf_rn(qs[4 * c4 + 1], r.y, dot);
            dot = __fmaf_rn(qs[4 * c4 + 2], r.z, dot);
            dot = __fmaf_rn(qs[4 * c4 + 3], r.w, dot);
            rn = c4 == 0 ? __fmul_rn(r.x, r.x) : __fadd_rn(rn, __fmul_rn(r.x, r.x));
            rn = __fadd_rn(rn, __fmul_rn(r.y, r.y));
            rn = __fadd_rn(rn, __fmul_rn(r.z, r.z));
            rn = __fadd_rn(rn, __fmul_rn(r.w, r.w));
        }
        const float d = sqdist_from_dot(dot, qn, rn);
        if (d < bd[K - 1]) topk_insert<K>(bd, bi, d, n);  // ascending n within a thread: strict < keeps the lower index
    }
    // K rounds: block-wide arg-min over the heads of the per-thread lists; the winning thread pops its head
    int head = 0;
    for (int k = 0; k < K; ++k) {
        float hd = __int_as_float(0x7f800000);
        int hn = 0x7fffffff;
#pragma unroll
        for (int j = 0; j < K; ++j)
            if (j == head) {
                hd = bd[j];
                hn = bi[j];
            }
        unsigned long long key = head < K ? few_key(hd, hn) : ~0ull;
        unsigned long long m = key;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const unsigned long long t = __shfl_xor_sync(0xffffffffu, m, o);
            m = t < m ? t : m;
        }
        if (lane == 0) wbest[warp] = m;
        __syncthreads();
        if (tid == 0) {
            unsigned long long g = wbest[0];
            for (int w2 = 1; w2 < KNN_FEW_THREADS / 32; ++w2) g = wbest[w2] < g ? wbest[w2] : g;
            round_best = g;
        }
        __syncthreads();
        if (key == round_best && head < K) {  // keys are unique (distinct indices): exactly one thread
            const size_t o = ((size_t)b * S + s) * K + k;
            if (dist_out) dist_out[o] = hd;
            idx_out[o] = hn;
            ++head;
        }
        __syncthreads();
    }
}

// Exact search for an indirect list of queries (see knn_tiled_kernel): the fallback pass of mpc_knn_tc_f32.
int launch_knn_tiled_indirect8(const float* ref, const float* qry, float* dist_out, int64_t* idx_out, const int* qlist,
                               const int* qcount, int B, int N, int S, int C, cudaStream_t st) {
    if (C % 64 != 0 || C > 256) return MPC_ERR_UNSUPPORTED;
    const size_t smem = ((size_t)C * KT_Q + KT_CK * KT_R + KT_Q * KT_DH + KT_Q + KT_R) * sizeof(float);
    auto kern = knn_tiled_kernel<8>;
    MPC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((unsigned)ceil_div(S, KT_Q), (unsigned)B);
    kern<<<grid, KT_THREADS, smem, st>>>(ref, qry, dist_out, idx_out, N, S, C, qlist, qcount);
    MPC_LAUNCH_CHECK();
    // (the two kernels partition the clouds by their listed-query count; each leaves the other's clouds at once)
    knn_few_kernel<8><<<dim3((unsigned)(S < KNN_FEW_MAX ? S : KNN_FEW_MAX), (unsigned)B), KNN_FEW_THREADS,
                        (size_t)C * sizeof(float), st>>>(ref, qry, dist_out, idx_out, N, S, C, qlist, qcount);
    MPC_LAUNCH_CHECK();
    return MPC_OK;
}

// ---- ball query ---------------------------------------------------------------------------------------------
constexpr int BQ_THREADS = 128;
constexpr int BQ_TILE = 1024;

__global__ void __launch_bounds__(BQ_THREADS)
ball_query3_kernel(const float* __restrict__ xyz, const float* __restrict__ new_xyz,
                   int64_t* __restrict__ idx_out, float r2, int N, int S, int nsample) {
    __shared__ float4 tile[BQ_TILE];
    const int b = blockIdx.y;
    const int s = blockIdx.x * BQ_THREADS + threadIdx.x;
    const bool active = s < S;
    const float* rb = xyz + (size_t)b * N * 3;
    float qx = 0.f, qy = 0.f, qz = 0.f;
    if (active) {
        const float* q = new_xyz + ((size_t)b * S + s) * 3;
        qx = q[0];
        qy = q[1];
        qz = q[2];
    }
    const float qn = sqnorm3(qx, qy, qz);
    int64_t* o = idx_out + ((size_t)b * S + (active ? s : 0)) * nsample;
    int cnt = 0;
    int64_t first = N;
    for (int t0 = 0; t0 < N; t0 += BQ_TILE) {
        const int tn = min(BQ_TILE, N - t0);
        const bool need = active && cnt < nsample;
        if (!__syncthreads_or(need)) break;  // every query of the CTA is full
        for (int j = threadIdx.x; j < tn; j += BQ_THREADS) {
            const float* r = rb + (size_t)(t0 + j) * 3;
            float x = r[0], y = r[1], z = r[2];
            tile[j] = make_float4(x, y, z, sqnorm3(x, y, z));
        }
        __syncthreads();
        if (need) {
            for (int j = 0; j < tn && cnt < nsample; ++j) {
                const float4 r = tile[j];
                float dot = __fmul_rn(qx, r.x);
                dot = __fmaf_rn(qy, r.y, dot);
                dot = __fmaf_rn(qz, r.z, dot);
                const float d = sqdist_from_dot(dot, qn, r.w);
                if (!(d > r2)) {
                    if (cnt == 0) first = t0 + j;
                    o[cnt++] = t0 + j;
                }
            }
        }
    }
    if (active)
        for (int k = cnt; k < nsample; ++k) o[k] = first;
}

// generic C: one thread per query straight from global memory (not on any live model path)
__global__ void __launch_bounds__(BQ_THREADS)
ball_query_generic_kernel(const float* __restrict__ xyz, const float* __restrict__ new_xyz,
                          int64_t* __restrict__ idx_out, float r2, int N, int S, int C, int nsample) {
    const int b = blockIdx.y;
    const int s = blockIdx.x * BQ_THREADS + threadIdx.x;
    if (s >= S) return;
    const float* q = new_xyz + ((size_t)b * S + s) * C;
    float qn = __fmul_rn(q[0], q[0]);
    for (int c = 1; c < C; ++c) qn = __fadd_rn(qn, __fmul_rn(q[c], q[c]));
    int64_t* o = idx_out + ((size_t)b * S + s) * nsample;
    int cnt = 0;
    int64_t first = N;
    for (int n = 0; n < N && cnt < nsample; ++n) {
        const float* r = xyz + ((size_t)b * N + n) * C;
        float dot = 0.f, rn = __fmul_rn(r[0], r[0]);
        for (int c = 0; c < C; ++c) dot = __fmaf_rn(q[c], r[c], dot);
        for (int c = 1; c < C; ++c) rn = __fadd_rn(rn, __fmul_rn(r[c], r[c]));
        const float d = sqdist_from_dot(dot, qn, rn);
        if (!(d > r2)) {
            if (cnt == 0) first = n;
            o[cnt++] = n;
        }
    }
    for (int k = cnt; k < nsample; ++k) o[k] = first;
}

}  // namespace mpc

MPC_API int mpc_knn_f32(const float* ref, const float* qry, float* dist_out, int64_t* idx_out, int64_t B,
                        int64_t N, int64_t S, int64_t C, int64_t K, mpc_stream_t stream) {
    using namespace mpc;
    if (B < 0 || N <= 0 || S < 0 || C <= 0 || K <= 0 || K > N) return MPC_ERR_INVALID;
    if (B == 0 || S == 0) return MPC_OK;  // nothing to do: empty tensors carry null pointers
    if (!ref || !qry || !idx_out) return MPC_ERR_INVALID;
    if (K > 32 || C > 1024 || N > INT32_MAX || S > INT32_MAX || B > 65535) return MPC_ERR_UNSUPPORTED;
    cudaStream_t st = (cudaStream_t)stream;
    const int b = (int)B, n = (int)N, s = (int)S, c = (int)C;
    if (K <= 3) {
        if (K == 3) return launch_knn<3>(ref, qry, dist_out, idx_out, b, n, s, c, st);
        if (K == 1) return launch_knn<1>(ref, qry, dist_out, idx_out, b, n, s, c, st);
    }
    if (K == 8) return launch_knn<8>(ref, qry, dist_out, idx_out, b, n, s, c, st);
    if (K == 9) return launch_knn<9>(ref, qry, dist_out, idx_out, b, n, s, c, st);
    if (K == 16) return launch_knn<16>(ref, qry, dist_out, idx_out, b, n, s, c, st);
    if (K == 32) return launch_knn<32>(ref, qry, dist_out, idx_out, b, n, s, c, st);
    return MPC_ERR_UNSUPPORTED;  // the host wrapper rounds K up to a supported list length and slices
}

MPC_API int mpc_ball_query_f32(const float* xyz, const float* new_xyz, int64_t* idx_out, float r2, int64_t B,
                               int64_t N, int64_t S, int64_t C, int64_t nsample, mpc_stream_t stream) {
    using namespace mpc;
    if (B < 0 || N <= 0 || S < 0 || C <= 0 || nsample <= 0) return MPC_ERR_INVALID;
    if (B == 0 || S == 0) return MPC_OK;
    if (!xyz || !new_xyz || !idx_out) return MPC_ERR_INVALID;
    if (N > INT32_MAX || S > INT32_MAX || B > 65535 || nsample > INT32_MAX) return MPC_ERR_UNSUPPORTED;
    cudaStream_t st = (cudaStream_t)stream;
    dim3 grid((unsigned)ceil_div(S, BQ_THREADS), (unsigned)B);
    if (C == 3)
        ball_query3_kernel<<<grid, BQ_THREADS, 0, st>>>(xyz, new_xyz, idx_out, r2, (int)N, (int)S, (int)nsample);
    else
        ball_query_generic_kernel<<<grid, BQ_THREADS, 0, st>>>(xyz, new_xyz, idx_out, r2, (int)N, (int)S, (int)C,
                                                               (int)nsample);
    MPC_LAUNCH_CHECK();
    return MPC_OK;
}
