// Difference-wise attention core of LocalTrans for sm_100a.  See include/mpc_b200.h (mpc_attn_*) for the
// contract and the reference lines replaced (pointnet2_utils.py:518-569).
//
// The reference materialises ~9 tensors of shape [B,S,K,C] per call (gathered k, gathered v, energy, scaled
// energy, softmax, row-sum, shifted attention, product, max).  Here one thread owns 4 channels (feature
// branch) or 1 channel (coordinate branch) of one centre point, keeps the K neighbour values in registers and
// writes only the [B,S,C] context; the backward recomputes the softmax instead of saving it.  Both are
// gather-bound: per centre point they read K rows of k and v (L2-resident at the live sizes) with 128-bit
// loads, so the roofline that applies is HBM/L2 bandwidth, not the tensor pipe.
#include <cuda_bf16.h>

#include "common.cuh"

namespace mpc {

constexpr int AT = 256;
constexpr int KMAX = 32;

static inline unsigned attn_grid(int64_t work_items, int threads) {
    int64_t g = ceil_div(work_items, threads);
    const int64_t cap = (int64_t)kNumSMs * 16;
    return (unsigned)(g < 1 ? 1 : (g > cap ? cap : g));
}

// One channel of the core.  e[j] holds q - k_j on entry.  Returns ctx; fills a[j] (softmax) and O = sum a.
// rs = 1/sqrt(C).  Scaling and normalisation multiply by reciprocals (one IEEE division per channel instead of
// 2K of them: the division sequence was ~half of the kernel's instructions); the result differs from the
// reference's divide-by-sqrt(C) by at most 1 ulp per term, far inside the 1e-5 feature tolerance.
template <int K>
__device__ __forceinline__ float attn_channel(const float (&e)[K], const float (&v)[K], float rs,
                                              float (&a)[K], float& O, int& jstar) {
    float s[K];
    float m = -__int_as_float(0x7f800000);
#pragma unroll
    for (int j = 0; j < K; ++j) {
        s[j] = e[j] * rs;
        m = fmaxf(m, s[j]);
    }
    float Z = 0.f;
#pragma unroll
    for (int j = 0; j < K; ++j) {
        a[j] = __expf(s[j] - m);
        Z += a[j];
    }
    const float rZ = __fdiv_rn(1.0f, Z);
    O = 0.f;
#pragma unroll
    for (int j = 0; j < K; ++j) {
        a[j] *= rZ;
        O += a[j];
    }
    float best = -__int_as_float(0x7f800000);
    jstar = 0;
#pragma unroll
    for (int j = 0; j < K; ++j) {
        const float c = (a[j] - O) * v[j];
        if (c > best) {  // first maximum, like torch.max(dim)
            best = c;
            jstar = j;
        }
    }
    return best;
}

// Backward of one channel: given g = dL/dctx, the recomputed a, O, jstar: returns de[j] (w.r.t. energy q-k_j)
// and dv (for v_{jstar}).  ctx = (a_{j*} - sum_i a_i) * v_{j*}.
template <int K>
__device__ __forceinline__ void attn_channel_bwd(float g, const float (&a)[K], float O, int jstar,
                                                 const float (&v)[K], float rs, float (&de)[K], float& dv) {
    float vstar = 0.f, astar = 0.f;
#pragma unroll
    for (int j = 0; j < K; ++j)
        if (j == jstar) {
            vstar = v[j];
            astar = a[j];
        }
    dv = g * (astar - O);
    const float gv = g * vstar;
    // dL/da_i = gv * (delta(i,j*) - 1);  softmax backward: ds_i = a_i * (da_i - sum_m a_m da_m)
    const float dot = gv * astar - gv * O;  // sum_m a_m da_m
#pragma unroll
    for (int j = 0; j < K; ++j) {
        const float da = (j == jstar ? gv : 0.f) - gv;
        de[j] = a[j] * (da - dot) * rs;
    }
}

// ---- feature branch ---------------------------------------------------------------------------------------------
template <int K>
__global__ void __launch_bounds__(AT)
attn_feat_fwd_kernel(const float* __restrict__ q, int64_t ldq, const float* __restrict__ kf,
                     const float* __restrict__ vf, int64_t ldkv, const int64_t* __restrict__ idx,
                     float* __restrict__ ctx, int S, int N, int CV, float sqrtc, int64_t total) {
    pdl_prologue();
    for (int64_t t = (int64_t)blockIdx.x * AT + threadIdx.x; t < total; t += (int64_t)gridDim.x * AT) {
        const int64_t row = t / CV;
        const int c4 = (int)(t - row * CV) * 4;
        // 64-bit integer division is a ~100-instruction software routine; rows < 2^32 whenever B < 2^32 / S
        const int64_t b = row < 0xffffffffll ? (int64_t)((unsigned)row / (unsigned)S) : row / S;
        const int64_t* irow = idx + row * K;
        float4 kk[K], vv[K];
#pragma unroll
        for (int j = 0; j < K; ++j) {
            const size_t src = ((size_t)b * N + clamp_index(__ldg(irow + j), N)) * ldkv + c4;
            kk[j] = __ldg(reinterpret_cast<const float4*>(kf + src));
            vv[j] = __ldg(reinterpret_cast<const float4*>(vf + src));
        }
        const float4 qq = __ldg(reinterpret_cast<const float4*>(q + row * ldq + c4));
        float4 out;
        float e[K], v[K], a[K], O;
        int js;
#define MPC_CH(comp)                                              \
    _Pragma("unroll") for (int j = 0; j < K; ++j) {               \
        e[j] = qq.comp - kk[j].comp;                              \
        v[j] = vv[j].comp;                                        \
    }                                                             \
    out.comp = attn_channel<K>(e, v, sqrtc, a, O, js);
        MPC_CH(x) MPC_CH(y) MPC_CH(z) MPC_CH(w)
#undef MPC_CH
        *reinterpret_cast<float4*>(ctx + row * (int64_t)CV * 4 + c4) = out;
    }
}

// bf16-I/O forward (inference path): rows travel as bf16 (half the gathered bytes), the arithmetic is the fp32
// expression of the kernel above, one rounding at the store.
__device__ __forceinline__ float4 ld_bf16x4(const __nv_bfloat16* p) {
    const uint2 w = __ldg(reinterpret_cast<const uint2*>(p));
    return make_float4(__uint_as_float(w.x << 16), __uint_as_float(w.x & 0xffff0000u), __uint_as_float(w.y << 16),
                       __uint_as_float(w.y & 0xffff0000u));
}

template <int K>
__global__ void __launch_bounds__(AT)
attn_feat_fwd_bf16_kernel(const __nv_bfloat16* __restrict__ q, int64_t ldq, const __nv_bfloat16* __restrict__ kf,
                          const __nv_bfloat16* __restrict__ vf, int64_t ldkv, const int64_t* __restrict__ idx,
                          __nv_bfloat16* __restrict__ ctx, int S, int N, int CV, float sqrtc, int64_t total) {
    pdl_prologue();
    for (int64_t t = (int64_t)blockIdx.x * AT + threadIdx.x; t < total; t += (int64_t)gridDim.x * AT) {
        const int64_t row = t / CV;
        const int c4 = (int)(t - row * CV) * 4;
        const int64_t b = row < 0xffffffffll ? (int64_t)((unsigned)row / (unsigned)S) : row / S;
        const int64_t* irow = idx + row * K;
        float4 kk[K], vv[K];
#pragma unroll
        for (int j = 0; j < K; ++j) {
            const size_t src = ((size_t)b * N + clamp_index(__ldg(irow + j), N)) * ldkv + c4;
            kk[j] = ld_bf16x4(kf + src);
            vv[j] = ld_bf16x4(vf + src);
        }
        const float4 qq = ld_bf16x4(q + row * ldq + c4);
        float4 out;
        float e[K], v[K], a[K], O;
        int js;
#define MPC_CH(comp)                                              \
    _Pragma("unroll") for (int j = 0; j < K; ++j) {               \
        e[j] = qq.comp - kk[j].comp;                              \
        v[j] = vv[j].comp;                                        \
    }                                                             \
    out.comp = attn_channel<K>(e, v, sqrtc, a, O, js);
        MPC_CH(x) MPC_CH(y) MPC_CH(z) MPC_CH(w)
#undef MPC_CH
        const __nv_bfloat162 h0 = __floats2bfloat162_rn(out.x, out.y), h1 = __floats2bfloat162_rn(out.z, out.w);
        uint2 o;
        o.x = *reinterpret_cast<const uint32_t*>(&h0);
        o.y = *reinterpret_cast<const uint32_t*>(&h1);
        *reinterpret_cast<uint2*>(ctx + row * (int64_t)CV * 4 + c4) = o;
    }
}

template <int K>
__global__ void __launch_bounds__(AT, 2)
attn_feat_bwd_kernel(const float* __restrict__ gctx, const float* __restrict__ q, int64_t ldq,
                     const float* __restrict__ kf, const float* __restrict__ vf, int64_t ldkv,
                     const int64_t* __restrict__ idx, float* __restrict__ gq, int64_t ldgq,
                     float* __restrict__ gkf, float* __restrict__ gvf, int64_t ldgkv, float* __restrict__ gbias,
                     int S, int N, int CV, float sqrtc, int64_t total) {
    pdl_prologue();
    // column sums of grad_q / grad_k / grad_v (= the bias gradients of the three projections): AT % CV == 0, so a
    // thread's channel group never changes across its grid-stride rows and the sums live in registers
    float4 sq = make_float4(0.f, 0.f, 0.f, 0.f), sk = sq, sv = sq;
    for (int64_t t = (int64_t)blockIdx.x * AT + threadIdx.x; t < total; t += (int64_t)gridDim.x * AT) {
        const int64_t row = t / CV;
        const int c4 = (int)(t - row * CV) * 4;
        // 64-bit integer division is a ~100-instruction software routine; rows < 2^32 whenever B < 2^32 / S
        const int64_t b = row < 0xffffffffll ? (int64_t)((unsigned)row / (unsigned)S) : row / S;
        const int64_t* irow = idx + row * K;
        int nb[K];
        float4 kk[K], vv[K];
#pragma unroll
        for (int j = 0; j < K; ++j) {
            nb[j] = clamp_index(__ldg(irow + j), N);
            const size_t src = ((size_t)b * N + nb[j]) * ldkv + c4;
            kk[j] = __ldg(reinterpret_cast<const float4*>(kf + src));
            vv[j] = __ldg(reinterpret_cast<const float4*>(vf + src));
        }
        const float4 qq = __ldg(reinterpret_cast<const float4*>(q + row * ldq + c4));
        const float4 gg = __ldg(reinterpret_cast<const float4*>(gctx + row * (int64_t)CV * 4 + c4));
        float4 dk[K];
        float4 dq = make_float4(0.f, 0.f, 0.f, 0.f);
        float e[K], v[K], a[K], de[K], O, dv;
        int js;
#define MPC_CH(comp, ci)                                                                   \
    _Pragma("unroll") for (int j = 0; j < K; ++j) {                                        \
        e[j] = qq.comp - kk[j].comp;                                                       \
        v[j] = vv[j].comp;                                                                 \
    }                                                                                      \
    attn_channel<K>(e, v, sqrtc, a, O, js);                                                \
    attn_channel_bwd<K>(gg.comp, a, O, js, v, sqrtc, de, dv);                              \
    _Pragma("unroll") for (int j = 0; j < K; ++j) {                                        \
        dq.comp += de[j];                                                                  \
        dk[j].comp = -de[j];                                                               \
    }                                                                                      \
    sq.comp += dq.comp;                                                                    \
    sk.comp -= dq.comp;                                                                    \
    sv.comp += dv;                                                                         \
    {                                                                                      \
        int n = nb[0];                                                                     \
        _Pragma("unroll") for (int j = 1; j < K; ++j) n = (j == js) ? nb[j] : n;           \
        red_add_f32(gvf + ((size_t)b * N + n) * ldgkv + c4 + ci, dv);      \
    }
        MPC_CH(x, 0) MPC_CH(y, 1) MPC_CH(z, 2) MPC_CH(w, 3)
#undef MPC_CH
        *reinterpret_cast<float4*>(gq + row * ldgq + c4) = dq;
#pragma unroll
        for (int j = 0; j < K; ++j) red_add_f32x4(gkf + ((size_t)b * N + nb[j]) * ldgkv + c4, dk[j]);
    }
    if (gbias) {  // gbias [3][C]: q, k, v
        __shared__ float4 red[3][AT];
        red[0][threadIdx.x] = sq;
        red[1][threadIdx.x] = sk;
        red[2][threadIdx.x] = sv;
        __syncthreads();
        if (threadIdx.x < CV) {
            const int C = CV * 4;
            for (int w = 0; w < 3; ++w) {
                float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
                for (int i = threadIdx.x; i < AT; i += CV) {
                    const float4 u = red[w][i];
                    acc.x += u.x; acc.y += u.y; acc.z += u.z; acc.w += u.w;
                }
                red_add_f32x4(gbias + w * C + threadIdx.x * 4, acc);
            }
        }
    }
}

// ---- coordinate branch --------------------------------------------------------------------------------------------
// thread = (centre row, channel c); blockDim.x is a multiple of C so a thread's channel (and its 3*CIN weights)
// never changes across its grid-stride rows.
template <int K, int CIN>
__global__ void __launch_bounds__(256)
attn_xyz_fwd_kernel(const float* __restrict__ feat, const int64_t* __restrict__ cidx,
                    const int64_t* __restrict__ idx, const float* __restrict__ wq, const float* __restrict__ bq,
                    const float* __restrict__ wk, const float* __restrict__ bk, const float* __restrict__ wv,
                    const float* __restrict__ bv, const float* __restrict__ wr, const float* __restrict__ br,
                    float* __restrict__ ctx, float* __restrict__ res_out, int64_t rows, int S, int N, int C,
                    int cin_rt, float sqrtc) {
    pdl_prologue();
    const int cin = CIN > 0 ? CIN : cin_rt;
    constexpr int CM = CIN > 0 ? CIN : 16;
    const int Cb = C < 256 ? C : 256;  // channels per CTA; C > 256 is split along gridDim.y (C % 256 == 0)
    const int c = blockIdx.y * Cb + threadIdx.x % Cb;
    const int rpb = blockDim.x / Cb;
    float Wq[CM], Wk[CM], Wv[CM], Wr[CM];
#pragma unroll
    for (int i = 0; i < CM; ++i)
        if (i < cin) {
            Wq[i] = wq[c * cin + i];
            Wk[i] = wk[c * cin + i];
            Wv[i] = wv[c * cin + i];
            Wr[i] = wr ? wr[c * cin + i] : 0.f;
        }
    const float Bq = bq[c], Bk = bk[c], Bv = bv[c], Br = wr ? br[c] : 0.f;
    for (int64_t row = (int64_t)blockIdx.x * rpb + threadIdx.x / Cb; row < rows; row += (int64_t)gridDim.x * rpb) {
        // 64-bit integer division is a ~100-instruction software routine; rows < 2^32 whenever B < 2^32 / S
        const int64_t b = row < 0xffffffffll ? (int64_t)((unsigned)row / (unsigned)S) : row / S;
        const int cn = cidx ? clamp_index(__ldg(cidx + row), N) : (int)(row - b * S);
        const float* cp = feat + ((size_t)b * N + cn) * cin;
        float cen[CM];
#pragma unroll
        for (int i = 0; i < CM; ++i)
            if (i < cin) cen[i] = __ldg(cp + i);
        float qv = 0.f, rv = 0.f;
#pragma unroll
        for (int i = 0; i < CM; ++i)
            if (i < cin) {
                qv = fmaf(cen[i], Wq[i], qv);
                rv = fmaf(cen[i], Wr[i], rv);
            }
        qv += Bq;
        if (res_out) res_out[row * C + c] = rv + Br;  // conv_res.linear(centre), pre-BatchNorm
        float e[K], v[K], a[K], O;
        int js;
#pragma unroll
        for (int j = 0; j < K; ++j) {
            const float* np = feat + ((size_t)b * N + clamp_index(__ldg(idx + row * K + j), N)) * cin;
            float kv = 0.f, vv = 0.f;
#pragma unroll
            for (int i = 0; i < CM; ++i)
                if (i < cin) {
                    const float d = __ldg(np + i) - cen[i];
                    kv = fmaf(d, Wk[i], kv);
                    vv = fmaf(d, Wv[i], vv);
                }
            e[j] = qv - (kv + Bk);
            v[j] = vv + Bv;
        }
        ctx[row * C + c] = attn_channel<K>(e, v, sqrtc, a, O, js);
    }
}

template <int K, int CIN>
__global__ void __launch_bounds__(256, 3)
attn_xyz_bwd_kernel(const float* __restrict__ gctx, const float* __restrict__ feat,
                    const int64_t* __restrict__ cidx, const int64_t* __restrict__ idx,
                    const float* __restrict__ wq, const float* __restrict__ bq, const float* __restrict__ wk,
                    const float* __restrict__ bk, const float* __restrict__ wv, const float* __restrict__ bv,
                    const float* __restrict__ wr, const float* __restrict__ gres,
                    float* __restrict__ gwq, float* __restrict__ gbq, float* __restrict__ gwk,
                    float* __restrict__ gbk, float* __restrict__ gwv, float* __restrict__ gbv,
                    float* __restrict__ gwr, float* __restrict__ gbr,
                    float* __restrict__ gfeat, int64_t rows, int S, int N, int C, int cin_rt, float sqrtc) {
    pdl_prologue();
    const int cin = CIN > 0 ? CIN : cin_rt;
    constexpr int CM = CIN > 0 ? CIN : 16;
    const int Cb = C < 256 ? C : 256;
    const int c = blockIdx.y * Cb + threadIdx.x % Cb;
    const int rpb = blockDim.x / Cb;
    const bool warp_uniform_row = (C % 32) == 0;  // all 32 lanes of a warp work on the same centre row
    float Wq[CM], Wk[CM], Wv[CM], Wr[CM], GWq[CM], GWk[CM], GWv[CM], GWr[CM];
#pragma unroll
    for (int i = 0; i < CM; ++i) {
        GWq[i] = GWk[i] = GWv[i] = GWr[i] = 0.f;
        if (i < cin) {
            Wq[i] = wq[c * cin + i];
            Wk[i] = wk[c * cin + i];
            Wv[i] = wv[c * cin + i];
            Wr[i] = gres ? wr[c * cin + i] : 0.f;
        }
    }
    const float Bq = bq[c], Bk = bk[c], Bv = bv[c];
    float GBq = 0.f, GBk = 0.f, GBv = 0.f, GBr = 0.f;
    for (int64_t row = (int64_t)blockIdx.x * rpb + threadIdx.x / Cb; row < rows; row += (int64_t)gridDim.x * rpb) {
        // 64-bit integer division is a ~100-instruction software routine; rows < 2^32 whenever B < 2^32 / S
        const int64_t b = row < 0xffffffffll ? (int64_t)((unsigned)row / (unsigned)S) : row / S;
        const int cn = cidx ? clamp_index(__ldg(cidx + row), N) : (int)(row - b * S);
        const float* cp = feat + ((size_t)b * N + cn) * cin;
        float cen[CM];
#pragma unroll
        for (int i = 0; i < CM; ++i)
            if (i < cin) cen[i] = __ldg(cp + i);
        float qv = 0.f;
#pragma unroll
        for (int i = 0; i < CM; ++i)
            if (i < cin) qv = fmaf(cen[i], Wq[i], qv);
        qv += Bq;
        float e[K], v[K], a[K], de[K], O, dv;
        int js, nb[K];
#pragma unroll
        for (int j = 0; j < K; ++j) {
            nb[j] = clamp_index(__ldg(idx + row * K + j), N);
            const float* np = feat + ((size_t)b * N + nb[j]) * cin;
            float kv = 0.f, vv = 0.f;
#pragma unroll
            for (int i = 0; i < CM; ++i)
                if (i < cin) {
                    const float d = __ldg(np + i) - cen[i];
                    kv = fmaf(d, Wk[i], kv);
                    vv = fmaf(d, Wv[i], vv);
                }
            e[j] = qv - (kv + Bk);
            v[j] = vv + Bv;
        }
        attn_channel<K>(e, v, sqrtc, a, O, js);
        attn_channel_bwd<K>(__ldg(gctx + row * C + c), a, O, js, v, sqrtc, de, dv);
        // q = Wq cen + bq ; k_j = Wk (x_j - cen) + bk ; v_j = Wv (x_j - cen) + bv ; e_j = q - k_j
        float dq = 0.f;
        float gcen[CM];
#pragma unroll
        for (int i = 0; i < CM; ++i) gcen[i] = 0.f;
#pragma unroll
        for (int j = 0; j < K; ++j) {
            dq += de[j];
            const float dkj = -de[j];
            const float dvj = (j == js) ? dv : 0.f;
            GBk += dkj;
            GBv += dvj;
            const float* np = feat + ((size_t)b * N + nb[j]) * cin;
#pragma unroll
            for (int i = 0; i < CM; ++i)
                if (i < cin) {
                    const float d = __ldg(np + i) - cen[i];
                    GWk[i] = fmaf(dkj, d, GWk[i]);
                    GWv[i] = fmaf(dvj, d, GWv[i]);
                    if (gfeat) {
                        float gd = dkj * Wk[i] + dvj * Wv[i];  // dL/d(x_j - cen)_i from this channel
                        gcen[i] -= gd;
                        if (warp_uniform_row) {
#pragma unroll
                            for (int o = 16; o > 0; o >>= 1) gd += __shfl_xor_sync(0xffffffffu, gd, o);
                            if ((threadIdx.x & 31) == 0) red_add_f32(gfeat + ((size_t)b * N + nb[j]) * cin + i, gd);
                        } else {
                            red_add_f32(gfeat + ((size_t)b * N + nb[j]) * cin + i, gd);
                        }
                    }
                }
        }
        GBq += dq;
        const float gr = gres ? __ldg(gres + row * C + c) : 0.f;  // grad of conv_res.linear(centre)
        GBr += gr;
#pragma unroll
        for (int i = 0; i < CM; ++i)
            if (i < cin) {
                GWq[i] = fmaf(dq, cen[i], GWq[i]);
                GWr[i] = fmaf(gr, cen[i], GWr[i]);
                if (gfeat) {
                    float gd = gcen[i] + dq * Wq[i] + gr * Wr[i];
                    if (warp_uniform_row) {
#pragma unroll
                        for (int o = 16; o > 0; o >>= 1) gd += __shfl_xor_sync(0xffffffffu, gd, o);
                        if ((threadIdx.x & 31) == 0) red_add_f32(gfeat + ((size_t)b * N + cn) * cin + i, gd);
                    } else {
                        red_add_f32(gfeat + ((size_t)b * N + cn) * cin + i, gd);
                    }
                }
            }
    }
    // reduce the per-thread partials over the CTA's row slots in shared memory first (threads tid, tid + Cb, ...
    // own the same channel), then one red.global per (CTA, channel, quantity): 4x-8x fewer same-address atomics
    if constexpr (CIN == 0) {  // generic input width (not on a live model path): direct reductions
#pragma unroll
        for (int i = 0; i < CM; ++i)
            if (i < cin) {
                red_add_f32(gwq + c * cin + i, GWq[i]);
                red_add_f32(gwk + c * cin + i, GWk[i]);
                red_add_f32(gwv + c * cin + i, GWv[i]);
                if (gres) red_add_f32(gwr + c * cin + i, GWr[i]);
            }
        red_add_f32(gbq + c, GBq);
        red_add_f32(gbk + c, GBk);
        red_add_f32(gbv + c, GBv);
        if (gres) red_add_f32(gbr + c, GBr);
        return;
    }
    constexpr int RQ = CIN > 0 ? 4 * (CIN + 1) : 1;
    __shared__ float red[RQ][256];
    {
        int q = 0;
#pragma unroll
        for (int i = 0; i < CM; ++i) {
            red[q++][threadIdx.x] = GWq[i];
            red[q++][threadIdx.x] = GWk[i];
            red[q++][threadIdx.x] = GWv[i];
            red[q++][threadIdx.x] = GWr[i];
        }
        red[q++][threadIdx.x] = GBq;
        red[q++][threadIdx.x] = GBk;
        red[q++][threadIdx.x] = GBv;
        red[q++][threadIdx.x] = GBr;
    }
    __syncthreads();
    if ((int)threadIdx.x < Cb) {
        auto total = [&](int q) {
            float a = 0.f;
            for (int t = threadIdx.x; t < (int)blockDim.x; t += Cb) a += red[q][t];
            return a;
        };
#pragma unroll
        for (int i = 0; i < CM; ++i)
            if (i < cin) {
                red_add_f32(gwq + c * cin + i, total(4 * i + 0));
                red_add_f32(gwk + c * cin + i, total(4 * i + 1));
                red_add_f32(gwv + c * cin + i, total(4 * i + 2));
                if (gres) red_add_f32(gwr + c * cin + i, total(4 * i + 3));
            }
        red_add_f32(gbq + c, total(4 * CM + 0));
        red_add_f32(gbk + c, total(4 * CM + 1));
        red_add_f32(gbv + c, total(4 * CM + 2));
        if (gres) red_add_f32(gbr + c, total(4 * CM + 3));
    }
}

static inline int xyz_threads(int C) {  // C <= 256: several rows per CTA; else 256 channels per CTA
    if (C >= 256) return 256;
    return C * (256 / C);
}

}  // namespace mpc

using namespace mpc;

#define MPC_DISPATCH_K(K, CALL)            \
    switch (K) {                           \
        case 3: { constexpr int KK = 3; CALL; } break;   \
        case 4: { constexpr int KK = 4; CALL; } break;   \
        case 8: { constexpr int KK = 8; CALL; } break;   \
        case 9: { constexpr int KK = 9; CALL; } break;   \
        case 16: { constexpr int KK = 16; CALL; } break; \
        case 32: { constexpr int KK = 32; CALL; } break; \
        default: return MPC_ERR_UNSUPPORTED;             \
    }

MPC_API int mpc_attn_feat_fwd_f32(const float* q, int64_t ldq, const float* kf, const float* vf, int64_t ldkv,
                                  const int64_t* idx, float* ctx_out, int64_t B, int64_t S, int64_t N, int64_t K,
                                  int64_t C, mpc_stream_t stream) {
    if (!q || !kf || !vf || !idx || !ctx_out || B < 0 || S < 0 || N <= 0 || K <= 0 || C <= 0) return MPC_ERR_INVALID;
    if (C % 4 || ldq % 4 || ldkv % 4 || ldq < C || ldkv < C || N > INT32_MAX) return MPC_ERR_INVALID;
    if (((uintptr_t)q | (uintptr_t)kf | (uintptr_t)vf | (uintptr_t)ctx_out) & 15u) return MPC_ERR_INVALID;
    if (K > KMAX || C > 1024) return MPC_ERR_UNSUPPORTED;
    if (B == 0 || S == 0) return MPC_OK;
    const int CV = (int)(C / 4);
    const int64_t total = B * S * CV;
    const float sqrtc = 1.0f / sqrtf((float)C);  // passed to the kernels as the reciprocal scale
    cudaStream_t st = (cudaStream_t)stream;
    MPC_DISPATCH_K(K, (pdl_launch(attn_feat_fwd_kernel<KK>, dim3(attn_grid(total, AT)), dim3(AT), 0, st, 
                          q, ldq, kf, vf, ldkv, idx, ctx_out, (int)S, (int)N, CV, sqrtc, total)));
    MPC_LAUNCH_CHECK();
    return MPC_OK;
}

MPC_API int mpc_attn_feat_fwd_bf16(const void* q, int64_t ldq, const void* kf, const void* vf, int64_t ldkv,
                                   const int64_t* idx, void* ctx_out, int64_t B, int64_t S, int64_t N, int64_t K,
                                   int64_t C, mpc_stream_t stream) {
    if (!q || !kf || !vf || !idx || !ctx_out || B < 0 || S < 0 || N <= 0 || K <= 0 || C <= 0) return MPC_ERR_INVALID;
    if (C % 4 || ldq % 4 || ldkv % 4 || ldq < C || ldkv < C || N > INT32_MAX) return MPC_ERR_INVALID;
    if (((uintptr_t)q | (uintptr_t)kf | (uintptr_t)vf | (uintptr_t)ctx_out) & 7u) return MPC_ERR_INVALID;
    if (K > KMAX || C > 1024) return MPC_ERR_UNSUPPORTED;
    if (B == 0 || S == 0) return MPC_OK;
    const int CV = (int)(C / 4);
    const int64_t total = B * S * CV;
    const float sqrtc = 1.0f / sqrtf((float)C);
    cudaStream_t st = (cudaStream_t)stream;
    MPC_DISPATCH_K(K, (pdl_launch(attn_feat_fwd_bf16_kernel<KK>, dim3(attn_grid(total, AT)), dim3(AT), 0, st,
                          static_cast<const __nv_bfloat16*>(q), ldq, static_cast<const __nv_bfloat16*>(kf),
                          static_cast<const __nv_bfloat16*>(vf), ldkv, idx, static_cast<__nv_bfloat16*>(ctx_out), (int)S,
                          (int)N, CV, sqrtc, total)));
    MPC_LAUNCH_CHECK();
    return MPC_OK;
}

MPC_API int mpc_attn_feat_bwd_f32(const float* grad_ctx, const float* q, int64_t ldq, const float* kf,
                                  const float* vf, int64_t ldkv, const int64_t* idx, float* grad_q, int64_t ldgq,
                                  float* grad_kf, float* grad_vf, int64_t ldgkv, float* grad_bias, int64_t B, int64_t S,
                                  int64_t N, int64_t K, int64_t C, mpc_stream_t stream) {
    if (!grad_ctx || !q || !kf || !vf || !idx || !grad_q || !grad_kf || !grad_vf) return MPC_ERR_INVALID;
    if (B < 0 || S < 0 || N <= 0 || K <= 0 || C <= 0 || C % 4 || ldq % 4 || ldkv % 4 || ldgq % 4 || ldgkv % 4)
        return MPC_ERR_INVALID;
    if (ldq < C || ldkv < C || ldgq < C || ldgkv < C || N > INT32_MAX) return MPC_ERR_INVALID;
    if (((uintptr_t)grad_ctx | (uintptr_t)q | (uintptr_t)kf | (uintptr_t)vf | (uintptr_t)grad_q |
         (uintptr_t)grad_kf | (uintptr_t)grad_vf) & 15u)
        return MPC_ERR_INVALID;
    if (K > KMAX || C > 1024) return MPC_ERR_UNSUPPORTED;
    if (grad_bias && (AT % (C / 4) != 0 || ((uintptr_t)grad_bias & 15u))) return MPC_ERR_UNSUPPORTED;
    if (B == 0 || S == 0) return MPC_OK;
    const int CV = (int)(C / 4);
    const int64_t total = B * S * CV;
    const float sqrtc = 1.0f / sqrtf((float)C);  // passed to the kernels as the reciprocal scale
    cudaStream_t st = (cudaStream_t)stream;
    MPC_DISPATCH_K(K, (pdl_launch(attn_feat_bwd_kernel<KK>, dim3(attn_grid(total, AT)), dim3(AT), 0, st, 
                          grad_ctx, q, ldq, kf, vf, ldkv, idx, grad_q, ldgq, grad_kf, grad_vf, ldgkv, grad_bias, (int)S,
                          (int)N, CV, sqrtc, total)));
    MPC_LAUNCH_CHECK();
    return MPC_OK;
}

MPC_API int mpc_attn_xyz_fwd_f32(const float* feat, const int64_t* center_idx, const int64_t* idx, const float* wq,
                                 const float* bq, const float* wk, const float* bk, const float* wv, const float* bv,
                                 const float* wr, const float* br, float* ctx_out, float* res_out, int64_t B,
                                 int64_t S, int64_t N, int64_t K, int64_t Cin, int64_t C, mpc_stream_t stream) {
    if (!feat || !idx || !wq || !bq || !wk || !bk || !wv || !bv || !ctx_out) return MPC_ERR_INVALID;
    if ((wr != nullptr) != (res_out != nullptr) || (wr != nullptr) != (br != nullptr)) return MPC_ERR_INVALID;
    if (B < 0 || S < 0 || N <= 0 || K <= 0 || Cin <= 0 || C <= 0 || N > INT32_MAX) return MPC_ERR_INVALID;
    if (!center_idx && S != N) return MPC_ERR_INVALID;
    if (K > KMAX || C > 1024 || Cin > 16) return MPC_ERR_UNSUPPORTED;
    if (B == 0 || S == 0) return MPC_OK;
    if (C > 256 && C % 256) return MPC_ERR_UNSUPPORTED;
    const int threads = xyz_threads((int)C);
    const int rpb = C >= 256 ? 1 : threads / (int)C;
    const int64_t rows = B * S;
    const dim3 grid(attn_grid(ceil_div(rows, rpb) * threads, threads), C > 256 ? (unsigned)(C / 256) : 1u);
    const float sqrtc = 1.0f / sqrtf((float)C);  // passed to the kernels as the reciprocal scale
    cudaStream_t st = (cudaStream_t)stream;
    if (Cin == 3) {
        MPC_DISPATCH_K(K, (pdl_launch(attn_xyz_fwd_kernel<KK, 3>, dim3(grid), dim3(threads), 0, st, 
                              feat, center_idx, idx, wq, bq, wk, bk, wv, bv, wr, br, ctx_out, res_out, rows, (int)S,
                              (int)N, (int)C, 3, sqrtc)));
    } else {
        MPC_DISPATCH_K(K, (pdl_launch(attn_xyz_fwd_kernel<KK, 0>, dim3(grid), dim3(threads), 0, st, 
                              feat, center_idx, idx, wq, bq, wk, bk, wv, bv, wr, br, ctx_out, res_out, rows, (int)S,
                              (int)N, (int)C, (int)Cin, sqrtc)));
    }
    MPC_LAUNCH_CHECK();
    return MPC_OK;
}

MPC_API int mpc_attn_xyz_bwd_f32(const float* grad_ctx, const float* feat, const int64_t* center_idx,
                                 const int64_t* idx, const float* wq, const float* bq, const float* wk,
                                 const float* bk, const float* wv, const float* bv, const float* wr,
                                 const float* grad_res, float* grad_wq, float* grad_bq, float* grad_wk, float* grad_bk,
                                 float* grad_wv, float* grad_bv, float* grad_wr, float* grad_br, float* grad_feat,
                                 int64_t B, int64_t S, int64_t N, int64_t K, int64_t Cin, int64_t C,
                                 mpc_stream_t stream) {
    if (!grad_ctx || !feat || !idx || !wq || !bq || !wk || !bk || !wv || !bv) return MPC_ERR_INVALID;
    if (grad_res && (!wr || !grad_wr || !grad_br)) return MPC_ERR_INVALID;
    if (!grad_wq || !grad_bq || !grad_wk || !grad_bk || !grad_wv || !grad_bv) return MPC_ERR_INVALID;
    if (B < 0 || S < 0 || N <= 0 || K <= 0 || Cin <= 0 || C <= 0 || N > INT32_MAX) return MPC_ERR_INVALID;
    if (!center_idx && S != N) return MPC_ERR_INVALID;
    if (K > KMAX || C > 1024 || Cin > 16) return MPC_ERR_UNSUPPORTED;
    if (B == 0 || S == 0) return MPC_OK;
    if (C > 256 && C % 256) return MPC_ERR_UNSUPPORTED;
    const int threads = xyz_threads((int)C);
    const int rpb = C >= 256 ? 1 : threads / (int)C;
    const int64_t rows = B * S;
    // fewer, longer-lived CTAs: every thread ends with 3*Cin+3 reductions into shared accumulators
    int64_t g = ceil_div(rows, rpb);
    const int64_t cap = (int64_t)kNumSMs * 4;
    const dim3 grid((unsigned)(g > cap ? cap : g), C > 256 ? (unsigned)(C / 256) : 1u);
    const float sqrtc = 1.0f / sqrtf((float)C);  // passed to the kernels as the reciprocal scale
    cudaStream_t st = (cudaStream_t)stream;
    if (Cin == 3) {
        MPC_DISPATCH_K(K, (pdl_launch(attn_xyz_bwd_kernel<KK, 3>, dim3(grid), dim3(threads), 0, st, 
                              grad_ctx, feat, center_idx, idx, wq, bq, wk, bk, wv, bv, wr, grad_res, grad_wq, grad_bq,
                              grad_wk, grad_bk, grad_wv, grad_bv, grad_wr, grad_br, grad_feat, rows, (int)S, (int)N,
                              (int)C, 3, sqrtc)));
    } else {
        MPC_DISPATCH_K(K, (pdl_launch(attn_xyz_bwd_kernel<KK, 0>, dim3(grid), dim3(threads), 0, st, 
                              grad_ctx, feat, center_idx, idx, wq, bq, wk, bk, wv, bv, wr, grad_res, grad_wq, grad_bq,
                              grad_wk, grad_bk, grad_wv, grad_bv, grad_wr, grad_br, grad_feat, rows, (int)S, (int)N,
                              (int)C, (int)Cin, sqrtc)));
    }
    MPC_LAUNCH_CHECK();
    return MPC_OK;
}
