// Tail of the shared-MLP block (`Linear`: nn.Linear -> BatchNorm1d over channels -> LeakyReLU(0.2)) for
// sm_100a.  See include/mpc_b200.h (mpc_bn_*) and pointnet2_utils.py:413-425.
//
// Works on the [M,C] view (M = every leading axis), so the two permute().contiguous() copies the reference
// makes around BatchNorm1d (:420) disappear.  All three kernels are pure HBM streams: 128-bit accesses, one
// warp per 128 consecutive channels of a row, per-channel partial sums in registers -> shared -> one fp64
// atomic per (CTA, channel).
#include "common.cuh"

namespace mpc {

// tuning knobs (mpc_debug_set_knob; 0 = built-in default): 0 = elementwise CTAs per SM, 1 = column-reduction CTAs
// per SM, 2 = elementwise float4 per thread the grid is sized for
int64_t g_knob[8] = {0, 0, 0, 0, 0, 0, 0, 0};  // 3 = force the grid-wide FPS variant for N > 8192 (tests); 6 = smallest N for the pruned FPS variant (< 0: never)
static inline int64_t knob(int i, int64_t dflt) { return g_knob[i] > 0 ? g_knob[i] : dflt; }

constexpr int BN_TX = 32;  // lanes along channels (float4 each => 128 channels per CTA column)
constexpr int BN_TY = 8;   // row slots per CTA

struct Dims {
    dim3 grid, block;
};
static inline Dims bn_dims(int64_t M, int CV) {
    Dims d;
    d.block = dim3(BN_TX, BN_TY);
    unsigned gx = (unsigned)ceil_div(CV, BN_TX);
    int64_t want = ((int64_t)kNumSMs * 8) / gx;  // ~8 CTAs per SM overall
    int64_t gy = ceil_div(M, BN_TY * 4);         // at least 4 rows per thread
    if (gy > want) gy = want;
    if (gy < 1) gy = 1;
    d.grid = dim3(gx, (unsigned)gy);
    return d;
}

__device__ __forceinline__ float4 ld4(const float* p, bool vec) {
    if (vec) return __ldg(reinterpret_cast<const float4*>(p));
    return make_float4(0.f, 0.f, 0.f, 0.f);
}

// Scratch contract (include/mpc_b200.h): the fp64 scratch [2C sums | ticket A | ticket B] is zero on entry and the
// kernel that consumes the sums leaves it zero again -- its last CTA to finish (ticket B) clears it -- so no memset
// launch is needed between uses.  Must be reached by every thread of every CTA (1-D blocks).
__device__ __forceinline__ void clear_scratch_when_last(double* scratch, int C) {
    __shared__ bool last_cta;
    __threadfence();  // this CTA's reads of the sums are done before its ticket is drawn
    __syncthreads();
    if (threadIdx.x == 0)
        last_cta = atomicAdd(reinterpret_cast<unsigned*>(scratch + 2 * C + 1), 1u) == gridDim.x * gridDim.y - 1;
    __syncthreads();
    if (last_cta)
        for (int i = threadIdx.x; i < 2 * C + 2; i += blockDim.x) scratch[i] = 0.0;
}
// Clears a buffer on behalf of a LATER launch in the stream (the weight-gradient GEMM's split-reduction target).
__device__ __forceinline__ void zero_service(float* buf, int64_t count) {
    if (!buf) return;
    float4* b4 = reinterpret_cast<float4*>(buf);
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count / 4; i += (int64_t)gridDim.x * blockDim.x)
        b4[i] = z;
}

// Reduce 4 channel partials (a: 4 values, b: 4 values) over the BN_TY row slots and add to the fp64 scratch.
__device__ __forceinline__ void block_reduce_to_scratch(float4 a, float4 b, double* __restrict__ sa,
                                                        double* __restrict__ sb, int cbase, int C) {
    __shared__ float4 ra[BN_TY][BN_TX], rb[BN_TY][BN_TX];
    ra[threadIdx.y][threadIdx.x] = a;
    rb[threadIdx.y][threadIdx.x] = b;
    __syncthreads();
    if (threadIdx.y == 0) {
        double A[4] = {0, 0, 0, 0}, Bv[4] = {0, 0, 0, 0};
#pragma unroll
        for (int y = 0; y < BN_TY; ++y) {
            float4 u = ra[y][threadIdx.x], w = rb[y][threadIdx.x];
            A[0] += u.x; A[1] += u.y; A[2] += u.z; A[3] += u.w;
            Bv[0] += w.x; Bv[1] += w.y; Bv[2] += w.z; Bv[3] += w.w;
        }
#pragma unroll
        for (int i = 0; i < 4; ++i)
            if (cbase + i < C) {
                atomicAdd(sa + cbase + i, A[i]);
                atomicAdd(sb + cbase + i, Bv[i]);
            }
    }
}

// per-channel sum and sum of squares (C % 4 == 0 path: float4; else scalar path with 1 channel per lane)
template <bool VEC4>
__global__ void __launch_bounds__(BN_TX* BN_TY)
bn_sums_kernel(const float* __restrict__ y, double* __restrict__ s1, double* __restrict__ s2, int64_t M, int C) {
    pdl_prologue();
    const int cw = VEC4 ? 4 : 1;
    const int cbase = (blockIdx.x * BN_TX + threadIdx.x) * cw;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
    if (cbase < C) {
        for (int64_t r = (int64_t)blockIdx.y * BN_TY + threadIdx.y; r < M; r += (int64_t)gridDim.y * BN_TY) {
            if (VEC4) {
                const float4 v = __ldg(reinterpret_cast<const float4*>(y + r * C + cbase));
                a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
                b.x = fmaf(v.x, v.x, b.x); b.y = fmaf(v.y, v.y, b.y);
                b.z = fmaf(v.z, v.z, b.z); b.w = fmaf(v.w, v.w, b.w);
            } else {
                const float v = __ldg(y + r * C + cbase);
                a.x += v;
                b.x = fmaf(v, v, b.x);
            }
        }
    }
    block_reduce_to_scratch(a, b, s1, s2, cbase, VEC4 ? C : min(C, cbase + 1));
}

// ---- fast path: C/4 is a power of two <= 256 (every BatchNorm width in both models) ------------------------
// 256 threads = (256/CV) row slots x CV float4 columns; each thread streams its rows with 4 independent
// 128-bit loads in flight, partials are reduced over the row slots in shared memory, one fp64 atomic per
// (CTA, channel, quantity); the last CTA to finish (atomic ticket) finalises, so no second launch is needed.
constexpr int RT = 256;

struct StatsFinal {
    float* stats;
    float* running_mean;
    float* running_var;
    int64_t* num_batches_tracked;
    float momentum;
};

__device__ __forceinline__ void finalize_channel(const double* s1, const double* s2, const StatsFinal& f, int64_t M,
                                                 int C, int c) {
    const double mean = __ldcg(s1 + c) / (double)M;
    double var = __ldcg(s2 + c) / (double)M - mean * mean;
    var = var < 0.0 ? 0.0 : var;
    f.stats[c] = (float)mean;
    f.stats[C + c] = (float)var;
    if (f.running_mean) f.running_mean[c] = (1.0f - f.momentum) * f.running_mean[c] + f.momentum * (float)mean;
    if (f.running_var) {
        const double unbiased = M > 1 ? var * ((double)M / (double)(M - 1)) : var;
        f.running_var[c] = (1.0f - f.momentum) * f.running_var[c] + f.momentum * (float)unbiased;
    }
}

// MODE 0: (y, y*y) for the forward statistics.  MODE 1: (dz, dz*xhat) for the backward reductions.
// MODE 2: plain column sums (bias gradient of a projection that is not followed by BatchNorm).
template <int MODE>
__global__ void __launch_bounds__(RT)
col_reduce_kernel(const float* __restrict__ y, const float* __restrict__ gout, const float* __restrict__ mean,
                  const float* __restrict__ var, const float* __restrict__ gamma, const float* __restrict__ beta,
                  float eps, float slope, double* __restrict__ s1, double* __restrict__ s2,
                  unsigned* __restrict__ ticket, StatsFinal fin, int64_t M, int C, int CV, int64_t gld4) {
    pdl_prologue();
    __shared__ float4 ra[RT], rb[RT];
    __shared__ bool last;
    const int v = threadIdx.x % CV, slot = threadIdx.x / CV, slots = RT / CV;
    const int c0 = v * 4;
    float4 mu = make_float4(0.f, 0.f, 0.f, 0.f), is = mu, g = mu, be = mu;
    if (MODE == 1) {
        mu = __ldg(reinterpret_cast<const float4*>(mean + c0));
        const float4 vr = __ldg(reinterpret_cast<const float4*>(var + c0));
        is = make_float4(1.0f / sqrtf(vr.x + eps), 1.0f / sqrtf(vr.y + eps), 1.0f / sqrtf(vr.z + eps),
                         1.0f / sqrtf(vr.w + eps));
        g = __ldg(reinterpret_cast<const float4*>(gamma + c0));
        be = __ldg(reinterpret_cast<const float4*>(beta + c0));
    }
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
    const int64_t stride = (int64_t)gridDim.x * slots;
    const float4* yp = reinterpret_cast<const float4*>(y);
    const float4* gp = reinterpret_cast<const float4*>(gout);
    auto accum = [&](const float4& yv, const float4& gv) {
        if (MODE == 0) {
            a.x += yv.x; a.y += yv.y; a.z += yv.z; a.w += yv.w;
            b.x = fmaf(yv.x, yv.x, b.x); b.y = fmaf(yv.y, yv.y, b.y);
            b.z = fmaf(yv.z, yv.z, b.z); b.w = fmaf(yv.w, yv.w, b.w);
        } else if (MODE == 2) {
            a.x += yv.x; a.y += yv.y; a.z += yv.z; a.w += yv.w;
        } else {
#define MPC_R(comp)                                                     \
    {                                                                   \
        const float xh = (yv.comp - mu.comp) * is.comp;                 \
        const float z = fmaf(xh, g.comp, be.comp);                      \
        const float dz = z > 0.f ? gv.comp : gv.comp * slope;           \
        a.comp += dz;                                                   \
        b.comp = fmaf(dz, xh, b.comp);                                  \
    }
            MPC_R(x) MPC_R(y) MPC_R(z) MPC_R(w)
#undef MPC_R
        }
    };
    int64_t r = (int64_t)blockIdx.x * slots + slot;
    for (; r + 3 * stride < M; r += 4 * stride) {  // 4 (8 in MODE 1) independent 128-bit loads in flight
        float4 y0 = __ldg(yp + r * CV + v), y1 = __ldg(yp + (r + stride) * CV + v);
        float4 y2 = __ldg(yp + (r + 2 * stride) * CV + v), y3 = __ldg(yp + (r + 3 * stride) * CV + v);
        float4 g0 = y0, g1 = y0, g2 = y0, g3 = y0;
        if (MODE == 1) {
            g0 = __ldg(gp + r * gld4 + v);  // (grad_out rows may be strided: a column slice of a wider gradient)
            g1 = __ldg(gp + (r + stride) * gld4 + v);
            g2 = __ldg(gp + (r + 2 * stride) * gld4 + v);
            g3 = __ldg(gp + (r + 3 * stride) * gld4 + v);
        }
        accum(y0, g0);
        accum(y1, g1);
        accum(y2, g2);
        accum(y3, g3);
    }
    for (; r < M; r += stride) {
        const float4 y0 = __ldg(yp + r * CV + v);
        const float4 g0 = MODE == 1 ? __ldg(gp + r * gld4 + v) : y0;
        accum(y0, g0);
    }
    ra[threadIdx.x] = a;
    rb[threadIdx.x] = b;
    __syncthreads();
    if (threadIdx.x < CV) {
        double A[4] = {0, 0, 0, 0}, Bv[4] = {0, 0, 0, 0};
        for (int sl = 0; sl < slots; ++sl) {
            const float4 u = ra[sl * CV + v], w = rb[sl * CV + v];
            A[0] += u.x; A[1] += u.y; A[2] += u.z; A[3] += u.w;
            Bv[0] += w.x; Bv[1] += w.y; Bv[2] += w.z; Bv[3] += w.w;
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            atomicAdd(s1 + c0 + i, A[i]);
            atomicAdd(s2 + c0 + i, Bv[i]);
        }
    }
    if (MODE == 0 || MODE == 2) {  // last CTA done: finalise in the same launch
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) last = atomicAdd(ticket, 1u) == gridDim.x - 1;
        __syncthreads();
        if (last) {
            __threadfence();
            if (MODE == 0) {
                for (int c = threadIdx.x; c < C; c += RT) finalize_channel(s1, s2, fin, M, C, c);
                if (threadIdx.x == 0 && fin.num_batches_tracked) *fin.num_batches_tracked += 1;
            } else {
                for (int c = threadIdx.x; c < C; c += RT) fin.stats[c] = (float)__ldcg(s1 + c);
            }
            // scratch contract: every other CTA has drawn its ticket, so nothing touches the sums any more
            for (int c = threadIdx.x; c < C; c += RT) s1[c] = s2[c] = 0.0;
            if (threadIdx.x == 0) *ticket = 0u;
        }
    }
}

static inline bool fast_cv(int64_t C) {
    if (C % 4) return false;
    const int64_t cv = C / 4;
    return cv >= 1 && cv <= 256 && (cv & (cv - 1)) == 0;
}
static inline unsigned col_reduce_grid(int64_t M, int CV) {
    const int slots = RT / CV;
    int64_t g = ceil_div(M, (int64_t)slots * 8);  // >= 8 rows per thread
    const int64_t cap = (int64_t)kNumSMs * knob(1, 2);  // few CTAs: the fp64 atomics per channel serialise in L2
    return (unsigned)(g < 1 ? 1 : (g > cap ? cap : g));
}

// mean / biased variance, plus nn.BatchNorm1d's running-statistics update (unbiased variance, momentum)
__global__ void bn_finalize_kernel(double* __restrict__ s1, double* __restrict__ s2,
                                   float* __restrict__ stats, float* __restrict__ running_mean,
                                   float* __restrict__ running_var, int64_t* __restrict__ num_batches_tracked,
                                   float momentum, int64_t M, int C) {
    pdl_prologue();
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c == 0 && num_batches_tracked) *num_batches_tracked += 1;
    if (c >= C) return;
    const double mean = s1[c] / (double)M;
    double var = s2[c] / (double)M - mean * mean;
    var = var < 0.0 ? 0.0 : var;
    stats[c] = (float)mean;
    stats[C + c] = (float)var;
    s1[c] = 0.0;  // scratch contract: each thread clears the channel it consumed
    s2[c] = 0.0;
    if (running_mean) running_mean[c] = (1.0f - momentum) * running_mean[c] + momentum * (float)mean;
    if (running_var) {
        const double unbiased = M > 1 ? var * ((double)M / (double)(M - 1)) : var;
        running_var[c] = (1.0f - momentum) * running_var[c] + momentum * (float)unbiased;
    }
}

template <bool VEC4>
__global__ void __launch_bounds__(256)
bn_act_fwd_kernel(const float* __restrict__ y, const float* __restrict__ mean, const float* __restrict__ var,
                  const float* __restrict__ gamma, const float* __restrict__ beta, float eps, float slope,
                  float* __restrict__ out, int C, int64_t total) {
    pdl_prologue();
    const int cw = VEC4 ? 4 : 1;
    for (int64_t t = (int64_t)blockIdx.x * 256 + threadIdx.x; t < total; t += (int64_t)gridDim.x * 256) {
        const int c = (int)((t * cw) % C);
        if (VEC4) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(y) + t);
            const float4 mu = __ldg(reinterpret_cast<const float4*>(mean + c));
            const float4 vr = __ldg(reinterpret_cast<const float4*>(var + c));
            const float4 g = __ldg(reinterpret_cast<const float4*>(gamma + c));
            const float4 be = __ldg(reinterpret_cast<const float4*>(beta + c));
            float4 o;
#define MPC_BN(comp)                                                        \
    {                                                                       \
        const float sc = g.comp * (1.0f / sqrtf(vr.comp + eps));            \
        const float z = fmaf(v.comp - mu.comp, sc, be.comp);                \
        o.comp = z > 0.f ? z : z * slope;                                   \
    }
            MPC_BN(x) MPC_BN(y) MPC_BN(z) MPC_BN(w)
#undef MPC_BN
            reinterpret_cast<float4*>(out)[t] = o;
        } else {
            const float sc = gamma[c] * (1.0f / sqrtf(var[c] + eps));
            const float z = fmaf(y[t] - mean[c], sc, beta[c]);
            out[t] = z > 0.f ? z : z * slope;
        }
    }
}

// ---- fast elementwise paths: 1024 % C == 0, so with 256 threads x float4 a thread's 4 channels never change across
// its grid-stride elements and every per-channel coefficient lives in registers (the generic kernels below
// re-derive them -- rsqrt, divisions, fp64 -> fp32 -- for every element: ~300 instructions per float4).
__global__ void __launch_bounds__(256)
bn_act_fwd_fast_kernel(const float4* __restrict__ y, const float* __restrict__ mean, const float* __restrict__ var,
                       const float* __restrict__ gamma, const float* __restrict__ beta, float eps, float slope,
                       const float4* __restrict__ residual, float4* __restrict__ out, int C, int64_t total) {
    pdl_prologue();
    const int c = (threadIdx.x * 4) % C;
    const float4 mu = __ldg(reinterpret_cast<const float4*>(mean + c));
    const float4 vr = __ldg(reinterpret_cast<const float4*>(var + c));
    const float4 g = __ldg(reinterpret_cast<const float4*>(gamma + c));
    const float4 be = __ldg(reinterpret_cast<const float4*>(beta + c));
    const float4 sc = make_float4(g.x * (1.0f / sqrtf(vr.x + eps)), g.y * (1.0f / sqrtf(vr.y + eps)),
                                  g.z * (1.0f / sqrtf(vr.z + eps)), g.w * (1.0f / sqrtf(vr.w + eps)));
    const int64_t stride = (int64_t)gridDim.x * 256;
    int64_t t = (int64_t)blockIdx.x * 256 + threadIdx.x;
    auto apply = [&](const float4& v) {
        float4 o;
        float z;
        z = fmaf(v.x - mu.x, sc.x, be.x); o.x = z > 0.f ? z : z * slope;
        z = fmaf(v.y - mu.y, sc.y, be.y); o.y = z > 0.f ? z : z * slope;
        z = fmaf(v.z - mu.z, sc.z, be.z); o.z = z > 0.f ? z : z * slope;
        z = fmaf(v.w - mu.w, sc.w, be.w); o.w = z > 0.f ? z : z * slope;
        return o;
    };
    if (residual) {  // out = act(BN(y)) + residual: the `residual + ffn(context)` of LocalTrans in the same pass
        for (; t < total; t += stride) {
            float4 o = apply(__ldg(y + t));
            const float4 r = __ldg(residual + t);
            o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w;
            out[t] = o;
        }
        return;
    }
    for (; t + 3 * stride < total; t += 4 * stride) {  // 4 independent 128-bit loads in flight
        const float4 v0 = __ldg(y + t), v1 = __ldg(y + t + stride), v2 = __ldg(y + t + 2 * stride),
                     v3 = __ldg(y + t + 3 * stride);
        out[t] = apply(v0);
        out[t + stride] = apply(v1);
        out[t + 2 * stride] = apply(v2);
        out[t + 3 * stride] = apply(v3);
    }
    for (; t < total; t += stride) out[t] = apply(__ldg(y + t));
}

__global__ void __launch_bounds__(256)
bn_bwd_apply_fast_kernel(const float4* __restrict__ gout, const float4* __restrict__ y,
                         const float* __restrict__ mean, const float* __restrict__ var,
                         const float* __restrict__ gamma, const float* __restrict__ beta, float eps, float slope,
                         int train, double* __restrict__ s1, double* __restrict__ s2,
                         float4* __restrict__ gy, float* __restrict__ ggamma, float* __restrict__ gbeta,
                         float* __restrict__ zero_buf, int64_t zero_count, int64_t M, int C, int64_t total,
                         int64_t gld4, int cv_shift) {
    pdl_prologue();
    zero_service(zero_buf, zero_count);
    if (blockIdx.x == 0)
        for (int i = threadIdx.x; i < C; i += 256) {
            gbeta[i] = (float)s1[i];
            ggamma[i] = (float)s2[i];
        }
    const int64_t cvm = ((int64_t)1 << cv_shift) - 1;  // C / 4 is a power of two on this path
    auto gat = [&](int64_t t) { return __ldg(gout + (t >> cv_shift) * gld4 + (t & cvm)); };  // strided grad_out rows
    const int c = (threadIdx.x * 4) % C;
    const float invM = train ? (float)(1.0 / (double)M) : 0.f;
    float mu[4], is[4], g[4], be[4], c1[4], c2[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        mu[i] = mean[c + i];
        is[i] = 1.0f / sqrtf(var[c + i] + eps);
        g[i] = gamma[c + i];
        be[i] = beta[c + i];
        c1[i] = invM * (float)s1[c + i];
        c2[i] = invM * (float)s2[c + i];
    }
    const int64_t stride = (int64_t)gridDim.x * 256;
    int64_t t = (int64_t)blockIdx.x * 256 + threadIdx.x;
    auto apply = [&](const float4& yv, const float4& gv) {
        const float ya[4] = {yv.x, yv.y, yv.z, yv.w}, ga[4] = {gv.x, gv.y, gv.z, gv.w};
        float o[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float xh = (ya[i] - mu[i]) * is[i];
            const float z = fmaf(xh, g[i], be[i]);
            const float dz = z > 0.f ? ga[i] : ga[i] * slope;
            o[i] = g[i] * is[i] * (dz - (c1[i] + xh * c2[i]));
        }
        return make_float4(o[0], o[1], o[2], o[3]);
    };
    for (; t + stride < total; t += 2 * stride) {  // 4 independent 128-bit loads in flight
        const float4 y0 = __ldg(y + t), g0 = gat(t), y1 = __ldg(y + t + stride), g1 = gat(t + stride);
        gy[t] = apply(y0, g0);
        gy[t + stride] = apply(y1, g1);
    }
    for (; t < total; t += stride) gy[t] = apply(__ldg(y + t), gat(t));
    clear_scratch_when_last(s1, C);
}

// Training forward straight from the GEMM epilogue's column sums: every thread derives mean / variance of its 4
// channels from the fp64 sums (so the separate finalise launch disappears); CTA 0 also publishes mean / biased
// variance for the backward pass and applies nn.BatchNorm1d's running-statistics update.
__global__ void __launch_bounds__(256)
bn_act_fwd_sums_kernel(const float4* __restrict__ y, double* __restrict__ s1, double* __restrict__ s2,
                       const float* __restrict__ gamma, const float* __restrict__ beta, float eps, float slope,
                       const float4* __restrict__ residual, float4* __restrict__ out, float* __restrict__ stats,
                       float* __restrict__ running_mean,
                       float* __restrict__ running_var, int64_t* __restrict__ num_batches_tracked, float momentum,
                       int64_t M, int C, int64_t total) {
    pdl_prologue();
    if (blockIdx.x == 0) {
        StatsFinal fin{stats, running_mean, running_var, num_batches_tracked, momentum};
        for (int i = threadIdx.x; i < C; i += 256) finalize_channel(s1, s2, fin, M, C, i);
        if (threadIdx.x == 0 && num_batches_tracked) *num_batches_tracked += 1;
    }
    const int c = (threadIdx.x * 4) % C;
    float mu[4], sc[4], be[4];
    const double invM = 1.0 / (double)M;  // one fp64 division per thread instead of eight
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const double mean = s1[c + i] * invM;
        double var = s2[c + i] * invM - mean * mean;
        var = var < 0.0 ? 0.0 : var;
        mu[i] = (float)mean;
        sc[i] = gamma[c + i] * (1.0f / sqrtf((float)var + eps));
        be[i] = beta[c + i];
    }
    const int64_t stride = (int64_t)gridDim.x * 256;
    int64_t t = (int64_t)blockIdx.x * 256 + threadIdx.x;
    auto apply = [&](const float4& v) {
        float4 o;
        float z;
        z = fmaf(v.x - mu[0], sc[0], be[0]); o.x = z > 0.f ? z : z * slope;
        z = fmaf(v.y - mu[1], sc[1], be[1]); o.y = z > 0.f ? z : z * slope;
        z = fmaf(v.z - mu[2], sc[2], be[2]); o.z = z > 0.f ? z : z * slope;
        z = fmaf(v.w - mu[3], sc[3], be[3]); o.w = z > 0.f ? z : z * slope;
        return o;
    };
    if (residual) {
        for (; t < total; t += stride) {
            float4 o = apply(__ldg(y + t));
            const float4 r = __ldg(residual + t);
            o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w;
            out[t] = o;
        }
    } else {
        for (; t + 3 * stride < total; t += 4 * stride) {
            const float4 v0 = __ldg(y + t), v1 = __ldg(y + t + stride), v2 = __ldg(y + t + 2 * stride),
                         v3 = __ldg(y + t + 3 * stride);
            out[t] = apply(v0);
            out[t + stride] = apply(v1);
            out[t + 2 * stride] = apply(v2);
            out[t + 3 * stride] = apply(v3);
        }
        for (; t < total; t += stride) out[t] = apply(__ldg(y + t));
    }
    clear_scratch_when_last(s1, C);
}

static inline bool fast_ew(int64_t C) { return C % 4 == 0 && C <= 1024 && 1024 % C == 0; }

// backward pass A: per channel sum(dz) and sum(dz * xhat)
template <bool VEC4>
__global__ void __launch_bounds__(BN_TX* BN_TY)
bn_bwd_sums_kernel(const float* __restrict__ gout, const float* __restrict__ y, const float* __restrict__ mean,
                   const float* __restrict__ var, const float* __restrict__ gamma, const float* __restrict__ beta,
                   float eps, float slope, double* __restrict__ s1, double* __restrict__ s2, int64_t M, int C) {
    pdl_prologue();
    const int cw = VEC4 ? 4 : 1;
    const int cbase = (blockIdx.x * BN_TX + threadIdx.x) * cw;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
    if (cbase < C) {
        float mu[4], is[4], g[4], be[4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
            if (i < cw) {
                mu[i] = mean[cbase + i];
                is[i] = 1.0f / sqrtf(var[cbase + i] + eps);
                g[i] = gamma[cbase + i];
                be[i] = beta[cbase + i];
            }
        for (int64_t r = (int64_t)blockIdx.y * BN_TY + threadIdx.y; r < M; r += (int64_t)gridDim.y * BN_TY) {
            float yv[4], gv[4];
            if (VEC4) {
                const float4 v = __ldg(reinterpret_cast<const float4*>(y + r * C + cbase));
                const float4 go = __ldg(reinterpret_cast<const float4*>(gout + r * C + cbase));
                yv[0] = v.x; yv[1] = v.y; yv[2] = v.z; yv[3] = v.w;
                gv[0] = go.x; gv[1] = go.y; gv[2] = go.z; gv[3] = go.w;
            } else {
                yv[0] = __ldg(y + r * C + cbase);
                gv[0] = __ldg(gout + r * C + cbase);
            }
            float da[4] = {0, 0, 0, 0}, db[4] = {0, 0, 0, 0};
#pragma unroll
            for (int i = 0; i < 4; ++i)
                if (i < cw) {
                    const float xh = (yv[i] - mu[i]) * is[i];
                    const float z = fmaf(xh, g[i], be[i]);
                    const float dz = z > 0.f ? gv[i] : gv[i] * slope;
                    da[i] = dz;
                    db[i] = dz * xh;
                }
            a.x += da[0]; a.y += da[1]; a.z += da[2]; a.w += da[3];
            b.x += db[0]; b.y += db[1]; b.z += db[2]; b.w += db[3];
        }
    }
    block_reduce_to_scratch(a, b, s1, s2, cbase, VEC4 ? C : min(C, cbase + 1));
}

// backward pass B: grad_y; also publishes grad_beta = s1, grad_gamma = s2 (CTA 0)
template <bool VEC4>
__global__ void __launch_bounds__(256)
bn_bwd_apply_kernel(const float* __restrict__ gout, const float* __restrict__ y, const float* __restrict__ mean,
                    const float* __restrict__ var, const float* __restrict__ gamma, const float* __restrict__ beta,
                    float eps, float slope, int train, double* __restrict__ s1, double* __restrict__ s2,
                    float* __restrict__ gy, float* __restrict__ ggamma, float* __restrict__ gbeta,
                    float* __restrict__ zero_buf, int64_t zero_count, int64_t M, int C, int64_t total) {
    pdl_prologue();
    zero_service(zero_buf, zero_count);
    if (blockIdx.x == 0)
        for (int c = threadIdx.x; c < C; c += 256) {
            gbeta[c] = (float)s1[c];
            ggamma[c] = (float)s2[c];
        }
    const int cw = VEC4 ? 4 : 1;
    const float invM = train ? (float)(1.0 / (double)M) : 0.f;
    for (int64_t t = (int64_t)blockIdx.x * 256 + threadIdx.x; t < total; t += (int64_t)gridDim.x * 256) {
        const int c0 = (int)((t * cw) % C);
        float yv[4], gv[4], o[4];
        if (VEC4) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(y) + t);
            const float4 go = __ldg(reinterpret_cast<const float4*>(gout) + t);
            yv[0] = v.x; yv[1] = v.y; yv[2] = v.z; yv[3] = v.w;
            gv[0] = go.x; gv[1] = go.y; gv[2] = go.z; gv[3] = go.w;
        } else {
            yv[0] = y[t];
            gv[0] = gout[t];
        }
#pragma unroll
        for (int i = 0; i < 4; ++i)
            if (i < cw) {
                const int c = c0 + i;
                const float is = 1.0f / sqrtf(var[c] + eps);
                const float g = gamma[c];
                const float xh = (yv[i] - mean[c]) * is;
                const float z = fmaf(xh, g, beta[c]);
                const float dz = z > 0.f ? gv[i] : gv[i] * slope;
                const float db = (float)s1[c], dg = (float)s2[c];
                o[i] = g * is * (dz - invM * (db + xh * dg));
            }
        if (VEC4)
            reinterpret_cast<float4*>(gy)[t] = make_float4(o[0], o[1], o[2], o[3]);
        else
            gy[t] = o[0];
    }
    clear_scratch_when_last(s1, C);
}

static inline unsigned ew_grid(int64_t total) {
    // sized for >= 4 float4 per thread (the unrolled loop keeps 4 loads in flight) and at most 4 CTAs per SM: a
    // per-thread prologue (coefficients from fp64 sums) amortised over 1-2 elements was the whole cost before
    int64_t g = ceil_div(total, 256 * knob(2, 4));
    const int64_t cap = (int64_t)kNumSMs * knob(0, 4);
    return (unsigned)(g < 1 ? 1 : (g > cap ? cap : g));
}
static inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace mpc

using namespace mpc;

MPC_API int mpc_bn_stats_f32(const float* y, float* stats, float* running_mean, float* running_var,
                             int64_t* num_batches_tracked, float momentum, double* scratch, int64_t M, int64_t C,
                             mpc_stream_t stream) {
    if (!y || !stats || !scratch || M <= 0 || C <= 0 || C > INT32_MAX / 4) return MPC_ERR_INVALID;
    cudaStream_t st = (cudaStream_t)stream;
    if (fast_cv(C) && al16(y)) {
        const int CV = (int)(C / 4);
        StatsFinal fin{stats, running_mean, running_var, num_batches_tracked, momentum};
        pdl_launch(col_reduce_kernel<0>, dim3(col_reduce_grid(M, CV)), dim3(RT), 0, st, 
            y, nullptr, nullptr, nullptr, nullptr, nullptr, 0.f, 0.f, scratch, scratch + C,
            reinterpret_cast<unsigned*>(scratch + 2 * C), fin, M, (int)C, CV, (int64_t)CV);
        MPC_LAUNCH_CHECK();
        return MPC_OK;
    }
    const bool v4 = C % 4 == 0 && al16(y);
    const Dims d = bn_dims(M, (int)(v4 ? C / 4 : C));
    if (v4)
        pdl_launch(bn_sums_kernel<true>, dim3(d.grid), dim3(d.block), 0, st, y, scratch, scratch + C, M, (int)C);
    else
        pdl_launch(bn_sums_kernel<false>, dim3(d.grid), dim3(d.block), 0, st, y, scratch, scratch + C, M, (int)C);
    MPC_LAUNCH_CHECK();
    pdl_launch(bn_finalize_kernel, dim3((unsigned)ceil_div(C, 128)), dim3(128), 0, st, scratch, scratch + C, stats, running_mean, running_var,
                                                                   num_batches_tracked, momentum, M, (int)C);
    MPC_LAUNCH_CHECK();
    return MPC_OK;
}

MPC_API int mpc_bn_act_fwd_f32(const float* y, const float* mean, const float* var, const float* gamma,
                               const float* beta, float eps, float slope, const float* residual, float* out, int64_t M,
                               int64_t C, mpc_stream_t stream) {
    if (!y || !mean || !var || !gamma || !beta || !out || M < 0 || C <= 0) return MPC_ERR_INVALID;
    if (residual && !(C % 4 == 0 && fast_ew(C) && al16(residual))) return MPC_ERR_UNSUPPORTED;
    if (M == 0) return MPC_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const bool v4 = C % 4 == 0 && al16(y) && al16(out) && al16(mean) && al16(var) && al16(gamma) && al16(beta);
    const int64_t total = M * (v4 ? C / 4 : C);
    if (v4 && fast_ew(C))
        pdl_launch(bn_act_fwd_fast_kernel, dim3(ew_grid(total)), dim3(256), 0, st, reinterpret_cast<const float4*>(y), mean, var, gamma, beta,
                                                              eps, slope, reinterpret_cast<const float4*>(residual),
                                                              reinterpret_cast<float4*>(out), (int)C, total);
    else if (v4)
        pdl_launch(bn_act_fwd_kernel<true>, dim3(ew_grid(total)), dim3(256), 0, st, y, mean, var, gamma, beta, eps, slope, out, (int)C, total);
    else
        pdl_launch(bn_act_fwd_kernel<false>, dim3(ew_grid(total)), dim3(256), 0, st, y, mean, var, gamma, beta, eps, slope, out, (int)C, total);
    MPC_LAUNCH_CHECK();
    return MPC_OK;
}

MPC_API int mpc_bn_act_bwd_f32(const float* grad_out, const float* y, const float* mean, const float* var,
                               const float* gamma, const float* beta, float eps, float slope, int train,
                               float* grad_y, float* grad_gamma, float* grad_beta, double* scratch, float* zero_buf,
                               int64_t zero_count, int64_t ld_gout, int64_t M, int64_t C, mpc_stream_t stream) {
    if (!grad_out || !y || !mean || !var || !gamma || !beta || !grad_y || !grad_gamma || !grad_beta || !scratch)
        return MPC_ERR_INVALID;
    if (M <= 0 || C <= 0) return MPC_ERR_INVALID;
    if (zero_buf && (!al16(zero_buf) || (zero_count & 3) || zero_count < 0)) return MPC_ERR_UNSUPPORTED;
    if (ld_gout < C) return MPC_ERR_INVALID;
    // strided grad_out rows only on the fast path (C a power of two in [4, 1024], 16-byte aligned rows)
    const bool fast = C % 4 == 0 && fast_cv(C) && fast_ew(C) && al16(y) && al16(grad_out) && al16(grad_y) && al16(mean) &&
                      al16(var) && al16(gamma) && al16(beta) && ld_gout % 4 == 0;
    if (ld_gout != C && !fast) return MPC_ERR_UNSUPPORTED;
    cudaStream_t st = (cudaStream_t)stream;
    const bool v4 = C % 4 == 0 && al16(y) && al16(grad_out) && al16(grad_y);
    const Dims d = bn_dims(M, (int)(v4 ? C / 4 : C));
    const int64_t total = M * (v4 ? C / 4 : C);
    if (v4) {
        if (fast_cv(C) && al16(mean) && al16(var) && al16(gamma) && al16(beta)) {
            const int CV = (int)(C / 4);
            pdl_launch(col_reduce_kernel<1>, dim3(col_reduce_grid(M, CV)), dim3(RT), 0, st, y, grad_out, mean, var, gamma, beta, eps, slope,
                                                                       scratch, scratch + C, nullptr, StatsFinal{}, M,
                                                                       (int)C, CV, ld_gout / 4);
        } else {
            pdl_launch(bn_bwd_sums_kernel<true>, dim3(d.grid), dim3(d.block), 0, st, grad_out, y, mean, var, gamma, beta, eps, slope,
                                                                scratch, scratch + C, M, (int)C);
        }
        MPC_LAUNCH_CHECK();
        if (fast_ew(C))
            pdl_launch(bn_bwd_apply_fast_kernel, dim3(ew_grid(total)), dim3(256), 0, st, 
                reinterpret_cast<const float4*>(grad_out), reinterpret_cast<const float4*>(y), mean, var, gamma, beta, eps,
                slope, train, scratch, scratch + C, reinterpret_cast<float4*>(grad_y), grad_gamma, grad_beta, zero_buf,
                zero_count, M, (int)C, total, ld_gout / 4, __builtin_ctz((unsigned)(C / 4)));
        else
            pdl_launch(bn_bwd_apply_kernel<true>, dim3(ew_grid(total)), dim3(256), 0, st, grad_out, y, mean, var, gamma, beta, eps, slope,
                                                                     train, scratch, scratch + C, grad_y, grad_gamma,
                                                                     grad_beta, zero_buf, zero_count, M, (int)C, total);
    } else {
        pdl_launch(bn_bwd_sums_kernel<false>, dim3(d.grid), dim3(d.block), 0, st, grad_out, y, mean, var, gamma, beta, eps, slope, scratch,
                                                             scratch + C, M, (int)C);
        MPC_LAUNCH_CHECK();
        pdl_launch(bn_bwd_apply_kernel<false>, dim3(ew_grid(total)), dim3(256), 0, st, grad_out, y, mean, var, gamma, beta, eps, slope,
                                                                  train, scratch, scratch + C, grad_y, grad_gamma,
                                                                  grad_beta, zero_buf, zero_count, M, (int)C, total);
    }
    MPC_LAUNCH_CHECK();
    return MPC_OK;
}

MPC_API int mpc_bn_finalize_f32(double* sums, float* stats, float* running_mean, float* running_var,
                                int64_t* num_batches_tracked, float momentum, int64_t M, int64_t C,
                                mpc_stream_t stream) {
    if (!sums || !stats || M <= 0 || C <= 0) return MPC_ERR_INVALID;
    pdl_launch(bn_finalize_kernel, dim3((unsigned)ceil_div(C, 128)), dim3(128), 0, (cudaStream_t)stream, 
        sums, sums + C, stats, running_mean, running_var, num_batches_tracked, momentum, M, (int)C);
    MPC_LAUNCH_CHECK();
    return MPC_OK;
}

MPC_API int mpc_bn_act_fwd_sums_f32(const float* y, double* sums, const float* gamma, const float* beta, float eps,
                                    float slope, const float* residual, float* out, float* stats, float* running_mean,
                                    float* running_var, int64_t* num_batches_tracked, float momentum, int64_t M,
                                    int64_t C, mpc_stream_t stream) {
    if (!y || !sums || !gamma || !beta || !out || !stats || M <= 0 || C <= 0) return MPC_ERR_INVALID;
    if (!fast_ew(C) || !al16(y) || !al16(out) || (residual && !al16(residual))) return MPC_ERR_UNSUPPORTED;
    const int64_t total = M * (C / 4);
    pdl_launch(bn_act_fwd_sums_kernel, dim3(ew_grid(total)), dim3(256), 0, (cudaStream_t)stream, 
        reinterpret_cast<const float4*>(y), sums, sums + C, gamma, beta, eps, slope,
        reinterpret_cast<const float4*>(residual), reinterpret_cast<float4*>(out), stats, running_mean, running_var,
        num_batches_tracked, momentum, M, (int)C, total);
    MPC_LAUNCH_CHECK();
    return MPC_OK;
}

MPC_API int mpc_col_sum_f32(const float* y, float* out, double* scratch, int64_t M, int64_t C, mpc_stream_t stream) {
    if (!y || !out || !scratch || M <= 0 || C <= 0) return MPC_ERR_INVALID;
    if (!fast_cv(C) || !al16(y)) return MPC_ERR_UNSUPPORTED;
    cudaStream_t st = (cudaStream_t)stream;
    const int CV = (int)(C / 4);
    StatsFinal fin{out, nullptr, nullptr, nullptr, 0.f};
    pdl_launch(col_reduce_kernel<2>, dim3(col_reduce_grid(M, CV)), dim3(RT), 0, st, y, nullptr, nullptr, nullptr, nullptr, nullptr, 0.f, 0.f,
                                                               scratch, scratch + C,
                                                               reinterpret_cast<unsigned*>(scratch + 2 * C), fin, M, (int)C,
                                                               CV, (int64_t)CV);
    MPC_LAUNCH_CHECK();
    return MPC_OK;
}

MPC_API int mpc_debug_set_knob(int id, int64_t value) {
    if (id < 0 || id >= 8) return MPC_ERR_INVALID;
    mpc::g_knob[id] = value;
    return MPC_OK;
}

MPC_API int mpc_version(void) { return 100; }  /* 0.1.0 */
MPC_API int mpc_compiled_arch(void) { return 1000; }

// ---------------------------------------------------------------------------------------------------------------
// Label-smoothed cross entropy of the part-seg head (get_loss, R/models/repsurf/pointnet2_part_seg_msg.py:159-180):
// one_hot*(1-eps) + (1-one_hot)*eps/(C-1) against log_softmax(pred), summed over classes, mean over rows.  The
// reference's chain is ~10 elementwise / reduce launches over [M,C]; here one warp per row does it in one pass
// (forward) and one pass (backward), rows may be strided (the 50-class head writes 52-float rows).
// ---------------------------------------------------------------------------------------------------------------
namespace mpc {

constexpr int CE_WARPS = 8;

__global__ void __launch_bounds__(CE_WARPS * 32)
smooth_ce_fwd_kernel(const float* __restrict__ pred, int64_t ld, const int64_t* __restrict__ target, float eps,
                     float* __restrict__ lse_out, float* __restrict__ loss, double* __restrict__ scratch, int64_t M,
                     int C) {
    pdl_prologue();
    __shared__ double part[CE_WARPS];
    __shared__ bool last;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double acc = 0.0;
    for (int64_t row = (int64_t)blockIdx.x * CE_WARPS + warp; row < M; row += (int64_t)gridDim.x * CE_WARPS) {
        const float* x = pred + row * ld;
        float mx = -INFINITY, sx = 0.f;
        for (int c = lane; c < C; c += 32) {
            const float v = x[c];
            mx = fmaxf(mx, v);
            sx += v;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
            sx += __shfl_xor_sync(0xffffffffu, sx, o);
        }
        float se = 0.f;
        for (int c = lane; c < C; c += 32) se += expf(x[c] - mx);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) se += __shfl_xor_sync(0xffffffffu, se, o);
        const float lse = mx + logf(se);
        if (lane == 0) {
            const int t = clamp_index(target[row], C);
            const float lp_t = x[t] - lse;
            const float sum_lp = sx - (float)C * lse;  // sum over classes of log p
            const float off = eps / (float)(C - 1);
            acc += (double)(-((1.0f - eps) * lp_t + off * (sum_lp - lp_t)));
            lse_out[row] = lse;
        }
    }
    if (lane == 0) part[warp] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int w = 0; w < CE_WARPS; ++w) s += part[w];
        atomicAdd(scratch, s);
        __threadfence();
        last = atomicAdd(reinterpret_cast<unsigned*>(scratch + 1), 1u) == gridDim.x - 1;
        if (last) {  // scratch contract: the last CTA publishes the mean and leaves the scratch zeroed
            __threadfence();
            *loss = (float)(__ldcg(scratch) / (double)M);
            scratch[0] = 0.0;
            scratch[1] = 0.0;
        }
    }
}

__global__ void __launch_bounds__(CE_WARPS * 32)
smooth_ce_bwd_kernel(const float* __restrict__ pred, int64_t ld, const int64_t* __restrict__ target, float eps,
                     const float* __restrict__ lse, const float* __restrict__ grad_loss, float* __restrict__ grad_pred,
                     int64_t ldg, int64_t M, int C) {
    pdl_prologue();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float g = __ldg(grad_loss) / (float)M;
    const float off = eps / (float)(C - 1);
    for (int64_t row = (int64_t)blockIdx.x * CE_WARPS + warp; row < M; row += (int64_t)gridDim.x * CE_WARPS) {
        const float* x = pred + row * ld;
        const float l = lse[row];
        const int t = clamp_index(target[row], C);
        for (int c = lane; c < C; c += 32) {
            const float w = c == t ? 1.0f - eps : off;
            grad_pred[row * ldg + c] = g * (expf(x[c] - l) - w);  // the smoothed target sums to 1
        }
    }
}

}  // namespace mpc

MPC_API int mpc_smooth_ce_fwd_f32(const float* pred, int64_t ld, const int64_t* target, float eps, float* lse,
                                  float* loss, double* scratch, int64_t M, int64_t C, mpc_stream_t stream) {
    if (!pred || !target || !lse || !loss || !scratch || M <= 0 || C <= 1 || ld < C) return MPC_ERR_INVALID;
    if (C > INT32_MAX) return MPC_ERR_UNSUPPORTED;
    int64_t g = ceil_div(M, CE_WARPS * 4);
    const int64_t cap = (int64_t)kNumSMs * 8;
    pdl_launch(smooth_ce_fwd_kernel, dim3((unsigned)(g > cap ? cap : g)), dim3(CE_WARPS * 32), 0, (cudaStream_t)stream,
               pred, ld, target, eps, lse, loss, scratch, M, (int)C);
    MPC_LAUNCH_CHECK();
    return MPC_OK;
}

MPC_API int mpc_smooth_ce_bwd_f32(const float* pred, int64_t ld, const int64_t* target, float eps, const float* lse,
                                  const float* grad_loss, float* grad_pred, int64_t ldg, int64_t M, int64_t C,
                                  mpc_stream_t stream) {
    if (!pred || !target || !lse || !grad_loss || !grad_pred || M <= 0 || C <= 1 || ld < C || ldg < C)
        return MPC_ERR_INVALID;
    int64_t g = ceil_div(M, CE_WARPS * 4);
    const int64_t cap = (int64_t)kNumSMs * 8;
    pdl_launch(smooth_ce_bwd_kernel, dim3((unsigned)(g > cap ? cap : g)), dim3(CE_WARPS * 32), 0, (cudaStream_t)stream,
               pred, ld, target, eps, lse, grad_loss, grad_pred, ldg, M, (int)C);
    MPC_LAUNCH_CHECK();
    return MPC_OK;
}
