// Umbrella surface features (RepSurf) for sm_100a -- SURVEY.md 8f row f1.  See include/mpc_b200.h
// (mpc_umbrella_features_f32) and R/modules/pointnet2_utils.py:310-378, R/modules/recons_utils.py (cal_normal,
// cal_center, cal_const, check_nan_umb), R/modules/polar_utils.py:10-31 (xyz2sphere).
//
// One thread per point: the G = k-1 neighbours (self excluded) are gathered relative to the point, sorted by
// azimuth in registers (stable insertion network on (phi, rank)), and the G triangles (origin, s_g, s_{g+1}) yield
// centroid / spherical coordinates / unit normal / plane constant without ever materialising the reference's
// [B,N,G,3,3] umbrella tensor, its argsort or its four advanced-indexing passes.  The 10 x G floats of a point are
// staged through shared memory so that the CTA writes its [128 points x G x C] slab with coalesced 128-bit stores.
#include <math_constants.h>

#include "common.cuh"

namespace mpc {

constexpr int UT = 128;  // points per CTA

__device__ __forceinline__ void sphere3(float x, float y, float z, float& rho, float& th, float& ph) {
    rho = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z)));
    th = rho == 0.f ? 0.f : acosf(z / rho) / CUDART_PI_F;
    ph = atan2f(y, x) / (2.0f * CUDART_PI_F) + 0.5f;
}

template <int G>
__global__ void __launch_bounds__(UT)
umbrella_kernel(const float* __restrict__ xyz, const int64_t* __restrict__ idx, int64_t ldk,
                const float* __restrict__ sign, float* __restrict__ out, int N, int C, int64_t total) {
    pdl_prologue();
    extern __shared__ float stage[];  // [UT][G*C + 1] (odd row pitch: conflict-free per-thread rows)
    const int pitch = G * C + 1;
    const int64_t p = (int64_t)blockIdx.x * UT + threadIdx.x;
    float* mine = stage + threadIdx.x * pitch;
    if (p < total) {
        const int b = (int)(p / N);
        const float* cloud = xyz + (size_t)b * N * 3;
        const float cx = cloud[(p - (int64_t)b * N) * 3 + 0], cy = cloud[(p - (int64_t)b * N) * 3 + 1],
                    cz = cloud[(p - (int64_t)b * N) * 3 + 2];
        float rx[G], ry[G], rz[G], key[G];
#pragma unroll
        for (int g = 0; g < G; ++g) {
            const int n = clamp_index(idx[p * ldk + 1 + g], N);  // column 0 is the point itself
            rx[g] = __fsub_rn(cloud[n * 3 + 0], cx);
            ry[g] = __fsub_rn(cloud[n * 3 + 1], cy);
            rz[g] = __fsub_rn(cloud[n * 3 + 2], cz);
            key[g] = atan2f(ry[g], rx[g]) / (2.0f * CUDART_PI_F) + 0.5f;
        }
        // stable ascending sort by azimuth (insertion network, fully unrolled: everything stays in registers)
#pragma unroll
        for (int i = 1; i < G; ++i) {
#pragma unroll
            for (int j = i; j > 0; --j) {
                const bool sw = key[j] < key[j - 1];
                const float k0 = sw ? key[j] : key[j - 1], k1 = sw ? key[j - 1] : key[j];
                const float x0 = sw ? rx[j] : rx[j - 1], x1 = sw ? rx[j - 1] : rx[j];
                const float y0 = sw ? ry[j] : ry[j - 1], y1 = sw ? ry[j - 1] : ry[j];
                const float z0 = sw ? rz[j] : rz[j - 1], z1 = sw ? rz[j - 1] : rz[j];
                key[j - 1] = k0; key[j] = k1;
                rx[j - 1] = x0; rx[j] = x1;
                ry[j - 1] = y0; ry[j] = y1;
                rz[j - 1] = z0; rz[j] = z1;
            }
        }
        const float flip = sign ? sign[b] : 1.0f;
        float nx[G], ny[G], nz[G], ex[G], ey[G], ez[G], cs[G];
        bool bad[G];
#pragma unroll
        for (int g = 0; g < G; ++g) {
            const int h = (g + 1) % G;
            // cross(s_g, s_h) with separately rounded products (torch.cross: a1*b2 - a2*b1 ...)
            const float ux = __fsub_rn(__fmul_rn(ry[g], rz[h]), __fmul_rn(rz[g], ry[h]));
            const float uy = __fsub_rn(__fmul_rn(rz[g], rx[h]), __fmul_rn(rx[g], rz[h]));
            const float uz = __fsub_rn(__fmul_rn(rx[g], ry[h]), __fmul_rn(ry[g], rx[h]));
            const float nrm = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(ux, ux), __fmul_rn(uy, uy)), __fmul_rn(uz, uz)));
            nx[g] = ux / nrm;
            ny[g] = uy / nrm;
            nz[g] = uz / nrm;
            ex[g] = __fadd_rn(__fadd_rn(0.f, rx[g]), rx[h]) / 3.0f;
            ey[g] = __fadd_rn(__fadd_rn(0.f, ry[g]), ry[h]) / 3.0f;
            ez[g] = __fadd_rn(__fadd_rn(0.f, rz[g]), rz[h]) / 3.0f;
        }
        const float pos = (nx[0] > 0.f ? 1.0f : -1.0f) * flip;  // first triangle's x made positive, then random_inv
        int first = 0;
        bool found = false;
#pragma unroll
        for (int g = 0; g < G; ++g) {
            nx[g] *= pos;
            ny[g] *= pos;
            nz[g] *= pos;
            cs[g] = __fadd_rn(__fadd_rn(__fmul_rn(nx[g], ex[g]), __fmul_rn(ny[g], ey[g])), __fmul_rn(nz[g], ez[g])) /
                    1.7320508075688772f;
            bad[g] = isnan(nx[g]) || isnan(ny[g]) || isnan(nz[g]);
            if (!bad[g] && !found) {
                found = true;
                first = g;
            }
        }
        // values of the first valid triangle (triangle 0 when none is valid: NaNs stay, as in the reference)
        float fx = 0.f, fy = 0.f, fz = 0.f, fex = 0.f, fey = 0.f, fez = 0.f, fc = 0.f;
#pragma unroll
        for (int g = 0; g < G; ++g)
            if (g == first) {
                fx = nx[g]; fy = ny[g]; fz = nz[g];
                fex = ex[g]; fey = ey[g]; fez = ez[g];
                fc = cs[g];
            }
#pragma unroll
        for (int g = 0; g < G; ++g) {
            float rho, th, ph;
            sphere3(ex[g], ey[g], ez[g], rho, th, ph);  // polar channels keep the triangle's own centroid
            float* o = mine + g * C;
            o[0] = bad[g] ? fex : ex[g];
            o[1] = bad[g] ? fey : ey[g];
            o[2] = bad[g] ? fez : ez[g];
            o[3] = rho;
            o[4] = th;
            o[5] = ph;
            o[6] = bad[g] ? fx : nx[g];
            o[7] = bad[g] ? fy : ny[g];
            o[8] = bad[g] ? fz : nz[g];
            if (C == 10) o[9] = bad[g] ? fc : cs[g];
        }
    }
    __syncthreads();
    // coalesced copy-out of this CTA's slab
    const int64_t p0 = (int64_t)blockIdx.x * UT;
    const int rows = (int)min((int64_t)UT, total - p0);
    const int per = G * C;
    float* dst = out + p0 * per;
    for (int i = threadIdx.x; i < rows * per; i += UT) dst[i] = stage[(i / per) * pitch + (i % per)];
}

}  // namespace mpc

using namespace mpc;

MPC_API int mpc_umbrella_features_f32(const float* xyz, const int64_t* idx, int64_t ldk, const float* sign, float* out,
                                      int64_t B, int64_t N, int64_t k, int64_t C, mpc_stream_t stream) {
    if (B < 0 || N <= 0 || (C != 9 && C != 10) || k < 3 || ldk < k) return MPC_ERR_INVALID;
    if (B == 0) return MPC_OK;
    if (!xyz || !idx || !out) return MPC_ERR_INVALID;
    if (N > INT32_MAX / 3) return MPC_ERR_UNSUPPORTED;
    const int64_t total = B * N;
    const unsigned grid = (unsigned)ceil_div(total, UT);
    cudaStream_t st = (cudaStream_t)stream;
    const int G = (int)k - 1;
    const size_t smem = (size_t)UT * (G * C + 1) * sizeof(float);
#define MPC_UMB(GG)                                                                                                   \
    case GG:                                                                                                          \
        if (smem > 48 * 1024)                                                                                         \
            MPC_CUDA(cudaFuncSetAttribute(umbrella_kernel<GG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        pdl_launch(umbrella_kernel<GG>, dim3(grid), dim3(UT), smem, st, xyz, idx, ldk, sign, out, (int)N, (int)C,    \
                   total);                                                                                            \
        break;
    switch (G) {
        MPC_UMB(2) MPC_UMB(3) MPC_UMB(4) MPC_UMB(5) MPC_UMB(6) MPC_UMB(7) MPC_UMB(8) MPC_UMB(9) MPC_UMB(10) MPC_UMB(11)
        MPC_UMB(12) MPC_UMB(15)
        default:
            return MPC_ERR_UNSUPPORTED;
    }
#undef MPC_UMB
    MPC_LAUNCH_CHECK();
    return MPC_OK;
}
