/*
 * pointset_oracle.c -- CPU restatement of the reference's point-set ops.  TEST INFRASTRUCTURE ONLY.
 *
 * This file is the checker for the CUDA path; it is never linked into, imported by, or executed from
 * the product package.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load liboracle.so.
 *
 * Parity status: the reference (ssr0512/Markov-Process-Analysis-on-Point-Cloud) ships NO tests and NO
 * golden vectors for this path ("parity unpinned" by the reference's own tests).  The
 * restatement is pinned instead against outputs of the reference itself, executed in the build
 * container by tests/golden/make_golden.py and committed under tests/golden/.
 *
 * R = Markov_Process_Analysis_on_Point_Cloud/ in the reference tree.  The arithmetic of every function
 * is fixed to one explicit IEEE-754 binary32 evaluation order so that the CUDA kernels can be
 * bit-identical to it:
 *   - FPS            R/modules/pointnet2_utils.py:84-109   (== R/modules/repsurface_utils.py:150-172)
 *   - square_distance R/modules/pointnet2_utils.py:190-209
 *   - knn_point       R/modules/pointnet2_utils.py:211-222
 *   - query_ball_point R/modules/pointnet2_utils.py:112-134
 *   - three_nn        R/modules/pointnet2_utils.py:899-901
 *   - upsample (Markov state transition) R/modules/pointnet2_utils.py:13-50
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off -fopenmp).  -ffp-contract=off matters: every
 * fused multiply-add below is written as an explicit fmaf() and nothing else may be contracted.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------------------------------
 * Farthest point sampling.  R/modules/pointnet2_utils.py:84-109.
 *   distance = 1e10 (:95); start index supplied by the caller (the reference draws it with
 *   torch.randint on the CPU generator, :96); per iteration: record `farthest` (:99), dist =
 *   sum_c (xyz - centroid)^2 evaluated as separate sub / mul / add ops, left to right, NO fma (:105);
 *   distance = dist where dist < distance (:106-107); farthest = argmax(distance), first (lowest)
 *   index on ties (:108).
 * xyz [B,N,C] fp32, start [B] int64, out [B,npoint] int64.
 * ---------------------------------------------------------------------------------------------- */
ORC_API int orc_fps_f32(const float* xyz, const int64_t* start, int64_t* out, int64_t B, int64_t N,
                        int64_t C, int64_t npoint) {
    if (B < 0 || N <= 0 || C <= 0 || npoint < 0) return 1;
#pragma omp parallel for schedule(dynamic, 1)
    for (int64_t b = 0; b < B; ++b) {
        const float* p = xyz + b * N * C;
        float* mind = (float*)malloc(sizeof(float) * (size_t)N);
        float* pt = (float*)malloc(sizeof(float) * (size_t)(N * C)); /* channel-major copy: unit-stride inner loops */
        for (int64_t n = 0; n < N; ++n) {
            mind[n] = 1e10f;
            for (int64_t k = 0; k < C; ++k) pt[k * N + n] = p[n * C + k];
        }
        float* acc = (float*)malloc(sizeof(float) * (size_t)N);
        int64_t far = start[b];
        for (int64_t i = 0; i < npoint; ++i) {
            out[b * npoint + i] = far;
            const float* c = p + far * C;
            /* the per-point arithmetic is the same scalar sequence as before (sub, mul, add: no fma), evaluated for
             * all points channel by channel so that the compiler can use SIMD lanes: lane-wise IEEE ops, same bits */
            {
                const float c0 = c[0];
                const float* x = pt;
#pragma omp simd
                for (int64_t n = 0; n < N; ++n) {
                    float d0 = x[n] - c0;
                    acc[n] = d0 * d0;
                }
            }
            for (int64_t k = 1; k < C; ++k) {
                const float ck = c[k];
                const float* x = pt + k * N;
#pragma omp simd
                for (int64_t n = 0; n < N; ++n) {
                    float dk = x[n] - ck;
                    acc[n] = acc[n] + dk * dk; /* -ffp-contract=off: mul then add, two roundings */
                }
            }
            float best = -INFINITY;
#pragma omp simd reduction(max : best)
            for (int64_t n = 0; n < N; ++n) {
                float m = mind[n];
                m = acc[n] < m ? acc[n] : m;
                mind[n] = m;
                best = m > best ? m : best;
            }
            int64_t besti = 0;
            for (int64_t n = 0; n < N; ++n) { /* first (lowest) index attaining the maximum: torch.max's tie rule */
                if (mind[n] == best) {
                    besti = n;
                    break;
                }
            }
            far = besti;
        }
        free(acc);
        free(pt);
        free(mind);
    }
    return 0;
}

/* Squared norm, sequential, non-fused: ((x0*x0 + x1*x1) + x2*x2) + ...  Bit-identical to
 * torch.sum(x ** 2, -1) on CPU for C = 3 (checked in tests/golden/make_golden.py). */
static inline float sqnorm_seq(const float* x, int64_t C) {
    float acc = x[0] * x[0];
    for (int64_t k = 1; k < C; ++k) acc = acc + x[k] * x[k];
    return acc;
}

/* Expanded-form squared distance of R/modules/pointnet2_utils.py:204-208:
 *   dist = -2 * (src . dst); dist += |src|^2; dist += |dst|^2     (that order)
 * The dot product is an fma chain over c = 0..C-1 starting from 0 (bit-identical to the CPU sgemm
 * the reference's torch.matmul reaches for C = 3). */
static inline float sqdist_expanded(const float* q, float qn, const float* r, float rn, int64_t C) {
    float dot = 0.0f;
    for (int64_t k = 0; k < C; ++k) dot = fmaf(q[k], r[k], dot);
    float d = -2.0f * dot;
    d = d + qn;
    d = d + rn;
    return d;
}

/* ------------------------------------------------------------------------------------------------
 * knn_point(nsample, xyz, new_xyz): R/modules/pointnet2_utils.py:211-222.
 *   sqrdists = square_distance(new_xyz, xyz); topk(k, largest=False, sorted=True).
 * ref [B,N,C] (the reference's `xyz`), qry [B,S,C] (`new_xyz`); dist_out [B,S,K] fp32 ascending,
 * idx_out [B,S,K] int64.  torch.topk's order among exactly equal distances is implementation
 * defined; this restatement (and the CUDA kernel) fix it to ascending index.  K <= N required.
 * ---------------------------------------------------------------------------------------------- */
#define ORC_KNN_LANES 32
ORC_API int orc_knn_f32(const float* ref, const float* qry, float* dist_out, int64_t* idx_out,
                        int64_t B, int64_t N, int64_t S, int64_t C, int64_t K) {
    if (K <= 0 || K > N || C <= 0) return 1;
    /* channel-major, lane-padded copy of the reference set, so that ORC_KNN_LANES pairs run their fma chains side
     * by side in SIMD lanes: every pair still sees exactly the scalar sequence of sqdist_expanded (same bits) */
    const int64_t Np = (N + ORC_KNN_LANES - 1) / ORC_KNN_LANES * ORC_KNN_LANES;
    float* rnorm = (float*)malloc(sizeof(float) * (size_t)(B * Np));
    float* rt = (float*)malloc(sizeof(float) * (size_t)(B * C * Np));
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < B * Np; ++i) {
        int64_t b = i / Np, n = i % Np;
        if (n < N) {
            rnorm[i] = sqnorm_seq(ref + (b * N + n) * C, C);
            for (int64_t k = 0; k < C; ++k) rt[(b * C + k) * Np + n] = ref[(b * N + n) * C + k];
        } else {
            rnorm[i] = 0.0f;
            for (int64_t k = 0; k < C; ++k) rt[(b * C + k) * Np + n] = 0.0f;
        }
    }
#pragma omp parallel for schedule(dynamic, 16)
    for (int64_t bs = 0; bs < B * S; ++bs) {
        int64_t b = bs / S;
        const float* q = qry + bs * C;
        float qn = sqnorm_seq(q, C);
        float* bd = dist_out + bs * K;
        int64_t* bi = idx_out + bs * K;
        int64_t cnt = 0;
        const float* rtb = rt + b * C * Np;
        const float* rnb = rnorm + b * Np;
        for (int64_t n0 = 0; n0 < N; n0 += ORC_KNN_LANES) {
            float dot[ORC_KNN_LANES], dd[ORC_KNN_LANES];
#pragma omp simd
            for (int j = 0; j < ORC_KNN_LANES; ++j) dot[j] = 0.0f;
            for (int64_t k = 0; k < C; ++k) {
                const float qk = q[k];
                const float* col = rtb + k * Np + n0;
#pragma omp simd
                for (int j = 0; j < ORC_KNN_LANES; ++j) dot[j] = fmaf(qk, col[j], dot[j]);
            }
#pragma omp simd
            for (int j = 0; j < ORC_KNN_LANES; ++j) {
                float d = -2.0f * dot[j];
                d = d + qn;
                dd[j] = d + rnb[n0 + j];
            }
            const int64_t lim = N - n0 < ORC_KNN_LANES ? N - n0 : ORC_KNN_LANES;
            for (int64_t j = 0; j < lim; ++j) {
                const float d = dd[j];
                if (cnt == K && !(d < bd[K - 1])) continue;
                int64_t pos = cnt < K ? cnt : K - 1;
                while (pos > 0 && d < bd[pos - 1]) { /* strict <: equal distances keep ascending index */
                    bd[pos] = bd[pos - 1];
                    bi[pos] = bi[pos - 1];
                    --pos;
                }
                bd[pos] = d;
                bi[pos] = n0 + j;
                if (cnt < K) ++cnt;
            }
        }
    }
    free(rt);
    free(rnorm);
    return 0;
}

/* ------------------------------------------------------------------------------------------------
 * query_ball_point(radius, nsample, xyz, new_xyz): R/modules/pointnet2_utils.py:112-134.
 * Indices n with NOT(sqdist > radius^2) in ascending n (:125-128), first nsample, padded with the
 * first hit (:129-131); a query with no hit keeps the out-of-range value N (latent reference bug,
 * preserved).  r2 is radius**2 rounded to fp32 by the caller (the comparison at :127 is fp32).
 * ---------------------------------------------------------------------------------------------- */
ORC_API int orc_ball_query_f32(const float* xyz, const float* new_xyz, int64_t* idx_out, float r2,
                               int64_t B, int64_t N, int64_t S, int64_t C, int64_t nsample) {
    if (nsample <= 0 || C <= 0) return 1;
    float* rnorm = (float*)malloc(sizeof(float) * (size_t)(B * N));
    for (int64_t i = 0; i < B * N; ++i) rnorm[i] = sqnorm_seq(xyz + i * C, C);
#pragma omp parallel for schedule(dynamic, 16)
    for (int64_t bs = 0; bs < B * S; ++bs) {
        int64_t b = bs / S;
        const float* q = new_xyz + bs * C;
        float qn = sqnorm_seq(q, C);
        int64_t* o = idx_out + bs * nsample;
        int64_t cnt = 0;
        for (int64_t n = 0; n < N && cnt < nsample; ++n) {
            float d = sqdist_expanded(q, qn, xyz + (b * N + n) * C, rnorm[b * N + n], C);
            if (!(d > r2)) o[cnt++] = n;
        }
        int64_t first = cnt > 0 ? o[0] : N;
        for (int64_t k = cnt; k < nsample; ++k) o[k] = first;
    }
    free(rnorm);
    return 0;
}

/* ------------------------------------------------------------------------------------------------
 * upsample(points, knn_idx, scale_ratio): the Markov state transition, R/modules/pointnet2_utils.py:13-50,
 * restated sparsely (the reference materialises a dense [B,S,N,C] tensor, :36-42).
 *   out[b,n,:] = (sum over s with n in knn_idx[b,s,:] of points[b,s,:]) / cnt[b,n]
 *   cnt[b,n]   = #{ s : n in knn_idx[b,s,:]  and  points[b,s,0] != 0 }  (count_nonzero of channel 0, :44),
 *                0 -> 1 (:45-46).  A repeated n inside one row s contributes once (scatter_ overwrites, :40).
 * points [B,S,C], idx [B,S,K] int64 with values in [0,N), out [B,N,C], cnt [B,N] fp32.
 * Summation runs over s ascending.
 * ---------------------------------------------------------------------------------------------- */
ORC_API int orc_transition_fwd_f32(const float* points, const int64_t* idx, float* out, float* cnt,
                                   int64_t B, int64_t S, int64_t K, int64_t C, int64_t N) {
    memset(out, 0, sizeof(float) * (size_t)(B * N * C));
    memset(cnt, 0, sizeof(float) * (size_t)(B * N));
#pragma omp parallel for schedule(dynamic, 1)
    for (int64_t b = 0; b < B; ++b) {
        for (int64_t s = 0; s < S; ++s) {
            const float* p = points + (b * S + s) * C;
            const int64_t* row = idx + (b * S + s) * K;
            for (int64_t k = 0; k < K; ++k) {
                int64_t n = row[k];
                if (n < 0 || n >= N) continue; /* caller validates; never read out of range */
                int dup = 0;
                for (int64_t j = 0; j < k; ++j) dup |= (row[j] == n);
                if (dup) continue;
                float* o = out + (b * N + n) * C;
                for (int64_t c = 0; c < C; ++c) o[c] = o[c] + p[c];
                if (p[0] != 0.0f) cnt[b * N + n] += 1.0f;
            }
        }
        for (int64_t n = 0; n < N; ++n) {
            float d = cnt[b * N + n];
            if (d == 0.0f) d = 1.0f;
            cnt[b * N + n] = d;
            float* o = out + (b * N + n) * C;
            for (int64_t c = 0; c < C; ++c) o[c] = o[c] / d;
        }
    }
    return 0;
}

/* Backward of the transition with respect to `points` (what autograd derives from :28-48; cnt is a
 * constant of the graph because count_nonzero is not differentiable):
 *   grad_points[b,s,:] = sum over distinct n in knn_idx[b,s,:] of grad_out[b,n,:] / cnt[b,n]. */
ORC_API int orc_transition_bwd_f32(const float* grad_out, const int64_t* idx, const float* cnt,
                                   float* grad_points, int64_t B, int64_t S, int64_t K, int64_t C,
                                   int64_t N) {
#pragma omp parallel for schedule(static)
    for (int64_t bs = 0; bs < B * S; ++bs) {
        int64_t b = bs / S;
        const int64_t* row = idx + bs * K;
        float* g = grad_points + bs * C;
        for (int64_t c = 0; c < C; ++c) g[c] = 0.0f;
        for (int64_t k = 0; k < K; ++k) {
            int64_t n = row[k];
            if (n < 0 || n >= N) continue;
            int dup = 0;
            for (int64_t j = 0; j < k; ++j) dup |= (row[j] == n);
            if (dup) continue;
            const float* go = grad_out + (b * N + n) * C;
            float d = cnt[b * N + n];
            for (int64_t c = 0; c < C; ++c) g[c] = g[c] + go[c] / d;
        }
    }
    return 0;
}

ORC_API int orc_version(void) { return 1; }
