/*
 * mpc_b200.h -- C ABI of libmpc_b200.so: the B200 (sm_100a) point-set hot path of
 * ssr0512/Markov-Process-Analysis-on-Point-Cloud.
 *
 * The reference has no FFI: its seam is the set of Python free functions and nn.Modules in
 * R/modules/pointnet2_utils.py and R/modules/repsurface_utils.py (R = Markov_Process_Analysis_on_Point_Cloud/).
 * Each entry point below names the reference function whose arithmetic it replaces (file:line).  The
 * Python host side (markov-process-analysis-on-point-cloud_b200/pointnet2_utils.py) mirrors the reference
 * signatures on top of this ABI; INTEGRATION.md shows the ctypes binding a maintainer would add.
 *
 * Conventions (every function):
 *   - all pointers are DEVICE pointers on the current device; tensors are dense row-major ("contiguous");
 *   - float data is IEEE binary32, indices are int64 (the reference's API dtype);
 *   - the caller owns every buffer (outputs and scratch included); nothing is allocated, nothing
 *     synchronises; work is enqueued on `stream` (a cudaStream_t passed as void*);
 *   - re-entrant and capturable into a CUDA graph; process-global state is limited to the tensor-map cache of
 *     the GEMMs (mutex-protected), the barrier words of the grid-wide FPS variant (N > 262 144: one such call in flight
 *     per device) and the debug knobs: safe for one process per GPU under torchrun;
 *   - reduction scratch buffers follow the "zero on entry, zero on exit" contract stated with the BatchNorm entry
 *     points: the caller allocates them zeroed once, no call memsets them;
 *   - returns 0 on success, a positive cudaError_t if a launch failed, or a negative MPC_ERR_* code if
 *     the arguments are rejected (nothing is launched in that case).
 *   - indices that are out of range are never dereferenced: gather-type kernels clamp them into range
 *     (the reference would raise from ATen; the Python wrappers can validate with MPC_CHECK_INDEX=1).
 */
#ifndef MPC_B200_H_
#define MPC_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MPC_OK 0
#define MPC_ERR_INVALID (-1)     /* bad sizes / null pointers */
#define MPC_ERR_UNSUPPORTED (-2) /* shape outside what the kernels cover (documented per function) */

typedef void* mpc_stream_t; /* cudaStream_t */

/* Library / build identification.  mpc_version() = 10000*major + 100*minor + patch. */
int mpc_version(void);
/* Compute capability the kernels were compiled for (1000 = sm_100a). */
int mpc_compiled_arch(void);

/* ---------------------------------------------------------------------------------------------------
 * Farthest point sampling.  Replaces farthest_point_sample, R/modules/pointnet2_utils.py:84-109
 * (== R/modules/repsurface_utils.py:150-172).
 *   xyz [B,N,C] f32, start [B] i64 (the reference draws it with torch.randint on the CPU generator, :96;
 *   the host passes it in), out [B,npoint] i64.
 * Bit-exact contract: running min distance starts at 1e10; dist = ((dx*dx + dy*dy) + dz*dz) with
 * separately rounded mul/add (no fma); update on strict <; next = argmax, lowest index on ties.
 * C == 3: one persistent CTA per cloud for N <= 8192, one thread-block cluster (<= 16 CTAs, DSMEM argmax
 * exchange) per cloud for N <= 262144, and one cooperative launch across the whole GPU (<= 148 CTAs x 8192 points in
 * registers, grid barrier per round through library-global barrier words: one such call in flight per device) for
 * N <= 1 212 416.  C != 3 (feature-space FPS): N * 4 bytes of shared memory, N <= 50000.
 * Otherwise MPC_ERR_UNSUPPORTED.
 * ------------------------------------------------------------------------------------------------- */
int mpc_fps_f32(const float* xyz, const int64_t* start, int64_t* out, int64_t B, int64_t N, int64_t C,
                int64_t npoint, mpc_stream_t stream);

/* ---------------------------------------------------------------------------------------------------
 * k nearest neighbours.  Replaces square_distance + knn_point, R/modules/pointnet2_utils.py:190-222, and
 * (with K = 3) the three_nn of PointNetFeaturePropagation, :899-901.
 *   ref [B,N,C] f32 (the reference's `xyz`), qry [B,S,C] f32 (`new_xyz`),
 *   dist_out [B,S,K] f32 ascending (may be NULL), idx_out [B,S,K] i64.
 * Bit-exact contract: d = ((-2*dot) + |q|^2) + |r|^2, dot = fma chain over c = 0..C-1 from 0, norms
 * sequential non-fused; ascending (d, index) order -- equal distances resolve to the lower index (torch.topk
 * leaves that order undefined).  1 <= K <= 32, K <= N, C <= 1024; else MPC_ERR_INVALID / UNSUPPORTED.
 * ------------------------------------------------------------------------------------------------- */
int mpc_knn_f32(const float* ref, const float* qry, float* dist_out, int64_t* idx_out, int64_t B,
                int64_t N, int64_t S, int64_t C, int64_t K, mpc_stream_t stream);

/* ---------------------------------------------------------------------------------------------------
 * k nearest neighbours in feature space with the distance GEMM on the tensor cores.  Same operands, same outputs
 * and the SAME bit-exact contract as mpc_knn_f32 (it replaces the same reference lines, R/modules/pointnet2_utils.py:
 * 190-222, whose square_distance is a torch.matmul): filter-and-refine.  The 128 x 128 dot products of a (query tile,
 * reference tile) pair are tcgen05 kind::tf32 MMAs with the 3xTF32 operand split (fp32-level products, fp32
 * accumulation in TMEM, operands by TMA); the epilogue thread of a query keeps its 16 best approximate distances;
 * every true neighbour provably lies within 2 eps of the K-th best approximate distance (eps = 2^-14 (|q|^2 +
 * max|r|^2)), so when the 16th entry is beyond that bound the candidates' EXACT distances (the contract's FP32 fma
 * chain) are evaluated and ranked by (distance, index); queries with more near-ties than the list holds (identical
 * feature vectors after a Markov transition) go through the exact brute-force kernel.  Results are therefore
 * bit-identical to mpc_knn_f32 for any input.
 *   C == 64, K == 8, N >= 16, operands 16-byte aligned; else MPC_ERR_UNSUPPORTED (callers fall back to mpc_knn_f32).
 *   workspace: mpc_knn_tc_workspace_bytes(...) bytes, 256-byte aligned, caller-owned, no initialisation needed.
 *   After the call its first words hold diagnostics: uint32[B] bit patterns of max|r|^2 per cloud, int32[B] queries
 *   per cloud that took the exact fallback, then one uint32 = bit pattern of the worst observed
 *   |approximate - exact| / (|q|^2 + max|r|^2) over all refined candidates (to be compared with eps' 2^-14).
 * ------------------------------------------------------------------------------------------------- */
int mpc_knn_tc_workspace_bytes(int64_t B, int64_t N, int64_t S, int64_t C, int64_t K, int64_t* bytes_out);
int mpc_knn_tc_f32(const float* ref, const float* qry, float* dist_out, int64_t* idx_out, void* workspace,
                   int64_t workspace_bytes, int64_t B, int64_t N, int64_t S, int64_t C, int64_t K, mpc_stream_t stream);

/* ---------------------------------------------------------------------------------------------------
 * k nearest neighbours of 3-D coordinates through a uniform grid.  Same operands (C = 3 implied), same outputs and the
 * SAME bit-exact contract as mpc_knn_f32 (it replaces the same reference lines, R/modules/pointnet2_utils.py:190-222,
 * for the coordinate searches of LocalMerge :448, Fuse :667-704 and three_nn :899-901): the reference set of each
 * cloud is counting-sorted into grid cells (about K/4 points per cell), a query evaluates only the cells its
 * neighbourhood can reach -- every candidate with the contract's fp32 expression, ranked by (distance, index) -- and
 * stops once every cell intersecting the ball of radius sqrt(d_K + eps) has been visited, eps = 64 * 2^-24 *
 * (|q|^2 + max|r|^2) bounding how far the expanded form can lie below the true squared distance.  Results are
 * therefore bit-identical to mpc_knn_f32 for any input; the work drops from B*S*N to about B*S*7K distance
 * evaluations.  K in {1, 3, 8, 9, 16, 32}, K <= N, N <= 2^30; else MPC_ERR_INVALID / UNSUPPORTED.
 *   workspace: mpc_knn3_grid_workspace_bytes(...) bytes, 256-byte aligned, caller-owned, no initialisation needed
 *   (cell table, counting-sort cursors, sorted copy of the reference set).
 * ------------------------------------------------------------------------------------------------- */
int mpc_knn3_grid_workspace_bytes(int64_t B, int64_t N, int64_t S, int64_t K, int64_t* bytes_out);
int mpc_knn3_grid_f32(const float* ref, const float* qry, float* dist_out, int64_t* idx_out, void* workspace,
                      int64_t workspace_bytes, int64_t B, int64_t N, int64_t S, int64_t K, mpc_stream_t stream);

/* ---------------------------------------------------------------------------------------------------
 * Ball query.  Replaces query_ball_point, R/modules/pointnet2_utils.py:112-134.
 *   xyz [B,N,C], new_xyz [B,S,C], idx_out [B,S,nsample] i64; r2 = radius^2 rounded to f32.
 * First `nsample` indices n (ascending) with NOT(d > r2), d as in mpc_knn_f32; padded with the first hit;
 * a query without any hit yields N in every slot (reference quirk, preserved).
 * ------------------------------------------------------------------------------------------------- */
int mpc_ball_query_f32(const float* xyz, const float* new_xyz, int64_t* idx_out, float r2, int64_t B,
                       int64_t N, int64_t S, int64_t C, int64_t nsample, mpc_stream_t stream);

/* ---------------------------------------------------------------------------------------------------
 * Gather / group.  Replaces index_points, R/modules/pointnet2_utils.py:64-81 (idx rank 2 "gathering" or
 * rank 3 "grouping": pass M = S or M = S*K).
 *   points [B,N,C] f32, idx [B,M] i64, out [B,M,C] f32:  out[b,m,:] = points[b, idx[b,m], :].
 * mpc_gather_bwd_f32 is what autograd derives (index_put_ accumulate): it zero-fills grad_points [B,N,C]
 * and adds grad_out rows into it with red.global (run-to-run fp32 summation order is not fixed).
 * mpc_gather_i64 is the same gather for an int64 payload with C = 1 (the composed FPS index chains of
 * Fuse.forward, :617-628).
 * ------------------------------------------------------------------------------------------------- */
int mpc_gather_f32(const float* points, const int64_t* idx, float* out, int64_t B, int64_t N, int64_t M,
                   int64_t C, mpc_stream_t stream);
int mpc_gather_bwd_f32(const float* grad_out, const int64_t* idx, float* grad_points, int64_t B, int64_t N,
                       int64_t M, int64_t C, mpc_stream_t stream);
/* bf16 payloads (half the HBM bytes): forward moves bf16 rows bit for bit, backward accumulates the scattered bf16
 * gradient rows in an F32 buffer (cleared by the call; C % 2 == 0), which the host casts back. */
int mpc_gather_bf16(const void* points, const int64_t* idx, void* out, int64_t B, int64_t N, int64_t M, int64_t C,
                    mpc_stream_t stream);
int mpc_gather_bwd_bf16(const void* grad_out, const int64_t* idx, float* grad_points, int64_t B, int64_t N, int64_t M,
                        int64_t C, mpc_stream_t stream);
/* Bytes of the reduction scratch (see the scratch contract) a layer of C channels needs: (2C + 2) doubles. */
int mpc_reduction_scratch_bytes(int64_t C);
int mpc_gather_i64(const int64_t* values, const int64_t* idx, int64_t* out, int64_t B, int64_t N, int64_t M,
                   mpc_stream_t stream);

/* ---------------------------------------------------------------------------------------------------
 * Markov state transition.  Replaces upsample, R/modules/pointnet2_utils.py:13-50 (dense [B,S,N,C]
 * scatter + sum + count_nonzero) with its sparse form D^-1 A^T X:
 *   out[b,n,:] = (sum over s with n in idx[b,s,:] of points[b,s,:]) / cnt[b,n]
 *   cnt[b,n]   = #{s : n in idx[b,s,:] and points[b,s,0] != 0}, 0 -> 1;  a repeated n inside one row counts once.
 *   points [B,S,C], idx [B,S,K] i64 (values in [0,N)), out [B,N,C], cnt [B,N] f32 (kept for backward).
 * Backward (cnt is a constant of the graph): grad_points[b,s,:] = sum_k grad_out[b,idx[b,s,k],:] / cnt[b,idx[b,s,k]].
 * K <= 32.
 * ------------------------------------------------------------------------------------------------- */
int mpc_transition_fwd_f32(const float* points, const int64_t* idx, float* out, float* cnt, int64_t B,
                           int64_t S, int64_t K, int64_t C, int64_t N, mpc_stream_t stream);
/* Same result through the gather form: per-cloud reverse-neighbour lists (CSR) are built from idx with S*K integer
 * atomics + a scan, then every output row is written once from its sorted source list: no float atomics, no
 * zero-fill / normalise passes, deterministic summation order (ascending s).  workspace: B*(2N+1) + B*S*K int32. */
int mpc_transition_fwd_csr_f32(const float* points, const int64_t* idx, float* out, float* cnt, int32_t* workspace,
                               int64_t B, int64_t S, int64_t K, int64_t C, int64_t N, mpc_stream_t stream);
/* The two halves of mpc_transition_fwd_csr_f32, for callers that apply the SAME neighbour table to several feature
 * tensors (Fuse.forward reuses the encoder's kNN indices in every stage, R/modules/pointnet2_utils.py:663-704): build
 * the reverse-neighbour lists once (depends on idx only), then apply them to any points [B,S,C]. */
int mpc_transition_csr_build(const int64_t* idx, int32_t* workspace, int64_t B, int64_t S, int64_t K, int64_t N,
                             mpc_stream_t stream);
int mpc_transition_csr_apply_f32(const float* points, const int32_t* workspace, float* out, float* cnt, int64_t B,
                                 int64_t S, int64_t K, int64_t C, int64_t N, mpc_stream_t stream);
int mpc_transition_bwd_f32(const float* grad_out, const int64_t* idx, const float* cnt, float* grad_points,
                           int64_t B, int64_t S, int64_t K, int64_t C, int64_t N, mpc_stream_t stream);

/* ---------------------------------------------------------------------------------------------------
 * three_interpolate.  Replaces R/modules/pointnet2_utils.py:903-906.
 *   dist [B,N,3] f32 and idx [B,N,3] i64 from mpc_knn_f32(K=3); points2 [B,S,C];
 *   weight_out [B,N,3] = (1/(d+1e-8)) / sum_i (1/(d_i+1e-8)); out [B,N,C] = sum_i w_i * points2[b,idx_i,:].
 * Backward: grad_points2 [B,S,C] zero-filled, += w_i * grad_out[b,n,:] at row idx_i.
 * ------------------------------------------------------------------------------------------------- */
int mpc_three_interpolate_fwd_f32(const float* points2, const float* dist, const int64_t* idx,
                                  float* weight_out, float* out, int64_t B, int64_t N, int64_t S, int64_t C,
                                  mpc_stream_t stream);
int mpc_three_interpolate_bwd_f32(const float* grad_out, const float* weight, const int64_t* idx,
                                  float* grad_points2, int64_t B, int64_t N, int64_t S, int64_t C,
                                  mpc_stream_t stream);

/* ---------------------------------------------------------------------------------------------------
 * Difference-wise attention core, feature branch.  Replaces the materialised [B,S,K,C] chain of
 * LocalTrans.forward with xyz=False, R/modules/pointnet2_utils.py:553-569 (gather k, gather v, q - k,
 * per-channel softmax over K of energy/sqrt(C), minus its own sum over K, times v, max over K).
 *   q   [B,S,C] rows with stride ldq floats; kf, vf [B,N,C] rows with stride ldkv floats (so q/k/v may be
 *   column slices of one fused projection buffer); idx [B,S,K] i64; ctx_out [B,S,C] dense.
 * Backward recomputes the softmax: given grad_ctx [B,S,C] it writes grad_q [B,S,C] (stride ldgq) and
 * ACCUMULATES (red.global) into grad_kf / grad_vf [B,N,C] (stride ldgkv), which the caller zero-fills.
 * grad_bias (optional, [3][C], ACCUMULATED) receives the column sums of grad_q / grad_k / grad_v, i.e. the bias
 * gradients of the q / k / v projections, so no separate reduction over the points is needed.
 * C % 4 == 0, C <= 1024, K <= 32 (K = 8 is the specialised fast path).
 * ------------------------------------------------------------------------------------------------- */
int mpc_attn_feat_fwd_f32(const float* q, int64_t ldq, const float* kf, const float* vf, int64_t ldkv,
                          const int64_t* idx, float* ctx_out, int64_t B, int64_t S, int64_t N, int64_t K,
                          int64_t C, mpc_stream_t stream);
/* bf16-I/O forward of the same core (inference path): q, kf, vf, ctx_out are bf16 (strides in elements, 8-byte
 * aligned rows), arithmetic in fp32, one rounding at the store. */
int mpc_attn_feat_fwd_bf16(const void* q, int64_t ldq, const void* kf, const void* vf, int64_t ldkv,
                           const int64_t* idx, void* ctx_out, int64_t B, int64_t S, int64_t N, int64_t K, int64_t C,
                           mpc_stream_t stream);
int mpc_attn_feat_bwd_f32(const float* grad_ctx, const float* q, int64_t ldq, const float* kf,
                          const float* vf, int64_t ldkv, const int64_t* idx, float* grad_q, int64_t ldgq,
                          float* grad_kf, float* grad_vf, int64_t ldgkv, float* grad_bias, int64_t B, int64_t S,
                          int64_t N, int64_t K, int64_t C, mpc_stream_t stream);

/* ---------------------------------------------------------------------------------------------------
 * Difference-wise attention core, coordinate branch.  Replaces LocalTrans.forward with xyz=True,
 * R/modules/pointnet2_utils.py:520-544: k, v = W (neighbour - centre) + b computed on the fly from the
 * Cin-channel differences (Cin = 3 in every live model, Cin <= 16 supported), q = Wq centre + bq, then the
 * same core as above.  Nothing of size [B,S,K,C] is ever written.
 *   feat [B,N,Cin]; center_idx [B,S] i64 or NULL (then S == N and centre s is point s); idx [B,S,K] i64;
 *   wq,wk,wv [C,Cin] (nn.Linear layout), bq,bk,bv [C]; ctx_out [B,S,C].
 * Backward: grad_w* [C,Cin] and grad_b* [C] are ACCUMULATED (caller zero-fills); grad_feat [B,N,Cin] may
 * be NULL (coordinates are data in the reference's training scripts) else it is ACCUMULATED too.
 * Optional fused residual projection (conv_res.linear of the same block, :515 -> :415): if wr [C,Cin] / br [C] are
 * given, res_out [B,S,C] = Wr centre + br (pre-BatchNorm) is produced by the same kernel; the backward then takes
 * grad_res [B,S,C] and accumulates grad_wr / grad_br (and the centre's share of grad_feat).  Pass NULL to skip.
 * ------------------------------------------------------------------------------------------------- */
int mpc_attn_xyz_fwd_f32(const float* feat, const int64_t* center_idx, const int64_t* idx, const float* wq,
                         const float* bq, const float* wk, const float* bk, const float* wv, const float* bv,
                         const float* wr, const float* br, float* ctx_out, float* res_out, int64_t B, int64_t S,
                         int64_t N, int64_t K, int64_t Cin, int64_t C, mpc_stream_t stream);
int mpc_attn_xyz_bwd_f32(const float* grad_ctx, const float* feat, const int64_t* center_idx,
                         const int64_t* idx, const float* wq, const float* bq, const float* wk,
                         const float* bk, const float* wv, const float* bv, const float* wr, const float* grad_res,
                         float* grad_wq, float* grad_bq, float* grad_wk, float* grad_bk, float* grad_wv,
                         float* grad_bv, float* grad_wr, float* grad_br, float* grad_feat, int64_t B, int64_t S,
                         int64_t N, int64_t K, int64_t Cin, int64_t C, mpc_stream_t stream);

/* ---------------------------------------------------------------------------------------------------
 * Shared-MLP block tail.  Replaces the BatchNorm1d-over-channels + LeakyReLU(0.2) of `Linear.forward`,
 * R/modules/pointnet2_utils.py:413-425, on the [M,C] view (M = all leading axes), without the two
 * permute().contiguous() copies of :420.
 * SCRATCH CONTRACT (all `scratch` / `sums` / `stat_scratch` arguments below): a caller-owned buffer of 2*C+2
 *   doubles that is ZERO ON ENTRY and is left ZERO ON EXIT by the call that consumes it (the last CTA to finish
 *   clears it), so a persistent per-layer buffer allocated zeroed once never needs a memset launch again.
 *   mpc_bn_stats_f32: stats[0:C] = mean, stats[C:2C] = biased variance of y over M rows (fp64 accumulation
 *     in `scratch`).  If non-NULL, running_mean / running_var [C] get
 *     nn.BatchNorm1d's update (x = (1-momentum)*x + momentum*stat, unbiased variance) and
 *     *num_batches_tracked (device int64) is incremented.
 *   mpc_bn_act_fwd_f32: out = lrelu(gamma * (y - mean) * rsqrt(var + eps) + beta, slope) [+ residual]; slope = 1 =>
 *     no act.  residual (optional, [M,C]) folds LocalTrans's `residual + ffn(context)` (:572) into the same pass
 *     (needs C % 4 == 0 and 1024 % C == 0).
 *   mpc_bn_act_bwd_f32: given grad_out and the saved pre-norm y, writes grad_y [M,C] and grad_gamma/beta [C]
 *     (train = 1: batch statistics take part in the gradient; train = 0: running statistics are constants).
 *     zero_buf (optional, zero_count floats, 16-byte aligned, count % 4 == 0) is cleared by the same launches on
 *     behalf of a later call in the stream (the split-reduction target of mpc_linear_wgrad_f32).
 *     ld_gout: row stride of grad_out in floats (C when contiguous); a wider stride -- grad_out is a column slice of
 *     the gradient of a concatenation -- needs C a power of two in [4, 1024] and 16-byte aligned rows.
 * ------------------------------------------------------------------------------------------------- */
int mpc_bn_stats_f32(const float* y, float* stats, float* running_mean, float* running_var,
                     int64_t* num_batches_tracked, float momentum, double* scratch, int64_t M, int64_t C,
                     mpc_stream_t stream);
/* Training-mode forward in one launch from the column sums the GEMM epilogue left in `sums` (sum, then sum of
 * squares, C doubles each): mean / variance are derived on the fly, out = lrelu(BN(y)), stats[0:C] / stats[C:2C]
 * receive mean / biased variance for the backward pass, running statistics are updated like nn.BatchNorm1d.
 * Requires C % 4 == 0, 1024 % C == 0; otherwise MPC_ERR_UNSUPPORTED (use mpc_bn_finalize_f32 + mpc_bn_act_fwd_f32).
 * Consumes `sums` (clears it, see the scratch contract). */
int mpc_bn_act_fwd_sums_f32(const float* y, double* sums, const float* gamma, const float* beta, float eps,
                            float slope, const float* residual, float* out, float* stats, float* running_mean,
                            float* running_var, int64_t* num_batches_tracked, float momentum, int64_t M, int64_t C,
                            mpc_stream_t stream);
/* out[c] = sum over the M rows of y[:,c] (fp64 accumulation; scratch contract above).  The bias gradient of a
 * projection that is not followed by BatchNorm (q / k / v of LocalTrans).  C/4 must be a power of two <= 256. */
int mpc_col_sum_f32(const float* y, float* out, double* scratch, int64_t M, int64_t C, mpc_stream_t stream);
int mpc_bn_act_fwd_f32(const float* y, const float* mean, const float* var, const float* gamma,
                       const float* beta, float eps, float slope, const float* residual, float* out, int64_t M,
                       int64_t C, mpc_stream_t stream);
int mpc_bn_act_bwd_f32(const float* grad_out, const float* y, const float* mean, const float* var,
                       const float* gamma, const float* beta, float eps, float slope, int train,
                       float* grad_y, float* grad_gamma, float* grad_beta, double* scratch, float* zero_buf,
                       int64_t zero_count, int64_t ld_gout, int64_t M, int64_t C, mpc_stream_t stream);

/* ---------------------------------------------------------------------------------------------------
 * Shared-MLP projection on the tensor cores.  Replaces the nn.Linear inside `Linear` and the q/k/v projections
 * of LocalTrans, R/modules/pointnet2_utils.py:408,415,489-491:   y[M,N] = x[M,K] w[N,K]^T + bias[N].
 *   x rows have stride ldx floats, w rows ldw, y rows ldy (so operands / results may be column slices).
 * tcgen05.mma kind::tf32 with the 3xTF32 operand split (hi/lo), fp32 accumulation in TMEM, TMA-fed: fp32-level
 * accuracy (the parity tolerance is rtol 1e-4).  Requires K % 32 == 0, ldx % 4 == ldw % 4 == 0, 16-byte aligned
 * x and w; other shapes return MPC_ERR_UNSUPPORTED and the host falls back to the library GEMM.
 * stat_scratch (optional, scratch contract above: zero on entry; this call only accumulates): the epilogue adds the
 * per-column sum and sum of squares of y over the M rows -- the BatchNorm batch statistics of the `Linear` block --
 * from the tile while it is still in shared memory, one atomic per column per CTA; mpc_bn_act_fwd_sums_f32 or
 * mpc_bn_finalize_f32 consume (and clear) them.
 * group_bias (optional, [ceil(M / rows_per_group), N] f32, N % 4 == 0): a second bias shared by
 * rows_per_group consecutive rows.  It carries the projection of input channels that are constant over a cloud's
 * points (the broadcast global-pool and label channels of the part-seg head, R/modules/pointnet2_utils.py:843-853):
 * y = x_a Wa^T + (g Wb^T)[cloud] + b equals the reference's Linear over cat(x_a, broadcast g) with 3.5x less work.
 * ------------------------------------------------------------------------------------------------- */
int mpc_linear_fwd_f32(const float* x, int64_t ldx, const float* w, int64_t ldw, const float* bias, float* y,
                       int64_t ldy, double* stat_scratch, const float* group_bias, int64_t rows_per_group, int64_t M,
                       int64_t K, int64_t N, mpc_stream_t stream);
/* stats[0:C] = mean, stats[C:2C] = biased variance from sums[0:C] = sum(y), sums[C:2C] = sum(y*y) over M rows;
 * running statistics / num_batches_tracked updated like nn.BatchNorm1d when non-NULL.  Consumes (clears) `sums`. */
int mpc_bn_finalize_f32(double* sums, float* stats, float* running_mean, float* running_var,
                        int64_t* num_batches_tracked, float momentum, int64_t M, int64_t C, mpc_stream_t stream);
/* Input gradient of the same layer:  gx[M,K] = gy[M,N] w[N,K].  The weight matrix is consumed as stored (MN-major
 * tensor-core operand), no transposed copy.  K % 32 == 0, ldg % 4 == 0.  zero_buf (optional, zero_count floats,
 * 16-byte aligned, count % 4 == 0) is cleared by the same launch on behalf of the mpc_linear_wgrad_f32 that follows. */
int mpc_linear_dgrad_f32(const float* gy, int64_t ldg, const float* w, int64_t ldw, float* gx, int64_t ldx,
                         int64_t M, int64_t K, int64_t N, float* zero_buf, int64_t zero_count, mpc_stream_t stream);

/* ---------------------------------------------------------------------------------------------------
 * fp32 inference form of the shared-MLP block (the reference's `Linear` in eval(), R/modules/pointnet2_utils.py:413-425
 * + the residual of :515,574,640-709): the 3xTF32 tcgen05 GEMM of mpc_linear_fwd_f32 with the BatchNorm (running
 * statistics) affine map, the LeakyReLU and the residual add in its epilogue -- no separate normalise pass:
 *     y[m,n] = LeakyReLU_slope(acc[m,n] * scale[n] + shift[n]) (+ residual[m,n])
 * Operand constraints as mpc_linear_fwd_f32; y rows 16-byte aligned (ldy % 4 == 0); residual [M,N] f32 (ldr) or NULL.
 * ------------------------------------------------------------------------------------------------- */
int mpc_linear_affine_act_f32(const float* x, int64_t ldx, const float* w, int64_t ldw, const float* scale,
                              const float* shift, float slope, const float* residual, int64_t ldr, float* y,
                              int64_t ldy, int64_t M, int64_t K, int64_t N, mpc_stream_t stream);

/* ---------------------------------------------------------------------------------------------------
 * bf16-I/O shared-MLP block for inference: the reference's whole `Linear` block in eval() -- nn.Linear -> BatchNorm1d
 * with running statistics -> LeakyReLU, R/modules/pointnet2_utils.py:413-425 -- plus the residual that LocalTrans /
 * Fuse add to it (:515,574,640-709), as ONE tcgen05 kernel (kind::f16, bf16 operands, fp32 accumulation in TMEM):
 *     out[m,n] = act(acc[m,n] * scale[n] + shift[n]) (+ residual[m,n]),  acc = x[m,:] . w[n,:],  act = LeakyReLU(slope)
 *   x [M,K] bf16 (row stride ldx elements), w [N,K] bf16 (ldw), scale / shift [N] f32 (NULL = 1 / 0; for a BatchNorm
 *   in eval(): scale = gamma / sqrt(var + eps), shift = beta + (bias - mean) * scale), slope = 1 for no activation,
 *   residual [M,N] bf16 (ldr) or NULL, out [M,N] bf16 or (out_is_f32 != 0) f32 with row stride ldo.
 *   K % 64 == 0, ldx / ldw % 8 == 0, x / w 16-byte aligned; else MPC_ERR_UNSUPPORTED (callers use the fp32 path).
 * mpc_f32_to_bf16: row-wise conversion x [rows, cols] f32 (ldx) -> y bf16 (ldy >= cols; pad columns are zeroed), used
 * for weights (once per layer) and for the first activation of a bf16 forward.
 * ------------------------------------------------------------------------------------------------- */
int mpc_linear_bf16(const void* x, int64_t ldx, const void* w, int64_t ldw, const float* scale, const float* shift,
                    float slope, const void* residual, int64_t ldr, void* out, int64_t ldo, int64_t out_is_f32,
                    int64_t M, int64_t K, int64_t N, mpc_stream_t stream);
int mpc_f32_to_bf16(const float* x, int64_t ldx, void* y, int64_t ldy, int64_t rows, int64_t cols, mpc_stream_t stream);
/* Weight gradient of the same layer (what autograd derives for nn.Linear):  gw[N,K] = gy[M,N]^T x[M,K].
 * Same 3xTF32 tcgen05 pipeline with MN-major operand descriptors (no transposed copies), the reduction over the M
 * points split across the SMs and combined with TMA reduce-add into gw.  gw_is_zero = 1: an earlier call in the
 * stream already cleared gw (zero_buf of mpc_bn_act_bwd_f32 / mpc_linear_dgrad_f32); 0: this call clears it with a
 * memset first.  K % 32 == 0. */
int mpc_linear_wgrad_f32(const float* gy, int64_t ldg, const float* x, int64_t ldx, float* gw, int64_t ldw,
                         int64_t M, int64_t K, int64_t N, int64_t gw_is_zero, mpc_stream_t stream);

/* ---------------------------------------------------------------------------------------------------
 * Label-smoothed cross entropy of the part-seg head.  Replaces get_loss.forward,
 * R/models/repsurf/pointnet2_part_seg_msg.py:159-180 (one_hot smoothing + log_softmax + sum + mean; ~10 launches).
 *   pred [M,C] f32 with row stride ld floats, target [M] i64, eps = smoothing (0.1 in the reference);
 *   fwd: lse [M] f32 (log-sum-exp per row, kept for backward), loss [1] f32 = mean over rows; scratch: 2 doubles under
 *   the scratch contract below (zero on entry, zero on exit).
 *   bwd: grad_pred [M,C] (row stride ldg) = grad_loss[0] / M * (softmax(pred) - smoothed target).
 * ------------------------------------------------------------------------------------------------- */
int mpc_smooth_ce_fwd_f32(const float* pred, int64_t ld, const int64_t* target, float eps, float* lse, float* loss,
                          double* scratch, int64_t M, int64_t C, mpc_stream_t stream);
int mpc_smooth_ce_bwd_f32(const float* pred, int64_t ld, const int64_t* target, float eps, const float* lse,
                          const float* grad_loss, float* grad_pred, int64_t ldg, int64_t M, int64_t C,
                          mpc_stream_t stream);

/* ---------------------------------------------------------------------------------------------------
 * Umbrella surface features.  Replaces group_by_umbrella + cal_normal(is_group=True) +
 * cal_center + xyz2sphere + cal_const + check_nan_umb as composed by UmbrellaSurfaceConstructor.forward,
 * R/modules/pointnet2_utils.py:310-378 (helpers in R/modules/recons_utils.py, R/modules/polar_utils.py:10-31).
 *   xyz [B,N,3] f32; idx [B,N,ldk] i64 = the k nearest neighbours of every point in its own cloud, column 0 the point
 *   itself (mpc_knn_f32 with ref = qry), columns 1..k-1 used; sign [B] f32 (+1/-1, the per-cloud random_inv draw of
 *   recons_utils.py:50) or NULL; out [B,N,k-1,C] f32 with C = 10 (centroid 3, spherical coordinates of the centroid
 *   3, unit normal 3, plane constant 1) or C = 9 (no constant).
 * Neighbours are sorted by azimuth (stable); triangle g = (point, s_g, s_{g+1 mod k-1}); triangles with a NaN normal
 * take normal / centroid / constant of the first valid triangle of the same point.  3 <= k <= 13 or k = 16.
 * ------------------------------------------------------------------------------------------------- */
int mpc_umbrella_features_f32(const float* xyz, const int64_t* idx, int64_t ldk, const float* sign, float* out,
                              int64_t B, int64_t N, int64_t k, int64_t C, mpc_stream_t stream);

/* Debug facility (not part of the data path): when set to a device buffer of 1024 int64, CTA 0 of the tensor-core
 * kernels records a clock64() timeline per warp role (producer / splitter / MMA / epilogue: 256 slots each).
 * Pass NULL to switch it off (the default).  Process-global. */
int mpc_debug_trace_buffer(void* device_buffer);
/* Debug facility: kernel-variant and launch-geometry knobs (0 restores the built-in default).  Every value gives the
 * same results except the two timing experiments of id 4 marked below.
 * id 0: elementwise CTAs per SM of the streaming BatchNorm kernels, 1: their column-reduction CTAs per SM, 2: float4
 *       per thread their elementwise grid is sized for (scratch/bench_bn.py picked the defaults);
 * id 3: non-zero forces the grid-wide FPS variant for every cloud above 8192 points (parity tests of that variant at
 *       sizes the CPU oracle finishes);
 * id 4: feature-space kNN: 1 = never the register-tiled kernel; >= 100 = mpc_knn_tc_f32 timing experiments that SKIP
 *       WORK AND RETURN WRONG RESULTS (101: no selection, 102: no MMAs; scratch/knn_tc_time.py only);
 * id 5: KB of shared-memory padding of the register-tiled kNN kernel (occupancy experiments) / CTAs per SM cap of the
 *       grouped transition gather;
 * id 6: smallest N that takes the bucket-pruned FPS variant (default 8193; < 0: never);
 * id 7: non-zero forces the ungrouped transition gather kernel.
 * Process-global. */
int mpc_debug_set_knob(int id, int64_t value);

#ifdef __cplusplus
}
#endif
#endif /* MPC_B200_H_ */
