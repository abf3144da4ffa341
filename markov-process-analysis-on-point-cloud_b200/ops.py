"""Point-set ops of the Markov encoder/decoder on B200: the reference's free functions, same names and
argument order, backed by the hand-written sm_100a kernels in csrc/ through the C ABI (include/mpc_b200.h).

R = Markov_Process_Analysis_on_Point_Cloud/ in the reference tree.  Mirrors R/modules/pointnet2_utils.py:13-222
(and the byte-identical copies in R/modules/repsurface_utils.py:129-204).  The upstream `cuda=` / `is_group=`
keyword arguments that reference call sites still pass are accepted and ignored.

Every op runs on the caller's current CUDA stream, allocates only its outputs through torch, never
synchronises, and raises if given CPU tensors: there is no CPU fallback.
"""
import contextlib
import ctypes
import os

import numpy as np
import torch

from . import _lib
from ._lib import call, on_tensor_device, ptr, require_cuda

_i64 = ctypes.c_int64
_CHECK_INDEX = os.environ.get("MPC_CHECK_INDEX", "0") == "1"
_KNN_LIST_LENGTHS = (1, 3, 8, 9, 16, 32)
_TRANSITION_IMPL = os.environ.get("MPC_TRANSITION", "csr")  # "atomic": red.global scatter form
# Feature-space searches (C = 64, k = 8) run their distance GEMM on the tensor cores (mpc_knn_tc_f32: filter on
# tcgen05, exact FP32 refinement, bit-identical results); MPC_KNN_TC=0 forces the FP32-SIMT brute-force kernels.
_KNN_TC = os.environ.get("MPC_KNN_TC", "1") == "1"
_KNN_TC_MIN_N = int(os.environ.get("MPC_KNN_TC_MIN_N", "512"))
# Coordinate searches (C = 3) in large clouds go through the uniform-grid kernel (mpc_knn3_grid_f32: bit-identical
# results, ~7K instead of N distance evaluations per query); MPC_KNN_GRID=0 forces brute force.
_KNN_GRID = os.environ.get("MPC_KNN_GRID", "1") == "1"
_KNN_GRID_MIN_N = int(os.environ.get("MPC_KNN_GRID_MIN_N", "2048"))
knn_tc_debug = None  # set to a list to collect (workspace, B) of every tensor-core search (tests read the diagnostics)


def knn_tc_available():
    return _KNN_TC


def set_knn_tc(on):
    global _KNN_TC
    _KNN_TC = bool(on)


def set_knn_grid(on, min_n=None):
    global _KNN_GRID, _KNN_GRID_MIN_N
    _KNN_GRID = bool(on)
    if min_n is not None:
        _KNN_GRID_MIN_N = int(min_n)


def _f32c(t):
    if t.dtype != torch.float32:
        raise TypeError("mpc_b200 kernels compute in float32, got %s" % t.dtype)
    return t.contiguous()


def _i64c(t):
    return t.to(torch.int64).contiguous()


# ------------------------------------------------------------------------------------------------------
# index injection / recording (test and debugging aid: feed another implementation's FPS / kNN indices
# into this path so that float outputs can be compared without neighbour flips -- SURVEY.md 4 / H1)
# ------------------------------------------------------------------------------------------------------
class _IndexTape:
    def __init__(self):
        self.inject = None  # list of index tensors consumed in call order
        self.record = None  # list that receives (kind, idx) in call order
        self.fps_starts = None  # list of [B] int64 start tensors consumed in call order
        self.audit = None  # list receiving (kind, own idx, own dist, injected idx, reference set, queries) per search


_tape = _IndexTape()


@contextlib.contextmanager
def index_tape(inject=None, record=None, fps_starts=None, audit=None):
    """inject: iterable of index tensors returned (in call order) by farthest_point_sample / knn_point
    instead of computing them; record: list receiving ("fps"|"knn"|"knnf", idx); fps_starts: iterable of [B] start
    indices used instead of drawing them from the CPU generator; audit: list receiving, for every neighbour search
    (with or without injection), (kind, own idx, own dist, injected idx or None, reference set, queries) -- the search
    itself always runs, so a test can compare this path's neighbours with the injected ones on the very operands this
    path saw ("teacher forcing": the tie audit of SURVEY.md 8c)."""
    old = (_tape.inject, _tape.record, _tape.fps_starts, _tape.audit)
    _tape.inject = list(inject) if inject is not None else None
    _tape.record = record
    _tape.fps_starts = list(fps_starts) if fps_starts is not None else None
    _tape.audit = audit
    try:
        yield
    finally:
        _tape.inject, _tape.record, _tape.fps_starts, _tape.audit = old


def _taped(kind, device):
    if _tape.inject is not None:
        idx = _tape.inject.pop(0)
        return idx.to(device=device, dtype=torch.int64).contiguous()
    return None


def _record(kind, idx):
    if _tape.record is not None:
        _tape.record.append((kind, idx))


# ------------------------------------------------------------------------------------------------------
# sampling / neighbour search (integer outputs, no gradient)
# ------------------------------------------------------------------------------------------------------
def draw_fps_start(B, N, device):
    """The reference draws the start index on the CPU default generator and moves it to the device
    (R/modules/pointnet2_utils.py:96); consuming the generator identically keeps runs comparable."""
    if _tape.fps_starts is not None:
        return _tape.fps_starts.pop(0).to(device=device, dtype=torch.int64)
    return torch.randint(0, N, (B,), dtype=torch.long).to(device)


# Sampling one batch ahead.  The FPS chain of a forward (xyz -> N/2 -> N/4 ...) depends on nothing but the input
# coordinates and is a pure latency chain (22 500 sequential rounds on a 24 000-point block: ~17 ms on 16 SMs), so a
# training / inference loop can compute it for batch i+1 while batch i is in flight: sampling_pyramid() runs the chain,
# sampled_ahead() feeds its result to the next forward, whose sampling steps then consume the indices in call order
# instead of launching (and instead of drawing start indices: the ahead computation drew them).
_fps_ahead = None


@contextlib.contextmanager
def sampled_ahead(indices):
    """indices: the FPS index tensors of every sampling step of ONE forward, in call order (sampling_pyramid)."""
    global _fps_ahead
    old = _fps_ahead
    _fps_ahead = list(indices)
    try:
        yield
    finally:
        _fps_ahead = old


@on_tensor_device
@torch.no_grad()
def sampling_pyramid(xyz, npoints, starts=None, levels=None, base=None):
    """The sampling chain of a forward on coordinates xyz [B,N,3]: FPS to npoints[0], gather, FPS to npoints[1], ...
    -> list of int64 [B,npoint] index tensors (level i indexes level i-1's points).  `starts`: one [B] start-index
    tensor per level (default: drawn like the reference draws them, R/modules/pointnet2_utils.py:96).
    A chain can be computed in segments (a deeper software pipeline: different batches' segments run side by side):
    levels = (first, last) restricts the call to levels first..last-1, `base` = the coordinates level `first` samples
    from (the third return value of the previous segment); returns (indices of these levels, next base)."""
    require_cuda(xyz if base is None else base)
    first, last = levels if levels is not None else (0, len(npoints))
    out, cur = [], (xyz if base is None else base)
    for i in range(first, last):
        B, N, _ = cur.shape
        start = starts[i] if starts is not None else draw_fps_start(B, N, cur.device)
        idx = _fps_launch(cur, npoints[i], start)
        out.append(idx)
        if i + 1 < len(npoints):
            cur = index_points(cur, idx)
    return out if levels is None else (out, cur)


@on_tensor_device
@torch.no_grad()
def farthest_point_sample(xyz, npoint, cuda=False, start=None):
    """R/modules/pointnet2_utils.py:84-109.  xyz [B,N,C] -> int64 [B,npoint]; bit-exact with the reference
    for C = 3 (same start index, same arithmetic order, lowest index on ties)."""
    require_cuda(xyz)
    B, N, C = xyz.shape
    if _fps_ahead is not None and start is None and _tape.inject is None:
        out = _fps_ahead.pop(0)
        if out.shape != (B, npoint):
            raise ValueError("sampled_ahead: expected FPS indices of shape %s, got %s" % ((B, npoint), tuple(out.shape)))
        _record("fps", out)
        return out
    if start is None:
        start = draw_fps_start(B, N, xyz.device)  # drawn even when injecting: RNG consumption stays identical
    taped = _taped("fps", xyz.device)
    if taped is not None:
        _record("fps", taped)
        return taped
    out = _fps_launch(xyz, npoint, start)
    _record("fps", out)
    return out


def _fps_launch(xyz, npoint, start):
    B, N, C = xyz.shape
    xyz = _f32c(xyz.detach())
    start = _i64c(start.to(xyz.device))
    if start.shape != (B,):
        raise ValueError("farthest_point_sample: start must hold one index per cloud, got shape %s" % (tuple(start.shape),))
    if _CHECK_INDEX and B:  # (the kernel clamps an out-of-range start into [0, N); the reference would raise)
        _check_index(start, N, "farthest_point_sample start")
    out = torch.empty(B, npoint, dtype=torch.int64, device=xyz.device)
    call("mpc_fps_f32", ptr(xyz), ptr(start), ptr(out), _i64(B), _i64(N), _i64(C), _i64(npoint),
         algo_bytes=B * (N * C * 4 + npoint * 8))
    return out


def _fps_compute(xyz, npoint):
    """Prefetch path: draws the start index like farthest_point_sample, no tape bookkeeping."""
    B, N, _ = xyz.shape
    if _fps_ahead is not None:
        out = _fps_ahead.pop(0)
        if out.shape != (B, npoint):
            raise ValueError("sampled_ahead: expected FPS indices of shape %s, got %s" % ((B, npoint), tuple(out.shape)))
        return out
    return _fps_launch(xyz, npoint, draw_fps_start(B, N, xyz.device))


@on_tensor_device
def sample(nsample, feature, cuda=False):
    """Data-side FPS of the reference's training scripts (`sample(args.num_point, points, cuda=...)`,
    R/tool/train_cls_scanobjectnn.py:244 -- called there, defined nowhere in the shipped tree; upstream RepSurf
    semantics): feature [B, C, N] with xyz in channels 0..2 -> [B, C, nsample], the columns picked by farthest
    point sampling on the coordinates."""
    require_cuda(feature)
    xyz = feature[:, :3, :].permute(0, 2, 1).contiguous()
    idx = farthest_point_sample(xyz, nsample)
    return index_points(feature.permute(0, 2, 1).contiguous(), idx).permute(0, 2, 1).contiguous()


def _list_length(k, N):
    """The kernels keep sorted neighbour lists of these compiled lengths; a longer list is computed and sliced (when
    the compiled length exceeds the number of reference points, _knn_compute pads the reference set)."""
    if k > _KNN_LIST_LENGTHS[-1]:
        raise ValueError("knn_point supports nsample <= %d, got %d" % (_KNN_LIST_LENGTHS[-1], k))
    for L in _KNN_LIST_LENGTHS:
        if k <= L:
            return L


@on_tensor_device
@torch.no_grad()
def knn_point(nsample, xyz, new_xyz):
    """R/modules/pointnet2_utils.py:211-222.  NOTE the reference's argument order: (k, reference set
    [B,N,C], queries [B,S,C]) -> (dist [B,S,k] ascending, idx int64 [B,S,k]).  Equal distances resolve to
    the lower index (torch.topk leaves it undefined)."""
    require_cuda(xyz, new_xyz)
    B, N, C = xyz.shape
    S = new_xyz.shape[1]
    if nsample > N:
        raise RuntimeError("selected index k out of range")  # what torch.topk raises in the reference
    taped = _taped("knn", xyz.device)
    # inside one forward the same coordinate search can be asked for twice (la0 and la1_up both search the full
    # cloud in itself) or be prefetched: the geometry scope remembers coordinate-space results by operand identity
    dist, idx = _knn_compute(nsample, xyz, new_xyz)
    if _tape.audit is not None:
        _tape.audit.append(("knn" if C == 3 else "knnf", idx, dist, taped, xyz, new_xyz))
    if taped is not None:
        idx = taped
    _record("knn" if C == 3 else "knnf", idx)  # "knnf": feature-space search (tie-prone, see tests)
    return dist, idx


@torch.no_grad()
def _knn_compute(nsample, xyz, new_xyz):
    """The search itself (no tape bookkeeping); coordinate-space results are remembered by the geometry scope."""
    B, N, C = xyz.shape
    S = new_xyz.shape[1]
    cache = _geo.cache if (_geo is not None and C == 3) else None
    key = (xyz.data_ptr(), new_xyz.data_ptr(), tuple(xyz.shape), tuple(new_xyz.shape), xyz.stride(), new_xyz.stride(),
           nsample)
    if cache is not None and key in cache:
        return cache[key][:2]
    if xyz.dtype == torch.bfloat16 or new_xyz.dtype == torch.bfloat16:  # bf16 inference: indices from fp32 arithmetic
        same = new_xyz is xyz
        xyz = xyz.float()
        new_xyz = xyz if same else new_xyz.float()
    xyz, new_xyz = _f32c(xyz.detach()), _f32c(new_xyz.detach())
    L = _list_length(nsample, N)
    if L > N:
        # k <= N < L (a handful of reference points): pad the reference set with L - N points far outside any cloud;
        # their distances exceed every real one, so they never enter the first k <= N entries that are returned
        pad = torch.full((B, L - N, C), 1e18, dtype=torch.float32, device=xyz.device)
        xyz_run, N_run = torch.cat((xyz, pad), 1), L
    else:
        xyz_run, N_run = xyz, N
    dist = torch.empty(B, S, L, dtype=torch.float32, device=xyz.device)
    idx = torch.empty(B, S, L, dtype=torch.int64, device=xyz.device)
    if (_KNN_TC and C == 64 and L == 8 and N_run >= _KNN_TC_MIN_N and xyz_run.data_ptr() % 16 == 0
            and new_xyz.data_ptr() % 16 == 0 and B * max(N_run, S) < 2 ** 31):
        need = ctypes.c_int64(0)
        _lib.load().mpc_knn_tc_workspace_bytes(_i64(B), _i64(N_run), _i64(S), _i64(C), _i64(L), ctypes.byref(need))
        ws = torch.empty(need.value, dtype=torch.uint8, device=xyz.device)
        call("mpc_knn_tc_f32", ptr(xyz_run), ptr(new_xyz), ptr(dist), ptr(idx), ptr(ws), _i64(need.value), _i64(B),
             _i64(N_run), _i64(S), _i64(C), _i64(L), algo_bytes=B * ((N + S) * C * 4 + S * L * 12))
        if knn_tc_debug is not None:
            knn_tc_debug.append((ws, B, N_run, S))
        if cache is not None:
            cache[key] = (dist, idx, xyz, new_xyz)
        return dist, idx
    if _KNN_GRID and C == 3 and N_run >= _KNN_GRID_MIN_N:
        need = ctypes.c_int64(0)
        _lib.load().mpc_knn3_grid_workspace_bytes(_i64(B), _i64(N_run), _i64(S), _i64(L), ctypes.byref(need))
        ws = torch.empty(need.value, dtype=torch.uint8, device=xyz.device)
        call("mpc_knn3_grid_f32", ptr(xyz_run), ptr(new_xyz), ptr(dist), ptr(idx), ptr(ws), _i64(need.value), _i64(B),
             _i64(N_run), _i64(S), _i64(L), algo_bytes=B * ((N + S) * C * 4 + S * L * 12))
    else:
        call("mpc_knn_f32", ptr(xyz_run), ptr(new_xyz), ptr(dist), ptr(idx), _i64(B), _i64(N_run), _i64(S), _i64(C),
             _i64(L), algo_bytes=B * ((N + S) * C * 4 + S * L * 12))
    if L != nsample:
        dist, idx = dist[:, :, :nsample].contiguous(), idx[:, :, :nsample].contiguous()
    if cache is not None:
        cache[key] = (dist, idx, xyz, new_xyz)  # the operands stay alive with the entry: their addresses are the key
    return dist, idx


def query_knn_point(k, xyz, new_xyz, cuda=False):
    """Upstream RepSurf name still used by reference call sites (R/modules/repsurface_utils.py:111,
    R/modules/recons_utils.py:19): indices only."""
    return knn_point(k, xyz, new_xyz)[1]


def square_distance(src, dst):
    """R/modules/pointnet2_utils.py:190-209: dense [B,N,M] expanded-form distances.  Provided for API
    completeness (the kernels never materialise this matrix); computed as a full-length sorted neighbour
    list scattered back would be wasteful, so this is the one function kept as a plain ATen expression."""
    require_cuda(src, dst)
    B, N, _ = src.shape
    M = dst.shape[1]
    dist = -2 * torch.matmul(src, dst.permute(0, 2, 1))
    dist += torch.sum(src ** 2, -1).view(B, N, 1)
    dist += torch.sum(dst ** 2, -1).view(B, 1, M)
    return dist


@on_tensor_device
@torch.no_grad()
def query_ball_point(radius, nsample, xyz, new_xyz, cuda=False):
    """R/modules/pointnet2_utils.py:112-134 -> int64 [B,S,nsample] (ascending index order, padded with the
    first hit, N where a query has no hit)."""
    require_cuda(xyz, new_xyz)
    B, N, C = xyz.shape
    S = new_xyz.shape[1]
    xyz, new_xyz = _f32c(xyz.detach()), _f32c(new_xyz.detach())
    out = torch.empty(B, S, nsample, dtype=torch.int64, device=xyz.device)
    r2 = ctypes.c_float(float(np.float32(radius ** 2)))
    call("mpc_ball_query_f32", ptr(xyz), ptr(new_xyz), ptr(out), r2, _i64(B), _i64(N), _i64(S), _i64(C),
         _i64(nsample))
    return out


# ------------------------------------------------------------------------------------------------------
# bf16 inference path (activations travel as bf16, arithmetic in fp32: mpc_linear_bf16 / mpc_attn_feat_fwd_bf16 /
# mpc_gather_bf16).  Indices still come from fp32 arithmetic: coordinate searches see the fp32 cloud, feature-space
# searches see the bf16 features widened to fp32.
# ------------------------------------------------------------------------------------------------------
_INFER_BF16 = False
# eval() forward without autograd: Linear -> BatchNorm(running statistics) -> LeakyReLU (+ residual) as ONE GEMM launch
# (affine map + activation + residual in the epilogue); MPC_FUSE_EVAL=0 keeps the separate normalise kernel (A/B runs)
_FUSE_EVAL = os.environ.get("MPC_FUSE_EVAL", "1") == "1"


def set_fuse_eval(on):
    global _FUSE_EVAL
    _FUSE_EVAL = bool(on)



@contextlib.contextmanager
def bf16_inference():
    """Inside this context the drop-in modules, when in eval(), run their bf16-I/O kernels: every `Linear` block
    (nn.Linear -> BatchNorm1d(running statistics) -> LeakyReLU [+ residual]) is ONE tcgen05 GEMM with the affine map
    and the activation in its epilogue, the attention core and the gathers move bf16 rows.  No autograd."""
    global _INFER_BF16
    old = _INFER_BF16
    _INFER_BF16 = True
    try:
        with torch.no_grad():
            yield
    finally:
        _INFER_BF16 = old


def bf16_active():
    return _INFER_BF16


def to_bf16_rows(x2d):
    """[M,K] f32 -> bf16 (one pass, mpc_f32_to_bf16); bf16 input is returned as is."""
    if x2d.dtype == torch.bfloat16:
        return x2d if x2d.stride(1) == 1 else x2d.contiguous()
    x2d = _f32c(x2d)
    M, K = x2d.shape
    out = torch.empty(M, K, dtype=torch.bfloat16, device=x2d.device)
    call("mpc_f32_to_bf16", ptr(x2d), _i64(K), ptr(out), _i64(K), _i64(M), _i64(K), algo_bytes=M * K * 6)
    return out


def _bf16_weight(*ws):
    """bf16 copy of a weight matrix (or of several stacked along the output dimension), rounded once and cached on
    the first parameter, keyed by the parameters' versions."""
    key = tuple((w.data_ptr(), w._version) for w in ws)
    hit = getattr(ws[0], "_mpc_bf16", None)
    if hit is not None and hit[0] == key:
        return hit[1]
    w = ws[0].detach() if len(ws) == 1 else torch.cat([w.detach() for w in ws], 0)
    w16 = to_bf16_rows(w.contiguous())
    ws[0]._mpc_bf16 = (key, w16)
    return w16


def _bn_affine(bias, bn):
    """BatchNorm1d in eval() after a Linear bias is the per-channel affine map y * scale + shift."""
    key = (bn.weight._version, bn.bias._version, bn.running_mean._version, bn.running_var._version,
           bias._version if bias is not None else -1, bn.weight.data_ptr())
    hit = getattr(bn, "_mpc_affine", None)
    if hit is not None and hit[0] == key:
        return hit[1], hit[2]
    scale = (bn.weight.detach() * torch.rsqrt(bn.running_var + bn.eps)).float().contiguous()
    b = bias.detach() if bias is not None else 0.0
    shift = (bn.bias.detach() + (b - bn.running_mean) * scale).float().contiguous()
    bn._mpc_affine = (key, scale, shift)
    return scale, shift


def linear_bf16(x2d, w16, scale, shift, slope=1.0, residual2d=None, out_f32=False):
    """out = LeakyReLU_slope(x2d @ w16^T * scale + shift) (+ residual2d): mpc_linear_bf16.  x2d [M,K] bf16 with
    K % 64 == 0; returns bf16 [M,N] (rows padded to 16 bytes when N % 8 != 0) or f32."""
    M, K = x2d.shape
    N = w16.shape[0]
    if out_f32:
        out = torch.empty(M, N, dtype=torch.float32, device=x2d.device)
        ldo = N
    else:
        Np = (N + 7) & ~7
        out = torch.empty(M, Np, dtype=torch.bfloat16, device=x2d.device)
        ldo = Np
        if Np != N:
            out = out[:, :N]
    if residual2d is not None:
        residual2d = to_bf16_rows(residual2d)
    call("mpc_linear_bf16", ptr(x2d), _i64(x2d.stride(0)), ptr(w16), _i64(w16.stride(0)), ptr(scale), ptr(shift),
         ctypes.c_float(slope), ptr(residual2d), _i64(residual2d.stride(0) if residual2d is not None else 0), ptr(out),
         _i64(ldo), _i64(1 if out_f32 else 0), _i64(M), _i64(K), _i64(N),
         algo_bytes=(M * K + N * K) * 2 + M * N * (4 if out_f32 else 2) * (2 if residual2d is not None else 1))
    return out


def _bf16_gemm_ok(x2d):
    return (x2d.is_cuda and x2d.shape[1] % 64 == 0 and x2d.shape[0] > 0
            and (x2d.dtype == torch.float32 or (x2d.stride(0) % 8 == 0 and x2d.data_ptr() % 16 == 0)))


def feat_attention_bf16(center, features, idx, wq, bq, wk, bk, wv, bv):
    """Feature branch of LocalTrans for the bf16 inference path: q projection, fused k|v projection (two
    mpc_linear_bf16 launches, biases in the epilogue) and the bf16 attention core."""
    B, S, Cin = center.shape
    N = features.shape[1]
    C = wq.shape[0]
    K = idx.shape[2]
    c2d, f2d = to_bf16_rows(center.reshape(-1, Cin)), to_bf16_rows(features.reshape(-1, Cin))
    q = linear_bf16(c2d, _bf16_weight(wq), None, bq.detach())
    kv = linear_bf16(f2d, _bf16_weight(wk, wv), None, _cat_cached(bk, bv))
    out = torch.empty(B, S, C, dtype=torch.bfloat16, device=center.device)
    call("mpc_attn_feat_fwd_bf16", ptr(q), _i64(q.stride(0)), ptr(kv), ctypes.c_void_p(kv.data_ptr() + 2 * C),
         _i64(kv.stride(0)), ptr(_i64c(idx)), ptr(out), _i64(B), _i64(S), _i64(N), _i64(K), _i64(C),
         algo_bytes=B * ((2 * N * C + 2 * S * C) * 2 + S * K * 8))
    return out


def _cat_cached(a, b):
    key = (a._version, b._version, a.data_ptr(), b.data_ptr())
    hit = getattr(a, "_mpc_cat", None)
    if hit is not None and hit[0] == key:
        return hit[1]
    c = torch.cat((a.detach(), b.detach()), 0).float().contiguous()
    a._mpc_cat = (key, c)
    return c


class SmoothCE(torch.autograd.Function):
    """Label-smoothed cross entropy of the part-seg head (R/models/repsurf/pointnet2_part_seg_msg.py:159-180) in one
    forward and one backward kernel; pred [M,C] may have padded rows."""

    @staticmethod
    def forward(ctx, pred, target, eps):
        M, C = pred.shape
        if pred.stride(1) != 1:
            pred = pred.contiguous()
        target = _i64c(target.reshape(-1))
        lse = torch.empty(M, dtype=torch.float32, device=pred.device)
        loss = torch.empty((), dtype=torch.float32, device=pred.device)
        scratch = _ce_scratch(pred.device)
        call("mpc_smooth_ce_fwd_f32", ptr(pred), _i64(pred.stride(0)), ptr(target), ctypes.c_float(eps), ptr(lse),
             ptr(loss), ptr(scratch), _i64(M), _i64(C), algo_bytes=M * (C * 4 + 12))
        ctx.save_for_backward(pred, target, lse)
        ctx.eps = eps
        return loss

    @staticmethod
    def backward(ctx, grad_loss):
        pred, target, lse = ctx.saved_tensors
        M, C = pred.shape
        Cp = (C + 3) & ~3
        g = torch.empty(M, Cp, dtype=torch.float32, device=pred.device)[:, :C]  # padded rows: TMA-ready for the head GEMM
        grad_loss = _f32c(grad_loss.reshape(1))
        call("mpc_smooth_ce_bwd_f32", ptr(pred), _i64(pred.stride(0)), ptr(target), ctypes.c_float(ctx.eps), ptr(lse),
             ptr(grad_loss), ptr(g), _i64(g.stride(0)), _i64(M), _i64(C), algo_bytes=2 * M * C * 4)
        return g, None, None


_ce_scratch_pool = {}


def _ce_scratch(device):
    buf = _ce_scratch_pool.get(device)
    if buf is None:
        buf = _ce_scratch_pool[device] = torch.zeros(2, dtype=torch.float64, device=device)
    return buf


@on_tensor_device
def smooth_cross_entropy(pred, target, eps=0.1):
    """mean_rows( -sum_c smooth_one_hot(target)[c] * log_softmax(pred)[c] ), pred [M,C] f32 on the device."""
    require_cuda(pred)
    return SmoothCE.apply(pred if pred.dtype == torch.float32 else pred.float(), target, float(eps))


def xyz2sphere(xyz, normalize=True):
    """R/modules/polar_utils.py:10-31: [..., 3] -> (rho, theta, phi), theta = 0 where rho = 0; normalised to
    [0, 1] when `normalize`.  Elementwise on the device (coordinates carry no gradient on this path)."""
    rho = torch.sqrt(torch.sum(torch.pow(xyz, 2), dim=-1, keepdim=True)).clamp(min=0)
    theta = torch.acos(xyz[..., 2, None] / rho)
    phi = torch.atan2(xyz[..., 1, None], xyz[..., 0, None])
    theta = torch.where(rho == 0, torch.zeros_like(theta), theta)
    if normalize:
        theta = theta / np.pi
        phi = phi / (2 * np.pi) + .5
    return torch.cat([rho, theta, phi], dim=-1)


@on_tensor_device
@torch.no_grad()
def umbrella_features(center, k=9, return_dist=True, sign=None):
    """The umbrella feature UmbrellaSurfaceConstructor feeds to its MLP (R/modules/pointnet2_utils.py:360-378):
    center [B,N,3] -> [B,N,k-1,10] (9 without the plane constant): centroid, spherical coordinates, unit normal,
    constant of the k-1 triangles around every point.  `sign` [B] (+1/-1) is the per-cloud random_inv flip.  One
    kNN launch + one fused kernel; coordinates carry no gradient on this path."""
    require_cuda(center)
    B, N, _ = center.shape
    center = _f32c(center.detach())
    _, idx = knn_point(k, center, center)
    C = 10 if return_dist else 9
    out = torch.empty(B, N, k - 1, C, dtype=torch.float32, device=center.device)
    if sign is not None:
        sign = _f32c(sign.to(center.device).reshape(B))
    call("mpc_umbrella_features_f32", ptr(center), ptr(idx), _i64(idx.stride(1)), ptr(sign), ptr(out), _i64(B),
         _i64(N), _i64(k), _i64(C), algo_bytes=B * N * (12 + k * 8 + (k - 1) * C * 4))
    return out


def _check_index(idx, n, what):
    if _CHECK_INDEX and idx.numel():
        lo, hi = int(idx.min()), int(idx.max())
        if lo < 0 or hi >= n:
            raise RuntimeError("%s: index out of range [0, %d): min %d max %d" % (what, n, lo, hi))


# ------------------------------------------------------------------------------------------------------
# gather / group
# ------------------------------------------------------------------------------------------------------
class _Gather(torch.autograd.Function):
    @staticmethod
    def forward(ctx, points, idx):
        B, N, C = points.shape
        M = idx.numel() // B if B else 0
        out = torch.empty(tuple(idx.shape) + (C,), dtype=torch.float32, device=points.device)
        call("mpc_gather_f32", ptr(points), ptr(idx), ptr(out), _i64(B), _i64(N), _i64(M), _i64(C),
             algo_bytes=B * M * (8 * C + 8))
        ctx.save_for_backward(idx)
        ctx.dims = (B, N, M, C)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        (idx,) = ctx.saved_tensors
        B, N, M, C = ctx.dims
        grad_out = _f32c(grad_out)
        grad_points = torch.empty(B, N, C, dtype=torch.float32, device=grad_out.device)
        call("mpc_gather_bwd_f32", ptr(grad_out), ptr(idx), ptr(grad_points), _i64(B), _i64(N), _i64(M), _i64(C),
             algo_bytes=B * (M * (8 * C + 8) + N * C * 4))
        return grad_points, None


class _GatherBF16(torch.autograd.Function):
    """index_points for bf16 feature rows: half the HBM bytes of the fp32 path.  Forward moves the rows bit for bit;
    backward scatters the bf16 gradient rows into an fp32 accumulator and casts the sum back."""

    @staticmethod
    def forward(ctx, points, idx):
        B, N, C = points.shape
        M = idx.numel() // B if B else 0
        out = torch.empty(tuple(idx.shape) + (C,), dtype=torch.bfloat16, device=points.device)
        call("mpc_gather_bf16", ptr(points), ptr(idx), ptr(out), _i64(B), _i64(N), _i64(M), _i64(C),
             algo_bytes=B * M * (4 * C + 8))
        ctx.save_for_backward(idx)
        ctx.dims = (B, N, M, C)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        (idx,) = ctx.saved_tensors
        B, N, M, C = ctx.dims
        grad_out = grad_out.to(torch.bfloat16).contiguous()
        acc = torch.empty(B, N, C, dtype=torch.float32, device=grad_out.device)
        if C % 2 == 0:
            call("mpc_gather_bwd_bf16", ptr(grad_out), ptr(idx), ptr(acc), _i64(B), _i64(N), _i64(M), _i64(C),
                 algo_bytes=B * (M * (2 * C + 8) + N * C * 4))
        else:
            g32 = grad_out.float()
            call("mpc_gather_bwd_f32", ptr(g32), ptr(idx), ptr(acc), _i64(B), _i64(N), _i64(M), _i64(C))
        return acc.to(torch.bfloat16), None


@on_tensor_device
def index_points(points, idx, cuda=False, is_group=False):
    """R/modules/pointnet2_utils.py:64-81.  points [B,N,C], idx [B,S] or [B,S,K] (int64) ->
    [B,S,C] / [B,S,K,C].  float32 payloads are differentiable (backward = scatter-add); int64 payloads
    (the composed FPS index chains of Fuse.forward) are moved bit-for-bit."""
    require_cuda(points, idx)
    idx = _i64c(idx)
    B, N, C = points.shape
    _check_index(idx, N, "index_points")
    if points.dtype == torch.float32:
        return _Gather.apply(_f32c(points), idx)
    if points.dtype == torch.bfloat16:
        return _GatherBF16.apply(points.contiguous(), idx)
    if points.dtype == torch.int64:
        points = points.contiguous()
        M = idx.numel() // B if B else 0
        out = torch.empty(tuple(idx.shape) + (C,), dtype=torch.int64, device=points.device)
        if C == 1:
            call("mpc_gather_i64", ptr(points), ptr(idx), ptr(out), _i64(B), _i64(N), _i64(M))
        else:  # an int64 row of C values is a float32 row of 2C values as far as a byte mover cares
            call("mpc_gather_f32", ptr(points), ptr(idx), ptr(out), _i64(B), _i64(N), _i64(M), _i64(2 * C))
        return out
    raise TypeError("index_points supports float32, bfloat16 and int64 payloads, got %s" % points.dtype)


# ------------------------------------------------------------------------------------------------------
# Markov state transition
# ------------------------------------------------------------------------------------------------------
def _csr_lists(idx, B, S, K, n_out):
    """Reverse-neighbour lists (CSR) of a kNN index tensor: workspace of mpc_transition_csr_build.  They depend on the
    indices only, and Fuse applies the same encoder kNN tensors in every decoder stage (R/modules/pointnet2_utils.py:
    663-704: 14 transitions over 9 distinct index tensors per forward), so inside a geometry scope they are built once
    per (index tensor, target size) and shared -- across streams through the event recorded after the build."""
    geo = _geo
    key = (idx.data_ptr(), idx._version, tuple(idx.shape), idx.stride(), n_out)
    if geo is not None:
        hit = geo.csr.get(key)
        if hit is not None:
            ws, ev, _ = hit
            torch.cuda.current_stream().wait_event(ev)
            ws.record_stream(torch.cuda.current_stream())
            return ws
    ws = torch.empty(B * (2 * n_out + 1) + B * S * K, dtype=torch.int32, device=idx.device)
    call("mpc_transition_csr_build", ptr(idx), ptr(ws), _i64(B), _i64(S), _i64(K), _i64(n_out),
         algo_bytes=B * S * K * 8)
    if geo is not None:
        ev = torch.cuda.Event()
        ev.record()
        geo.csr[key] = (ws, ev, idx)  # idx stays alive with the entry: its address is the key
    return ws


class _Transition(torch.autograd.Function):
    @staticmethod
    def forward(ctx, points, idx, n_out):
        B, S, C = points.shape
        K = idx.shape[2]
        out = torch.empty(B, n_out, C, dtype=torch.float32, device=points.device)
        cnt = torch.empty(B, n_out, dtype=torch.float32, device=points.device)
        if _TRANSITION_IMPL == "csr":  # gather form over reverse-neighbour lists (deterministic, no float atomics)
            ws = _csr_lists(idx, B, S, K, n_out)
            call("mpc_transition_csr_apply_f32", ptr(points), ptr(ws), ptr(out), ptr(cnt), _i64(B), _i64(S), _i64(K),
                 _i64(C), _i64(n_out), algo_bytes=B * ((S + n_out) * C * 4 + S * K * 8))
        else:
            call("mpc_transition_fwd_f32", ptr(points), ptr(idx), ptr(out), ptr(cnt), _i64(B), _i64(S), _i64(K),
                 _i64(C), _i64(n_out), algo_bytes=B * ((S + n_out) * C * 4 + S * K * 8))
        ctx.save_for_backward(idx, cnt)
        ctx.dims = (B, S, K, C, n_out)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        idx, cnt = ctx.saved_tensors
        B, S, K, C, n_out = ctx.dims
        grad_out = _f32c(grad_out)
        g = torch.empty(B, S, C, dtype=torch.float32, device=grad_out.device)
        call("mpc_transition_bwd_f32", ptr(grad_out), ptr(idx), ptr(cnt), ptr(g), _i64(B), _i64(S), _i64(K),
             _i64(C), _i64(n_out), algo_bytes=B * ((S + n_out) * C * 4 + S * K * 8 + n_out * 4))
        return g, None, None


@on_tensor_device
def upsample(points, knn_idx, scale_ratio=2, dist=None, n_out=None):
    """The Markov state transition, R/modules/pointnet2_utils.py:13-50: out = D^-1 A^T points with A the
    S x N kNN incidence (K ones per row) and D the count-normalisation of :44-48.  points [B,S,C], knn_idx
    [B,S,K] with values < S*scale_ratio -> [B, S*scale_ratio, C].  `dist` is accepted and ignored, as in the
    reference (:30-34).  n_out overrides S*scale_ratio (sizes that are not an integer multiple)."""
    require_cuda(points, knn_idx)
    S = points.shape[1]
    if n_out is None:
        n_out = S * scale_ratio
    knn_idx = _i64c(knn_idx)
    _check_index(knn_idx, n_out, "upsample")
    if points.dtype == torch.bfloat16:  # bf16 inference path: the sparse product accumulates in fp32
        return _Transition.apply(points.float().contiguous(), knn_idx, int(n_out)).to(torch.bfloat16)
    return _Transition.apply(_f32c(points), knn_idx, int(n_out))


# ------------------------------------------------------------------------------------------------------
# three_nn / three_interpolate
# ------------------------------------------------------------------------------------------------------
@torch.no_grad()
def three_nn(xyz1, xyz2):
    """R/modules/pointnet2_utils.py:899-901: for every xyz1 point the 3 nearest xyz2 points (expanded-form
    distances, ascending) -> (dist [B,N,3], idx int64 [B,N,3])."""
    return knn_point(3, xyz2, xyz1)


class ThreeInterpolate(torch.autograd.Function):
    """R/modules/pointnet2_utils.py:903-906 as an autograd Function: inverse-distance weights
    1/(d+1e-8) normalised over the 3 neighbours, weighted sum of the gathered rows."""

    @staticmethod
    def forward(ctx, points2, dist, idx):
        B, S, C = points2.shape
        N = idx.shape[1]
        weight = torch.empty(B, N, 3, dtype=torch.float32, device=points2.device)
        out = torch.empty(B, N, C, dtype=torch.float32, device=points2.device)
        call("mpc_three_interpolate_fwd_f32", ptr(points2), ptr(dist), ptr(idx), ptr(weight), ptr(out), _i64(B),
             _i64(N), _i64(S), _i64(C), algo_bytes=B * (4 * N * C * 4 + N * 36))
        ctx.save_for_backward(idx, weight)
        ctx.dims = (B, N, S, C)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        idx, weight = ctx.saved_tensors
        B, N, S, C = ctx.dims
        grad_out = _f32c(grad_out)
        g = torch.empty(B, S, C, dtype=torch.float32, device=grad_out.device)
        call("mpc_three_interpolate_bwd_f32", ptr(grad_out), ptr(weight), ptr(idx), ptr(g), _i64(B), _i64(N),
             _i64(S), _i64(C), algo_bytes=B * (4 * N * C * 4 + N * 36 + S * C * 4))
        return g, None, None


@on_tensor_device
def three_interpolate(points2, dist, idx):
    require_cuda(points2, dist, idx)
    return ThreeInterpolate.apply(_f32c(points2), _f32c(dist.detach()), _i64c(idx))


# ------------------------------------------------------------------------------------------------------
# difference-wise attention core
# ------------------------------------------------------------------------------------------------------
class AttnFeat(torch.autograd.Function):
    """LocalTrans core, feature branch (R/modules/pointnet2_utils.py:553-569).  q [B,S,C]; kv [B,N,2C] holds
    the projected keys in columns [0,C) and values in [C,2C) (one fused projection); idx [B,S,K]."""

    @staticmethod
    def forward(ctx, q, kv, idx):
        B, S, C = q.shape
        N = kv.shape[1]
        K = idx.shape[2]
        out = torch.empty(B, S, C, dtype=torch.float32, device=q.device)
        call("mpc_attn_feat_fwd_f32", ptr(q), _i64(C), ptr(kv), ctypes.c_void_p(kv.data_ptr() + 4 * C),
             _i64(2 * C), ptr(idx), ptr(out), _i64(B), _i64(S), _i64(N), _i64(K), _i64(C),
             algo_bytes=B * ((2 * N * C + 2 * S * C) * 4 + S * K * 8))
        ctx.save_for_backward(q, kv, idx)
        return out

    @staticmethod
    def backward(ctx, grad_ctx):
        q, kv, idx = ctx.saved_tensors
        B, S, C = q.shape
        N = kv.shape[1]
        K = idx.shape[2]
        grad_ctx = _f32c(grad_ctx)
        gq = torch.empty_like(q)
        gkv = torch.zeros_like(kv)
        call("mpc_attn_feat_bwd_f32", ptr(grad_ctx), ptr(q), _i64(C), ptr(kv),
             ctypes.c_void_p(kv.data_ptr() + 4 * C), _i64(2 * C), ptr(idx), ptr(gq), _i64(C), ptr(gkv),
             ctypes.c_void_p(gkv.data_ptr() + 4 * C), _i64(2 * C), ptr(None), _i64(B), _i64(S), _i64(N), _i64(K),
             _i64(C), algo_bytes=B * ((4 * N * C + 3 * S * C) * 4 + S * K * 8))
        return gq, gkv, None


def feat_attention_fusable(center, wq):
    """FeatAttention needs the tensor-core GEMMs (Cin % 32 == 0) and a channel count the bias-sum path covers
    (C/4 divides 256)."""
    Cin, C = center.shape[-1], wq.shape[0]
    return (_GEMM_IMPL == "tcgen05" and center.is_cuda and center.dtype == torch.float32 and Cin % 32 == 0
            and C % 4 == 0 and 256 % (C // 4) == 0 and C <= 1024)


class FeatAttention(torch.autograd.Function):
    """One autograd node for the feature branch of LocalTrans (R/modules/pointnet2_utils.py:548-569): q projection
    of the centres, fused k|v projection of all points (tcgen05 GEMMs), and the difference-wise attention core.
    The backward's attention kernel also emits the three bias gradients (column sums of grad_q / grad_k / grad_v),
    so nothing reduces over the points a second time."""

    @staticmethod
    def forward(ctx, center, features, idx, wq, bq, wk, bk, wv, bv):
        B, S, Cin = center.shape
        N = features.shape[1]
        C = wq.shape[0]
        K = idx.shape[2]
        dev = center.device
        c2d, f2d = center.reshape(-1, Cin), features.reshape(-1, Cin)
        wkv = torch.cat((wk, wv), 0)
        bkv = torch.cat((bk, bv), 0)
        q = torch.empty(B * S, C, dtype=torch.float32, device=dev)
        kv = torch.empty(B * N, 2 * C, dtype=torch.float32, device=dev)
        _tc_gemm(c2d, wq, bq, q)
        _tc_gemm(f2d, wkv, bkv, kv)
        out = torch.empty(B, S, C, dtype=torch.float32, device=dev)
        call("mpc_attn_feat_fwd_f32", ptr(q), _i64(C), ptr(kv), ctypes.c_void_p(kv.data_ptr() + 4 * C),
             _i64(2 * C), ptr(idx), ptr(out), _i64(B), _i64(S), _i64(N), _i64(K), _i64(C),
             algo_bytes=B * ((2 * N * C + 2 * S * C) * 4 + S * K * 8))
        ctx.save_for_backward(c2d, f2d, idx, wq, wkv, q, kv)
        ctx.shapes = (B, S, N, Cin, C, K, center.data_ptr() == features.data_ptr() and S == N)
        ctx.wparams = (wq, wk, wv)  # the parameter objects (the saved wkv is a concatenated copy)
        return out

    @staticmethod
    def backward(ctx, grad_ctx):
        c2d, f2d, idx, wq, wkv, q, kv = ctx.saved_tensors
        B, S, N, Cin, C, K, same = ctx.shapes
        dev = q.device
        grad_ctx = _f32c(grad_ctx)
        gq = torch.empty_like(q)
        gkv = torch.zeros_like(kv)
        gbias = torch.zeros(3 * C, dtype=torch.float32, device=dev)
        call("mpc_attn_feat_bwd_f32", ptr(grad_ctx), ptr(q), _i64(C), ptr(kv),
             ctypes.c_void_p(kv.data_ptr() + 4 * C), _i64(2 * C), ptr(idx), ptr(gq), _i64(C), ptr(gkv),
             ctypes.c_void_p(gkv.data_ptr() + 4 * C), _i64(2 * C), ptr(gbias), _i64(B), _i64(S), _i64(N), _i64(K),
             _i64(C), algo_bytes=B * ((4 * N * C + 3 * S * C) * 4 + S * K * 8))
        g_center = g_feat = None
        gwq = torch.empty(C, Cin, dtype=torch.float32, device=dev)
        gwkv = torch.empty(2 * C, Cin, dtype=torch.float32, device=dev)
        zeroed = ctx.needs_input_grad[0] or ctx.needs_input_grad[1]
        if zeroed:  # each dgrad launch also clears the split-reduction target of the wgrad GEMM that follows
            g_center = _tc_dgrad(gq, wq, c2d, zero=gwq).view(B, S, Cin)
            g_feat = _tc_dgrad(gkv, wkv, f2d, zero=gwkv).view(B, N, Cin)
        def wgrads():
            call("mpc_linear_wgrad_f32", ptr(gq), _i64(C), ptr(c2d), _i64(Cin), ptr(gwq), _i64(Cin), _i64(B * S),
                 _i64(Cin), _i64(C), _i64(1 if zeroed else 0), algo_bytes=(B * S * (Cin + C) + C * Cin) * 4)
            call("mpc_linear_wgrad_f32", ptr(gkv), _i64(2 * C), ptr(f2d), _i64(Cin), ptr(gwkv), _i64(Cin),
                 _i64(B * N), _i64(Cin), _i64(2 * C), _i64(1 if zeroed else 0),
                 algo_bytes=(B * N * (Cin + 2 * C) + 2 * C * Cin) * 4)

        if zeroed and _DEFER_WGRAD and _STREAMS_ENABLED and _wgrad_deferrable(*ctx.wparams):
            _defer_wgrad(wgrads, (gq, c2d, gkv, f2d, gwq, gwkv))  # off the critical path, see _defer_wgrad
        else:
            wgrads()
        return (g_center, g_feat, None, gwq, gbias[:C], gwkv[:C], gbias[C:2 * C], gwkv[C:], gbias[2 * C:])


class AttnXyz(torch.autograd.Function):
    """LocalTrans core, coordinate branch (R/modules/pointnet2_utils.py:520-544) with the q/k/v projections
    of the Cin-channel differences computed inside the kernel.  If the residual projection's weights (wr, br =
    conv_res.linear of the same block, :515) are given, the kernel also returns res = Wr centre + br (pre-BatchNorm),
    so the 3-channel GEMM, its weight-gradient reduction over all points and its bias reduction never run as
    separate library calls.  Returns (ctx, res or None)."""

    @staticmethod
    def forward(ctx, feat, center_idx, idx, wq, bq, wk, bk, wv, bv, wr, br):
        B, N, Cin = feat.shape
        S, K = idx.shape[1], idx.shape[2]
        C = wq.shape[0]
        out = torch.empty(B, S, C, dtype=torch.float32, device=feat.device)
        res = torch.empty(B, S, C, dtype=torch.float32, device=feat.device) if wr is not None else None
        call("mpc_attn_xyz_fwd_f32", ptr(feat), ptr(center_idx), ptr(idx), ptr(wq), ptr(bq), ptr(wk), ptr(bk),
             ptr(wv), ptr(bv), ptr(wr), ptr(br), ptr(out), ptr(res), _i64(B), _i64(S), _i64(N), _i64(K), _i64(Cin),
             _i64(C), algo_bytes=B * (N * Cin * 4 + S * K * 8 + S * C * 4 * (2 if wr is not None else 1))
             + 4 * C * (Cin + 1) * 4)
        ctx.save_for_backward(feat, idx, wq, bq, wk, bk, wv, bv, wr)
        ctx.center_idx = center_idx
        return out, res

    @staticmethod
    def backward(ctx, grad_ctx, grad_res):
        feat, idx, wq, bq, wk, bk, wv, bv, wr = ctx.saved_tensors
        center_idx = ctx.center_idx
        B, N, Cin = feat.shape
        S, K = idx.shape[1], idx.shape[2]
        C = wq.shape[0]
        dev = feat.device
        grad_ctx = _f32c(grad_ctx) if grad_ctx is not None else torch.zeros(B, S, C, device=dev)
        if wr is not None:
            grad_res = _f32c(grad_res) if grad_res is not None else torch.zeros(B, S, C, device=dev)
        else:
            grad_res = None
        gw = torch.zeros(4, C, Cin, dtype=torch.float32, device=dev)
        gb = torch.zeros(4, C, dtype=torch.float32, device=dev)
        gfeat = torch.zeros_like(feat) if ctx.needs_input_grad[0] else None
        call("mpc_attn_xyz_bwd_f32", ptr(grad_ctx), ptr(feat), ptr(center_idx), ptr(idx), ptr(wq), ptr(bq), ptr(wk),
             ptr(bk), ptr(wv), ptr(bv), ptr(wr), ptr(grad_res), ptr(gw[0]), ptr(gb[0]), ptr(gw[1]), ptr(gb[1]),
             ptr(gw[2]), ptr(gb[2]), ptr(gw[3]) if wr is not None else ptr(None),
             ptr(gb[3]) if wr is not None else ptr(None), ptr(gfeat), _i64(B), _i64(S), _i64(N), _i64(K), _i64(Cin),
             _i64(C), algo_bytes=B * (N * Cin * 4 + S * K * 8 + S * C * 4 * (2 if wr is not None else 1))
             + 8 * C * (Cin + 1) * 4)
        return (gfeat, None, None, gw[0], gb[0], gw[1], gb[1], gw[2], gb[2],
                gw[3] if wr is not None else None, gb[3] if wr is not None else None)


# ------------------------------------------------------------------------------------------------------
# BatchNorm1d-over-channels + LeakyReLU on the [M,C] view
# ------------------------------------------------------------------------------------------------------
_scratch_pool = {}


def _scratch(owner, role, C):
    """Persistent fp64 reduction scratch of one layer (2C+2 doubles), allocated zeroed once.  The C ABI's scratch
    contract (include/mpc_b200.h): zero on entry, and the kernel that consumes the sums leaves it zero again, so no
    memset launch sits between the producer and its predecessor in the stream.  Keyed by the layer's parameter
    storage and the role (a layer never runs concurrently with itself)."""
    key = (owner.device, owner.data_ptr(), role, C)
    buf = _scratch_pool.get(key)
    if buf is None:
        buf = _scratch_pool[key] = torch.zeros(2 * C + 2, dtype=torch.float64, device=owner.device)
    return buf


def reset_scratch():
    """Re-zero every pooled scratch buffer (only needed after a CUDA error interrupted a producer/consumer pair)."""
    for buf in list(_scratch_pool.values()) + list(_ce_scratch_pool.values()):
        buf.zero_()


_lib.on_error = reset_scratch


def _grad_rows(grad_out, C):
    """grad_out [M,C] as the BatchNorm-backward kernels can read it: a column slice of a wider gradient (what
    torch.cat's backward hands to each branch) is consumed in place when rows are 16-byte aligned and C is a power of
    two in [4, 1024]; anything else is made contiguous.  Returns (tensor, row stride in floats)."""
    if (grad_out.dim() == 2 and grad_out.dtype == torch.float32 and grad_out.stride(1) == 1
            and grad_out.stride(0) >= C and grad_out.stride(0) % 4 == 0 and grad_out.data_ptr() % 16 == 0
            and 4 <= C <= 1024 and (C & (C - 1)) == 0):
        return grad_out, grad_out.stride(0)
    return _f32c(grad_out), C


class BNAct(torch.autograd.Function):
    """Tail of the reference's `Linear` block (R/modules/pointnet2_utils.py:417-423) with bn=False (=> BatchNorm1d)
    on the [M,C] view.  Running statistics are updated in place exactly like nn.BatchNorm1d (momentum 0.1,
    unbiased variance into running_var)."""

    @staticmethod
    def forward(ctx, y, gamma, beta, running_mean, running_var, num_batches_tracked, training, momentum, eps,
                slope):
        M, C = y.shape
        dev = y.device
        if training:
            if M <= 1:
                raise ValueError("Expected more than 1 value per channel when training, got input size %s"
                                 % (tuple(y.shape),))
            scratch = _scratch(gamma, "bn_stats", C)
            stats = torch.empty(2 * C, dtype=torch.float32, device=dev)
            call("mpc_bn_stats_f32", ptr(y), ptr(stats), ptr(running_mean), ptr(running_var),
                 ptr(num_batches_tracked), ctypes.c_float(momentum), ptr(scratch), _i64(M), _i64(C),
                 algo_bytes=M * C * 4)
            mean, var = stats[:C], stats[C:]
        else:
            mean, var = running_mean, running_var
        out = torch.empty_like(y)
        call("mpc_bn_act_fwd_f32", ptr(y), ptr(mean), ptr(var), ptr(gamma), ptr(beta), ctypes.c_float(eps),
             ctypes.c_float(slope), ptr(None), ptr(out), _i64(M), _i64(C), algo_bytes=2 * M * C * 4)
        ctx.save_for_backward(y, mean, var, gamma, beta)
        ctx.cfg = (training, eps, slope)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        y, mean, var, gamma, beta = ctx.saved_tensors
        training, eps, slope = ctx.cfg
        M, C = y.shape
        grad_out = _f32c(grad_out)
        gy = torch.empty_like(y)
        gg = torch.empty(C, dtype=torch.float32, device=y.device)
        gb = torch.empty(C, dtype=torch.float32, device=y.device)
        scratch = _scratch(gamma, "bn_bwd", C)
        call("mpc_bn_act_bwd_f32", ptr(grad_out), ptr(y), ptr(mean), ptr(var), ptr(gamma), ptr(beta),
             ctypes.c_float(eps), ctypes.c_float(slope), ctypes.c_int(1 if training else 0), ptr(gy), ptr(gg),
             ptr(gb), ptr(scratch), ptr(None), _i64(0), _i64(C), _i64(M), _i64(C), algo_bytes=3 * M * C * 4)
        return gy, gg, gb, None, None, None, None, None, None, None


@on_tensor_device
def bn_act(y2d, gamma, beta, running_mean, running_var, num_batches_tracked, training, momentum=0.1, eps=1e-5,
           slope=0.2):
    """slope = 1.0 means no activation (Linear(..., act=False))."""
    require_cuda(y2d)
    return BNAct.apply(_f32c(y2d), gamma, beta, running_mean, running_var, num_batches_tracked, bool(training),
                       float(momentum), float(eps), float(slope))


# ------------------------------------------------------------------------------------------------------
# shared-MLP projection on the tensor cores
# ------------------------------------------------------------------------------------------------------
# Weight gradients off the backward critical path: only the optimiser (after backward) reads them, so the wgrad GEMM
# of a Linear+BN block is issued on a dedicated stream and joined once, by an autograd end-of-backward callback.
# OPT-IN (set_defer_wgrad(True) or MPC_DEFER_WGRAD=1): safe whenever gradients are only read after backward() has
# returned (this repo's training steps: all-reduce and optimiser follow backward); NOT safe under hooks that consume a
# gradient the moment autograd accumulates it (torch DDP's reducer), hence off by default in the drop-in modules.
_DEFER_WGRAD = os.environ.get("MPC_DEFER_WGRAD", "0") == "1"


def set_defer_wgrad(on):
    """Run the weight-gradient GEMMs of the shared-MLP blocks off the backward critical path (see above)."""
    global _DEFER_WGRAD
    _DEFER_WGRAD = bool(on)


_wgrad_streams = {}
_wgrad_pending = {}  # device -> id of the backward pass (autograd graph task) whose join callback is queued


def _wgrad_deferrable(*params):
    """Deferring is only sound when AccumulateGrad will STEAL the gradient buffer (p.grad is None: nothing reads
    the buffer before the end-of-backward join) and nothing observes the gradient earlier (tensor hooks /
    post-accumulate hooks, e.g. a DDP-style reducer).  With gradient accumulation (p.grad already set),
    zero_grad(set_to_none=False) or a weight used twice in one backward, autograd reads the buffer on the backward
    stream right away: those launches run inline -- after joining whatever the weight-gradient stream still has in
    flight for this device, because p.grad may be a buffer it is still writing."""
    ok = True
    for p in params:
        if p is None:
            continue
        if p.grad is not None or getattr(p, "_backward_hooks", None) or getattr(p, "_post_accumulate_grad_hooks", None):
            ok = False
    if not ok:
        cur = torch.cuda.current_stream()
        st = _wgrad_streams.get(cur.device)
        if st is not None:
            cur.wait_stream(st)
    return ok


def _defer_wgrad(fn, tensors):
    cur = torch.cuda.current_stream()
    st = _wgrad_streams.get(cur.device)
    if st is None:
        st = _wgrad_streams[cur.device] = torch.cuda.Stream(device=cur.device)
    st.wait_stream(cur)  # everything issued so far (the operands, the cleared output) is visible
    with torch.cuda.stream(st):
        fn()
    for t in tensors:
        if t is not None:
            t.record_stream(st)
    # one join per backward pass: keyed by the autograd graph task, so a backward that raised (its callback never ran)
    # cannot leave the device marked as pending for the next one
    task = torch._C._current_graph_task_id()
    if _wgrad_pending.get(cur.device) != task or task < 0:
        dev = cur.device
        _wgrad_pending[dev] = task

        def join():
            _wgrad_pending.pop(dev, None)
            torch.cuda.current_stream(dev).wait_stream(st)

        if task >= 0:
            torch.autograd.Variable._execution_engine.queue_callback(join)
        else:  # not inside a backward pass (direct call of a Function's backward): join at once
            join()


# dgrad || wgrad of a Linear+BN block on two streams: measured 2 % SLOWER on the part-seg step (both GEMMs want every
# SM), so off by default
_BWD_PAIR = os.environ.get("MPC_BWD_PAIR", "0") == "1"
_GEMM_IMPL = os.environ.get("MPC_GEMM", "tcgen05")  # "cublas" forces the library GEMM (A/B comparisons)


def _tc_ok(x2d, w):
    K = x2d.shape[1]
    return (_GEMM_IMPL == "tcgen05" and x2d.is_cuda and x2d.dtype == torch.float32 and K % 32 == 0
            and x2d.shape[0] > 0)


def _tc_gemm(x2d, w, bias, out, stat_scratch=None, group_bias=None, rows_per_group=0):
    """out[M,N] = x2d[M,K] @ w[N,:K]^T (+ bias) (+ group_bias[row // rows_per_group]) through mpc_linear_fwd_f32
    (3xTF32 tcgen05).  stat_scratch (2N+2 doubles) additionally receives the per-column sum / sum of squares of `out`
    from the epilogue.  K is taken from x2d, so `w` may be wider (a column slice of it is used)."""
    M, K = x2d.shape
    N = w.shape[0]
    call("mpc_linear_fwd_f32", ptr(x2d), _i64(x2d.stride(0)), ptr(w), _i64(w.stride(0)), ptr(bias), ptr(out),
         _i64(out.stride(0)), ptr(stat_scratch), ptr(group_bias), _i64(rows_per_group), _i64(M), _i64(K), _i64(N),
         algo_bytes=(M * K + M * N + N * K) * 4)


def _tc_dgrad(gy, w, x_like, zero=None):
    """grad_x[M,K] = gy[M,N] @ w[N,K] on the tensor cores (weights consumed as stored); library GEMM when the
    shape is outside the kernel's reach.  `zero` (optional tensor) is cleared by the same launch for the
    weight-gradient GEMM that follows."""
    M, N = gy.shape
    K = w.shape[1]
    if _GEMM_IMPL == "tcgen05" and K % 32 == 0 and gy.stride(0) % 4 == 0 and M > 0:
        gx = torch.empty(M, K, dtype=torch.float32, device=gy.device)
        call("mpc_linear_dgrad_f32", ptr(gy), _i64(gy.stride(0)), ptr(w), _i64(w.stride(0)), ptr(gx), _i64(K),
             _i64(M), _i64(K), _i64(N), ptr(zero), _i64(zero.numel() if zero is not None else 0),
             algo_bytes=(M * K + M * N + N * K) * 4)
        return gx
    if zero is not None:
        zero.zero_()
    return gy.mm(w)


class LinearTC(torch.autograd.Function):
    """y = x W^T + b with forward, grad-input and grad-weight all on the tcgen05 kernel.  An output width that is
    not a multiple of 4 (the 50-class head) is computed into / read from a buffer whose rows are padded to 16 bytes,
    which is all TMA needs; the pad columns are never read."""

    @staticmethod
    def forward(ctx, x2d, w, bias):
        M, K = x2d.shape
        N = w.shape[0]
        Np = (N + 3) & ~3
        y = torch.empty(M, Np, dtype=torch.float32, device=x2d.device)
        if Np != N:
            y = y[:, :N]
        _tc_gemm(x2d, w, bias, y)
        ctx.save_for_backward(x2d, w)
        ctx.has_bias = bias is not None
        return y

    @staticmethod
    def backward(ctx, gy):
        x2d, w = ctx.saved_tensors
        M, K = x2d.shape
        N = w.shape[0]
        if N % 4:
            gp = torch.empty(M, (N + 3) & ~3, dtype=torch.float32, device=gy.device)[:, :N]
            gp.copy_(gy)
            gy = gp
        else:
            gy = _f32c(gy)
        gx = gw = gb = None
        tc_wgrad = ctx.needs_input_grad[1] and _GEMM_IMPL == "tcgen05" and K % 32 == 0
        if tc_wgrad:
            gw = torch.empty(N, K, dtype=torch.float32, device=gy.device)
        if ctx.needs_input_grad[0]:  # the dgrad launch also clears gw (split-reduction target of the wgrad GEMM)
            gx = _tc_dgrad(gy, w, x2d, zero=gw)
        if ctx.needs_input_grad[1]:
            if tc_wgrad:
                call("mpc_linear_wgrad_f32", ptr(gy), _i64(gy.stride(0)), ptr(x2d), _i64(x2d.stride(0)), ptr(gw),
                     _i64(K), _i64(M), _i64(K), _i64(N), _i64(1 if ctx.needs_input_grad[0] else 0),
                     algo_bytes=(M * K + M * N + N * K) * 4)
            else:
                gw = gy.t().mm(x2d)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            cv = N // 4
            if N % 4 == 0 and 1 <= cv <= 256 and (cv & (cv - 1)) == 0:
                gb = torch.empty(N, dtype=torch.float32, device=gy.device)
                scratch = _scratch(w, "col_sum", N)
                call("mpc_col_sum_f32", ptr(gy), ptr(gb), ptr(scratch), _i64(M), _i64(N), algo_bytes=M * N * 4)
            else:
                gb = gy.sum(0)
        return gx, gw, gb


class LinearBNAct(torch.autograd.Function):
    """The whole shared-MLP block of the reference (`Linear.forward`, R/modules/pointnet2_utils.py:413-425, with
    bn=False => BatchNorm1d): y = x W^T + b on the tensor cores, batch statistics, normalise + LeakyReLU -- one
    autograd node.  Backward: BN/activation backward (two streaming kernels), grad-input and grad-weight on the
    tensor cores.  The Linear bias feeds a BatchNorm, so its gradient is known in closed form: exactly 0 with
    batch statistics (BatchNorm removes any per-channel constant) and gamma * rsqrt(var + eps) * grad_beta with
    running statistics -- no reduction over the M rows is needed."""

    @staticmethod
    def forward(ctx, x2d, w, bias, gamma, beta, running_mean, running_var, num_batches_tracked, training, momentum,
                eps, slope, residual):
        M, K = x2d.shape
        N = w.shape[0]
        dev = x2d.device
        fuse_res = residual is not None and N % 4 == 0 and N <= 1024 and 1024 % N == 0
        y = torch.empty(M, N, dtype=torch.float32, device=dev)
        if training:
            if M <= 1:
                raise ValueError("Expected more than 1 value per channel when training, got input size %s"
                                 % ((M, N),))
            # batch statistics come out of the GEMM epilogue (the tile is summed while still in shared memory)
            scratch = _scratch(w, "fwd_stats", N)
            stats = torch.empty(2 * N, dtype=torch.float32, device=dev)
            _tc_gemm(x2d, w, bias, y, stat_scratch=scratch)
            mean, var = stats[:N], stats[N:]
            out = torch.empty_like(y)
            if N % 4 == 0 and N <= 1024 and 1024 % N == 0:
                # finalise + normalise + activation + running-statistics update in one launch
                call("mpc_bn_act_fwd_sums_f32", ptr(y), ptr(scratch), ptr(gamma), ptr(beta), ctypes.c_float(eps),
                     ctypes.c_float(slope), ptr(residual if fuse_res else None), ptr(out), ptr(stats),
                     ptr(running_mean), ptr(running_var), ptr(num_batches_tracked), ctypes.c_float(momentum), _i64(M),
                     _i64(N), algo_bytes=(3 if fuse_res else 2) * M * N * 4)
            else:
                call("mpc_bn_finalize_f32", ptr(scratch), ptr(stats), ptr(running_mean), ptr(running_var),
                     ptr(num_batches_tracked), ctypes.c_float(momentum), _i64(M), _i64(N))
                call("mpc_bn_act_fwd_f32", ptr(y), ptr(mean), ptr(var), ptr(gamma), ptr(beta), ctypes.c_float(eps),
                     ctypes.c_float(slope), ptr(None), ptr(out), _i64(M), _i64(N), algo_bytes=2 * M * N * 4)
        else:
            _tc_gemm(x2d, w, bias, y)
            mean, var = running_mean, running_var
            out = torch.empty_like(y)
            call("mpc_bn_act_fwd_f32", ptr(y), ptr(mean), ptr(var), ptr(gamma), ptr(beta), ctypes.c_float(eps),
                 ctypes.c_float(slope), ptr(residual if fuse_res else None), ptr(out), _i64(M), _i64(N),
                 algo_bytes=(3 if fuse_res else 2) * M * N * 4)
        if residual is not None and not fuse_res:
            out = out + residual
        ctx.save_for_backward(x2d, w, y, mean, var, gamma, beta)
        ctx.cfg = (training, eps, slope, bias is not None)
        ctx.has_residual = residual is not None
        return out

    @staticmethod
    def backward(ctx, grad_out):
        x2d, w, y, mean, var, gamma, beta = ctx.saved_tensors
        training, eps, slope, has_bias = ctx.cfg
        M, K = x2d.shape
        N = w.shape[0]
        dev = y.device
        grad_out, ld_gout = _grad_rows(grad_out, N)
        gy = torch.empty_like(y)
        gg = torch.empty(N, dtype=torch.float32, device=dev)
        gb = torch.empty(N, dtype=torch.float32, device=dev)
        scratch = _scratch(w, "bn_bwd", N)
        gx = gw = gbias = None
        tc_wgrad = ctx.needs_input_grad[1] and K % 32 == 0 and N % 4 == 0
        zero_bias = has_bias and ctx.needs_input_grad[2] and training and N % 4 == 0
        flat = None
        if tc_wgrad or zero_bias:
            # one buffer cleared by the BatchNorm-backward launch: the weight gradient (split-reduction target of the
            # wgrad GEMM: no memset node in front of it) and the Linear bias gradient, which is exactly zero under
            # batch statistics (no fill launch)
            flat = torch.empty((N * K if tc_wgrad else 0) + (N if zero_bias else 0), dtype=torch.float32, device=dev)
            if tc_wgrad:
                gw = flat[:N * K].view(N, K)
            if zero_bias:
                gbias = flat[flat.numel() - N:]
        call("mpc_bn_act_bwd_f32", ptr(grad_out), ptr(y), ptr(mean), ptr(var), ptr(gamma), ptr(beta),
             ctypes.c_float(eps), ctypes.c_float(slope), ctypes.c_int(1 if training else 0), ptr(gy), ptr(gg),
             ptr(gb), ptr(scratch), ptr(flat), _i64(flat.numel() if flat is not None else 0), _i64(ld_gout), _i64(M),
             _i64(N), algo_bytes=3 * M * N * 4)
        def wgrad():
            call("mpc_linear_wgrad_f32", ptr(gy), _i64(N), ptr(x2d), _i64(K), ptr(gw), _i64(K), _i64(M), _i64(K),
                 _i64(N), _i64(1), algo_bytes=(M * K + M * N + N * K) * 4)
            return gw

        if ctx.needs_input_grad[0] and tc_wgrad and _BWD_PAIR:
            # grad-input and grad-weight only share their input: side by side (two graph branches)
            gx, _ = parallel(lambda: _tc_dgrad(gy, w, x2d), wgrad)
        elif (ctx.needs_input_grad[0] and tc_wgrad and _DEFER_WGRAD and _STREAMS_ENABLED
              and _wgrad_deferrable(w)):
            # nothing upstream waits for a weight gradient: it runs on the weight-gradient stream, ordered after the
            # BatchNorm backward above, and rejoins at the end of the backward pass (see _defer_wgrad)
            _defer_wgrad(wgrad, (gy, x2d, flat))
            gx = _tc_dgrad(gy, w, x2d)
        else:
            if ctx.needs_input_grad[0]:
                gx = _tc_dgrad(gy, w, x2d)
            if tc_wgrad:
                wgrad()
        if ctx.needs_input_grad[1] and not tc_wgrad:
            gw = gy.t().mm(x2d)
        if has_bias and ctx.needs_input_grad[2] and gbias is None:
            gbias = torch.zeros_like(gb) if training else gamma * torch.rsqrt(var + eps) * gb
        # out = act(BN(y)) + residual: the residual's gradient is the incoming gradient itself
        gres = grad_out if ctx.has_residual else None
        return gx, gw, gbias, gg, gb, None, None, None, None, None, None, None, gres


class LinearBNActSplit(torch.autograd.Function):
    """Linear -> BatchNorm -> LeakyReLU over cat(x_a, broadcast(g)) WITHOUT building the concatenation: the channels
    in g [G,Kb] are constant over the `rows` consecutive rows of a cloud, so  y = x_a Wa^T + (g Wb^T)[cloud] + b  with
    w = [Wa | Wb].  The per-cloud term enters the tcgen05 GEMM's epilogue as a row-group bias (BatchNorm statistics
    then come out of the same epilogue), and backward needs one per-cloud column sum of grad_y plus two tiny
    matmuls for the g / Wb gradients.  The part-seg head's 896 -> 512 layer (640 broadcast channels,
    R/modules/pointnet2_utils.py:843-853 + R/models/repsurf/pointnet2_part_seg_msg.py:137) shrinks 3.5x."""

    @staticmethod
    def forward(ctx, x2d, g, w, bias, gamma, beta, running_mean, running_var, num_batches_tracked, training, momentum,
                eps, slope, rows):
        M, Ka = x2d.shape
        N = w.shape[0]
        dev = x2d.device
        wb = w[:, Ka:]
        v = g.mm(wb.t()).contiguous()  # [G,N] per-cloud bias
        y = torch.empty(M, N, dtype=torch.float32, device=dev)
        out = torch.empty_like(y)
        if training:
            scratch = _scratch(w, "fwd_stats", N)
            stats = torch.empty(2 * N, dtype=torch.float32, device=dev)
            _tc_gemm(x2d, w, bias, y, stat_scratch=scratch, group_bias=v, rows_per_group=rows)
            mean, var = stats[:N], stats[N:]
            call("mpc_bn_act_fwd_sums_f32", ptr(y), ptr(scratch), ptr(gamma), ptr(beta), ctypes.c_float(eps),
                 ctypes.c_float(slope), ptr(None), ptr(out), ptr(stats), ptr(running_mean), ptr(running_var),
                 ptr(num_batches_tracked), ctypes.c_float(momentum), _i64(M), _i64(N), algo_bytes=2 * M * N * 4)
        else:
            _tc_gemm(x2d, w, bias, y, group_bias=v, rows_per_group=rows)
            mean, var = running_mean, running_var
            call("mpc_bn_act_fwd_f32", ptr(y), ptr(mean), ptr(var), ptr(gamma), ptr(beta), ctypes.c_float(eps),
                 ctypes.c_float(slope), ptr(None), ptr(out), _i64(M), _i64(N), algo_bytes=2 * M * N * 4)
        ctx.save_for_backward(x2d, g, w, y, mean, var, gamma, beta)
        ctx.cfg = (training, eps, slope, bias is not None, rows)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        x2d, g, w, y, mean, var, gamma, beta = ctx.saved_tensors
        training, eps, slope, has_bias, rows = ctx.cfg
        M, Ka = x2d.shape
        N, Kt = w.shape
        dev = y.device
        grad_out = _f32c(grad_out)
        gy = torch.empty_like(y)
        gg = torch.empty(N, dtype=torch.float32, device=dev)
        gb = torch.empty(N, dtype=torch.float32, device=dev)
        scratch = _scratch(w, "bn_bwd", N)
        flat = torch.empty(N * Kt + N, dtype=torch.float32, device=dev)  # [grad_w | grad_bias], cleared by the launch below
        gw = flat[:N * Kt].view(N, Kt)
        call("mpc_bn_act_bwd_f32", ptr(grad_out), ptr(y), ptr(mean), ptr(var), ptr(gamma), ptr(beta),
             ctypes.c_float(eps), ctypes.c_float(slope), ctypes.c_int(1 if training else 0), ptr(gy), ptr(gg),
             ptr(gb), ptr(scratch), ptr(flat), _i64(flat.numel()), _i64(N), _i64(M), _i64(N),
             algo_bytes=3 * M * N * 4)

        def wgrad():  # grad_w[:, :Ka] = gy^T x_a, written into the column slice of the full-width gradient
            call("mpc_linear_wgrad_f32", ptr(gy), _i64(N), ptr(x2d), _i64(Ka), ptr(gw), _i64(Kt), _i64(M), _i64(Ka),
                 _i64(N), _i64(1), algo_bytes=(M * Ka + M * N + N * Ka) * 4)

        deferred = _DEFER_WGRAD and _STREAMS_ENABLED and _wgrad_deferrable(w)
        if deferred:
            _defer_wgrad(wgrad, (gy, x2d, flat))
        else:
            wgrad()
        gx = None
        if ctx.needs_input_grad[0]:
            gx = torch.empty(M, Ka, dtype=torch.float32, device=dev)
            call("mpc_linear_dgrad_f32", ptr(gy), _i64(N), ptr(w), _i64(Kt), ptr(gx), _i64(Ka), _i64(M), _i64(Ka),
                 _i64(N), ptr(None), _i64(0), algo_bytes=(M * Ka + M * N + N * Ka) * 4)
        colsum = gy.view(-1, rows, N).sum(1)  # [G,N]: per-cloud column sums of grad_y
        gg_ = colsum.mm(w[:, Ka:]) if ctx.needs_input_grad[1] else None
        gwb = colsum.t().mm(g)  # [N,Kb]
        if deferred:
            # the deferred wgrad owns `flat` on its stream: the Wb slice is written there too, after it
            st = _wgrad_streams[dev]
            st.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(st):
                gw[:, Ka:].copy_(gwb)
            gwb.record_stream(st)
        else:
            gw[:, Ka:].copy_(gwb)
        gbias = None
        if has_bias and ctx.needs_input_grad[3]:
            gbias = flat[N * Kt:] if training else gamma * torch.rsqrt(var + eps) * gb
        return gx, gg_, gw, gbias, gg, gb, None, None, None, None, None, None, None, None


@on_tensor_device
def linear_bn_act_split(x_a, g, weight, bias, bn, training, slope):
    """x_a [B,Np,Ka] (per-point channels), g [B,Kb] (per-cloud channels) -> act(BN(Linear(cat(x_a, broadcast g)))),
    [B,Np,N]; see LinearBNActSplit.  Caller guarantees Ka % 32 == 0 (split_supported)."""
    require_cuda(x_a, g)
    B, Np, Ka = x_a.shape
    N = weight.shape[0]
    out = LinearBNActSplit.apply(_f32c(x_a.reshape(-1, Ka)), _f32c(g), weight.contiguous(), bias, bn.weight, bn.bias,
                                 bn.running_mean, bn.running_var, bn.num_batches_tracked, bool(training),
                                 _momentum(bn), float(bn.eps), float(slope), int(Np))
    return out.view(B, Np, N)


def _momentum(bn):
    """nn.BatchNorm's momentum=None means a cumulative moving average: factor 1 / num_batches_tracked (after the
    increment this forward performs)."""
    if bn.momentum is not None:
        return float(bn.momentum)
    return 1.0 / float(int(bn.num_batches_tracked) + 1)


def split_supported(n_points, Ka, N):
    return (_GEMM_IMPL == "tcgen05" and n_points > 0 and Ka % 32 == 0 and N % 4 == 0 and N <= 1024
            and 1024 % N == 0)


@on_tensor_device
def linear_bn_act(x, weight, bias, bn, training, slope, residual=None):
    """Linear -> BatchNorm1d(channels) -> LeakyReLU(slope) [+ residual] on any [..., K] input (slope = 1: no
    activation).  `bn` is the nn.BatchNorm1d holding gamma / beta / running statistics.  `residual` ([..., N]) is
    added to the result inside the normalise kernel (LocalTrans's `residual + ffn(context)`, Fuse's `conv(acc) + f`)."""
    require_cuda(x)
    shape = x.shape
    x2d = x.reshape(-1, shape[-1])
    N = weight.shape[0]
    if _INFER_BF16 and not training:
        scale, shift = _bn_affine(bias, bn)
        res2d = residual.reshape(-1, N) if residual is not None else None
        if _bf16_gemm_ok(x2d):
            out = linear_bf16(to_bf16_rows(x2d), _bf16_weight(weight), scale, shift, float(slope), res2d)
        else:  # K = 3 / 16 layers: a few kFLOP per point, library GEMM in fp32, one rounding at the end
            y = torch.nn.functional.linear(x2d.float(), weight) * scale + shift
            y = torch.nn.functional.leaky_relu(y, float(slope))
            out = (y if res2d is None else y + res2d.float()).to(torch.bfloat16)
        return out.view(*shape[:-1], N)
    if (not training and not torch.is_grad_enabled() and _FUSE_EVAL and _tc_ok(x2d, weight) and N % 4 == 0
            and x2d.data_ptr() % 16 == 0 and x2d.stride(0) % 4 == 0):
        # inference: BatchNorm (running statistics) is a per-channel affine map -- it, the LeakyReLU and the residual
        # add run in the GEMM's epilogue (mpc_linear_affine_act_f32): one launch, no normalise pass over HBM
        scale, shift = _bn_affine(bias, bn)
        x2d = _f32c(x2d)
        w = weight.contiguous()
        M, K = x2d.shape
        res2d = _f32c(residual.reshape(-1, N)) if residual is not None else None
        out = torch.empty(M, N, dtype=torch.float32, device=x2d.device)
        call("mpc_linear_affine_act_f32", ptr(x2d), _i64(x2d.stride(0)), ptr(w), _i64(w.stride(0)), ptr(scale),
             ptr(shift), ctypes.c_float(float(slope)), ptr(res2d), _i64(N if res2d is not None else 0), ptr(out), _i64(N),
             _i64(M), _i64(K), _i64(N), algo_bytes=(M * K + M * N * (2 if res2d is not None else 1) + N * K) * 4)
        return out.view(*shape[:-1], N)
    if _tc_ok(x2d, weight):
        res2d = _f32c(residual.reshape(-1, N)) if residual is not None else None
        out = LinearBNAct.apply(_f32c(x2d), weight.contiguous(), bias, bn.weight, bn.bias, bn.running_mean,
                                bn.running_var, bn.num_batches_tracked, bool(training), _momentum(bn),
                                float(bn.eps), float(slope), res2d)
    else:
        y = torch.nn.functional.linear(x2d, weight, bias)
        out = bn_act(y, bn.weight, bn.bias, bn.running_mean, bn.running_var, bn.num_batches_tracked, training,
                     momentum=_momentum(bn), eps=bn.eps, slope=slope)
        if residual is not None:
            out = out + residual.reshape(-1, N)
    return out.view(*shape[:-1], N)


@on_tensor_device
def linear(x, weight, bias=None):
    """nn.Linear's arithmetic on any [..., K] input: the tcgen05 3xTF32 kernel when K % 32 == 0, else the
    library GEMM (K = 3 / 16 layers: a few kFLOP per point)."""
    require_cuda(x)
    shape = x.shape
    x2d = x.reshape(-1, shape[-1])
    if _INFER_BF16 and not torch.is_grad_enabled() and _bf16_gemm_ok(x2d):
        y = linear_bf16(to_bf16_rows(x2d), _bf16_weight(weight), None, bias.detach() if bias is not None else None)
        return y.view(*shape[:-1], weight.shape[0])
    if _tc_ok(x2d, weight):
        x2d = _f32c(x2d)
        y = LinearTC.apply(x2d, weight.contiguous(), bias)
    else:
        y = torch.nn.functional.linear(x2d, weight, bias)
    return y.view(*shape[:-1], weight.shape[0])


# ------------------------------------------------------------------------------------------------------
# stream-level concurrency for independent branches
# ------------------------------------------------------------------------------------------------------
_STREAMS_ENABLED = os.environ.get("MPC_STREAMS", "1") == "1"
_stream_pool = {}


def _side_streams(device, n, avoid):
    pool = _stream_pool.setdefault(device, [])
    while len(pool) < 8:
        pool.append(torch.cuda.Stream(device=device))
    out = [st for st in pool if st != avoid]
    return out[:n]


def _tensors_of(obj):
    if isinstance(obj, torch.Tensor):
        yield obj
    elif isinstance(obj, (tuple, list)):
        for o in obj:
            yield from _tensors_of(o)


def parallel(*thunks):
    """Run independent branches concurrently: thunk 0 on the current stream, the others each on a side stream
    forked from it and joined back before returning (under CUDA-graph capture these become parallel graph
    branches).  The many small kernels of one Markov stage (three attention branches, the four source
    projections of a Fuse) are each latency-bound on their own; side by side they fill the machine.  Autograd
    runs every backward op on its forward op's stream, so the backward pass overlaps the same way.
    Results are returned in thunk order.  MPC_STREAMS=0 runs them one after the other."""
    if not _STREAMS_ENABLED or len(thunks) < 2 or not torch.cuda.is_available():
        return [t() for t in thunks]
    cur = torch.cuda.current_stream()
    streams = [st for st in _side_streams(cur.device, 8, cur)[:-1]
               if _geo is None or st != _geo.stream][:len(thunks) - 1]
    if len(streams) < len(thunks) - 1:
        return [t() for t in thunks]
    results = [None] * len(thunks)
    for st in streams:
        st.wait_stream(cur)
    for i, st in enumerate(streams):
        with torch.cuda.stream(st):
            results[i + 1] = thunks[i + 1]()
    results[0] = thunks[0]()
    for st in streams:
        cur.wait_stream(st)
    for r in results[1:]:
        for t in _tensors_of(r):
            if t.is_cuda:
                t.record_stream(cur)  # produced on a side stream, consumed (and later freed) on this one
    return results


class _GeoScope:
    """Coordinate-only work (FPS, coordinate-space kNN, coordinate gathers) depends on nothing but the input
    cloud, so it runs ahead of the feature pipeline that consumes it, on two lanes: the sampling chain
    (FPS -> gather -> FPS -> ...; strictly sequential, one CTA per cloud) on a high-priority stream of its own, and
    the neighbour searches on the geometry stream.  Every result carries the event recorded right after it was
    produced; a consumer waits for exactly that event, not for everything queued on the lane."""

    def __init__(self):
        cur = torch.cuda.current_stream()
        self.stream = _side_streams(cur.device, 8, cur)[-1]  # the last pool stream: never handed to parallel()
        self.fps_stream = _fps_streams.get(cur.device)
        if self.fps_stream is None:
            self.fps_stream = _fps_streams[cur.device] = torch.cuda.Stream(device=cur.device, priority=-1)
        self.stream.wait_stream(cur)  # fork: everything issued so far (the input cloud) is visible
        self.fps_stream.wait_stream(cur)
        self.cache = {}      # coordinate-space kNN results of this forward, by operand identity
        self.csr = {}        # reverse-neighbour lists of the transitions of this forward, by index-tensor identity
        self.fps_cache = {}  # prefetched (FPS indices, sampled coordinates), by operand identity
        self.events = {}     # tensor storage address -> event recorded after its producer
        self.keep = []       # keeps those tensors (hence their addresses) alive for the scope
        self.used_fps_lane = False

    def call(self, fn, lane="knn", deps=()):
        st = self.fps_stream if lane == "fps" else self.stream
        if lane == "fps":
            self.used_fps_lane = True
        for d in deps:  # operands produced on the other lane
            ev = self.events.get(d.data_ptr())
            if ev is not None:
                st.wait_event(ev)
        with torch.cuda.stream(st):
            out = fn()
        ev = torch.cuda.Event()
        ev.record(st)
        for t in _tensors_of(out):
            if t.is_cuda and t.data_ptr() not in self.events:  # a cached result keeps its producer's event
                self.events[t.data_ptr()] = ev
                self.keep.append(t)
                t.record_stream(self.stream)
                t.record_stream(self.fps_stream)
        return out

    def join(self, outputs=None):
        cur = torch.cuda.current_stream()
        tensors = [t for t in _tensors_of(outputs) if t.is_cuda]
        evs = [self.events.get(t.data_ptr()) for t in tensors]
        if not tensors or any(e is None for e in evs):
            cur.wait_stream(self.stream)
            cur.wait_stream(self.fps_stream)
        else:
            for e in set(evs):
                cur.wait_event(e)
        for t in tensors:
            t.record_stream(cur)


_geo = None
_fps_streams = {}


@contextlib.contextmanager
def geometry_scope():
    """Inside this scope geo_call() runs its thunk on a geometry lane; geo_join() makes the current stream wait
    for the producers of the given results.  Outside a scope (or with MPC_STREAMS=0) both are no-ops."""
    global _geo
    old = _geo
    _geo = _GeoScope() if (_STREAMS_ENABLED and torch.cuda.is_available()) else None
    try:
        yield
    finally:
        if _geo is not None:
            torch.cuda.current_stream().wait_stream(_geo.stream)
            torch.cuda.current_stream().wait_stream(_geo.fps_stream)
        _geo = old


def geo_call(fn, lane="knn", deps=()):
    return _geo.call(fn, lane, deps) if _geo is not None else fn()


def geo_join(outputs=None):
    if _geo is not None:
        _geo.join(outputs)
    return outputs


@torch.no_grad()
def fps_and_gather(points, npoint):
    """One sampling step of the encoders (R/modules/pointnet2_utils.py:771-772 and the like): FPS indices and the
    sampled coordinates.  Inside a geometry scope a prefetched result (geo_prefetch_pyramid) is returned."""
    if _geo is not None:
        hit = _geo.fps_cache.get((points.data_ptr(), tuple(points.shape), points.stride(), npoint))
        if hit is not None:
            _record("fps", hit[0])
            return hit[0], hit[1]
    idx = farthest_point_sample(points, npoint)
    return idx, index_points(points, idx)


def geo_prefetch_pyramid(xyz, npoints, k, self_levels=(0,), cross=()):
    """Issue the whole coordinate pyramid of one forward up front: the sampling chain xyz -> npoints[0] -> npoints[1]
    ... on the sampling lane, and on the neighbour-search lane the k-NN of every sampled state in its parent state
    (LocalMerge encoder stages), of the states in `self_levels` in themselves (level-0 stage, decoder stages) and of
    the (target, source) state pairs in `cross` (Fuse's non-adjacent transitions).  The consumers then find their
    results in the scope's caches (knn_point / fps_and_gather) in their own call order, so index recording and the
    CPU-generator draws for the FPS start indices keep the reference's order.  Returns the list of states' coordinates.
    No-op (returns None) outside a geometry scope or while indices are being injected."""
    if _geo is None or _tape.inject is not None:
        return None
    geo = _geo
    levels = [xyz]
    done = set()

    def knn(ref, qry):
        key = (ref.data_ptr(), qry.data_ptr())
        if key not in done:
            done.add(key)
            geo.call(lambda: _knn_compute(k, ref, qry), "knn", deps=(ref, qry))

    if 0 in self_levels:
        knn(xyz, xyz)
    for npoint in npoints:
        base = levels[-1]

        def sample(base=base, npoint=npoint):
            idx = _fps_compute(base, npoint)
            sub = index_points(base, idx)
            geo.fps_cache[(base.data_ptr(), tuple(base.shape), base.stride(), npoint)] = (idx, sub, base)
            return idx, sub

        _, sub = geo.call(sample, "fps", deps=(base,))
        levels.append(sub)
        knn(base, sub)
    for lv in self_levels:
        knn(levels[lv], levels[lv])
    for t, j in cross:
        knn(levels[t], levels[j])
    return levels


def launches():
    """C-ABI calls issued so far in this process (each enqueues one or more of our kernels)."""
    return _lib.launch_count


def kernels_launched():
    """Kernels of ours launched so far in this process."""
    return _lib.kernel_count
