"""CPU oracle for the point-set hot path.  TEST INFRASTRUCTURE ONLY -- never imported by the product.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs may import
this module.  The product package (markov-process-analysis-on-point-cloud_b200/) must work, and fail
loudly, without it.

What this is: a functional restatement (parameters live in a plain dict keyed exactly like the
reference's state_dict) of the reference's Markov encoder/decoder hot path, evaluated with CPU ATen
ops -- which is where the reference's own arithmetic lives (it has no native code) --
plus the C restatement in pointset_oracle.c (liboracle.so) for the index ops, whose binary32
evaluation order is fixed so that the CUDA kernels can be bit-identical to it.

Parity status: the reference ships no tests or golden vectors ("parity unpinned" by its own tests).
This oracle is pinned against outputs of the reference itself, executed in the build container by
tests/golden/make_golden.py (fixtures committed under tests/golden/, checked by
tests/test_oracle_golden.py).

R = Markov_Process_Analysis_on_Point_Cloud/ in the reference tree; every function cites the lines it
follows.
"""
import ctypes
import math
import os
import subprocess

import numpy as np
import torch
import torch.nn.functional as F

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force=False):
    """Compile liboracle.so with the committed Makefile (gcc).  Building the checker is not using it."""
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "pointset_oracle.c")
    if force or (not os.path.exists(so)) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "liboracle.so"])
    return so


def _lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build())
    return _LIB


def _p(t):
    return ctypes.c_void_p(t.data_ptr())


def _f32(t):
    return t.detach().to(torch.float32).contiguous()


# --------------------------------------------------------------------------------------------------
# index ops (C restatement)
# --------------------------------------------------------------------------------------------------
def draw_fps_start(B, N):
    """The reference draws the FPS start on the CPU default generator, R/modules/pointnet2_utils.py:96."""
    return torch.randint(0, N, (B,), dtype=torch.long)


def farthest_point_sample(xyz, npoint, start=None):
    """R/modules/pointnet2_utils.py:84-109.  xyz [B,N,C] -> int64 [B,npoint]."""
    xyz = _f32(xyz)
    B, N, C = xyz.shape
    if start is None:
        start = draw_fps_start(B, N)
    start = start.to(torch.int64).contiguous()
    out = torch.empty(B, npoint, dtype=torch.int64)
    rc = _lib().orc_fps_f32(_p(xyz), _p(start), _p(out), ctypes.c_int64(B), ctypes.c_int64(N),
                            ctypes.c_int64(C), ctypes.c_int64(npoint))
    assert rc == 0
    return out


def knn_point(nsample, xyz, new_xyz):
    """R/modules/pointnet2_utils.py:211-222 (argument order: k, reference set, queries).
    -> (dist fp32 [B,S,K] ascending, idx int64 [B,S,K]); equal distances in ascending index."""
    xyz, new_xyz = _f32(xyz), _f32(new_xyz)
    B, N, C = xyz.shape
    S = new_xyz.shape[1]
    dist = torch.empty(B, S, nsample, dtype=torch.float32)
    idx = torch.empty(B, S, nsample, dtype=torch.int64)
    rc = _lib().orc_knn_f32(_p(xyz), _p(new_xyz), _p(dist), _p(idx), ctypes.c_int64(B),
                            ctypes.c_int64(N), ctypes.c_int64(S), ctypes.c_int64(C),
                            ctypes.c_int64(nsample))
    assert rc == 0, "orc_knn_f32 rejected the arguments (K must be <= N)"
    return dist, idx


def square_distance_torch(src, dst):
    """R/modules/pointnet2_utils.py:190-209 with the reference's own ATen op chain (matmul, then the two
    in-place broadcast adds in that order)."""
    B, N, _ = src.shape
    M = dst.shape[1]
    dist = -2 * torch.matmul(src, dst.permute(0, 2, 1))
    dist += torch.sum(src ** 2, -1).view(B, N, 1)
    dist += torch.sum(dst ** 2, -1).view(B, 1, M)
    return dist


def knn_point_torch(nsample, xyz, new_xyz):
    """knn_point through ATen's topk exactly as the reference calls it (:220-221); its order among
    exactly equal distances is implementation defined."""
    d = square_distance_torch(new_xyz, xyz)
    return torch.topk(d, nsample, dim=-1, largest=False, sorted=True)


def query_ball_point(radius, nsample, xyz, new_xyz):
    """R/modules/pointnet2_utils.py:112-134."""
    xyz, new_xyz = _f32(xyz), _f32(new_xyz)
    B, N, C = xyz.shape
    S = new_xyz.shape[1]
    out = torch.empty(B, S, nsample, dtype=torch.int64)
    r2 = float(np.float32(radius ** 2))
    rc = _lib().orc_ball_query_f32(_p(xyz), _p(new_xyz), _p(out), ctypes.c_float(r2),
                                   ctypes.c_int64(B), ctypes.c_int64(N), ctypes.c_int64(S),
                                   ctypes.c_int64(C), ctypes.c_int64(nsample))
    assert rc == 0
    return out


def three_nn(xyz1, xyz2):
    """R/modules/pointnet2_utils.py:899-901: the 3 smallest expanded-form distances from each xyz1
    point to the xyz2 set, ascending.  -> (dist [B,N,3], idx [B,N,3])."""
    return knn_point(3, xyz2, xyz1)


# --------------------------------------------------------------------------------------------------
# float ops (differentiable through CPU autograd)
# --------------------------------------------------------------------------------------------------
def index_points(points, idx):
    """R/modules/pointnet2_utils.py:64-81: batched row gather; idx [B,S] or [B,S,K]."""
    B = points.shape[0]
    bidx = torch.arange(B, dtype=torch.long).view([B] + [1] * (idx.dim() - 1)).expand_as(idx)
    return points[bidx, idx.long()]


class _TransitionC(torch.autograd.Function):
    @staticmethod
    def forward(ctx, points, idx, n_out):
        points = _f32(points)
        idx = idx.to(torch.int64).contiguous()
        B, S, C = points.shape
        K = idx.shape[2]
        out = torch.empty(B, n_out, C, dtype=torch.float32)
        cnt = torch.empty(B, n_out, dtype=torch.float32)
        rc = _lib().orc_transition_fwd_f32(_p(points), _p(idx), _p(out), _p(cnt), ctypes.c_int64(B),
                                           ctypes.c_int64(S), ctypes.c_int64(K), ctypes.c_int64(C),
                                           ctypes.c_int64(n_out))
        assert rc == 0
        ctx.save_for_backward(idx, cnt)
        ctx.dims = (B, S, K, C, n_out)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        idx, cnt = ctx.saved_tensors
        B, S, K, C, n_out = ctx.dims
        grad_out = _f32(grad_out)
        g = torch.empty(B, S, C, dtype=torch.float32)
        rc = _lib().orc_transition_bwd_f32(_p(grad_out), _p(idx), _p(cnt), _p(g), ctypes.c_int64(B),
                                           ctypes.c_int64(S), ctypes.c_int64(K), ctypes.c_int64(C),
                                           ctypes.c_int64(n_out))
        assert rc == 0
        return g, None, None


def upsample_torch(points, knn_idx, n_out):
    """The same sparse restatement in ATen ops of any float dtype (the float64 "exact" runs that the full-size
    gradient tests measure both fp32 implementations against): out = D^-1 A^T points, D counting the sources
    whose channel-0 value is non-zero (:44), zero counts replaced by one (:45-46)."""
    B, S, C = points.shape
    K = knn_idx.shape[2]
    flat = (knn_idx + n_out * torch.arange(B).view(B, 1, 1)).reshape(-1)
    src = points.unsqueeze(2).expand(B, S, K, C).reshape(-1, C)
    out = torch.zeros(B * n_out, C, dtype=points.dtype).index_add(0, flat, src)
    cnt = torch.zeros(B * n_out, dtype=points.dtype).index_add(
        0, flat, (src[:, 0] != 0).to(points.dtype).detach())
    cnt = torch.where(cnt == 0, torch.ones_like(cnt), cnt)
    return (out / cnt.unsqueeze(1)).view(B, n_out, C)


def upsample(points, knn_idx, scale_ratio=2, dist=None, n_out=None):
    """The Markov state transition, R/modules/pointnet2_utils.py:13-50, restated sparsely (see
    pointset_oracle.c).  `dist` is accepted and ignored like the reference does (:30-34)."""
    S = points.shape[1]
    if n_out is None:
        n_out = S * scale_ratio
    if int(knn_idx.max()) >= n_out or int(knn_idx.min()) < 0:
        raise RuntimeError("index out of range in transition")  # ATen's scatter_ raises here too
    if points.dtype != torch.float32:
        return upsample_torch(points, knn_idx, int(n_out))
    return _TransitionC.apply(points, knn_idx, int(n_out))


def three_interpolate(points2, dist, idx):
    """R/modules/pointnet2_utils.py:903-906: inverse-distance weights 1/(d+1e-8), normalised, weighted
    sum of the three gathered rows."""
    recip = 1.0 / (dist + 1e-8)
    weight = recip / torch.sum(recip, dim=2, keepdim=True)
    B, N, _ = idx.shape
    return torch.sum(index_points(points2, idx) * weight.view(B, N, 3, 1), dim=2)


# --------------------------------------------------------------------------------------------------
# blocks.  P = dict of tensors keyed like the reference's state_dict; `pre` = key prefix ("a.b.").
# --------------------------------------------------------------------------------------------------
class Ctx:
    """Evaluation context: train/eval, which kNN implementation ranks neighbours, and the index tape
    (every FPS / kNN result in call order) used for index injection into the CUDA path."""

    def __init__(self, train=False, knn_impl="c", fps_starts=None, eval_blocks=(), inject=None):
        self.train = train
        # index tensors returned (in call order) instead of searching / sampling: a float64 run of the same network
        # on the fp32 run's neighbourhoods (tests/test_gpu_fullsize.py)
        self.inject = list(inject) if inject is not None else None
        self.eval_blocks = tuple(eval_blocks)  # Linear blocks (key prefixes) whose BatchNorm uses running statistics
        self.knn_impl = knn_impl
        self.fps_starts = list(fps_starts) if fps_starts is not None else None
        self.tape = []

    def knn(self, k, xyz, new_xyz, space="xyz"):
        """space: "xyz" (coordinates; tie-free on real clouds) or "feat" (feature space, where the
        transition leaves many points with IDENTICAL features and the reference's topk order among the
        exact ties is arbitrary -- taped as "knnf" so tests can audit instead of demanding equality)."""
        if self.inject is not None:
            i = self.inject.pop(0)
            self.tape.append(("knn" if space == "xyz" else "knnf", i))
            return None, i
        with torch.no_grad():
            if self.knn_impl == "torch":
                d, i = knn_point_torch(k, xyz.detach(), new_xyz.detach())
            else:
                d, i = knn_point(k, xyz, new_xyz)
        self.tape.append(("knn" if space == "xyz" else "knnf", i))
        return d, i

    def fps(self, xyz, npoint):
        B, N, _ = xyz.shape
        start = self.fps_starts.pop(0) if self.fps_starts is not None else draw_fps_start(B, N)
        i = self.inject.pop(0) if self.inject is not None else farthest_point_sample(xyz, npoint, start)
        self.tape.append(("fps", i))
        return i


def linear_block(P, pre, x, ctx, act=True):
    """`Linear(in, out, bn=False, act)`, R/modules/pointnet2_utils.py:401-425: nn.Linear, then -- the
    flag name is inverted, every call site passes bn=False -- BatchNorm1d over the channel axis with
    all leading axes as samples (:420), then LeakyReLU(0.2) if act."""
    y = F.linear(x, P[pre + "linear.weight"], P[pre + "linear.bias"])
    shp = y.shape
    y2 = y.reshape(-1, shp[-1])
    train = ctx.train and pre not in ctx.eval_blocks
    y2 = F.batch_norm(y2, P[pre + "norm2.running_mean"], P[pre + "norm2.running_var"],
                      P[pre + "norm2.weight"], P[pre + "norm2.bias"], training=train,
                      momentum=0.1, eps=1e-5)
    if train and (pre + "norm2.num_batches_tracked") in P:
        P[pre + "norm2.num_batches_tracked"] += 1
    y = y2.reshape(shp)
    if act:
        y = F.leaky_relu(y, 0.2)
    return y


def _attention_core(q, k, v):
    """R/modules/pointnet2_utils.py:532-544 (== :558-569): energy = q - k; per-channel softmax over the
    K neighbours of energy / sqrt(C); minus its own sum over K (:541-542); times v; max over K."""
    energy = q.unsqueeze(-2) - k
    attention = F.softmax(energy / np.sqrt(k.size(-1)), dim=-2)
    attention = attention - torch.sum(attention, dim=2, keepdim=True)
    return torch.max(attention * v, 2)[0]


def local_trans(P, pre, features, idx, ctx, FPS_idx=None, xyz=False, residual=False):
    """LocalTrans.forward, R/modules/pointnet2_utils.py:499-574 (== R/modules/repsurface_utils.py:470-540)."""
    center = index_points(features, FPS_idx) if FPS_idx is not None else features
    res = linear_block(P, pre + "conv_res.", center, ctx) if residual else center
    q = F.linear(center, P[pre + "q.weight"], P[pre + "q.bias"])
    if xyz:  # keys / values from (neighbour - centre) differences (:522-529)
        diff = index_points(features, idx) - center.unsqueeze(-2)
        k = F.linear(diff, P[pre + "k.weight"], P[pre + "k.bias"])
        v = F.linear(diff, P[pre + "v.weight"], P[pre + "v.bias"])
    else:  # keys / values projected first, then gathered (:552-556)
        k = index_points(F.linear(features, P[pre + "k.weight"], P[pre + "k.bias"]), idx)
        v = index_points(F.linear(features, P[pre + "v.weight"], P[pre + "v.bias"]), idx)
    context = _attention_core(q, k, v)
    return res + linear_block(P, pre + "ffn.", context, ctx)


def local_merge(P, pre, xyz, base_xyz, ctx, knn, residual, variant, normal=None, feature=None,
                FPS_idx=None):
    """LocalMerge.forward.  variant "seg": R/modules/pointnet2_utils.py:444-477 (three branches);
    variant "cls": R/modules/repsurface_utils.py:422-446 (two branches).  xyz_Trans always has
    residual=True (:439); feature branches use `residual`."""
    dist, idx = ctx.knn(knn, base_xyz, xyz)
    idx_feature = None
    if feature is not None:
        fs = index_points(feature, FPS_idx) if FPS_idx is not None else feature
        _, idx_feature = ctx.knn(knn, feature, fs, space="feat")
    if feature is None:
        out = local_trans(P, pre + "xyz_Trans.", xyz, idx, ctx, FPS_idx=FPS_idx, xyz=True, residual=True)
    elif variant == "seg":
        a = local_trans(P, pre + "xyz_Trans.", base_xyz, idx, ctx, FPS_idx=FPS_idx, xyz=True, residual=True)
        b = local_trans(P, pre + "feature_Trans1.", feature, idx, ctx, FPS_idx=FPS_idx, residual=residual)
        c = local_trans(P, pre + "feature_Trans2.", feature, idx_feature, ctx, FPS_idx=FPS_idx, residual=residual)
        out = linear_block(P, pre + "fc2.", torch.cat((a, b, c), dim=2), ctx)
    else:
        b = local_trans(P, pre + "feature_Trans.", feature, idx, ctx, FPS_idx=FPS_idx, residual=residual)
        c = local_trans(P, pre + "feature_Trans2.", feature, idx_feature, ctx, FPS_idx=FPS_idx, residual=residual)
        out = linear_block(P, pre + "fc2.", torch.cat((b, c), dim=2), ctx)
    if variant == "seg" and FPS_idx is not None and normal is not None:
        normal = index_points(normal, FPS_idx)  # :472-473 (carried, never consumed)
    return out, normal, idx, dist


def fuse(P, pre, t, f, fps, knn_enc, xyzs, ctx, knn=8):
    """Fuse.forward, R/modules/pointnet2_utils.py:612-709, for target state t (the reference selects it
    by the literal point count 2048/1024/512/256/128 = state 0..4).  f: 5 state features; fps[j]: FPS
    indices taking state j+1 points into state j; knn_enc[j]: encoder kNN of state j points inside state
    j-1; xyzs: 5 state coordinates.  Finer states come down through composed FPS indices (:617-632),
    coarser ones come up through the transition (`upsample`), adjacent states reusing the encoder's kNN
    and non-adjacent ones a fresh coordinate kNN (:667,681,685,696,700,704)."""
    acc = f[t]
    for j in range(5):
        if j == t:
            continue
        if j < t:
            comp = fps[t - 1]
            for m in range(t - 2, j - 1, -1):
                comp = index_points(fps[m].unsqueeze(-1), comp).squeeze(-1)
            src = index_points(f[j], comp)
        elif j == t + 1:
            src = upsample(f[j], knn_enc[j], n_out=f[t].shape[1])
        else:
            _, kidx = ctx.knn(knn, xyzs[t], xyzs[j])
            src = upsample(f[j], kidx, n_out=f[t].shape[1])
        acc = acc + linear_block(P, pre + "conv%d%d." % (j, t), src, ctx)
    return linear_block(P, pre + "conv%d." % t, acc, ctx) + f[t]


def stage_sizes(N):
    """State sizes of the part-seg encoder: the reference hard-codes 2048/1024/512/256/128
    (R/modules/pointnet2_utils.py:768-787); generalised to (N, N/2, N/4, N/8, N/16)."""
    return [N, N // 2, N // 4, N // 8, N // 16]


def keep_high_partseg(P, pre, xyz, normal, label, ctx):
    """KeepHighResolutionModulePartSeg.forward, R/modules/pointnet2_utils.py:758-858."""
    xyz = xyz.permute(0, 2, 1).contiguous()
    normal = normal.permute(0, 2, 1).contiguous()
    N = xyz.shape[1]
    n = stage_sizes(N)
    lm = lambda name, res, **kw: local_merge(P, pre + name + ".", ctx=ctx, knn=8, residual=res, variant="seg", **kw)
    # encoder (:765-791)
    e0, nrm0, knn0, _ = lm("la0", True, xyz=xyz, base_xyz=xyz, normal=normal)
    F0 = ctx.fps(xyz, n[1]); x1 = index_points(xyz, F0)
    e1, nrm1, knn1, _ = lm("la1", False, xyz=x1, base_xyz=xyz, normal=nrm0, feature=e0, FPS_idx=F0)
    F1 = ctx.fps(x1, n[2]); x2 = index_points(x1, F1)
    e2, nrm2, knn2, _ = lm("la2", False, xyz=x2, base_xyz=x1, normal=nrm1, feature=e1, FPS_idx=F1)
    F2 = ctx.fps(x2, n[3]); x3 = index_points(x2, F2)
    e3, nrm3, knn3, _ = lm("la3", True, xyz=x3, base_xyz=x2, normal=nrm2, feature=e2, FPS_idx=F2)
    F3 = ctx.fps(x3, n[4]); x4 = index_points(x3, F3)
    e4, nrm4, knn4, _ = lm("la4", True, xyz=x4, base_xyz=x3, normal=nrm3, feature=e3, FPS_idx=F3)
    fps = [F0, F1, F2, F3]
    knn_enc = [knn0, knn1, knn2, knn3, knn4]
    xyzs = [xyz, x1, x2, x3, x4]
    fz = lambda name, t, f: fuse(P, pre + name + ".", t, f, fps, knn_enc, xyzs, ctx)
    # decoder (:795-840)
    d4 = linear_block(P, pre + "mlp.", e4, ctx)
    d4 = fz("fuse1", 4, [e0, e1, e2, e3, d4])
    d3, _, _, _ = lm("la4_up", False, xyz=x3, base_xyz=x3, normal=nrm3,
                     feature=linear_block(P, pre + "up_conv4.", upsample(d4, knn4, n_out=n[3]), ctx))
    d3 = fz("fuse2", 3, [e0, e1, e2, d3, e4])
    d2, _, _, _ = lm("la3_up", False, xyz=x2, base_xyz=x2, normal=nrm2,
                     feature=linear_block(P, pre + "up_conv3.", upsample(d3, knn3, n_out=n[2]), ctx))
    d2 = fz("fuse3", 2, [e0, e1, d2, e3, e4])
    d1, _, _, _ = lm("la2_up", False, xyz=x1, base_xyz=x1, normal=nrm1,
                     feature=linear_block(P, pre + "up_conv2.", upsample(d2, knn2, n_out=n[1]), ctx))
    d1 = fz("fuse4", 1, [e0, d1, e2, e3, e4])
    d0, _, _, _ = lm("la1_up", False, xyz=xyz, base_xyz=xyz, normal=nrm0,
                     feature=linear_block(P, pre + "up_conv1.", upsample(d1, knn1, n_out=n[0]), ctx))
    d0 = fz("fuse5", 0, [d0, e1, e2, e3, e4])
    # head input (:843-853)
    glob = torch.cat([t.max(dim=1, keepdim=True)[0] for t in (d0, d1, d2, d3, d4)], dim=2)
    glob = glob.repeat(1, N, 1)
    lab = linear_block(P, pre + "conv7.", label, ctx).repeat(1, N, 1)
    head = linear_block(P, pre + "conv5.", d0, ctx)
    return xyz, torch.cat((head, glob, lab), 2)


def partseg_model(P, xyz, cls_label, ctx):
    """get_model.forward, R/models/repsurf/pointnet2_part_seg_msg.py:135-156 (dropout is identity in
    eval(); parity runs in train() set p=0, SURVEY.md 8a RNG note)."""
    _, final = keep_high_partseg(P, "keepHigh.", xyz, xyz, cls_label, ctx)
    x = linear_block(P, "conv8.", final, ctx)
    x = linear_block(P, "conv9.", x, ctx)
    x = linear_block(P, "conv10.", x, ctx)
    return F.linear(x, P["conv11.weight"], P["conv11.bias"])


def partseg_loss(pred, target):
    """get_loss.forward, R/models/repsurf/pointnet2_part_seg_msg.py:159-180: label-smoothed CE, eps 0.1."""
    target = target.contiguous().view(-1)
    n_class = pred.size(1)
    one_hot = torch.zeros_like(pred).scatter(1, target.view(-1, 1), 1)
    one_hot = one_hot * 0.9 + (1 - one_hot) * 0.1 / (n_class - 1)
    return -(one_hot * F.log_softmax(pred, dim=1)).sum(dim=1).mean()


def keep_high_cls(P, pre, xyz, normal, ctx):
    """KeepHighResolutionModule.forward, R/modules/repsurface_utils.py:572-639.  The reference samples
    to the literals 512/256/128/64/32 (:581-619)."""
    xyz = xyz.permute(0, 2, 1).contiguous()
    lm = lambda name, res, **kw: local_merge(P, pre + name + ".", ctx=ctx, knn=8, residual=res, variant="cls", **kw)
    feat, _, _, _ = lm("la0", True, xyz=xyz, base_xyz=xyz)
    base = xyz
    for name, res, npoint in (("la1", False, 512), ("la2", False, 256), ("la3", True, 128),
                              ("la4", True, 64), ("la5", True, 32)):
        Fi = ctx.fps(base, npoint)
        sub = index_points(base, Fi)
        feat, _, _, _ = lm(name, res, xyz=sub, base_xyz=base, feature=feat, FPS_idx=Fi)
        base = sub
    final = linear_block(P, pre + "conv4.", linear_block(P, pre + "conv3.", feat, ctx), ctx)
    x = torch.cat((final.max(dim=1)[0], final.mean(dim=1)), 1)  # adaptive max / avg pool (:632-634)
    x = F.linear(x, P[pre + "final_class.weight"], P[pre + "final_class.bias"])
    x = _bn_plain(P, pre + "bn.", x, ctx)
    return F.leaky_relu(x, 0.2)


def _bn_plain(P, pre, x, ctx):
    y = F.batch_norm(x, P[pre + "running_mean"], P[pre + "running_var"], P[pre + "weight"],
                     P[pre + "bias"], training=ctx.train, momentum=0.1, eps=1e-5)
    if ctx.train and (pre + "num_batches_tracked") in P:
        P[pre + "num_batches_tracked"] += 1
    return y


def cls_model(P, points, ctx):
    """Model.forward, R/models/repsurf/repsurf_ssg_umb.py:55-70 (dropout identity / p=0)."""
    center = points[:, :3, :]
    x = keep_high_cls(P, "keepHigh.", center, center, ctx)
    x = F.leaky_relu(_bn_plain(P, "bn1.", F.linear(x, P["fc1.weight"], P["fc1.bias"]), ctx), 0.2)
    x = F.leaky_relu(_bn_plain(P, "bn2.", F.linear(x, P["fc2.weight"], P["fc2.bias"]), ctx), 0.2)
    x = F.linear(x, P["fc3.weight"], P["fc3.bias"])
    return F.log_softmax(x, -1)


def smooth_cls_loss(pred, target, eps=0.1):
    """SmoothClsLoss.forward, R/util/utils.py:74-88 (pred is already log-softmax)."""
    n_class = pred.size(1)
    one_hot = torch.zeros_like(pred).scatter(1, target.view(-1, 1), 1)
    one_hot = one_hot * (1 - eps) + (1 - one_hot) * eps / (n_class - 1)
    return -(one_hot * pred).sum(dim=1).mean()


def feature_propagation(P, pre, xyz1, xyz2, points1, points2, ctx, act=False):
    """PointNetFeaturePropagation.forward, R/modules/pointnet2_utils.py:877-912 (three_nn +
    three_interpolate + Linear).  points1 is unused by the reference."""
    B, N, _ = xyz1.shape
    S = xyz2.shape[1]
    if S == 1:
        interp = points2.repeat(1, N, 1)
    else:
        with torch.no_grad():
            d, i = three_nn(xyz1, xyz2)
        ctx.tape.append(("three_nn", i))
        interp = three_interpolate(points2, d, i)
    return linear_block(P, pre + "conv.", interp, ctx, act=act)


# --------------------------------------------------------------------------------------------------
# umbrella surface features (RepSurf), R/modules/pointnet2_utils.py:310-399
# --------------------------------------------------------------------------------------------------
def sphere_coords(v):
    """xyz2sphere with normalize=True, R/modules/polar_utils.py:10-31: (rho, theta/pi, phi/(2 pi) + 0.5);
    theta = 0 where rho = 0."""
    rho = torch.sqrt((v * v).sum(-1, keepdim=True)).clamp(min=0)
    theta = torch.acos(v[..., 2:3] / rho)
    theta = torch.where(rho == 0, torch.zeros_like(theta), theta)
    phi = torch.atan2(v[..., 1:2], v[..., 0:1])
    return torch.cat([rho, theta / math.pi, phi / (2 * math.pi) + 0.5], -1)


def umbrella_features(center, k=9, return_dist=True, sign=None):
    """The per-point umbrella feature the reference feeds to UmbrellaSurfaceConstructor.mlps
    (R/modules/pointnet2_utils.py:360-378): for every point, its k-1 nearest neighbours (self excluded) relative to
    the point, sorted counter-clockwise by azimuth (:310-334), consecutive pairs (s_g, s_{g+1}) with the point itself
    (the origin) form G = k-1 triangles; per triangle: centroid (recons_utils.py:cal_center), its spherical
    coordinates, unit normal with the first triangle's x-component made positive (cal_normal, is_group=True), optional
    per-cloud sign flip `sign` [B] (random_inv, drawn by the caller like recons_utils.py:50), plane constant
    normal.centroid / sqrt(3) (cal_const); triangles with a NaN normal take normal / centroid / constant of the first
    valid triangle of the same point (check_nan_umb; the polar channels keep their own values).
    center [B,N,3] -> [B,N,G,10] (or 9 without the constant): centroid, polar, normal, constant."""
    B, N, _ = center.shape
    _, idx = knn_point(k, center, center)
    rel = index_points(center, idx)[:, :, 1:] - center.unsqueeze(2)  # [B,N,G,3]
    phi = sphere_coords(rel)[..., 2]
    order = torch.argsort(phi, dim=-1, stable=True)
    srt = torch.gather(rel, 2, order.unsqueeze(-1).expand(-1, -1, -1, 3))
    nxt = torch.roll(srt, -1, dims=2)
    nor = torch.cross(srt, nxt, dim=-1)  # (s_g - 0) x (s_{g+1} - 0)
    unit = nor / torch.norm(nor, dim=-1, keepdim=True)
    pos = (unit[:, :, 0:1, 0] > 0).float() * 2.0 - 1.0
    unit = unit * pos.unsqueeze(-1)
    if sign is not None:
        unit = unit * sign.view(B, 1, 1, 1).to(unit.dtype)
    cen = (torch.zeros_like(srt) + srt + nxt) / 3.0
    polar = sphere_coords(cen)
    const = (unit * cen).sum(-1, keepdim=True) / math.sqrt(3.0)
    bad = torch.isnan(unit).any(-1)  # [B,N,G]
    first = torch.argmax((~bad).int(), dim=-1)  # first valid triangle (0 when none is)
    pick = first.view(B, N, 1, 1)
    unit = torch.where(bad.unsqueeze(-1), torch.gather(unit, 2, pick.expand(-1, -1, 1, 3)).expand_as(unit), unit)
    cen = torch.where(bad.unsqueeze(-1), torch.gather(cen, 2, pick.expand(-1, -1, 1, 3)).expand_as(cen), cen)
    const = torch.where(bad.unsqueeze(-1), torch.gather(const, 2, pick).expand_as(const), const)
    parts = [cen, polar, unit] + ([const] if return_dist else [])
    return torch.cat(parts, -1)


def umbrella_constructor(P, pre, center_b3n, ctx, k=9, aggr="sum", sign=None):
    """UmbrellaSurfaceConstructor.forward, R/modules/pointnet2_utils.py:360-399 (return_dist=True): umbrella features
    -> 1x1 Conv2d/BatchNorm2d/ReLU x2 -> Conv2d -> aggregate over the G triangles.  center [B,3,N] -> [B,10,N]."""
    feat = umbrella_features(center_b3n.permute(0, 2, 1).contiguous(), k=k, return_dist=True, sign=sign)
    x = feat.permute(0, 3, 2, 1)  # [B,C,G,N]
    for conv, bn in (("0", "1"), ("3", "4")):
        x = F.conv2d(x, P[pre + "mlps.%s.weight" % conv], P.get(pre + "mlps.%s.bias" % conv))
        x = F.batch_norm(x, P[pre + "mlps.%s.running_mean" % bn], P[pre + "mlps.%s.running_var" % bn],
                         P[pre + "mlps.%s.weight" % bn], P[pre + "mlps.%s.bias" % bn], ctx.train, 0.1, 1e-5)
        x = F.relu(x)
    x = F.conv2d(x, P[pre + "mlps.6.weight"], P[pre + "mlps.6.bias"])
    if aggr == "max":
        return x.max(2)[0]
    if aggr == "avg":
        return x.mean(2)
    return x.sum(2)


# --------------------------------------------------------------------------------------------------
# RepSurf set abstraction, R/modules/repsurface_utils.py:12-84, 206-319
# --------------------------------------------------------------------------------------------------
def sample_and_group(npoint, radius, nsample, center, normal, feature, ctx, return_normal=True, return_polar=False):
    """R/modules/repsurface_utils.py:12-59 (channel-last tensors)."""
    fps_idx = ctx.fps(center, npoint)
    new_center = index_points(center, fps_idx)
    new_normal = index_points(normal, fps_idx)
    idx = query_ball_point(radius, nsample, center, new_center)
    ctx.tape.append(("ball", idx))
    group_normal = index_points(normal, idx)
    rel = index_points(center, idx) - new_center.unsqueeze(2)
    if return_polar:
        rel = torch.cat([rel, sphere_coords(rel)], -1)
    parts = [rel] + ([group_normal] if (return_normal or feature is None) else [])
    if feature is not None:
        parts.append(index_points(feature, idx))
    return new_center, new_normal, torch.cat(parts, -1)


def _conv_bn_relu(P, conv, bn, x, ctx, relu=True):
    x = F.conv2d(x, P[conv + ".weight"], P[conv + ".bias"])
    x = F.batch_norm(x, P[bn + ".running_mean"], P[bn + ".running_var"], P[bn + ".weight"], P[bn + ".bias"],
                     ctx.train, 0.1, 1e-5)
    return F.relu(x) if relu else x


def surface_abstraction(P, pre, center, normal, feature, ctx, npoint, radius, nsample, n_layers, return_polar=True,
                        return_normal=True, pos_channel=None):
    """SurfaceAbstraction.forward (:229-254) or, with pos_channel, SurfaceAbstractionCD.forward (:287-319); group_all
    = False.  Channel-first inputs [B,C,N] like the reference; returns (new_center [B,3,S], new_normal, new_feature)."""
    c, n = center.permute(0, 2, 1), normal.permute(0, 2, 1)
    f = feature.permute(0, 2, 1) if feature is not None else None
    new_center, new_normal, g = sample_and_group(npoint, radius, nsample, c, n, f, ctx, return_normal, return_polar)
    x = g.permute(0, 3, 2, 1)  # [B,C,K,S]
    if pos_channel is not None:
        loc = _conv_bn_relu(P, pre + "mlp_l0", pre + "bn_l0", x[:, :pos_channel], ctx, relu=False)
        feat = _conv_bn_relu(P, pre + "mlp_f0", pre + "bn_f0", x[:, pos_channel:], ctx, relu=False)
        x = F.relu(loc + feat)
    for i in range(n_layers):
        x = _conv_bn_relu(P, pre + "mlp_convs.%d" % i, pre + "mlp_bns.%d" % i, x, ctx)
    return new_center.permute(0, 2, 1), new_normal.permute(0, 2, 1), x.max(2)[0]


# --------------------------------------------------------------------------------------------------
# deterministic synthetic parameters: a function of (key, shape) only, so the reference, the oracle
# and the CUDA modules can all be loaded with identical weights without shipping a checkpoint.
# --------------------------------------------------------------------------------------------------
def synthetic_state_dict(spec, seed=0):
    """spec: iterable of (key, shape, dtype-string).  Linear weights U(-1/sqrt(fan_in), +), biases
    U(-0.1, 0.1), norm weights U(0.75, 1.25), running_mean U(-0.1, 0.1), running_var U(0.75, 1.25),
    num_batches_tracked 0."""
    import zlib

    out = {}
    for key, shape, dtype in spec:
        g = torch.Generator().manual_seed((zlib.crc32(key.encode()) ^ (seed * 2654435761)) & 0x7FFFFFFF)
        shape = tuple(shape)
        if key.endswith("num_batches_tracked"):
            t = torch.zeros(shape, dtype=torch.int64)
        elif key.endswith("running_var") or (key.endswith(".weight") and len(shape) == 1):
            t = torch.rand(shape, generator=g) * 0.5 + 0.75
        elif key.endswith("running_mean") or key.endswith(".bias"):
            t = torch.rand(shape, generator=g) * 0.2 - 0.1
        else:
            bound = 1.0 / math.sqrt(shape[-1]) if len(shape) > 1 else 0.1
            t = (torch.rand(shape, generator=g) * 2 - 1) * bound
        assert str(t.dtype).replace("torch.", "") == dtype, (key, t.dtype, dtype)
        out[key] = t
    return out
