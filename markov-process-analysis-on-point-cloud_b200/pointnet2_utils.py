"""Drop-in mirror of R/modules/pointnet2_utils.py (R = Markov_Process_Analysis_on_Point_Cloud/ in the
reference tree): the same free functions and nn.Module classes, same constructor / forward signatures and the
same state_dict keys and shapes (2189 keys for the part-seg model), running on the sm_100a kernels.

What changed underneath, not at the interface:
  * `Linear` applies BatchNorm1d + LeakyReLU on the [M,C] view in one fused kernel pair (no permute copies);
  * `LocalTrans` never materialises a [B,S,K,C] tensor: keys and values are projected once per point (one fused
    GEMM) and gathered inside the attention kernel; the coordinate branch projects the 3-channel differences
    inside the kernel;
  * `upsample` (the Markov state transition) is the sparse D^-1 A^T X product instead of a dense [B,S,N,C] scatter;
  * `Fuse` / `KeepHighResolutionModulePartSeg` take their state sizes from the input (N, N/2, N/4, N/8, N/16)
    instead of the literals 2048/1024/512/256/128 (R/modules/pointnet2_utils.py:615-707,768-787); at N = 2048
    they are the reference's network.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops
from ._lib import on_tensor_device
from .ops import (farthest_point_sample, index_points, knn_point, query_ball_point, query_knn_point,  # noqa: F401
                  sample, square_distance, three_interpolate, three_nn, upsample)


class Linear(nn.Module):
    """R/modules/pointnet2_utils.py:401-425.  nn.Linear -> norm -> LeakyReLU(0.2) if act.  The flag name is
    inverted in the reference and kept: bn=True selects LayerNorm (norm1), bn=False selects BatchNorm1d over the
    channel axis (norm2); every call site passes bn=False.  Both norms are always constructed (state_dict)."""

    def __init__(self, in_channels, out_channels, bn=True, act=True):
        super().__init__()
        self.act_flag = act
        self.bn_flag = bn
        self.linear = nn.Linear(in_channels, out_channels)
        self.norm1 = nn.LayerNorm(out_channels)
        self.norm2 = nn.BatchNorm1d(out_channels)
        self.act = nn.LeakyReLU(negative_slope=0.2)

    def forward(self, input, residual=None):
        """`residual` (extension, same shape as the output) is added to the result inside the normalise kernel;
        Linear(x) + r with one pass less.  The reference signature is forward(input)."""
        if self.bn_flag is True:  # LayerNorm branch: never taken by a shipped model
            y = self.norm1(ops.linear(input, self.linear.weight, self.linear.bias))
            y = self.act(y) if self.act_flag is True else y
            return y if residual is None else y + residual
        return ops.linear_bn_act(input, self.linear.weight, self.linear.bias, self.norm2, self.training,
                                 0.2 if self.act_flag is True else 1.0, residual=residual)

    def forward_split(self, x_a, g):
        """forward(cat(x_a [B,N,Ka], g [B,Kb] broadcast over the N points)) without building the concatenation
        (extension; ops.LinearBNActSplit).  BatchNorm branch only."""
        return ops.linear_bn_act_split(x_a, g, self.linear.weight, self.linear.bias, self.norm2, self.training,
                                       0.2 if self.act_flag is True else 1.0)


class LocalTrans(nn.Module):
    """Difference-wise attention, R/modules/pointnet2_utils.py:479-574."""

    def __init__(self, in_c, out_c, patch_num, usetanh=False, residual=False):
        super().__init__()
        self.patchNum = patch_num
        self.residual = residual
        self.usetanh = usetanh
        self.out_c = out_c
        self.q = nn.Linear(in_c, out_c)
        self.k = nn.Linear(in_c, out_c)
        self.v = nn.Linear(in_c, out_c)
        self.conv_res = Linear(in_c, out_c, bn=False)
        self.ffn = Linear(out_c, out_c, bn=False)
        self.tanh = nn.Tanh()

    @on_tensor_device
    def forward(self, features, idx, pos, FPS_idx=None, xyz=False):
        if self.usetanh is True:
            # the reference's tanh branch (:535-537) multiplies [B,S,K,C] by [B,S,K,C] with matmul, which
            # cannot run for K != C; no shipped configuration enables it.
            raise NotImplementedError("LocalTrans(usetanh=True) is dead code in the reference")
        if xyz is True:
            # coordinate branch: q/k/v projections of the differences AND the residual projection
            # conv_res.linear(centre) are computed inside one kernel; BatchNorm + LeakyReLU of conv_res follow
            fused_res = self.residual is True and self.conv_res.bn_flag is not True
            lin = self.conv_res.linear
            context, res = ops.AttnXyz.apply(features.float().contiguous(),
                                             FPS_idx.contiguous() if FPS_idx is not None else None,
                                             idx.contiguous(), self.q.weight, self.q.bias, self.k.weight, self.k.bias,
                                             self.v.weight, self.v.bias, lin.weight if fused_res else None,
                                             lin.bias if fused_res else None)
            if fused_res:
                n = self.conv_res.norm2
                shape = res.shape
                residual = ops.bn_act(res.view(-1, shape[-1]), n.weight, n.bias, n.running_mean, n.running_var,
                                      n.num_batches_tracked, self.training, momentum=ops._momentum(n), eps=n.eps,
                                      slope=0.2 if self.conv_res.act_flag is True else 1.0).view(shape)
            else:
                center = index_points(features, FPS_idx) if FPS_idx is not None else features
                residual = self.conv_res(center) if self.residual is True else center
            return self.ffn(context, residual=residual)
        center = index_points(features, FPS_idx) if FPS_idx is not None else features
        residual = self.conv_res(center) if self.residual is True else center
        if ops.bf16_active() and not self.training and center.shape[-1] % 64 == 0 and self.out_c % 8 == 0:
            context = ops.feat_attention_bf16(center, features, idx, self.q.weight, self.q.bias, self.k.weight,
                                              self.k.bias, self.v.weight, self.v.bias)
            return self.ffn(context, residual=residual)
        if ops.feat_attention_fusable(center, self.q.weight):
            context = ops.FeatAttention.apply(center.contiguous(), features.contiguous(), idx.contiguous(),
                                              self.q.weight, self.q.bias, self.k.weight, self.k.bias,
                                              self.v.weight, self.v.bias)
        else:
            q = ops.linear(center, self.q.weight, self.q.bias)
            kv = ops.linear(features, torch.cat((self.k.weight, self.v.weight), 0),
                            torch.cat((self.k.bias, self.v.bias), 0))
            context = ops.AttnFeat.apply(q.contiguous(), kv.contiguous(), idx.contiguous())
        return self.ffn(context, residual=residual)


class LocalMerge(nn.Module):
    """Encoder / decoder stage, R/modules/pointnet2_utils.py:427-477 (three-branch part-seg variant)."""

    def __init__(self, in_channels, out_channels, knn, usetanh=False, residual=False):
        super().__init__()
        self.knn = knn
        self.usetanh = usetanh
        self.residual = residual
        self.fc2 = Linear(out_channels * 3, out_channels, bn=False)
        self.xyz_Trans = LocalTrans(3, out_channels, knn, usetanh=self.usetanh, residual=True)
        self.normal_Trans = LocalTrans(10, out_channels, knn, usetanh=self.usetanh, residual=True)
        self.feature_Trans1 = LocalTrans(in_channels, out_channels, knn, usetanh=self.usetanh, residual=self.residual)
        self.feature_Trans2 = LocalTrans(in_channels, out_channels, knn, usetanh=self.usetanh, residual=self.residual)

    @on_tensor_device
    def forward(self, xyz, base_xyz, normal=None, feature=None, FPS_idx=None, xyz_flag=True):
        # coordinate-space neighbour search: geometry stream (it only depends on the cloud), joined here
        dist, idx = ops.geo_join(ops.geo_call(lambda: knn_point(self.knn, base_xyz, xyz)))
        if feature is None:
            merge_features = self.xyz_Trans(features=xyz, idx=idx, pos=base_xyz, FPS_idx=FPS_idx, xyz=True)
        else:
            # three independent branches (the feature-space kNN belongs to the third): run side by side
            def branch_xyz():
                return self.xyz_Trans(features=base_xyz, idx=idx, pos=base_xyz, FPS_idx=FPS_idx, xyz=True)

            def branch_f1():
                return self.feature_Trans1(features=feature, idx=idx, pos=base_xyz, FPS_idx=FPS_idx)

            def branch_f2():
                fs = index_points(feature, FPS_idx) if FPS_idx is not None else feature
                _, idx_feature = knn_point(self.knn, feature, fs)
                return self.feature_Trans2(features=feature, idx=idx_feature, pos=base_xyz, FPS_idx=FPS_idx)

            features2, features1, xyz_f = ops.parallel(branch_f2, branch_f1, branch_xyz)
            merge_features = self.fc2(torch.cat((xyz_f, features1, features2), dim=2))
        if FPS_idx is not None and normal is not None:
            normal = index_points(normal, FPS_idx)
        return merge_features, normal, idx, dist


class Fuse(nn.Module):
    """State-to-state transition fusion, R/modules/pointnet2_utils.py:576-709.  forward keeps the reference's
    signature; `num_point` selects the target state by its point count (2048/1024/512/256/128 in the reference;
    here: whichever of f0..f4 has that many points)."""

    def __init__(self, c0, c1, c2, c3, c4):
        super().__init__()
        self.knn = 8
        c = (c0, c1, c2, c3, c4)
        for t in (4, 3, 2, 1, 0):  # constructor order of the reference (:582-610)
            for j in range(5):
                if j != t:
                    setattr(self, "conv%d%d" % (j, t), Linear(c[j], c[t], bn=False))
            setattr(self, "conv%d" % t, Linear(c[t], c[t], bn=False))

    def forward(self, num_point, f0=None, f1=None, f2=None, f3=None, f4=None, FPS_0=None, FPS_1=None, FPS_2=None,
                FPS_3=None, knn_0=None, knn_1=None, knn_2=None, knn_3=None, knn_4=None, xyz0=None, xyz1=None,
                xyz2=None, xyz3=None, xyz4=None):
        f = [f0, f1, f2, f3, f4]
        fps = [FPS_0, FPS_1, FPS_2, FPS_3]
        knn_enc = [knn_0, knn_1, knn_2, knn_3, knn_4]
        xyzs = [xyz0, xyz1, xyz2, xyz3, xyz4]
        # the state sizes strictly decrease, so the point count identifies the state
        targets = [t for t in range(5) if f[t] is not None and f[t].shape[1] == num_point]
        if len(targets) != 1:
            return f0, f1, f2, f3, f4  # the reference's `if num_point == ...` chain falls through unchanged
        t = targets[0]
        n_t = f[t].shape[1]

        def source(j):
            def run():
                if j < t:  # finer state -> target through the composed FPS indices (:617-632)
                    comp = fps[t - 1]
                    for m in range(t - 2, j - 1, -1):
                        comp = index_points(fps[m].unsqueeze(-1), comp).squeeze(-1)
                    src = index_points(f[j], comp)
                elif j == t + 1:  # adjacent coarser state: reuse the encoder's kNN (:650,663,678,693)
                    src = upsample(f[j], knn_enc[j], n_out=n_t)
                else:  # non-adjacent coarser state: fresh coordinate kNN (:667,681,685,696,700,704)
                    _, kidx = ops.geo_join(ops.geo_call(lambda: knn_point(self.knn, xyzs[t], xyzs[j])))
                    src = upsample(f[j], kidx, n_out=n_t)
                return getattr(self, "conv%d%d" % (j, t))(src)
            return run

        # the four source states are independent of each other: transition / gather + projection side by side;
        # the sum keeps the reference's order f_t + f_0t + f_1t + ... (ascending source state)
        parts = ops.parallel(*[source(j) for j in range(5) if j != t])
        acc = f[t]
        for part in parts:
            acc = acc + part
        f[t] = getattr(self, "conv%d" % t)(acc, residual=f[t])
        return tuple(f)


class KeepHighResolutionModulePartSeg(nn.Module):
    """The 5-state Markov encoder / decoder of the part-seg model, R/modules/pointnet2_utils.py:711-858.
    forward(xyz [B,3,N], normal [B,3,N], label [B,1,16]) -> (xyz [B,N,3], final [B,N,896])."""

    def __init__(self, data_C, b1_C, b2_C, b3_C, b4_C, cuda=False):
        super().__init__()
        self.neighbour = 16
        self.cuda_ops = cuda  # the reference stores this as `self.cuda`, shadowing nn.Module.cuda()
        self.start = Linear(3, 32, bn=False)
        self.la0 = LocalMerge(32, 64, 8, usetanh=False, residual=True)
        self.la1 = LocalMerge(64, 64, 8, usetanh=False, residual=False)
        self.la2 = LocalMerge(64, 64, 8, usetanh=False, residual=False)
        self.la3 = LocalMerge(64, 128, 8, usetanh=False, residual=True)
        self.la4 = LocalMerge(128, 256, 8, usetanh=False, residual=True)
        self.la4_up = LocalMerge(128, 128, 8, usetanh=False, residual=False)
        self.la3_up = LocalMerge(64, 64, 8, usetanh=False, residual=False)
        self.la2_up = LocalMerge(64, 64, 8, usetanh=False, residual=False)
        self.la1_up = LocalMerge(64, 64, 8, usetanh=False, residual=False)
        self.up_conv4 = Linear(256, 128, bn=False)
        self.up_conv3 = Linear(128, 64, bn=False)
        self.up_conv2 = Linear(64, 64, bn=False)
        self.up_conv1 = Linear(64, 64, bn=False)
        self.mlp = Linear(256, 256, bn=False)
        self.conv5 = Linear(64, 256, bn=False)
        self.conv6 = Linear(64, 128, bn=False)
        self.conv7 = Linear(16, 64, bn=False)
        self.conv8 = Linear(64, 256, bn=False)
        self.fuse1 = Fuse(64, 64, 64, 128, 256)
        self.fuse2 = Fuse(64, 64, 64, 128, 256)
        self.fuse3 = Fuse(64, 64, 64, 128, 256)
        self.fuse4 = Fuse(64, 64, 64, 128, 256)
        self.fuse5 = Fuse(64, 64, 64, 128, 256)
        self.lrelu = nn.LeakyReLU(negative_slope=0.2)

    @on_tensor_device
    def forward(self, xyz, normal, label):
        xyz = xyz.permute(0, 2, 1).contiguous()
        normal = normal.permute(0, 2, 1).contiguous()
        with ops.geometry_scope():
            return self._forward(xyz, normal, label)

    @on_tensor_device
    def forward_parts(self, xyz, normal, label):
        """Same computation as forward, but the 896-channel head input is returned in its two parts instead of
        concatenated: (xyz [B,N,3], per-point channels [B,N,256], per-cloud channels [B,640] = global max pools + label
        embedding, which forward broadcasts over the N points).  cat(a, g[:, None].expand(-1, N, -1)) == forward()[1]."""
        xyz = xyz.permute(0, 2, 1).contiguous()
        normal = normal.permute(0, 2, 1).contiguous()
        with ops.geometry_scope():
            return self._forward(xyz, normal, label, parts=True)

    @staticmethod
    def _sample(points, npoint):
        """FPS + gather of the sampled coordinates, issued on the geometry stream (joined by the consumer's
        coordinate kNN, which follows it on that stream)."""
        return ops.geo_call(lambda: ops.fps_and_gather(points, npoint))

    def _forward(self, xyz, normal, label, parts=False):
        N = xyz.shape[1]
        n = [N, N // 2, N // 4, N // 8, N // 16]
        # the whole coordinate pyramid (4 sampling steps, 5 + 4 + 6 neighbour searches) only depends on xyz: issue
        # it now on the geometry lanes; the stages below pick their results up in the reference's call order
        ops.geo_prefetch_pyramid(xyz, n[1:], 8, self_levels=(0, 3, 2, 1),
                                 cross=((2, 4), (1, 3), (1, 4), (0, 2), (0, 3), (0, 4)))
        # encoder (:765-791)
        e0, nrm0, knn0, dist0 = self.la0(xyz=xyz, base_xyz=xyz, normal=normal, xyz_flag=True)
        F0, x1 = self._sample(xyz, n[1])
        e1, nrm1, knn1, dist1 = self.la1(xyz=x1, base_xyz=xyz, normal=nrm0, feature=e0, FPS_idx=F0, xyz_flag=True)
        F1, x2 = self._sample(x1, n[2])
        e2, nrm2, knn2, dist2 = self.la2(xyz=x2, base_xyz=x1, normal=nrm1, feature=e1, FPS_idx=F1, xyz_flag=False)
        F2, x3 = self._sample(x2, n[3])
        e3, nrm3, knn3, dist3 = self.la3(xyz=x3, base_xyz=x2, normal=nrm2, feature=e2, FPS_idx=F2, xyz_flag=True)
        F3, x4 = self._sample(x3, n[4])
        e4, nrm4, knn4, dist4 = self.la4(xyz=x4, base_xyz=x3, normal=nrm3, feature=e3, FPS_idx=F3, xyz_flag=False)
        kw = dict(FPS_0=F0, FPS_1=F1, FPS_2=F2, FPS_3=F3, knn_0=knn0, knn_1=knn1, knn_2=knn2, knn_3=knn3,
                  knn_4=knn4, xyz0=xyz, xyz1=x1, xyz2=x2, xyz3=x3, xyz4=x4)
        # decoder (:795-840): transition up, refine, fuse every state into the target state
        d4 = self.mlp(e4)
        d4 = self.fuse1(n[4], f0=e0, f1=e1, f2=e2, f3=e3, f4=d4, **kw)[4]
        d3, _, _, _ = self.la4_up(xyz=x3, base_xyz=x3, normal=nrm3,
                                  feature=self.up_conv4(upsample(d4, knn4, dist=dist4, n_out=n[3])))
        d3 = self.fuse2(n[3], f0=e0, f1=e1, f2=e2, f3=d3, f4=e4, **kw)[3]
        d2, _, _, _ = self.la3_up(xyz=x2, base_xyz=x2, normal=nrm2,
                                  feature=self.up_conv3(upsample(d3, knn3, dist=dist3, n_out=n[2])))
        d2 = self.fuse3(n[2], f0=e0, f1=e1, f2=d2, f3=e3, f4=e4, **kw)[2]
        d1, _, _, _ = self.la2_up(xyz=x1, base_xyz=x1, normal=nrm1,
                                  feature=self.up_conv2(upsample(d2, knn2, dist=dist2, n_out=n[1])))
        d1 = self.fuse4(n[1], f0=e0, f1=d1, f2=e2, f3=e3, f4=e4, **kw)[1]
        d0, _, _, _ = self.la1_up(xyz=xyz, base_xyz=xyz, normal=nrm0,
                                  feature=self.up_conv1(upsample(d1, knn1, dist=dist1, n_out=n[0])))
        d0 = self.fuse5(n[0], f0=d0, f1=e1, f2=e2, f3=e3, f4=e4, **kw)[0]
        # head input (:843-853): per-state global max pool, label embedding, concat -> 896 channels
        global_rep = torch.cat([t.max(dim=1, keepdim=True)[0] for t in (d0, d1, d2, d3, d4)], dim=2)
        label = self.conv7(label)
        if parts:  # (per-point channels [B,N,256], per-cloud channels [B,640]): the caller projects them separately
            return xyz, self.conv5(d0), torch.cat((global_rep, label), 2).squeeze(1)
        final = torch.cat((self.conv5(d0), global_rep.expand(-1, N, -1), label.expand(-1, N, -1)), 2)
        return xyz, final


class UmbrellaSurfaceConstructor(nn.Module):
    """Umbrella-based surface abstraction, R/modules/pointnet2_utils.py:336-399: same constructor,
    forward(center [B,3,N]) -> [B,in_channel,N] and state_dict (mlps.0 / .1 / .3 / .4 / .6).  The geometry -- kNN,
    azimuth sort, the k-1 triangles' centroid / polar / normal / constant, NaN repair -- is one kNN launch plus one
    fused kernel (ops.umbrella_features) instead of the reference's ~40 ATen ops on [B,N,G,3,3] tensors; the 10-channel
    1x1 convolutions stay torch modules.  random_inv draws its per-cloud sign on the CPU generator exactly like
    R/modules/recons_utils.py:50."""

    def __init__(self, k, in_channel, aggr_type='sum', return_dist=False, random_inv=True, cuda=False):
        super().__init__()
        self.k = k
        self.return_dist = return_dist
        self.random_inv = random_inv
        self.aggr_type = aggr_type
        self.cuda_ops = cuda  # the reference stores this as `self.cuda`, shadowing nn.Module.cuda()
        self.mlps = nn.Sequential(
            nn.Conv2d(in_channel, in_channel, 1, bias=False),
            nn.BatchNorm2d(in_channel),
            nn.ReLU(True),
            nn.Conv2d(in_channel, in_channel, 1, bias=True),
            nn.BatchNorm2d(in_channel),
            nn.ReLU(True),
            nn.Conv2d(in_channel, in_channel, 1, bias=True),
        )

    @on_tensor_device
    def forward(self, center):
        center = center.permute(0, 2, 1).contiguous()
        sign = None
        if self.random_inv:
            sign = torch.randint(0, 2, (center.size(0), 1, 1)).float() * 2. - 1.
        feat = ops.umbrella_features(center, self.k, self.return_dist, sign)  # [B,N,G,C]
        B, N, G, C = feat.shape
        # mlps on the [B*N*G, C] row view: a 1x1 Conv2d is a Linear over channels and BatchNorm2d over (B,G,N) is
        # BatchNorm over the rows.  fp32 GEMMs (cuDNN's convolution would default to TF32: 1e-3 off the reference)
        # and this repo's BatchNorm+ReLU kernels (LeakyReLU with slope 0)
        x = feat.reshape(-1, C)
        conv0, bn1, _, conv3, bn4, _, conv6 = self.mlps
        for conv, bn in ((conv0, bn1), (conv3, bn4)):
            x = F.linear(x, conv.weight.view(conv.out_channels, conv.in_channels), conv.bias)
            x = ops.bn_act(x, bn.weight, bn.bias, bn.running_mean, bn.running_var, bn.num_batches_tracked,
                           bn.training, momentum=ops._momentum(bn), eps=bn.eps, slope=0.0)
        x = F.linear(x, conv6.weight.view(conv6.out_channels, conv6.in_channels), conv6.bias).view(B, N, G, -1)
        if self.aggr_type == 'max':
            x = torch.max(x, 2)[0]
        elif self.aggr_type == 'avg':
            x = torch.mean(x, 2)
        else:
            x = torch.sum(x, 2)
        return x.permute(0, 2, 1)  # [B,C,N]


class PointNetFeaturePropagation(nn.Module):
    """three_nn + three_interpolate + Linear, R/modules/pointnet2_utils.py:860-912.  Inputs are [B,N,C]
    (the reference's docstring says [B,C,N] but its permutes are commented out, :887-892); points1 is unused
    and mlp_convs / mlp_bns are constructed but never called, as in the reference."""

    def __init__(self, in_channel, mlp, act=False):
        super().__init__()
        self.mlp_convs = nn.ModuleList()
        self.mlp_bns = nn.ModuleList()
        last_channel = in_channel
        for out_channel in mlp:
            self.mlp_convs.append(nn.Conv1d(last_channel, out_channel, 1))
            self.mlp_bns.append(nn.BatchNorm1d(out_channel))
            last_channel = out_channel
        self.act = act
        self.conv = Linear(in_channel, out_channel, bn=False, act=self.act)

    @on_tensor_device
    def forward(self, xyz1, xyz2, points1, points2):
        B, N, C = xyz1.shape
        S = xyz2.shape[1]
        if S == 1:
            interpolated_points = points2.repeat(1, N, 1)
        else:
            dists, idx = three_nn(xyz1, xyz2)
            interpolated_points = three_interpolate(points2, dists, idx)
        return self.conv(interpolated_points)
