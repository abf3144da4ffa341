"""Drop-in mirror of the Markov blocks in R/modules/repsurface_utils.py:380-639 (R =
Markov_Process_Analysis_on_Point_Cloud/ in the reference tree): the classifier's two-branch `LocalMerge` and
the six-stage encoder `KeepHighResolutionModule`, same signatures and state_dict keys (743 keys for the
classifier), on the sm_100a kernels.  `Linear` and `LocalTrans` are byte-identical between the reference's two
files, so they are shared with pointnet2_utils.py here.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops
from ._lib import on_tensor_device
from .ops import (farthest_point_sample, index_points, knn_point, query_ball_point, query_knn_point,  # noqa: F401
                  square_distance, xyz2sphere)
from .pointnet2_utils import Linear, LocalTrans, UmbrellaSurfaceConstructor  # noqa: F401


# ------------------------------------------------------------------------------------------------------
# RepSurf set abstraction (R/modules/repsurface_utils.py:12-84, 206-319)
# ------------------------------------------------------------------------------------------------------
def sample_and_group(npoint, radius, nsample, center, normal, feature, return_normal=True, return_polar=False,
                     cuda=False):
    """R/modules/repsurface_utils.py:12-59: FPS -> ball query -> grouped (relative coordinates [+ polar], normal,
    feature).  center [B,N,3], normal [B,N,Cn], feature [B,N,D] or None ->
    (new_center [B,S,3], new_normal [B,S,Cn], new_feature [B,S,nsample,C'])."""
    fps_idx = farthest_point_sample(center, npoint)
    new_center = index_points(center, fps_idx)
    new_normal = index_points(normal, fps_idx)
    idx = query_ball_point(radius, nsample, center, new_center, cuda=cuda)
    group_normal = index_points(normal, idx)
    group_center_norm = index_points(center, idx) - new_center.unsqueeze(2)
    if return_polar:
        group_center_norm = torch.cat([group_center_norm, xyz2sphere(group_center_norm)], dim=-1)
    if feature is not None:
        group_feature = index_points(feature, idx)
        new_feature = torch.cat([group_center_norm, group_normal, group_feature], dim=-1) if return_normal \
            else torch.cat([group_center_norm, group_feature], dim=-1)
    else:
        new_feature = torch.cat([group_center_norm, group_normal], dim=-1)
    return new_center, new_normal, new_feature


def sample_and_group_all(center, normal, feature, return_normal=True, return_polar=False):
    """R/modules/repsurface_utils.py:61-84: one group holding the whole cloud."""
    B, N, C = normal.shape
    new_center = torch.zeros(B, 1, 3, device=center.device)
    new_normal = new_center
    group_normal = normal.view(B, 1, N, C)
    group_center = center.view(B, 1, N, 3)
    if return_polar:
        group_center = torch.cat([group_center, xyz2sphere(group_center)], dim=-1)
    new_feature = torch.cat([group_center, group_normal, feature.view(B, 1, N, -1)], dim=-1) if return_normal \
        else torch.cat([group_center, feature.view(B, 1, N, -1)], dim=-1)
    return new_center, new_normal, new_feature


def _conv_bn(x, conv, bn, slope):
    """Conv2d(1x1) + BatchNorm2d [+ ReLU] of the reference on the channel-last row view [..., Cin]: the same
    arithmetic as Linear -> BatchNorm over all leading axes -> LeakyReLU(slope) (slope 0 = ReLU, 1 = none), i.e. the
    shared-MLP node of this repo (tcgen05 GEMM when Cin % 32 == 0, BatchNorm statistics from its epilogue)."""
    return ops.linear_bn_act(x, conv.weight.view(conv.out_channels, conv.in_channels), conv.bias, bn, bn.training,
                             slope)


class SurfaceAbstraction(nn.Module):
    """R/modules/repsurface_utils.py:206-254, same constructor / forward / state_dict (mlp_convs.i, mlp_bns.i).
    forward(center [B,3,N], normal [B,Cn,N], feature [B,D,N] or None) -> (new_center [B,3,S], new_normal [B,Cn,S],
    new_feature [B,mlp[-1],S])."""

    def __init__(self, npoint, radius, nsample, in_channel, mlp, group_all, return_polar=True, return_normal=True,
                 cuda=False):
        super().__init__()
        self.npoint = npoint
        self.radius = radius
        self.nsample = nsample
        self.return_normal = return_normal
        self.return_polar = return_polar
        self.cuda_ops = cuda  # the reference stores this as `self.cuda`, shadowing nn.Module.cuda()
        self.group_all = group_all
        self.mlp_convs = nn.ModuleList()
        self.mlp_bns = nn.ModuleList()
        last_channel = in_channel
        for out_channel in mlp:
            self.mlp_convs.append(nn.Conv2d(last_channel, out_channel, 1))
            self.mlp_bns.append(nn.BatchNorm2d(out_channel))
            last_channel = out_channel

    def _group(self, center, normal, feature):
        normal = normal.permute(0, 2, 1).contiguous()
        center = center.permute(0, 2, 1).contiguous()
        if feature is not None:
            feature = feature.permute(0, 2, 1).contiguous()
        if self.group_all:
            return sample_and_group_all(center, normal, feature, return_polar=self.return_polar,
                                        return_normal=self.return_normal)
        return sample_and_group(self.npoint, self.radius, self.nsample, center, normal, feature,
                                return_polar=self.return_polar, return_normal=self.return_normal, cuda=self.cuda_ops)

    @on_tensor_device
    def forward(self, center, normal, feature):
        new_center, new_normal, x = self._group(center, normal, feature)  # x [B,S,K,C] channel-last
        for conv, bn in zip(self.mlp_convs, self.mlp_bns):
            x = _conv_bn(x, conv, bn, 0.0)
        new_feature = torch.max(x, 2)[0].permute(0, 2, 1)  # max over the group -> [B,C,S]
        return new_center.permute(0, 2, 1), new_normal.permute(0, 2, 1), new_feature


class SurfaceAbstractionCD(SurfaceAbstraction):
    """R/modules/repsurface_utils.py:256-319: the first layer is split into a position branch (mlp_l0 / bn_l0 over
    the first pos_channel channels) and a feature branch (mlp_f0 / bn_f0), summed before the ReLU."""

    def __init__(self, npoint, radius, nsample, feat_channel, pos_channel, mlp, group_all, return_normal=True,
                 return_polar=False, cuda=False):
        nn.Module.__init__(self)
        self.npoint = npoint
        self.radius = radius
        self.nsample = nsample
        self.return_normal = return_normal
        self.return_polar = return_polar
        self.cuda_ops = cuda
        self.mlp_convs = nn.ModuleList()
        self.mlp_bns = nn.ModuleList()
        self.pos_channel = pos_channel
        self.group_all = group_all
        self.mlp_l0 = nn.Conv2d(self.pos_channel, mlp[0], 1)
        self.mlp_f0 = nn.Conv2d(feat_channel, mlp[0], 1)
        self.bn_l0 = nn.BatchNorm2d(mlp[0])
        self.bn_f0 = nn.BatchNorm2d(mlp[0])
        last_channel = mlp[0]
        for out_channel in mlp[1:]:
            self.mlp_convs.append(nn.Conv2d(last_channel, out_channel, 1))
            self.mlp_bns.append(nn.BatchNorm2d(out_channel))
            last_channel = out_channel

    @on_tensor_device
    def forward(self, center, normal, feature):
        new_center, new_normal, x = self._group(center, normal, feature)
        loc = _conv_bn(x[..., :self.pos_channel].contiguous(), self.mlp_l0, self.bn_l0, 1.0)
        feat = _conv_bn(x[..., self.pos_channel:].contiguous(), self.mlp_f0, self.bn_f0, 1.0)
        x = F.relu(loc + feat)
        for conv, bn in zip(self.mlp_convs, self.mlp_bns):
            x = _conv_bn(x, conv, bn, 0.0)
        new_feature = torch.max(x, 2)[0].permute(0, 2, 1)
        return new_center.permute(0, 2, 1), new_normal.permute(0, 2, 1), new_feature


class LocalMerge(nn.Module):
    """R/modules/repsurface_utils.py:406-446 (two feature branches; xyz_Trans only when there is no feature).
    fc1 and normal_Trans are constructed (state_dict) and never called, as in the reference."""

    def __init__(self, in_channels, out_channels, knn, usetanh=False, residual=False):
        super().__init__()
        self.knn = knn
        self.usetanh = usetanh
        self.residual = residual
        self.fc1 = Linear(out_channels * 2, out_channels, bn=False)
        self.fc2 = Linear(out_channels * 2, out_channels, bn=False)
        self.xyz_Trans = LocalTrans(3, out_channels, knn, usetanh=self.usetanh, residual=True)
        self.normal_Trans = LocalTrans(10, out_channels, knn, usetanh=self.usetanh, residual=True)
        self.feature_Trans = LocalTrans(in_channels, out_channels, knn, usetanh=self.usetanh, residual=self.residual)
        self.feature_Trans2 = LocalTrans(in_channels, out_channels, knn, usetanh=self.usetanh, residual=self.residual)

    @on_tensor_device
    def forward(self, xyz, base_xyz, normal=None, feature=None, FPS_idx=None, xyz_flag=True):
        dist, idx = ops.geo_join(ops.geo_call(lambda: knn_point(self.knn, base_xyz, xyz)))
        if feature is None:
            merge_features = self.xyz_Trans(features=xyz, idx=idx, pos=base_xyz, FPS_idx=FPS_idx, xyz=True)
        else:
            def branch_a():
                return self.feature_Trans(features=feature, idx=idx, pos=base_xyz, FPS_idx=FPS_idx)

            def branch_b():
                fs = index_points(feature, FPS_idx) if FPS_idx is not None else feature
                _, idx_feature = knn_point(self.knn, feature, fs)
                return self.feature_Trans2(features=feature, idx=idx_feature, pos=base_xyz, FPS_idx=FPS_idx)

            b, a = ops.parallel(branch_b, branch_a)  # two independent branches side by side
            merge_features = self.fc2(torch.cat((a, b), dim=2))
        return merge_features, normal, idx, dist


class KeepHighResolutionModule(nn.Module):
    """Classifier encoder, R/modules/repsurface_utils.py:542-639: la0 on the full cloud, then five
    (FPS -> LocalMerge) stages down to 512/256/128/64/32 points (the reference's literals, :581-619), conv3,
    conv4, max+avg pooling and the 2048->1024 projection.  forward(xyz [B,3,N], normal [B,3,N]) -> [B,1024]."""

    STAGES = (("la1", 512), ("la2", 256), ("la3", 128), ("la4", 64), ("la5", 32))

    def __init__(self, data_C, b1_C, b2_C, b3_C, b4_C, cuda=False):
        super().__init__()
        self.cuda_ops = cuda
        self.drop = nn.Dropout(0.5)
        self.la0 = LocalMerge(32, 64, 8, usetanh=False, residual=True)
        self.la1 = LocalMerge(64, 64, 8, usetanh=False, residual=False)
        self.la2 = LocalMerge(64, 64, 8, usetanh=False, residual=False)
        self.la3 = LocalMerge(64, 128, 8, usetanh=False, residual=True)
        self.la4 = LocalMerge(128, 256, 8, usetanh=False, residual=True)
        self.la5 = LocalMerge(256, 512, 8, usetanh=False, residual=True)
        self.start = Linear(3, 32, bn=False)
        self.conv3 = Linear(512, 512, bn=False)
        self.conv4 = Linear(512, 1024, bn=False)
        self.final = Linear(512, 1024, bn=False)
        self.final_class = nn.Linear(2048, 1024)
        self.bn = nn.BatchNorm1d(1024)
        self.lrelu = nn.LeakyReLU(negative_slope=0.2)

    @on_tensor_device
    def forward(self, xyz, normal):
        xyz = xyz.permute(0, 2, 1).contiguous()
        normal = normal.permute(0, 2, 1).contiguous()
        with ops.geometry_scope():  # FPS / coordinate kNN run ahead on the geometry lanes
            ops.geo_prefetch_pyramid(xyz, [npoint for _, npoint in self.STAGES], 8)
            feat, normal, _, _ = self.la0(xyz=xyz, base_xyz=xyz, normal=normal, xyz_flag=True)
            base = xyz
            for name, npoint in self.STAGES:
                fps_idx, sub = ops.geo_call(lambda points=base, npoint=npoint: ops.fps_and_gather(points, npoint))
                feat, normal, _, _ = getattr(self, name)(xyz=sub, base_xyz=base, normal=normal, feature=feat,
                                                        FPS_idx=fps_idx)
                base = sub
        final = self.conv4(self.conv3(feat))  # [B,32,1024]
        pooled = torch.cat((final.max(dim=1)[0], final.float().mean(dim=1).to(final.dtype)), 1)  # max / avg pool (:632-634)
        pooled = pooled.float()  # (bf16 inference path: the small head runs in fp32)
        return self.lrelu(self.bn(self.final_class(pooled)))
