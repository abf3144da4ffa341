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
from .ops import farthest_point_sample, index_points, knn_point, query_knn_point, square_distance  # noqa: F401
from .pointnet2_utils import Linear, LocalTrans


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
        pooled = torch.cat((final.max(dim=1)[0], final.mean(dim=1)), 1)  # adaptive max / avg pool (:632-634)
        return self.lrelu(self.bn(self.final_class(pooled)))
