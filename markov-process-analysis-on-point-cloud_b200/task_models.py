"""The two task models that call the hot path -- the boundary callers of SURVEY.md 2 (#12, #13) -- with the
reference's constructor / forward signatures and state_dict layout, so reference checkpoints load unchanged.

R = Markov_Process_Analysis_on_Point_Cloud/ in the reference tree.
  * Model(args)                       R/models/repsurf/repsurf_ssg_umb.py:35-70      (classifier, 743 keys)
  * get_model(num_classes, normal_channel) / get_loss()
                                      R/models/repsurf/pointnet2_part_seg_msg.py:33-180 (part-seg, 2189 keys)
  * SmoothClsLoss                     R/util/utils.py:74-88
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops
from ._lib import on_tensor_device
from .pointnet2_utils import KeepHighResolutionModulePartSeg, Linear
from .repsurface_utils import KeepHighResolutionModule, SurfaceAbstractionCD, UmbrellaSurfaceConstructor


class Model(nn.Module):
    """Classifier: keepHigh -> fc1/bn1/lrelu/drop -> fc2/bn2/lrelu/drop -> fc3 -> log_softmax."""

    def __init__(self, args):
        super().__init__()
        self.init_nsample = args.num_point
        self.return_dist = args.return_dist
        self.keepHigh = KeepHighResolutionModule(3, 64, 64, 64, 64, cuda=args.cuda_ops)
        self.fc1 = nn.Linear(1024, 512)
        self.bn1 = nn.BatchNorm1d(512)
        self.drop1 = nn.Dropout(0.5)
        self.fc2 = nn.Linear(512, 256)
        self.bn2 = nn.BatchNorm1d(256)
        self.drop2 = nn.Dropout(0.5)
        self.fc3 = nn.Linear(256, args.num_class)
        self.lrelu = nn.LeakyReLU(negative_slope=0.2)

    @on_tensor_device
    def forward(self, points):
        center = points[:, :3, :]
        normal = center  # the shipped model feeds the coordinates as `normal` (:59); it is never consumed
        x = self.keepHigh(center, normal)
        x = self.drop1(self.lrelu(self.bn1(self.fc1(x))))
        x = self.drop2(self.lrelu(self.bn2(self.fc2(x))))
        return F.log_softmax(self.fc3(x), -1)


class Model2x(nn.Module):
    """RepSurf-U 2x classifier, R/models/repsurf/repsurf_ssg_umb_2x.py:12-61 (`Model` there): umbrella surface
    constructor -> four SurfaceAbstractionCD stages -> MLP head.  Same constructor arguments (args.return_center,
    return_polar, num_point, return_dist, group_size, umb_pool, cuda_ops, num_class) and state_dict layout; it is the
    end-to-end consumer of the umbrella-feature and ball-query kernels."""

    def __init__(self, args):
        super().__init__()
        center_channel = 0 if not args.return_center else (6 if args.return_polar else 3)
        repsurf_channel = 10
        self.init_nsample = args.num_point
        self.return_dist = args.return_dist
        self.surface_constructor = UmbrellaSurfaceConstructor(args.group_size + 1, repsurf_channel,
                                                              return_dist=args.return_dist, aggr_type=args.umb_pool,
                                                              cuda=args.cuda_ops)
        self.sa1 = SurfaceAbstractionCD(npoint=512, radius=0.1, nsample=24, feat_channel=repsurf_channel,
                                        pos_channel=center_channel, mlp=[128, 128, 256], group_all=False,
                                        return_polar=args.return_polar, cuda=args.cuda_ops)
        self.sa2 = SurfaceAbstractionCD(npoint=128, radius=0.2, nsample=24, feat_channel=256 + repsurf_channel,
                                        pos_channel=center_channel, mlp=[256, 256, 512], group_all=False,
                                        return_polar=args.return_polar, cuda=args.cuda_ops)
        self.sa3 = SurfaceAbstractionCD(npoint=32, radius=0.4, nsample=24, feat_channel=512 + repsurf_channel,
                                        pos_channel=center_channel, mlp=[512, 512, 1024], group_all=False,
                                        return_polar=args.return_polar, cuda=args.cuda_ops)
        self.sa4 = SurfaceAbstractionCD(npoint=None, radius=None, nsample=None, feat_channel=1024 + repsurf_channel,
                                        pos_channel=center_channel, mlp=[1024, 1024, 2048], group_all=True,
                                        return_polar=args.return_polar, cuda=args.cuda_ops)
        self.classfier = nn.Sequential(  # (sic) the reference's attribute name, part of the state_dict keys
            nn.Linear(2048, 512), nn.BatchNorm1d(512), nn.ReLU(True), nn.Dropout(0.4),
            nn.Linear(512, 256), nn.BatchNorm1d(256), nn.ReLU(True), nn.Dropout(0.4),
            nn.Linear(256, args.num_class))

    @on_tensor_device
    def forward(self, points):
        center = points[:, :3, :]
        normal = self.surface_constructor(center)
        center, normal, feature = self.sa1(center, normal, None)
        center, normal, feature = self.sa2(center, normal, feature)
        center, normal, feature = self.sa3(center, normal, feature)
        center, normal, feature = self.sa4(center, normal, feature)
        feature = self.classfier(feature.reshape(-1, 2048))
        return F.log_softmax(feature, -1)


class get_model(nn.Module):
    """Part segmentation: keepHigh(xyz, normal=xyz, label) -> conv8+drop -> conv9 -> conv10 -> conv11."""

    def __init__(self, num_classes, normal_channel=False):
        super().__init__()
        self.normal_channel = normal_channel
        self.umb_pool = 'sum'
        self.group_size = 8
        self.return_dist = True
        self.keepHigh = KeepHighResolutionModulePartSeg(3, 64, 128, 256, 512, cuda=True)
        self.conv8 = Linear(896, 512, bn=False)
        self.conv9 = Linear(512, 256, bn=False)
        self.conv10 = Linear(256, 128, bn=False)
        self.conv11 = nn.Linear(128, num_classes)
        self.drop1 = nn.Dropout(0.5)
        self.drop2 = nn.Dropout(0.5)

    @on_tensor_device
    def forward(self, xyz, cls_label):
        if (xyz.is_cuda and ops.split_supported(xyz.shape[2], 256, self.conv8.linear.out_features)
                and not (ops.bf16_active() and not self.training)):
            # 640 of conv8's 896 input channels are constant over a cloud's points (global pools, label embedding):
            # project them once per cloud instead of once per point (ops.LinearBNActSplit) -- same arithmetic
            branch1_xyz, per_point, per_cloud = self.keepHigh.forward_parts(xyz, normal=xyz, label=cls_label)
            x = self.drop1(self.conv8.forward_split(per_point, per_cloud))
        else:
            branch1_xyz, final_points = self.keepHigh(xyz, normal=xyz, label=cls_label)
            x = self.drop1(self.conv8(final_points))
        x = self.conv9(x)
        x = self.conv10(x)
        x = ops.linear(x, self.conv11.weight, self.conv11.bias)  # conv11's arithmetic on the tcgen05 kernel
        return x, xyz


class get_loss(nn.Module):
    """Label-smoothed cross entropy, eps = 0.1 (R/models/repsurf/pointnet2_part_seg_msg.py:159-180)."""

    def forward(self, pred, target, trans_feat):
        if pred.is_cuda and pred.dim() == 2:  # one fused forward / backward kernel instead of ~10 elementwise launches
            return ops.smooth_cross_entropy(pred, target, 0.1)
        target = target.contiguous().view(-1)
        eps = 0.1
        n_class = pred.size(1)
        one_hot = torch.zeros_like(pred).scatter(1, target.view(-1, 1), 1)
        one_hot = one_hot * (1 - eps) + (1 - one_hot) * eps / (n_class - 1)
        log_prb = F.log_softmax(pred, dim=1)
        return -(one_hot * log_prb).sum(dim=1).mean()


class SmoothClsLoss(nn.Module):
    """R/util/utils.py:74-88 (pred is already log-softmax)."""

    def __init__(self, smoothing_ratio=0.1):
        super().__init__()
        self.smoothing_ratio = smoothing_ratio

    def forward(self, pred, target):
        eps = self.smoothing_ratio
        n_class = pred.size(1)
        one_hot = torch.zeros_like(pred).scatter(1, target.view(-1, 1), 1)
        one_hot = one_hot * (1 - eps) + (1 - one_hot) * eps / (n_class - 1)
        return -(one_hot * pred).sum(dim=1).mean()
