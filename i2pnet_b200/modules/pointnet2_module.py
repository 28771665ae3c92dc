"""Set up-convolution of the small-range model (mirror of src/modules/pointnet2_module.py: SetUpconvModule :7-81).
Same constructor / forward / state_dict keys (`mlp_conv.i.conv.*`, `mlp2_conv.i.bn_linear.*`); kNN kernel + row
gather + fused shared-MLP chain with the max over the neighbours underneath."""
import torch
import torch.nn as nn

from ..projectPN.PPBackbone_center import Conv2d, run_mlp
from .point_utils import grouping


class SetUpconvModule(nn.Module):
    def __init__(self, nsample, in_channels, mlp, mlp2, is_training, bn_decay=None, bn=True, pooling='max', radius=None,
                 knn=True):
        super().__init__()
        self.nsample, self.mlp, self.mlp2 = nsample, mlp, mlp2
        self.is_training, self.bn_decay, self.bn, self.pooling, self.radius, self.knn = is_training, bn_decay, bn, pooling, radius, knn
        self.last_channel = in_channels[-1] + 3
        self.mlp_conv, self.mlp2_conv = nn.ModuleList(), nn.ModuleList()
        kw = dict(stride=[1, 1], bn=True, use_bn_input=False)      # nn.BatchNorm2d with running statistics, as basicConv.py:39
        if mlp is not None:
            for c in mlp:
                self.mlp_conv.append(Conv2d(self.last_channel, c, [1, 1], **kw))
                self.last_channel = c
        self.last_channel = (mlp[-1] if len(mlp) > 0 else self.last_channel) + in_channels[0]
        if mlp2 is not None:
            for c in mlp2:
                self.mlp2_conv.append(Conv2d(self.last_channel, c, [1, 1], **kw))
                self.last_channel = c

    def forward(self, xyz1, xyz2, feat1, feat2, raw_feat_point=False, raw_xyz1=None, raw_xyz2=None):
        """xyz1 (B,n1,3) dense level, xyz2 (B,n2,3) coarse level, feat1 (B,n1,c1) or None, feat2 (B,n2,c2)
        -> (B,n1,mlp2[-1]): every dense point gathers its nsample nearest coarse points."""
        xyz2_grouped, xyz_diff, feat2_grouped, _, _ = grouping(feat2, self.nsample, xyz2, xyz1, raw_feat_point=raw_feat_point,
                                                               raw_xyz1=raw_xyz2, raw_xyz2=raw_xyz1)
        net = torch.cat([feat2_grouped, xyz_diff], dim=3)                  # B,n1,K,c2+3
        if self.pooling == 'max':
            if len(self.mlp_conv) > 0:
                feat1_new = run_mlp(self.mlp_conv, net, reduce_k=True)
            else:
                feat1_new = torch.max(net, dim=2)[0]
        else:
            feat1_new = torch.mean(run_mlp(self.mlp_conv, net), dim=2)
        if feat1 is not None:
            feat1_new = torch.cat([feat1_new, feat1], dim=2)
        return run_mlp(self.mlp2_conv, feat1_new.unsqueeze(2)).squeeze(2)
