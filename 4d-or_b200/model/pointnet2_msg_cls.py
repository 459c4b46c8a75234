"""PointNet++ MSG classification backbone used as the per-cloud encoder.

Mirrors ``PointNet2ClassificationMSG`` ("max we can run" variant,
PN2/models/pointnet2_msg_cls.py:45-78) and the parts of its base class that the scene-graph model
touches (``PointNet2ClassificationSSG.__init__/_break_up_pc/forward``,
PN2/models/pointnet2_ssg_cls.py:55-124): three SA levels N->512->128->1 and the never-executed
``fc_layer`` head, which is kept because its 667 176 parameters are part of the reference
``state_dict`` (keys ``backbone.fc_layer.{0,1,3,4,7}.*``).
"""
import torch
import torch.nn as nn

from ..pointnet2_ops.pointnet2_modules import PointnetSAModule, PointnetSAModuleMSG


class PointNet2ClassificationMSG(nn.Module):
    def __init__(self, input_dim):
        super().__init__()
        self.input_dim = input_dim
        self._build_model()

    def _build_model(self):
        f = self.input_dim - 3
        self.SA_modules = nn.ModuleList([
            PointnetSAModuleMSG(npoint=512, radii=[0.1, 0.2], nsamples=[16, 32],
                                mlps=[[f, 64, 64], [f, 64, 128]], use_xyz=True),
            PointnetSAModuleMSG(npoint=128, radii=[0.2, 0.4], nsamples=[32, 64],
                                mlps=[[64 + 128, 128, 128], [64 + 128, 128, 128]], use_xyz=True),
            PointnetSAModule(mlp=[128 + 128, 256, 256], use_xyz=True),
        ])
        # dead weight kept for checkpoint compatibility (pointnet2_ssg_cls.py:87-96)
        self.fc_layer = nn.Sequential(
            nn.Linear(1024, 512, bias=False), nn.BatchNorm1d(512), nn.ReLU(True),
            nn.Linear(512, 256, bias=False), nn.BatchNorm1d(256), nn.ReLU(True),
            nn.Dropout(0.5), nn.Linear(256, 40),
        )

    def forward_rows(self, pointcloud):
        """pointcloud (B, N, 3 + F) contiguous -> global feature (B, 256).  No split/transpose copies:
        SA1 reads xyz and the F input channels straight from the rows."""
        pc = pointcloud.contiguous()
        f = pc.shape[2] - 3
        xyz, feats = self.SA_modules[0].forward_rows(pc, pc if f > 0 else None, 3, f)
        for module in self.SA_modules[1:]:
            xyz_next, feats = module.forward_rows(xyz, feats, 0, feats.shape[2])
            xyz = xyz_next
        return feats[:, 0, :]

    def forward(self, pointcloud, return_features=False):
        """Reference contract: (B, N, 3 + F) -> features (B, 256, 1) when ``return_features``."""
        feats = self.forward_rows(pointcloud).unsqueeze(-1)
        if return_features:
            return feats
        return self.fc_layer(feats.squeeze(-1))
