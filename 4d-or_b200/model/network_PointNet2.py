"""``PointNetfeat`` (SGH/model/pointnets/network_PointNet2.py:14-25): the encoder wrapper the model
instantiates twice (objects: 6 input channels, edges: 7)."""
import torch.nn as nn

from .pointnet2_msg_cls import PointNet2ClassificationMSG


class PointNetfeat(nn.Module):
    def __init__(self, input_dim=6, out_size=1024, input_dropout=0.0):
        super().__init__()
        self.name = 'pnetenc'
        self.backbone = PointNet2ClassificationMSG(input_dim=input_dim)
        self.out_size = out_size
        self.input_dropout = input_dropout  # stored, never applied -- as in the reference (:19)

    def forward(self, x):
        """x (B, C, N) -- the collate's permuted view of (B, N, C) rows (or_dataset.py:66) -> (B, 256)."""
        assert x.ndim > 2
        return self.backbone.forward_rows(x.transpose(1, 2))
