"""Classifier heads (SGH/model/pointnets/network_PointNet.py:188-271): ``PointNetCls`` for objects and
``PointNetRelCls`` for relations (late fusion of the optional image embedding and of the two 6-way
object-type one-hots before the last linear).  Both end in ``log_softmax``."""
import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import dense


def _init_head(module):
    # BaseNetwork.init_weights('xavier_normal', 1) as called by the reference heads
    # (networks_base.py:13-56 via network_PointNet.py:206-208): xavier-normal weights, zero biases
    for m in module.children():
        if isinstance(m, nn.Linear):
            nn.init.xavier_normal_(m.weight.data, gain=1)
            nn.init.constant_(m.bias.data, 0.0)
        elif isinstance(m, nn.BatchNorm1d):
            nn.init.constant_(m.weight.data, 1)
            nn.init.constant_(m.bias.data, 0.0)


class PointNetCls(nn.Module):
    def __init__(self, k=2, in_size=1024, batch_norm=True, drop_out=True, init_weights=True):
        super().__init__()
        self.name = 'pnetcls'
        self.in_size, self.k = in_size, k
        self.use_batch_norm, self.use_drop_out = batch_norm, drop_out
        self.fc1 = nn.Linear(in_size, 512)
        self.fc2 = nn.Linear(512, 256)
        self.fc3 = nn.Linear(256, k)
        if drop_out:
            self.dropout = nn.Dropout(p=0.3)
        if batch_norm:
            self.bn1 = nn.BatchNorm1d(512)
            self.bn2 = nn.BatchNorm1d(256)
        self.relu = nn.ReLU()
        if init_weights:
            _init_head(self)

    def forward(self, x):
        if self.use_batch_norm:
            raise NotImplementedError("sg4d heads: batch_norm=True has no kernel path (the reference builds its heads "
                                      "with batch_norm=False, scene_graph_prediction_model.py:65-72)")
        x = dense.linear(x, self.fc1)
        x = dense.linear(x, self.fc2, pre_relu=True)          # the ReLU is fused into the operand stager
        if self.use_drop_out:
            x = self.dropout(x)
        return F.log_softmax(dense.linear(x, self.fc3, pre_relu=True), dim=1)


class PointNetRelCls(nn.Module):
    def __init__(self, k=2, in_size=1024, batch_norm=True, drop_out=True, init_weights=True,
                 image_embedding_size=None, n_object_types=None):
        super().__init__()
        self.name = 'pnetcls'
        self.in_size = in_size
        self.use_bn, self.use_drop_out = batch_norm, drop_out
        self.fc1 = nn.Linear(in_size, 512)
        self.fc2 = nn.Linear(512, 256)
        self.fc3 = nn.Linear(256 + (image_embedding_size or 0) + n_object_types * 2, k)
        if drop_out:
            self.dropout = nn.Dropout(p=0.3)
        if batch_norm:
            self.bn1 = nn.BatchNorm1d(512)
            self.bn2 = nn.BatchNorm1d(256)
        self.relu = nn.ReLU()
        if init_weights:
            _init_head(self)

    def forward(self, x, relation_objects_one_hot=None, image_embeddings=None):
        if self.use_bn:
            raise NotImplementedError("sg4d heads: batch_norm=True has no kernel path (the reference builds its heads "
                                      "with batch_norm=False, scene_graph_prediction_model.py:65-72)")
        x = dense.linear(x, self.fc1)
        x = dense.linear(x, self.fc2, pre_relu=True)          # the ReLU is fused into the operand stager
        if self.use_drop_out:
            x = self.dropout(x)
        x = self.relu(x)
        if image_embeddings is not None:  # late fusion (:265-267)
            if image_embeddings.dim() == 1:  # one scene: the same embedding for every edge
                image_embeddings = image_embeddings.unsqueeze(0).expand(len(x), -1)
            x = torch.cat([x, image_embeddings], dim=1)
        if relation_objects_one_hot is not None:
            x = torch.cat([x, relation_objects_one_hot], dim=1)
        return F.log_softmax(dense.linear(x, self.fc3), dim=1)
