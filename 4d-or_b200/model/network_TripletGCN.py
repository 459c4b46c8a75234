"""Triplet graph convolution (SGH/model/gcns/network_TripletGCN.py:11-80).

Same parameters (``nn1``: 768->512->1280 with BatchNorm1d(track_running_stats=False)+ReLU after both
linears; ``nn2``: 512->512 (+BN+ReLU) ->256) and the same maths as the reference's PyG
``MessagePassing`` subclass, with the gather / concat and the scatter-add done by sg4d kernels
(``rows.triplet_gather`` / ``rows.message_aggregate``).  ``aggr`` is always 'add' in the reference
(the config's GCN_AGGR is never forwarded, scene_graph_prediction_model.py:59-62).
"""
import torch
import torch.nn as nn

from .. import dense, rows


def build_mlp(dim_list, activation='relu', do_bn=False, dropout=0, on_last=False):
    layers = []
    last = len(dim_list) - 2
    for i, (d_in, d_out) in enumerate(zip(dim_list[:-1], dim_list[1:])):
        layers.append(nn.Linear(d_in, d_out))
        if i != last or on_last:
            if do_bn:
                layers.append(nn.BatchNorm1d(d_out, track_running_stats=False))
            if activation == 'relu':
                layers.append(nn.ReLU())
            elif activation == 'leakyrelu':
                layers.append(nn.LeakyReLU())
        if dropout > 0:
            layers.append(nn.Dropout(p=dropout))
    return nn.Sequential(*layers)


class TripletGCN(nn.Module):
    def __init__(self, dim_node, dim_edge, dim_hidden, aggr='add', use_bn=True):
        super().__init__()
        assert aggr == 'add'
        self.aggr = aggr
        self.dim_node, self.dim_edge, self.dim_hidden = dim_node, dim_edge, dim_hidden
        self.nn1 = build_mlp([dim_node * 2 + dim_edge, dim_hidden, dim_hidden * 2 + dim_edge],
                             do_bn=use_bn, on_last=True)
        self.nn2 = build_mlp([dim_hidden, dim_hidden, dim_node], do_bn=use_bn)

    @staticmethod
    def _run(seq, x):
        """A ``build_mlp`` stack on the tensor-core engine (sg4d.dense): Linear -> BatchNorm1d -> ReLU blocks as fused
        GEMM + statistics + backward-in-the-stager kernels, a trailing plain Linear with its bias in the GEMM epilogue."""
        layers = list(seq)
        i = 0
        while i < len(layers):
            lin = layers[i]
            if not isinstance(lin, nn.Linear):
                raise NotImplementedError(f"TripletGCN MLP: unexpected layer {type(lin).__name__}")
            if i + 2 < len(layers) and isinstance(layers[i + 1], nn.BatchNorm1d) and isinstance(layers[i + 2], nn.ReLU):
                x = dense.linear_bn_relu(x, lin, layers[i + 1])
                i += 3
            elif i + 1 == len(layers):
                x = dense.linear(x, lin)
                i += 1
            else:
                raise NotImplementedError("TripletGCN MLP: only [Linear, BatchNorm1d, ReLU] blocks and a final Linear "
                                          "(use_bn=True, activation='relu', dropout=0: what the reference builds)")
        return x

    def forward(self, x, edge_feature, edge_index, csr=None):
        if csr is None:
            csr = rows.EdgeCSR(edge_index, x.shape[0])
        h = self._run(self.nn1, rows.triplet_gather(x, edge_feature, csr))     # (E, 2*hidden + edge)
        new_e = h[:, self.dim_hidden:self.dim_hidden + self.dim_edge]
        m = rows.message_aggregate(h, self.dim_hidden, self.dim_edge, csr)      # sum over incoming edges
        return self._run(self.nn2, m), new_e


class TripletGCNModel(nn.Module):
    """A stack of TripletGCN layers with ReLU on node and edge features in between (:72-80)."""

    def __init__(self, num_layers, **kwargs):
        super().__init__()
        self.num_layers = num_layers
        self.gconvs = nn.ModuleList(TripletGCN(**kwargs) for _ in range(num_layers))

    def forward(self, node_feature, edge_feature, edges_indices):
        csr = rows.EdgeCSR(edges_indices, node_feature.shape[0])
        for i, gconv in enumerate(self.gconvs):
            node_feature, edge_feature = gconv(node_feature, edge_feature, edges_indices, csr)
            if i < self.num_layers - 1:
                node_feature = torch.relu(node_feature)
                edge_feature = torch.relu(edge_feature)
        return node_feature, edge_feature
