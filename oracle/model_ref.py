"""oracle/model_ref.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Functional CPU restatement (stock PyTorch fp32 ops + the C kernels of ``oracle/pn2_oracle.c``) of the
reference's scene-graph forward pass, driven directly by a reference-layout ``state_dict``:

  SA level      OPS/pointnet2_modules.py:29-74  (FPS -> gather -> per scale: ball query, group xyz,
                recentre, group features, cat, [conv1x1 -> BN2d -> ReLU]x2, max-pool) and
                OPS/pointnet2_utils.py:300-337, 353-383 (QueryAndGroup / GroupAll)
  backbone      PN2/models/pointnet2_msg_cls.py:45-78, PN2/models/pointnet2_ssg_cls.py:98-124
  encoder       SGH/model/pointnets/network_PointNet2.py:21-25
  TripletGCN    SGH/model/gcns/network_TripletGCN.py:36-58, 72-80 (PyG propagate = index_select on
                edge_index[1]/[0]; torch_scatter 'add' = index_add_)
  heads         SGH/model/pointnets/network_PointNet.py:210-224, 250-271
  wrapper/loss  SGH/model/scene_graph_prediction_model.py:87-109, 139-141

It keeps the reference's channel-major layout and op sequence on purpose: it is the checker for the
point-major product path, and is itself checked against the reference's own Python in
``tests/golden/make_golden.py`` (fixtures in tests/golden/).  Third-party pieces (PyG 2.0.2,
torch-scatter 2.0.9) are restated from their documented semantics: parity for them is pinned only
by the reference's call sites.

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this module.
"""
import torch
import torch.nn.functional as F

from . import pn2_ext_cpu as ext

SA_SPECS = [  # (npoint, radii, nsamples) -- pointnet2_msg_cls.py:50-78
    (512, [0.1, 0.2], [16, 32]),
    (128, [0.2, 0.4], [32, 64]),
    (None, [None], [None]),
]


class _Group(torch.autograd.Function):  # OPS/pointnet2_utils.py:196-241
    @staticmethod
    def forward(ctx, features, idx):
        ctx.save_for_backward(idx)
        ctx.n = features.size(2)
        return ext.group_points(features.contiguous(), idx)

    @staticmethod
    def backward(ctx, g):
        (idx,) = ctx.saved_tensors
        return ext.group_points_grad(g.contiguous(), idx, ctx.n), None


class _Gather(torch.autograd.Function):  # OPS/pointnet2_utils.py:70-100
    @staticmethod
    def forward(ctx, features, idx):
        ctx.save_for_backward(idx)
        ctx.n = features.size(2)
        return ext.gather_points(features.contiguous(), idx)

    @staticmethod
    def backward(ctx, g):
        (idx,) = ctx.saved_tensors
        return ext.gather_points_grad(g.contiguous(), idx, ctx.n), None


def _bn(sd, key, x, training, track=True):
    rm = sd.get(key + ".running_mean") if track else None
    rv = sd.get(key + ".running_var") if track else None
    if training and track and (key + ".num_batches_tracked") in sd:
        sd[key + ".num_batches_tracked"] += 1
    return F.batch_norm(x, rm, rv, sd[key + ".weight"], sd[key + ".bias"], training or not track, 0.1, 1e-5)


def sa_level(sd, prefix, spec, xyz, features, training, probe=None, pins=None):
    """xyz (B,N,3), features (B,C,N) -> (new_xyz (B,npoint,3)|None, (B,sum C_out,npoint)).

    ``pins``: {f"{prefix}.{scale}": (h1_mask (B,C1,m,ns) bool, garg (B,C2,m,1) long, out_mask (B,C2,m) bool)} -- the ReLU
    active sets and the max-pool row choice of ANOTHER evaluation of the same scale (the product under test).  With them
    pinned, the scale is a smooth function of its inputs and parameters: gradients can be compared at 1e-4 without the
    outliers that selections flipping between near-ties would cause (tests/test_gpu_model.py).  Values are unaffected
    wherever the two evaluations agree on the selection, and differ by at most the near-tie gap elsewhere."""
    npoint, radii, nsamples = spec
    new_xyz = None
    if npoint is not None:
        fps_idx = ext.furthest_point_sampling(xyz.contiguous(), npoint)
        xyz_flipped = xyz.transpose(1, 2).contiguous()
        new_xyz = _Gather.apply(xyz_flipped, fps_idx).transpose(1, 2).contiguous()
        if probe is not None:
            probe[prefix + ".fps_idx"] = fps_idx
    outs = []
    for i, (r, ns) in enumerate(zip(radii, nsamples)):
        if npoint is not None:
            idx = ext.ball_query(new_xyz, xyz.contiguous(), r, ns)
            if probe is not None:
                probe[f"{prefix}.ball_idx.{i}"] = idx
            gx = _Group.apply(xyz.transpose(1, 2).contiguous(), idx)
            gx = gx - new_xyz.transpose(1, 2).unsqueeze(-1)
            x = torch.cat([gx, _Group.apply(features, idx)], dim=1)
        else:
            x = torch.cat([xyz.transpose(1, 2).unsqueeze(2), features.unsqueeze(2)], dim=1)
        pin = pins.get(f"{prefix}.{i}") if pins is not None else None
        for j in (0, 3):  # [conv, bn, relu] x 2
            x = F.conv2d(x, sd[f"{prefix}.mlps.{i}.{j}.weight"])
            x = _bn(sd, f"{prefix}.mlps.{i}.{j + 1}", x, training)
            if pin is None:
                x = F.relu(x)
            elif j == 0:
                x = x * pin[0]                                      # pinned ReLU of the first block
        if pin is None:
            outs.append(F.max_pool2d(x, kernel_size=[1, x.size(3)]).squeeze(-1))
        else:                                                      # pinned max-pool row + pinned ReLU of the pooled value
            outs.append(torch.gather(x, 3, pin[1]).squeeze(-1) * pin[2])
    return new_xyz, torch.cat(outs, dim=1)


def encoder(sd, prefix, points, training, probe=None, pins=None):
    """points (B, C, N) as the dataset collate emits them -> (B, 256)."""
    pc = points.transpose(1, 2)
    xyz = pc[..., 0:3].contiguous()
    features = pc[..., 3:].transpose(1, 2).contiguous()
    for lvl, spec in enumerate(SA_SPECS):
        xyz, features = sa_level(sd, f"{prefix}.backbone.SA_modules.{lvl}", spec, xyz, features, training, probe, pins)
    return features[:, :, 0]


def _lin(sd, key, x):
    return F.linear(x, sd[key + ".weight"], sd[key + ".bias"])


def triplet_gcn(sd, prefix, x, e, edge_index, dim_hidden=512, dim_edge=256):
    x_j = x.index_select(0, edge_index[0])
    x_i = x.index_select(0, edge_index[1])
    h = torch.cat([x_i, e, x_j], dim=1)
    h = F.relu(_bn(sd, prefix + ".nn1.1", _lin(sd, prefix + ".nn1.0", h), True, track=False))
    h = F.relu(_bn(sd, prefix + ".nn1.4", _lin(sd, prefix + ".nn1.3", h), True, track=False))
    msg = h[:, :dim_hidden] + h[:, dim_hidden + dim_edge:]
    new_e = h[:, dim_hidden:dim_hidden + dim_edge]
    agg = torch.zeros(x.size(0), dim_hidden, dtype=x.dtype).index_add_(0, edge_index[1], msg)
    y = F.relu(_bn(sd, prefix + ".nn2.1", _lin(sd, prefix + ".nn2.0", agg), True, track=False))
    return _lin(sd, prefix + ".nn2.3", y), new_e


def head(sd, prefix, x, training, extra=(), dropout_mask=None):
    x = F.relu(_lin(sd, prefix + ".fc1", x))
    x = _lin(sd, prefix + ".fc2", x)
    if training:
        x = F.dropout(x, 0.3, True) if dropout_mask is None else x * dropout_mask
    x = F.relu(x)
    if extra:
        x = torch.cat([x] + list(extra), dim=1)
    return F.log_softmax(_lin(sd, prefix + ".fc3", x), dim=1)


def forward(sd, batch, training=True, n_layers=2, probe=None, image=False, dropout=True, pins=None):
    """Returns the reference's return_meta_data tuple (without the trailing None).  ``sd`` is a dict
    of tensors keyed like the reference state_dict; BN running statistics in it are updated in place
    when ``training``."""
    obj_feature = encoder(sd, "obj_encoder", batch["obj_points"], training, probe, pins)
    rel_feature = encoder(sd, "rel_encoder", batch["rel_points"], training, probe, pins)
    x, e = obj_feature, rel_feature
    for i in range(n_layers):
        x, e = triplet_gcn(sd, f"gcn.gconvs.{i}", x, e, batch["edge_indices"])
        if i < n_layers - 1:
            x, e = F.relu(x), F.relu(e)
    obj_cls = head(sd, "obj_predictor", x, training and dropout)
    extra = []
    if image:
        feats = _lin(sd, "full_image_feature_reduction", batch["full_image_features"])
        emb = feats.flatten() if feats.dim() == 2 else feats.flatten(1).index_select(0, batch["edge_scene"])
        extra.append(emb.unsqueeze(0).repeat(len(e), 1) if emb.dim() == 1 else emb)
    extra.append(batch["relation_objects_one_hot"])
    rel_cls = head(sd, "rel_predictor", e, training and dropout, extra)
    return obj_cls, rel_cls, obj_feature, rel_feature, x, e


def loss_fn(obj_cls, rel_cls, batch, weights_obj, weights_rel, lambda_o=1e-6):
    return lambda_o * F.nll_loss(obj_cls, batch["gt_class"], weight=weights_obj) + \
        F.nll_loss(rel_cls, batch["gt_rels"], weight=weights_rel)


def clone_state(sd, requires_grad=True):
    """Detached copy of a state_dict; floating-point parameters become autograd leaves."""
    out = {}
    for k, v in sd.items():
        t = v.detach().clone().cpu()
        is_param = t.is_floating_point() and not k.endswith(("running_mean", "running_var"))
        out[k] = t.requires_grad_(True) if (requires_grad and is_param) else t
    return out
