"""oracle/ref_harness.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Works where ``/root/reference`` is mounted
(the build container) or where its Python has been staged under the git-ignored ``baseline/_ref/reference``
(oracle/build_ref_ext.py:stage_reference_python; that copy travels to the GPU box for bench.py's reference arm).

Imports the reference's OWN Python for the hot path -- ``pointnet2_ops.pointnet2_utils`` /
``pointnet2_modules`` (OPS/), ``PointNet2ClassificationMSG`` (PN2/models/pointnet2_msg_cls.py),
``TripletGCNModel`` (SGH/model/gcns/network_TripletGCN.py), the heads
(SGH/model/pointnets/network_PointNet.py) and ``SGPNModelWrapper``
(SGH/model/scene_graph_prediction_model.py) -- unmodified, by pre-seeding ``sys.modules`` with tiny
stand-ins for the third-party packages that are not installed here (pytorch_lightning, h5py, lmdb,
msgpack_numpy, timm, torch_geometric, torch_scatter) and by registering a chosen extension object as
``pointnet2_ops._ext``.  It is used (a) to validate ``oracle/model_ref.py`` and
``oracle/pn2_oracle.c`` and (b) by ``tests/golden/make_golden.py`` to generate the committed
fixtures.

The torch_geometric / torch_scatter stand-ins restate what torch-geometric 2.0.2's
``MessagePassing.propagate`` and torch-scatter 2.0.9's ``scatter(reduce='add')`` do for a dense
``edge_index`` (flow source_to_target: x_j = x[edge_index[0]], x_i = x[edge_index[1]], aggregate over
edge_index[1]); those packages are un-vendored dependencies of the reference (README.md:87), so that
part of the oracle is "parity unpinned" beyond the reference's own call sites
(network_TripletGCN.py:40-58).
"""
import importlib
import json
import os
import sys
import types

import torch

# the mounted reference (build container) or its staged, git-ignored copy (GPU box; oracle/build_ref_ext.py)
_STAGED = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline", "_ref", "reference")
REF_ROOT = "/root/reference" if os.path.isdir("/root/reference/scene_graph_prediction") else _STAGED
OPS_LIB = os.path.join(REF_ROOT, "scene_graph_prediction/pointnet2_dir/pointnet2_ops_lib")


def available():
    return os.path.isdir(os.path.join(REF_ROOT, "scene_graph_prediction"))


def _mod(name, **attrs):
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


class _MessagePassing(torch.nn.Module):
    """What PyG 2.0.2 MessagePassing does for this call pattern (network_TripletGCN.py:32,41)."""

    def __init__(self, aggr="add", flow="source_to_target", node_dim=-2):
        super().__init__()
        self.aggr, self.flow, self.node_dim = aggr, flow, node_dim

    def propagate(self, edge_index, x=None, edge_feature=None):
        x_j = x.index_select(0, edge_index[0])
        x_i = x.index_select(0, edge_index[1])
        out = self.message(x_i=x_i, x_j=x_j, edge_feature=edge_feature)
        return self.aggregate(out, edge_index[1], None, x.size(0))


def _scatter(src, index, dim=-1, out=None, dim_size=None, reduce="sum"):
    assert reduce in ("add", "sum") and dim in (0, -2)
    res = torch.zeros((dim_size,) + tuple(src.shape[1:]), dtype=src.dtype, device=src.device)
    return res.index_add_(0, index, src)


def install(ext_module):
    """Seed sys.modules; ``ext_module`` becomes ``pointnet2_ops._ext``.  Idempotent per process
    (the reference binds _ext at import time, so one process = one ext)."""
    if not available():
        raise RuntimeError("the reference's Python is neither mounted at /root/reference nor staged under baseline/_ref")
    if "pointnet2_ops._ext" in sys.modules and sys.modules["pointnet2_ops._ext"] is not ext_module:
        raise RuntimeError("reference already imported with another _ext in this process")
    _mod("pytorch_lightning", LightningModule=torch.nn.Module, seed_everything=lambda s: None)
    for n in ("h5py", "lmdb", "msgpack_numpy"):
        _mod(n)
    _mod("timm", create_model=None)
    _mod("timm.data", resolve_data_config=None, create_transform=None)
    _mod("torch_geometric")
    _mod("torch_geometric.nn")
    _mod("torch_geometric.nn.conv", MessagePassing=_MessagePassing)
    _mod("torch_scatter", scatter=_scatter)
    for p in (REF_ROOT, OPS_LIB):
        if p not in sys.path:
            sys.path.insert(0, p)
    sys.modules["pointnet2_ops._ext"] = ext_module
    import pointnet2_ops  # noqa: F401  (the reference package, found via OPS_LIB)
    pointnet2_ops._ext = ext_module


def ref_utils():
    return importlib.import_module("pointnet2_ops.pointnet2_utils")


def ref_modules():
    return importlib.import_module("pointnet2_ops.pointnet2_modules")


def ref_backbone_cls():
    m = importlib.import_module("scene_graph_prediction.pointnet2_dir.pointnet2.models.pointnet2_msg_cls")
    return m.PointNet2ClassificationMSG


def ref_config(name="no_gt.json"):
    with open(os.path.join(REF_ROOT, "scene_graph_prediction/scene_graph_helpers/configs", name)) as f:
        return json.load(f)


def ref_wrapper_cls():
    m = importlib.import_module(
        "scene_graph_prediction.scene_graph_helpers.model.scene_graph_prediction_model")
    return m.SGPNModelWrapper


def build_ref_model(config=None, num_class=12, num_rel=15, seed=0):
    config = config or ref_config()
    torch.manual_seed(seed)
    names = [f"rel{i}" for i in range(num_rel - 1)] + ["none"]
    return ref_wrapper_cls()(config, num_class, num_rel, torch.ones(num_class), torch.ones(num_rel), names)
