"""CPU tests of the host-side mirror: state_dict layout, error behaviour, synthetic data, edge CSR."""
import json
import os

import pytest
import torch

import sg4d
from sg4d import rows, synthetic
from sg4d.model import SGPNModelWrapper
from sg4d.pointnet2_ops import _ext, pointnet2_modules, pointnet2_utils

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CFG = json.load(open(os.path.join(ROOT, "tests", "golden", "no_gt.json")))


def _model(image=False):
    cfg = json.loads(json.dumps(CFG))
    if image:
        cfg["IMAGE_INPUT"] = "full"
    return SGPNModelWrapper(cfg, 12, 15, torch.ones(12), torch.ones(15), [f"r{i}" for i in range(14)] + ["none"])


def test_state_dict_matches_reference_layout(golden_dir):
    want = json.load(open(os.path.join(golden_dir, "state_dict_keys.json")))["no_gt"]
    got = {k: list(v.shape) for k, v in _model().state_dict().items()}
    assert list(got) == list(want)          # same keys in the same order (188 entries)
    assert got == want
    assert sum(p.numel() for p in _model().parameters()) == 5225887


def test_image_config_head_width():
    m = _model(image=True)
    assert tuple(m.rel_predictor.fc3.weight.shape) == (15, 1036)
    assert tuple(m.full_image_feature_reduction.weight.shape) == (128, 2048)


def test_operator_api_surface():
    for name in ("furthest_point_sample", "gather_operation", "ball_query", "grouping_operation", "three_nn",
                 "three_interpolate", "QueryAndGroup", "GroupAll"):
        assert hasattr(pointnet2_utils, name)
    for name in ("PointnetSAModuleMSG", "PointnetSAModule", "PointnetFPModule", "build_shared_mlp"):
        assert hasattr(pointnet2_modules, name)
    for name in ("gather_points", "gather_points_grad", "furthest_point_sampling", "three_nn", "three_interpolate",
                 "three_interpolate_grad", "ball_query", "group_points", "group_points_grad"):
        assert hasattr(_ext, name)     # the nine names of bindings.cpp:6-19
    spec = [3, 64, 64]
    pointnet2_modules.PointnetSAModuleMSG(npoint=8, radii=[0.1], nsamples=[4], mlps=[spec])
    assert spec[0] == 6                # use_xyz bumps the caller's list in place, like the reference


def test_cpu_tensors_are_rejected_like_the_reference():
    xyz = torch.rand(1, 16, 3)
    with pytest.raises(RuntimeError, match="CPU not supported"):
        _ext.furthest_point_sampling(xyz, 4)
    with pytest.raises(RuntimeError, match="CPU not supported"):
        _ext.ball_query(xyz[:, :4].contiguous(), xyz, 0.1, 4)
    with pytest.raises(RuntimeError, match="contiguous"):
        _ext.furthest_point_sampling(torch.rand(1, 3, 16).transpose(1, 2), 4)
    with pytest.raises(RuntimeError, match="int tensor"):
        _ext.gather_points(torch.rand(1, 3, 16), torch.zeros(1, 4, dtype=torch.int64))
    with pytest.raises(RuntimeError, match="CPU not supported"):
        _ext.three_nn(xyz, xyz)


def test_synthetic_scene_contract():
    sc = synthetic.make_scene(3, n_obj=4, n_points_obj=256, n_points_rel=128)
    assert sc["obj_points"].shape == (4, 6, 256) and sc["rel_points"].shape == (6, 7, 128)
    assert sc["obj_points"].transpose(1, 2).is_contiguous()          # collate's permuted view
    assert sc["edge_indices"].shape == (2, 6) and sc["edge_indices"].dtype == torch.int64
    assert sc["relation_objects_one_hot"].sum(1).eq(2).all()
    xyz = sc["obj_points"][:, :3]
    assert float(xyz.pow(2).sum(1).max()) <= 1.0 + 1e-5
    assert synthetic.edge_list(12).shape[1] == 66 and synthetic.edge_list(12, "ordered").shape[1] == 132
    again = synthetic.make_scene(3, n_obj=4, n_points_obj=256, n_points_rel=128)
    assert torch.equal(sc["rel_points"], again["rel_points"])
    b = synthetic.make_batch(0, 3, n_obj=4, n_points_obj=64, n_points_rel=64)
    assert b["obj_points"].shape == (12, 6, 64) and b["rel_points"].shape == (18, 7, 64)
    assert int(b["edge_indices"].max()) == 11 and b["edge_scene"].tolist() == [0] * 6 + [1] * 6 + [2] * 6


def test_edge_csr():
    ei = torch.tensor([[0, 0, 0, 1, 1, 2, 3], [1, 2, 3, 2, 3, 3, 1]])
    csr = rows.EdgeCSR(ei, 4)
    order, ptr = csr.by_dst
    assert ptr.tolist() == [0, 0, 2, 4, 7]
    assert order.tolist() == [0, 6, 1, 3, 2, 4, 5]    # stable: ascending edge id within a destination
    order, ptr = csr.by_src
    assert ptr.tolist() == [0, 3, 5, 6, 7] and order.tolist() == [0, 1, 2, 3, 4, 5, 6]


def test_shard_range():
    from sg4d import parallel
    spans = [parallel.shard_range(10, r, 4) for r in range(4)]
    assert spans == [(0, 3), (3, 6), (6, 8), (8, 10)]


def test_bench_reference_arm_prints_one_json_line():
    """`bench.py --impl reference` (the CPU arm the driver runs beside ours): exactly one JSON line on stdout with the
    contract's keys, whatever the native libraries print."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                        "--cpu-clouds", "4", "--points", "1024"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "scenes/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and "sample" in d["cpu_baseline"]
    assert d["e2e"] == {"value": d["value"], "unit": "scenes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["metric"] == "scenes/sec fwd+bwd" and "workload" in d["config"]


def test_edge_csr_matches_bincount_reference():
    """EdgeCSR (sort + searchsorted, no host sync) == the obvious bincount/cumsum construction, including nodes
    without edges and an empty edge list"""
    from sg4d.rows import EdgeCSR
    g = torch.Generator().manual_seed(0)
    for n_nodes, n_edges in [(7, 30), (12, 66), (5, 0), (3, 4)]:
        ei = torch.randint(0, n_nodes, (2, n_edges), generator=g)
        csr = EdgeCSR(ei, n_nodes)
        for key, (order, ptr) in ((ei[1], csr.by_dst), (ei[0], csr.by_src)):
            counts = torch.bincount(key, minlength=n_nodes)
            want = torch.zeros(n_nodes + 1, dtype=torch.int64)
            want[1:] = torch.cumsum(counts, 0)
            assert ptr.dtype == torch.int32 and torch.equal(ptr.long(), want)
            assert torch.equal(key[order.long()], torch.sort(key, stable=True)[0])
            assert torch.equal(order.long(), torch.sort(key, stable=True)[1])          # stable: ascending edge id per node


def test_spatial_index_workspace_contract():
    """host-side entry points of the spatial index: supported range and workspace size (no GPU needed)"""
    lib = sg4d._lib.load()
    assert lib.sg4d_spatial_index_supported(1023) == 0 and lib.sg4d_spatial_index_supported(1024) == 1
    assert lib.sg4d_spatial_index_supported(80000) == 1 and lib.sg4d_spatial_index_supported(6144 * 64 + 1) == 0
    one, two = lib.sg4d_spatial_index_bytes(1, 80000), lib.sg4d_spatial_index_bytes(2, 80000)
    assert one >= 80000 * 20 and one % 4 == 0 and two > one
    per_cloud = two - one                       # one more cloud: its index + its todo list
    assert per_cloud % 128 == 0 or (per_cloud - 4 * 1025) % 128 == 0
    assert lib.sg4d_spatial_index_bytes(0, 80000) in (0, 128) and lib.sg4d_spatial_index_bytes(3, 0) == 0


def test_classification_report_matches_sklearn():
    """sg4d.metrics.classification_report == sklearn's (what scene_graph_prediction_model.py:213-229 logs), incl. empty classes"""
    import numpy as np
    import torch
    from sg4d import metrics
    sk = __import__("pytest").importorskip("sklearn.metrics")
    rng = np.random.RandomState(0)
    gts, preds = rng.randint(0, 12, 500), rng.randint(0, 15, 500)       # labels 12..14 never occur as ground truth
    names = [f"r{i}" for i in range(15)]
    want = sk.classification_report(gts, preds, labels=list(range(15)), target_names=names, output_dict=True, zero_division=0)
    got = metrics.classification_report(torch.from_numpy(gts), torch.from_numpy(preds), 15, names)
    for key in names + ["macro avg", "weighted avg"]:
        for f in ("precision", "recall", "f1-score"):
            assert abs(got[key][f] - want[key][f]) < 1e-12, (key, f)
    m = metrics.RelationMetrics(names)
    for take in (3, 1):
        m.update({"gt_rels": torch.from_numpy(gts[:100]), "take_idx": take}, torch.nn.functional.one_hot(torch.from_numpy(preds[:100]), 15).float())
    ev = m.evaluate("train")
    assert sorted(ev["takes"]) == [1, 3] and abs(ev["macro_f1"] - ev["takes"][1]["macro avg"]["f1-score"]) < 1e-12


def test_loss_scaler_follows_grad_scaler_rule():
    """LossScaler: unscaled gradients reach the optimizer, a non-finite gradient skips the step and halves the scale, the scale
    doubles after `growth_interval` finite steps (torch.cuda.amp.GradScaler semantics, what Lightning's precision=16 applies)"""
    import torch
    from sg4d.trainer import LossScaler
    w = torch.nn.Parameter(torch.tensor([1.0, -2.0]))
    opt = torch.optim.SGD([w], lr=0.5)
    sc = LossScaler(init_scale=8.0, growth_interval=2)
    sc.scale((w * torch.tensor([3.0, 4.0])).sum()).backward()
    assert torch.equal(w.grad, torch.tensor([24.0, 32.0]))
    assert sc.step(opt) and torch.allclose(w.detach(), torch.tensor([1.0 - 1.5, -2.0 - 2.0])) and sc.scale_value == 8.0
    w.grad = torch.tensor([float("inf"), 1.0])
    before = w.detach().clone()
    assert not sc.step(opt) and torch.equal(w.detach(), before) and sc.scale_value == 4.0
    for _ in range(2):
        w.grad = torch.ones(2)
        sc.step(opt)
    assert sc.scale_value == 8.0


def test_fused_scale_dispatch():
    """which fused kernel family a set-abstraction scale runs on (host logic of sg4d.mlp.sa_scale_kind): the model's SA1 scales
    recompute their first layer, its SA2 scales go through the first layer's linearity, anything else materialises rows"""
    from sg4d import mlp
    from sg4d.pointnet2_ops.pointnet2_modules import build_shared_mlp
    sa1 = build_shared_mlp([3 + 3, 64, 64])
    sa1b = build_shared_mlp([3 + 4, 64, 128])
    sa2 = build_shared_mlp([3 + 192, 128, 128])
    assert mlp.sa_scale_kind(sa1, 3, 16, False, 6, 3) == "sa1"
    assert mlp.sa_scale_kind(sa1b, 4, 32, False, 7, 3) == "sa1"
    assert mlp.sa_scale_kind(sa2, 192, 64, True, 192, 0) == "sa2"
    assert mlp.sa_scale_kind(sa2, 192, 64, False, 192, 0) == "sa2"            # no feature gradient needed: same kernels, no dFeats
    assert mlp.sa_scale_kind(build_shared_mlp([3 + 36, 64, 128]), 36, 32, True, 36, 0) == "sa2"
    assert mlp.sa_scale_kind(build_shared_mlp([3 + 4, 64, 128]), 4, 32, True, 4, 0) == "sa2"   # gradient into 4 features
    assert mlp.sa_scale_kind(build_shared_mlp([3 + 5, 64, 128]), 5, 32, True, 5, 0) is None    # 5 features: not a multiple of 4
    assert mlp.sa_scale_kind(sa2, 192, 48, True, 192, 0) is None              # nsample must be a power of two in 8..128
    assert mlp.sa_scale_kind(build_shared_mlp([3 + 192, 128, 256]), 192, 64, True, 192, 0) is None   # pooled width > 128
    assert mlp.sa_scale_kind(build_shared_mlp([3 + 192, 128, 64]), 192, 128, True, 192, 0) is None   # 64-wide pool: groups <= 64
    assert mlp.sa_scale_kind(build_shared_mlp([3 + 192, 128, 128, 128]), 192, 64, True, 192, 0) is None  # three layers
