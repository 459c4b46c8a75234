"""GPU crop / sample front-end (sg4d.frontend, csrc/frontend.cu) against the numpy restatement of the reference's
data_preparation (oracle/frontend_ref.py; SGH/dataset/data_preparation_utils.py:104-125, 178-218, 12-18, 37-39) on the same
scene and the SAME random draws: selected point indices, member counts and bounding boxes bit-exact; the normalised clouds
within 1e-6 (the centroid is an fp32 mean of up to 80 000 terms on both sides)."""
import numpy as np
import pytest
import torch

from oracle import frontend_ref

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n_obj,P,n,n_rel,pairs,seed", [(4, 5000, 512, 700, "ordered", 0), (12, 60000, 2048, 3000, "unordered", 1),
                                                        (3, 1500, 4000, 4000, "ordered", 2), (12, 200000, 20000, 20000, "unordered", 3)])
def test_frontend_matches_restatement(cuda, n_obj, P, n, n_rel, pairs, seed):
    from sg4d import frontend, synthetic
    points, masks = synthetic.make_raw_scene(seed, n_obj=n_obj, n_points=P)
    edges = synthetic.edge_list(n_obj, pairs)
    g = torch.Generator().manual_seed(seed)
    u_obj, u_rel = torch.rand(n_obj, n, generator=g), torch.rand(edges.shape[1], n_rel, generator=g)
    u_obj[0, 0], u_rel[0, 0] = 0.0, 0.99999994            # the extremes of [0, 1)
    want = frontend_ref.prepare_scene(points.numpy(), masks.numpy(), n_obj, edges.numpy(), u_obj.numpy(), u_rel.numpy())
    got = frontend.prepare_scene(points.to(cuda), masks.to(cuda), n_obj, n, n_rel, pairs=pairs, u_obj=u_obj.to(cuda),
                                 u_rel=u_rel.to(cuda), return_debug=True)
    dbg = got["_debug"]
    assert torch.equal(got["edge_indices"].cpu(), edges)
    np.testing.assert_array_equal(dbg["obj_box"].cpu().numpy(), want["obj_box"])
    np.testing.assert_array_equal(dbg["edge_totals"].cpu().numpy(), want["edge_totals"])
    np.testing.assert_array_equal(dbg["obj_picked"].cpu().numpy(), want["obj_picked"])          # index selection: bit-exact
    np.testing.assert_array_equal(dbg["rel_picked"].cpu().numpy(), want["rel_picked"])
    obj = got["obj_points"].permute(0, 2, 1).cpu()
    rel = got["rel_points"].permute(0, 2, 1).cpu()
    assert obj.shape == (n_obj, n, 6) and rel.shape == (edges.shape[1], n_rel, 7)
    torch.testing.assert_close(obj, want["obj_points"], rtol=0, atol=1e-6)
    torch.testing.assert_close(rel, want["rel_points"], rtol=0, atol=1e-6)
    assert torch.equal(obj[:, :, 3:], want["obj_points"][:, :, 3:])                             # colours / mask channel: copies
    assert torch.equal(rel[:, :, 3:], want["rel_points"][:, :, 3:])
    assert set(rel[:, :, 6].unique().tolist()) <= {0.0, 1.0, 2.0}


def test_frontend_feeds_the_model(cuda):
    """scene -> front-end -> SGPNModelWrapper forward / backward: the batch dict has the layout the model path expects"""
    import json, os
    from oracle import weights
    from sg4d import frontend, synthetic
    from sg4d.model import SGPNModelWrapper
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cfg = json.load(open(os.path.join(root, "tests", "golden", "no_gt.json")))
    m = SGPNModelWrapper(cfg, 12, 15, torch.ones(12), torch.ones(15), [f"r{i}" for i in range(14)] + ["none"])
    m.load_state_dict(weights.synth_state_dict(seed=0))
    m.to(cuda).train()
    points, masks = synthetic.make_raw_scene(5, n_obj=4, n_points=20000)
    batch = frontend.prepare_scene(points.to(cuda), masks.to(cuda), 4, 1024, 1500, pairs="ordered")
    batch["relation_objects_one_hot"] = torch.zeros(12, 12, device=cuda)
    batch["gt_class"] = torch.zeros(4, dtype=torch.long, device=cuda)
    batch["gt_rels"] = torch.zeros(12, dtype=torch.long, device=cuda)
    loss = m.training_step(batch)
    loss.backward()
    assert torch.isfinite(loss)


def test_frontend_writes_into_batch_slices(cuda):
    """out_obj / out_rel: two scenes cropped straight into their slices of one batch tensor = the per-scene results"""
    from sg4d import frontend, synthetic
    n_obj, P, n, n_rel = 4, 6000, 600, 900
    E = synthetic.edge_list(n_obj, "ordered").shape[1]
    obj_all = torch.full((2 * n_obj, n, 6), float("nan"), device=cuda)
    rel_all = torch.full((2 * E, n_rel, 7), float("nan"), device=cuda)
    for k in range(2):
        points, masks = synthetic.make_raw_scene(10 + k, n_obj=n_obj, n_points=P)
        g = torch.Generator().manual_seed(k)
        u_obj, u_rel = torch.rand(n_obj, n, generator=g).to(cuda), torch.rand(E, n_rel, generator=g).to(cuda)
        want = frontend.prepare_scene(points.to(cuda), masks.to(cuda), n_obj, n, n_rel, u_obj=u_obj, u_rel=u_rel)
        got = frontend.prepare_scene(points.to(cuda), masks.to(cuda), n_obj, n, n_rel, u_obj=u_obj, u_rel=u_rel,
                                     out_obj=obj_all[k * n_obj:(k + 1) * n_obj], out_rel=rel_all[k * E:(k + 1) * E])
        assert got["obj_points"].data_ptr() == obj_all[k * n_obj].data_ptr()
        assert torch.equal(obj_all[k * n_obj:(k + 1) * n_obj], want["obj_points"].permute(0, 2, 1))
        assert torch.equal(rel_all[k * E:(k + 1) * E], want["rel_points"].permute(0, 2, 1))
    with pytest.raises(RuntimeError):
        frontend.prepare_scene(points.to(cuda), masks.to(cuda), n_obj, n, n_rel, out_obj=obj_all[:n_obj, :, :5])
