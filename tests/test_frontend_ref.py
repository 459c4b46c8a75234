"""CPU: the numpy restatement of the reference's crop / sample step (oracle/frontend_ref.py) on a small scene -- invariants the
reference's data_preparation guarantees (SGH/dataset/data_preparation_utils.py:104-125, 178-218, 12-18)."""
import numpy as np
import torch

from oracle import frontend_ref


def test_restatement_invariants():
    from sg4d import synthetic
    points, masks = synthetic.make_raw_scene(7, n_obj=5, n_points=4000)
    edges = synthetic.edge_list(5, "ordered")
    g = torch.Generator().manual_seed(0)
    u_obj, u_rel = torch.rand(5, 300, generator=g), torch.rand(20, 400, generator=g)
    out = frontend_ref.prepare_scene(points.numpy(), masks.numpy(), 5, edges.numpy(), u_obj.numpy(), u_rel.numpy())
    m = masks.numpy()
    for i in range(5):
        assert (m[out["obj_picked"][i]] == i + 1).all()                       # object crops hold only the object's points
    xyz = points.numpy()[:, :3]
    for e in range(20):
        a, b = int(edges[0, e]), int(edges[1, e])
        lo = np.minimum(out["obj_box"][a][:3], out["obj_box"][b][:3])
        hi = np.maximum(out["obj_box"][a][3:], out["obj_box"][b][3:])
        sel = xyz[out["rel_picked"][e]]
        assert (sel > lo).all() and (sel < hi).all()                          # strictly inside the union box
        want_mask = (m[out["rel_picked"][e]] == a + 1) * 1 + (m[out["rel_picked"][e]] == b + 1) * 2
        assert np.array_equal(out["rel_points"][e][:, 6].numpy(), want_mask.astype(np.float32))
        # every member of both objects lies inside the padded union box, so it is a candidate
        assert out["edge_totals"][e] >= (m == a + 1).sum() + (m == b + 1).sum()
    for cl in list(out["obj_points"]) + list(out["rel_points"]):              # zero_mean: centroid 0, unit sphere
        assert float(cl[:, :3].mean(0).abs().max()) < 1e-5
        assert abs(float(cl[:, :3].pow(2).sum(1).sqrt().max()) - 1.0) < 1e-5
