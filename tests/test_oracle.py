"""CPU tests of the oracle itself: the plain-torch/C restatement (oracle/) must reproduce the golden
fixtures that were produced by running the REFERENCE's own Python (tests/golden/make_golden.py)."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import model_ref, pn2_ext_cpu as ext, weights


def _npz(golden_dir, name):
    return {k: v for k, v in np.load(os.path.join(golden_dir, name), allow_pickle=False).items()}


def test_opt_n_threads_is_floor_log2():
    # cuda_utils.h:15-19 evaluates int(log(n)/log(2.0)); the CUDA side uses an exact integer log2
    for n in list(range(1, 5000)) + [8000, 8191, 8192, 65535, 65536, 80000, 131072, 200000]:
        assert ext.opt_n_threads(n) == min(1 << (n.bit_length() - 1), 512)


def _fps_rule(xyz, m):
    """Independent statement of the FPS selection rule: arg-max of temp over
    the non-skipped points; ties -> smallest (bitreverse_L(k mod T), k div T)."""
    n = xyz.shape[0]
    L = min(n.bit_length() - 1, 9)
    T = 1 << L
    k = np.arange(n)
    rev = np.zeros(n, dtype=np.int64)
    r = k % T
    for bit in range(L):
        rev |= ((r >> bit) & 1) << (L - 1 - bit)
    prio = rev * (n // T + 2) + k // T
    x = xyz.astype(np.float32)

    def sq(d):  # same fp32 op order as the kernels
        t = np.float32(d[:, 1] * d[:, 1])
        t = np.float32(np.float64(d[:, 0]) * np.float64(d[:, 0]) + np.float64(t)).astype(np.float32)
        return np.float32(np.float64(d[:, 2]) * np.float64(d[:, 2]) + np.float64(t)).astype(np.float32)

    valid = sq(x).astype(np.float64) > 1e-3
    temp = np.full(n, 1e10, dtype=np.float32)
    out = [0]
    old = 0
    for _ in range(1, m):
        d = sq(x - x[old])
        temp[valid] = np.minimum(d[valid], temp[valid])
        if not valid.any():
            old = 0
        else:
            best = temp[valid].max()
            cand = np.where(valid & (temp == best))[0]
            old = int(cand[np.argmin(prio[cand])])
        out.append(old)
    return np.array(out, dtype=np.int32)


@pytest.mark.parametrize("n,m", [(5, 5), (33, 9), (100, 40), (513, 64), (700, 96), (1500, 128)])
def test_fps_literal_simulation_matches_priority_rule(n, m):
    g = torch.Generator().manual_seed(n)
    xyz = (torch.rand(1, n, 3, generator=g) * 4).round() / 4 - 0.5   # lattice -> many exact ties
    xyz[0, n // 3] = 0.0
    got = ext.furthest_point_sampling(xyz.contiguous(), m)[0].numpy()
    # fma via float64 is exact for these lattice values, so the numpy statement is bit-faithful here
    np.testing.assert_array_equal(got, _fps_rule(xyz[0].numpy(), m))


def test_ops_fixture(golden_dir):
    """reference pointnet2_utils (FPS, gather, ball_query, QueryAndGroup fwd+bwd) == oracle composition"""
    fx = _npz(golden_dir, "ops_small.npz")
    for tag in "abc":
        xyz = torch.from_numpy(fx[f"{tag}_xyz"])
        m, r, ns = int(fx[f"{tag}_m"]), float(fx[f"{tag}_r"]), int(fx[f"{tag}_ns"])
        fps = ext.furthest_point_sampling(xyz, m)
        np.testing.assert_array_equal(fps.numpy(), fx[f"{tag}_fps"])
        new_xyz = ext.gather_points(xyz.transpose(1, 2).contiguous(), fps).transpose(1, 2).contiguous()
        np.testing.assert_array_equal(new_xyz.numpy(), fx[f"{tag}_new_xyz"])
        ball = ext.ball_query(new_xyz, xyz, r, ns)
        np.testing.assert_array_equal(ball.numpy(), fx[f"{tag}_ball"])
        feats = torch.from_numpy(fx[f"{tag}_feats"])
        gx = ext.group_points(xyz.transpose(1, 2).contiguous(), ball) - new_xyz.transpose(1, 2).unsqueeze(-1)
        qg = torch.cat([gx, ext.group_points(feats, ball)], 1)
        np.testing.assert_array_equal(qg.numpy(), fx[f"{tag}_qg"])
        w = torch.from_numpy(fx[f"{tag}_w"])
        dfe = ext.group_points_grad(w[:, 3:].contiguous(), ball, feats.shape[2])
        np.testing.assert_allclose(dfe.numpy(), fx[f"{tag}_dfeats"], rtol=0, atol=1e-5)


def test_sa_module_fixture(golden_dir):
    fx = _npz(golden_dir, "sa_msg.npz")
    shapes = json.loads(str(fx["shapes"]))
    sd = model_ref.clone_state({"sa." + k: v for k, v in weights.synth_state_dict(shapes, seed=5).items()})
    xyz = torch.from_numpy(fx["xyz"])
    feats = torch.from_numpy(fx["feats"]).requires_grad_(True)
    new_xyz, out = model_ref.sa_level(sd, "sa", (64, [0.25, 0.5], [8, 16]), xyz, feats, training=True)
    np.testing.assert_array_equal(new_xyz.detach().numpy(), fx["new_xyz"])
    np.testing.assert_allclose(out.detach().numpy(), fx["out"], rtol=0, atol=1e-5)
    (out * torch.from_numpy(fx["w"])).sum().backward()
    np.testing.assert_allclose(feats.grad.numpy(), fx["dfeats"], rtol=0, atol=2e-5)
    for k in fx:
        if k.startswith("grad."):
            np.testing.assert_allclose(sd["sa." + k[5:]].grad.numpy(), fx[k], rtol=1e-4, atol=2e-5)
        if k.startswith("after."):
            np.testing.assert_allclose(sd["sa." + k[6:]].detach().numpy(), fx[k], rtol=1e-6, atol=1e-6)


def test_model_fixture(golden_dir):
    """oracle/model_ref.forward == reference SGPNModelWrapper (config-1 shaped scene)"""
    import sg4d.synthetic as syn
    fx = _npz(golden_dir, "model_cfg1.npz")
    batch = syn.make_scene(0, n_obj=4, n_points_obj=2048, n_points_rel=2048)
    sd = model_ref.clone_state(weights.synth_state_dict(seed=0))
    outs = model_ref.forward(sd, batch, training=True, dropout=False)
    for name, t in zip(("obj_cls", "rel_cls", "obj_feature", "rel_feature", "gcn_obj", "gcn_rel"), outs):
        np.testing.assert_allclose(t.detach().numpy(), fx[name], rtol=0, atol=1e-5, err_msg=name)
    loss = model_ref.loss_fn(outs[0], outs[1], batch, torch.from_numpy(fx["w_obj"]), torch.from_numpy(fx["w_rel"]),
                             float(fx["lambda_o"]))
    np.testing.assert_allclose(float(loss), float(fx["loss"]), rtol=1e-5)
    loss.backward()
    norms = json.loads(str(fx["grad_norms"]))
    for k, ref in norms.items():
        g = sd[k].grad
        if ref is None:
            assert g is None or float(g.norm()) == 0.0, k
        else:
            assert abs(float(g.double().norm()) - ref) <= 1e-4 * max(1.0, ref), k
    for k in fx:
        if k.startswith("grad."):
            np.testing.assert_allclose(sd[k[5:]].grad.numpy(), fx[k], rtol=1e-3, atol=2e-5, err_msg=k)
        if k.startswith("after."):
            np.testing.assert_allclose(sd[k[6:]].detach().numpy(), fx[k], rtol=1e-5, atol=1e-6, err_msg=k)
    sd_eval = model_ref.clone_state(weights.synth_state_dict(seed=0), requires_grad=False)
    # the fixture's eval pass ran AFTER one training step had updated the running statistics
    for k in sd_eval:
        if "running" in k or "num_batches" in k:
            sd_eval[k] = sd[k].detach().clone()
    with torch.no_grad():
        eo = model_ref.forward(sd_eval, batch, training=False, dropout=False)
    np.testing.assert_allclose(eo[0].numpy(), fx["eval_obj_cls"], rtol=0, atol=1e-5)
    np.testing.assert_allclose(eo[1].numpy(), fx["eval_rel_cls"], rtol=0, atol=1e-5)


def test_three_nn_and_interpolate_restatement():
    """the FP-module ops: oracle vs plain torch on small clouds"""
    from oracle import pn2_ext_cpu as ora
    g = torch.Generator().manual_seed(11)
    unknown, known = torch.rand(2, 50, 3, generator=g), torch.rand(2, 17, 3, generator=g)
    known[0, 5] = known[0, 2]                                  # exact tie: the lower index comes first
    dist2, idx = ora.three_nn(unknown, known)
    d = ((unknown[:, :, None, :] - known[:, None, :, :]) ** 2).sum(-1)
    want_d, want_i = torch.sort(d, dim=2, stable=True)
    assert torch.equal(idx.long(), want_i[:, :, :3])
    torch.testing.assert_close(dist2, want_d[:, :, :3], rtol=1e-6, atol=1e-7)
    few = ora.three_nn(unknown, known[:, :2].contiguous())     # m < 3: missing slots are index 0 / +inf
    assert torch.isinf(few[0][:, :, 2]).all() and int(few[1][:, :, 2].abs().sum()) == 0
    feats = torch.randn(2, 4, 17, generator=g)
    w = torch.rand(2, 50, 3, generator=g)
    out = ora.three_interpolate(feats, idx, w)
    gathered = torch.gather(feats[:, :, None, :].expand(-1, -1, 50, -1), 3, idx.long()[:, None].expand(-1, 4, -1, -1))
    torch.testing.assert_close(out, (gathered * w[:, None]).sum(-1), rtol=1e-6, atol=1e-6)
    go = torch.randn(2, 4, 50, generator=g)
    grad = ora.three_interpolate_grad(go, idx, w, 17)
    f2 = feats.clone().requires_grad_(True)
    g2 = torch.gather(f2[:, :, None, :].expand(-1, -1, 50, -1), 3, idx.long()[:, None].expand(-1, 4, -1, -1))
    ((g2 * w[:, None]).sum(-1) * go).sum().backward()
    torch.testing.assert_close(grad, f2.grad, rtol=1e-5, atol=1e-5)


def test_fp_fixture_from_reference_python(golden_dir):
    """fp_module.npz was produced by the reference's own pointnet2_utils.three_nn / three_interpolate wrappers; the
    oracle ext called directly must reproduce it (pins the restatement's argument order / sqrt / weight handling)"""
    import os
    import numpy as np
    from oracle import pn2_ext_cpu as ora
    fx = np.load(os.path.join(golden_dir, "fp_module.npz"))
    unknown, known = torch.from_numpy(fx["unknown"]), torch.from_numpy(fx["known"])
    dist2, idx = ora.three_nn(unknown, known)
    np.testing.assert_array_equal(idx.numpy(), fx["idx"])
    np.testing.assert_array_equal(torch.sqrt(dist2).numpy(), fx["dist"])
    out = ora.three_interpolate(torch.from_numpy(fx["known_feats"]), idx, torch.from_numpy(fx["weight"]))
    np.testing.assert_array_equal(out.numpy(), fx["interp"])
    g = ora.three_interpolate_grad(torch.from_numpy(fx["w_interp"]), idx, torch.from_numpy(fx["weight"]), known.shape[1])
    np.testing.assert_allclose(g.numpy(), fx["d_known_feats"], rtol=1e-6, atol=1e-6)
