"""GPU tests of the fused set-abstraction scales (csrc/mlp.cu, include/sg4d.h section 4): ball-query indices ->
pooled features without the grouped tensor.  Reference = tests/sa_ref.py: an fp64 PyTorch evaluation of the reference
op sequence with sg4d's max-pool / ReLU SELECTIONS PINNED; values and every gradient are compared at north_star's
1e-4 (relative L2 and worst element) with no outliers allowed."""
import pytest
import torch
import torch.nn as nn

import sa_ref

pytestmark = pytest.mark.gpu


def _mlp(cin, c1, c2, seed):
    from sg4d.pointnet2_ops.pointnet2_modules import build_shared_mlp
    torch.manual_seed(seed)
    m = build_shared_mlp([cin, c1, c2])
    with torch.no_grad():
        for layer in m:
            if isinstance(layer, nn.BatchNorm2d):
                layer.weight.copy_(torch.randn_like(layer.weight))      # mixed signs: exercises the min branch
                layer.bias.copy_(0.2 * torch.randn_like(layer.bias))
    return m


def _scene(b, n, m, ns, c_pts, c_feat, seed):
    """random clouds; centres are points of the cloud; idx = random neighbours with first-hit padding like ball query"""
    g = torch.Generator().manual_seed(seed)
    pts = torch.rand(b, n, 3 + c_pts, generator=g)
    pts[:, :, :3] = pts[:, :, :3] * 2 - 1
    feats = torch.randn(b, n, c_feat, generator=g) if c_feat else None
    centers = torch.stack([pts[i, torch.randperm(n, generator=g)[:m], :3] for i in range(b)]).contiguous()
    # ball-query rows: cnt DISTINCT ascending hits, remaining slots repeat the first hit (ball_query_gpu.cu:33-41)
    idx = torch.stack([torch.stack([torch.randperm(n, generator=g)[:ns] for _ in range(m)]) for _ in range(b)])
    cnt = torch.randint(1, ns + 1, (b, m), generator=g, dtype=torch.int32)
    srt, _ = torch.sort(idx, dim=2)
    slot = torch.arange(ns).view(1, 1, ns)
    idx = torch.where(slot < cnt.unsqueeze(-1), srt, srt[:, :, :1]).to(torch.int32).contiguous()
    return pts, feats, centers, idx, cnt


def _run(cuda, kind, b, n, m, ns, c_pts, c_feat, n1, n2, seed):
    from sg4d import mlp
    pts, feats, centers, idx, cnt = _scene(b, n, m, ns, c_pts, c_feat, seed)
    c = c_feat if c_feat else c_pts
    foff = 0 if c_feat else 3
    net = _mlp(3 + c, n1, n2, seed).to(cuda).train()
    dp, dc, di, dn = pts.to(cuda), centers.to(cuda), idx.to(cuda), cnt.to(cuda)
    df = feats.to(cuda) if feats is not None else None
    fsrc = df if df is not None else dp
    assert mlp.sa_scale_kind(net, c, ns, kind == "sa2", fsrc.shape[2], foff) == kind
    out, x, params, (m1, v1, m2, v2) = sa_ref.scale_parity(kind, dp, df, foff, c, dc, di, dn, net, seed)
    # BatchNorm running statistics (momentum 0.1, unbiased variance) after one training step
    rows = x.shape[0]
    for bn_, mean, var in ((net[1], m1, v1), (net[4], m2, v2)):
        torch.testing.assert_close(bn_.running_mean.double(), 0.1 * mean.detach(), rtol=1e-5, atol=1e-6)
        torch.testing.assert_close(bn_.running_var.double(), 0.9 + 0.1 * var.detach() * rows / (rows - 1), rtol=1e-5, atol=1e-6)
        assert int(bn_.num_batches_tracked) == 1
    # eval mode: running statistics, no updates
    net.eval()
    with torch.no_grad():
        out_e = mlp.fused_sa_scale(kind, dp, df if df is not None else dp, foff, c, dc, di, dn, net)
        rm = [net[1].running_mean.double(), net[1].running_var.double(), net[4].running_mean.double(), net[4].running_var.double()]
        w1, g1, b1, w2, g2, b2 = [p_.detach() for p_ in params]
        h1 = torch.relu((x.detach() @ w1.t() - rm[0]) / torch.sqrt(rm[1] + sa_ref.EPS) * g1 + b1)
        z2 = torch.relu((h1 @ w2.t() - rm[2]) / torch.sqrt(rm[3] + sa_ref.EPS) * g2 + b2).view(-1, ns, n2).max(1).values
    sa_ref.check("eval out", out_e, z2)
    fails = list(sa_ref.FAILS)
    sa_ref.FAILS.clear()
    assert not fails, fails


@pytest.mark.parametrize("b,n,m,ns,c_pts,n2,seed", [(2, 500, 64, 16, 3, 64, 1), (3, 700, 40, 32, 4, 128, 2), (1, 300, 37, 8, 3, 64, 3),
                                                    (2, 4000, 512, 32, 4, 128, 4), (1, 200, 19, 64, 0, 128, 5)])
def test_fused_sa1_scale_matches_pinned_fp64(cuda, b, n, m, ns, c_pts, n2, seed):
    """SA1-style scale (K = 3 + c <= 7): first layer recomputed in the operand stagers, single-pass backward"""
    _run(cuda, "sa1", b, n, m, ns, c_pts, 0, 64, n2, seed)


@pytest.mark.parametrize("b,n,m,ns,c_feat,n1,n2,seed", [(2, 512, 128, 32, 192, 128, 128, 6), (1, 300, 50, 64, 192, 128, 128, 7),
                                                        (2, 128, 33, 16, 192, 64, 64, 8), (2, 400, 70, 32, 36, 64, 128, 9)])
def test_fused_sa2_scale_matches_pinned_fp64(cuda, b, n, m, ns, c_feat, n1, n2, seed):
    """SA2-style scale (195 inputs): rows gathered by the operand stagers, gradient back into the source features"""
    _run(cuda, "sa2", b, n, m, ns, 0, c_feat, n1, n2, seed)


def test_fused_sa1_many_tiles_per_cta(cuda):
    """148 CTAs x several tiles each + a ragged last tile: exercises the persistent mbarrier phases of the gather path"""
    _run(cuda, "sa1", 9, 2000, 333, 32, 4, 0, 64, 128, 11)


@pytest.mark.parametrize("b,n,m,ns,c1,seed", [(2, 512, 128, 32, 128, 1), (3, 200, 33, 16, 64, 2), (1, 64, 5, 64, 128, 3)])
def test_linearity_kernels_match_torch(cuda, b, n, m, ns, c1, seed):
    """sg4d_gather_y1 (+ its BatchNorm partial sums), sg4d_group_rows_grad_dy and sg4d_group_sum_dy against plain torch"""
    from sg4d import _lib
    _, _, _, idx, cnt = _scene(b, n, m, ns, 0, 0, seed)
    g = torch.Generator().manual_seed(seed)
    rows = b * m * ns
    z, cc = torch.randn(b * n, c1, generator=g).to(cuda), torch.randn(b * m, c1, generator=g).to(cuda)
    di, dn = idx.to(cuda), cnt.to(cuda)
    y1 = torch.empty(rows, c1, device=cuda)
    part = torch.empty(2 * _lib.load().sg4d_gather_y1_parts(c1), dtype=torch.float64, device=cuda)
    _lib.call("sg4d_gather_y1", z, rows, n, m, ns, c1, z.data_ptr(), cc.data_ptr(), di.data_ptr(), y1.data_ptr(), part.data_ptr())
    flat = (di.long() + (torch.arange(b, device=cuda) * n).view(b, 1, 1)).reshape(-1)
    want = z[flat] - cc.repeat_interleave(ns, dim=0)
    assert torch.equal(y1, want)
    sums = part.view(-1, c1, 2).sum(0)
    torch.testing.assert_close(sums[:, 0], want.double().sum(0), rtol=1e-6, atol=1e-5)
    torch.testing.assert_close(sums[:, 1], want.double().pow(2).sum(0), rtol=1e-6, atol=1e-5)

    dz1 = torch.randn(rows, c1, generator=g).to(cuda)
    p1, q1, u1 = [torch.randn(c1, generator=g).to(cuda) for _ in range(3)]
    dy = (p1 * dz1 - (q1 * y1 + u1)).double()
    h = torch.empty(b * m, c1, device=cuda)
    _lib.call("sg4d_group_sum_dy", z, b * m, ns, c1, y1.data_ptr(), dz1.data_ptr(), p1.data_ptr(), q1.data_ptr(), u1.data_ptr(),
              h.data_ptr())
    torch.testing.assert_close(h.double(), dy.view(b * m, ns, c1).sum(1), rtol=1e-5, atol=1e-5)
    gs = torch.empty(b * n, c1, device=cuda)
    _lib.call("sg4d_group_rows_grad_dy", z, b, n, m, ns, c1, y1.data_ptr(), dz1.data_ptr(), p1.data_ptr(), q1.data_ptr(),
              u1.data_ptr(), di.data_ptr(), dn.data_ptr(), gs.data_ptr())
    want_g = torch.zeros(b * n, c1, dtype=torch.float64, device=cuda).index_add_(0, flat, dy)
    torch.testing.assert_close(gs.double(), want_g, rtol=1e-5, atol=2e-5)


@pytest.mark.parametrize("b,n,m,ns,c_pts,n2,seed", [(2, 900, 96, 16, 4, 64, 21), (2, 700, 50, 32, 3, 128, 22), (1, 400, 9, 128, 4, 128, 23)])
def test_sa1_dw2_gram_form_equals_operand_form(cuda, b, n, m, ns, c_pts, n2, seed):
    """sg4d_sa1_bwd_dw2_gram (T1 - diag(a2) W2 h1^T h1 - b2 (x) colsum h1) against sg4d_sa1_bwd_dw2 (dY2^T h1 with the dY2 operand
    built from y2) on the same saved tensors and arbitrary dsel / a2 / b2"""
    from sg4d import _lib, mlp
    pts, _, centers, idx, cnt = _scene(b, n, m, ns, c_pts, 0, seed)
    net = _mlp(3 + c_pts, 64, n2, seed).to(cuda).train()
    dp, dc, di, dn = pts.to(cuda), centers.to(cuda), idx.to(cuda), cnt.to(cuda)
    mlp.CAPTURE = []
    try:
        with torch.no_grad():
            mlp.fused_sa_scale("sa1", dp, dp, 3, c_pts, dc, di, dn, net)
        cap = [c for c in mlp.CAPTURE if c.get("kind") == "sa1" and "y2" in c][0]
    finally:
        mlp.CAPTURE = None
    g = torch.Generator().manual_seed(seed)
    groups = b * m
    dsel = torch.randn(groups, n2, generator=g).to(cuda)
    a2, b2 = (1e-3 * torch.randn(n2, generator=g)).to(cuda), (1e-3 * torch.randn(n2, generator=g)).to(cuda)
    src = mlp._src_args(dp, None, 3, c_pts, dc, di)
    rows = src[0]
    w2 = net[3].weight.detach().view(n2, 64).contiguous()
    t1 = cap["stats1"][1].contiguous()
    lib = _lib.load()
    want = torch.empty(n2, 64, device=cuda)
    _lib.call("sg4d_sa1_bwd_dw2", dp, *src, cap["w1s"].data_ptr(), t1.data_ptr(), n2, cap["y2"].data_ptr(), a2.data_ptr(), b2.data_ptr(),
              dsel.data_ptr(), cap["garg"].data_ptr(), mlp._wgrad_partial(rows, 64, cuda).data_ptr(), want.data_ptr())
    got = torch.empty(n2, 64, device=cuda)
    ws = torch.empty(lib.sg4d_sa1_bwd_dw2_gram_ws_floats(rows, n2), device=cuda)
    _lib.call("sg4d_sa1_bwd_dw2_gram", dp, *src, cap["w1s"].data_ptr(), t1.data_ptr(), n2, w2.data_ptr(), a2.data_ptr(), b2.data_ptr(),
              dsel.data_ptr(), cap["garg"].data_ptr(), ws.data_ptr(), got.data_ptr())
    err = float((got - want).double().norm() / want.double().norm())
    assert err < 1e-5, err
    assert float((got - want).abs().max() / want.abs().max()) < 1e-5
