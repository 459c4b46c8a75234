"""GPU tests of the tensor-core shared-MLP kernels (csrc/mlp.cu) against fp64 / fp32 PyTorch references of
the same op.  The 3xTF32 product must be at fp32 accuracy (plain TF32 would be ~1e-3)."""
import copy

import pytest
import torch
import torch.nn as nn

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("rows,k,lda,n", [(128, 32, 32, 128), (256, 8, 8, 64), (1000, 64, 64, 128), (4096, 196, 196, 128),
                                          (20000, 128, 128, 64), (148 * 128 * 3 + 77, 200, 200, 128)])
def test_linear_fwd_matches_fp64(cuda, rows, k, lda, n):
    from sg4d import mlp
    g = torch.Generator().manual_seed(rows + k)
    a = torch.randn(rows, lda, generator=g).to(cuda)
    w = (torch.randn(n, k, generator=g) / k ** 0.5).to(cuda)
    y, partial, _, _ = mlp.linear_fwd(a, k, mlp.pack_weight(w), n)
    want = (a[:, :k].double() @ w.double().t())
    err = (y.double() - want).abs().max().item()
    ref32 = (a[:, :k] @ w.t()).double()
    err32 = (ref32 - want).abs().max().item()
    assert err <= max(4 * err32, 2e-6), (err, err32)       # as accurate as an fp32 GEMM
    stats = partial.view(-1, 2)                            # one fp64 pair per epilogue thread; thread e owns column e % n
    s = stats[:, 0].view(-1, n).sum(0)
    q = stats[:, 1].view(-1, n).sum(0)
    # the statistics are sums over the kernel's OWN outputs (fp32 per 128-row tile, fp64 across tiles)
    yd = y.double()
    torch.testing.assert_close(s, yd.sum(0), rtol=1e-5, atol=1e-5 * rows ** 0.5)
    torch.testing.assert_close(q, (yd * yd).sum(0), rtol=1e-5, atol=1e-5 * rows ** 0.5)


@pytest.mark.parametrize("rows,k,n,group", [(1024, 64, 64, 16), (2048, 64, 128, 32), (4096, 128, 128, 64), (512, 128, 128, 128), (384, 64, 64, 32),
                                            (640, 128, 64, 8)])
def test_linear_fwd_prologue_and_group_reduce(cuda, rows, k, n, group):
    from sg4d import mlp
    g = torch.Generator().manual_seed(rows + group)
    a = torch.randn(rows, k, generator=g).to(cuda)
    w = (torch.randn(n, k, generator=g) / k ** 0.5).to(cuda)
    scale = (torch.randn(k, generator=g)).to(cuda)
    shift = (0.3 * torch.randn(k, generator=g)).to(cuda)
    gamma = torch.randn(n, generator=g).to(cuda)
    y, _, gsel, garg = mlp.linear_fwd(a, k, mlp.pack_weight(w), n, scale=scale, shift=shift, group=group, gamma=gamma)
    act = torch.relu(a.double() * scale.double() + shift.double())
    want = act @ w.double().t()
    assert (y.double() - want).abs().max().item() < 5e-6
    yg = y.view(-1, group, n)
    sel_max, _ = yg.max(1)
    sel_min, _ = yg.min(1)
    pos = gamma >= 0
    assert torch.equal(gsel, torch.where(pos, sel_max, sel_min))
    picked = torch.gather(yg, 1, garg.long().unsqueeze(1)).squeeze(1)
    assert torch.equal(picked, gsel)                        # garg points at a row holding the selected value


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


@pytest.mark.parametrize("cin,kp,c1,c2,group,groups,dx", [(6, 8, 64, 64, 16, 512, False), (7, 8, 64, 128, 32, 300, False),
                                                          (195, 196, 128, 128, 32, 256, True), (195, 196, 128, 128, 64, 130, True),
                                                          (67, 68, 64, 64, 8, 3000, True), (131, 132, 128, 64, 16, 999, True)])
def test_fused_shared_mlp_matches_unfused_modules(cuda, cin, kp, c1, c2, group, groups, dx):
    """fused tensor-core forward+backward vs the same nn.Sequential evaluated with stock fp32 PyTorch ops"""
    from sg4d import mlp
    from sa_ref import shared_mlp_rows
    m_f = _mlp(cin, c1, c2, 3).to(cuda).train()
    m_r = copy.deepcopy(m_f)
    assert mlp.supported(m_f, kp, group)
    rows = groups * group
    g = torch.Generator().manual_seed(cin + group)
    x = torch.zeros(rows, kp)
    x[:, :cin] = torch.randn(rows, cin, generator=g)
    x[group:2 * group] = x[group].clone()               # one group made of duplicates of a row (padding case)
    xr = x.to(cuda).requires_grad_(dx)                  # reference layout [xyz | feats | 0]
    if dx:                                              # feature-first layout [feats | xyz | 0]
        xf = torch.cat([x[:, 3:cin], x[:, :3], x[:, cin:]], 1).to(cuda).requires_grad_(True)
    else:
        xf = x.to(cuda)
    out_f = mlp.fused_shared_mlp(xf, cin, group, m_f, xyz_last=dx)
    out_r = shared_mlp_rows(m_r, xr).view(groups, group, c2).amax(1)
    torch.testing.assert_close(out_f, out_r, rtol=0, atol=2e-5)
    wgt = torch.randn(groups, c2, generator=g).to(cuda)
    (out_f * wgt).sum().backward()
    (out_r * wgt).sum().backward()
    for (n_f, p_f), (_, p_r) in zip(m_f.named_parameters(), m_r.named_parameters()):
        # gradients are piecewise smooth (an arg-max flipping between two near-tied rows re-routes one channel's
        # gradient): relative L2 error + at most 2 % of entries outside the element-wise tolerance
        scale = max(1.0, p_r.grad.abs().max().item())
        diff = (p_f.grad - p_r.grad).abs()
        rel_l2 = (diff.double().norm() / p_r.grad.double().norm()).item()
        frac_bad = (diff > 2e-4 * scale).float().mean().item()
        assert rel_l2 < 2e-3 and frac_bad < 0.02, (n_f, rel_l2, frac_bad, diff.max().item())
    if dx:
        gf = torch.cat([xf.grad[:, cin - 3:cin], xf.grad[:, :cin - 3], xf.grad[:, cin:]], 1)   # back to [xyz|feats|0]
        gr = xr.grad
        tol = 2e-5 * max(1.0, gr.abs().max().item())
        keep = torch.ones(rows, dtype=torch.bool, device=cuda)
        keep[group:2 * group] = False
        # xyz columns carry no downstream gradient on the model path and are not computed
        bad = ((gf[keep][:, 3:cin] - gr[keep][:, 3:cin]).abs() > tol).float().mean().item()
        assert bad < 1e-3, bad    # an arg-max flip between near-tied rows changes the gradient of those rows only
        # duplicated rows: the reference's amax splits the gradient among ties, max_pool2d (and sg4d) route it to
        # one row -- the sum over the duplicates (all that reaches the source point) must agree
        torch.testing.assert_close(gf[~keep][:, 3:cin].sum(0), gr[~keep][:, 3:cin].sum(0), rtol=0, atol=10 * tol)
    for (n_f, b_f), (_, b_r) in zip(m_f.named_buffers(), m_r.named_buffers()):
        torch.testing.assert_close(b_f.float(), b_r.float(), rtol=1e-5, atol=1e-6, msg=lambda s_: n_f + ": " + s_)
    # eval mode (running statistics)
    m_f.eval(), m_r.eval()
    with torch.no_grad():
        torch.testing.assert_close(mlp.fused_shared_mlp(xf, cin, group, m_f, xyz_last=dx),
                                   shared_mlp_rows(m_r, xr).view(groups, group, c2).amax(1), rtol=0, atol=2e-5)
