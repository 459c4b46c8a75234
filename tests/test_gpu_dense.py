"""GPU tests of the dense layers on the tensor-core engine (sg4d/dense.py, include/sg4d.h section 5) -- the GroupAll
level SA3, the TripletGCN MLPs and the classifier heads -- against fp64 PyTorch evaluations of the same modules
(nn.Linear / BatchNorm1d / ReLU: SGH/model/gcns/network_TripletGCN.py:11-27, SGH/model/pointnets/network_PointNet.py:
210-224).  Tolerance 1e-4 (relative L2 and worst element relative to the largest), ReLU / max-pool selections pinned."""
import pytest
import torch
import torch.nn as nn

import sa_ref

pytestmark = pytest.mark.gpu


def _close(name, got, want, tol=1e-4):
    e, w = sa_ref._errs(got, want)
    assert e <= tol and w <= tol, (name, e, w)


@pytest.mark.parametrize("rows,k,n,pre", [(528, 768, 512, False), (96, 512, 256, True), (66, 268, 15, False), (5000, 260, 256, False),
                                          (7, 256, 12, False), (300, 1036, 15, False), (12, 2048, 128, False), (129, 64, 64, True)])
def test_linear_matches_fp64(cuda, rows, k, n, pre):
    from sg4d import dense
    torch.manual_seed(rows + n)
    lin = nn.Linear(k, n).to(cuda)
    x = torch.randn(rows, k, device=cuda, requires_grad=True)
    y = dense.linear(x, lin, pre_relu=pre)
    w = torch.randn(rows, n, device=cuda)
    (y * w).sum().backward()
    xd = x.detach().double().requires_grad_(True)
    wd, bd = lin.weight.detach().double().requires_grad_(True), lin.bias.detach().double().requires_grad_(True)
    yd = (torch.relu(xd) if pre else xd) @ wd.t() + bd
    (yd * w.double()).sum().backward()
    _close("y", y.detach(), yd.detach())
    _close("dx", x.grad, xd.grad)
    _close("dw", lin.weight.grad, wd.grad)
    _close("db", lin.bias.grad, bd.grad)


@pytest.mark.parametrize("rows,k,n,pre", [(528, 768, 512, False), (528, 512, 1280, False), (96, 512, 512, False), (9, 512, 512, False),
                                          (3000, 128, 64, True)])
def test_linear_bn_relu_matches_fp64(cuda, rows, k, n, pre):
    from sg4d import dense
    torch.manual_seed(rows + n)
    lin = nn.Linear(k, n).to(cuda)
    bn = nn.BatchNorm1d(n, track_running_stats=False).to(cuda)
    with torch.no_grad():
        bn.weight.copy_(torch.randn(n))
        bn.bias.copy_(0.3 * torch.randn(n))
    x = torch.randn(rows, k, device=cuda, requires_grad=True)
    h = dense.linear_bn_relu(x, lin, bn, pre_relu=pre)
    w = torch.randn(rows, n, device=cuda)
    (h * w).sum().backward()
    xd = x.detach().double().requires_grad_(True)
    p = [t.detach().double().requires_grad_(True) for t in (lin.weight, lin.bias, bn.weight, bn.bias)]
    yd = (torch.relu(xd) if pre else xd) @ p[0].t() + p[1]
    z = (yd - yd.mean(0)) / torch.sqrt(yd.var(0, unbiased=False) + bn.eps) * p[2] + p[3]
    mask = h.detach() > 0                                   # pinned ReLU selection (near-zero activations)
    assert float((z.detach() * ((z.detach() > 0) != mask)).abs().max()) <= 1e-5
    hd = z * mask
    (hd * w.double()).sum().backward()
    _close("h", h.detach(), hd.detach())
    _close("dx", x.grad, xd.grad)
    _close("dw", lin.weight.grad, p[0].grad)
    _close("dgamma", bn.weight.grad, p[2].grad)
    _close("dbeta", bn.bias.grad, p[3].grad)
    assert lin.bias.grad is not None and float(lin.bias.grad.abs().max()) == 0.0      # cancels inside the BatchNorm


@pytest.mark.parametrize("cin,c1,c2,group,groups,dx", [(259, 256, 256, 128, 12, True), (259, 256, 256, 128, 150, True),
                                                       (8, 16, 24, 8, 64, True), (8, 16, 32, 16, 64, False), (67, 128, 256, 32, 40, True)])
def test_pooled_shared_mlp_matches_pinned_fp64(cuda, cin, c1, c2, group, groups, dx):
    """two [conv1x1 -> BN -> ReLU] blocks + max over `group` rows at widths outside the fused SA kernels (SA3: 259 -> 256 -> 256
    over 128 points; the operator API's narrow test shapes run zero-padded)"""
    from sg4d import dense, mlp
    from sg4d.pointnet2_ops.pointnet2_modules import build_shared_mlp
    torch.manual_seed(cin + c2)
    net = build_shared_mlp([cin, c1, c2]).to(cuda).train()
    with torch.no_grad():
        for layer in net:
            if isinstance(layer, nn.BatchNorm2d):
                layer.weight.copy_(torch.randn_like(layer.weight))
                layer.bias.copy_(0.2 * torch.randn_like(layer.bias))
    rows = groups * group
    kp = (cin + 3) // 4 * 4
    x = torch.zeros(rows, kp, device=cuda)
    x[:, :cin] = torch.randn(rows, cin, device=cuda)
    xr = x.clone()                                      # reference column order [xyz | feats | 0]
    xf = (torch.cat([x[:, 3:cin], x[:, :3], x[:, cin:]], 1) if dx else x).contiguous().requires_grad_(dx)
    mlp.CAPTURE = []
    try:
        out = dense.pooled_shared_mlp(xf, cin, group, net, xyz_last=dx)
        cap = [q for q in mlp.CAPTURE if "garg" in q][0]
    finally:
        mlp.CAPTURE = None
    w = torch.randn(out.shape, device=cuda)
    (out * w).sum().backward()
    xd = xr[:, :cin].double().requires_grad_(dx)
    params = [p_.detach().double().reshape(p_.shape[0], -1).squeeze(-1).requires_grad_(True) for p_ in
              (net[0].weight, net[1].weight, net[1].bias, net[3].weight, net[4].weight, net[4].bias)]
    h1m = sa_ref.h1_mask(cap, None)[:, :c1]
    pinned, free, _ = sa_ref.ref_scale(xd, params, group, cap["garg"][:, :c2], out.detach() > 0, h1m)
    assert float((free - pinned).detach().abs().max()) <= 1e-5
    (pinned * w.double()).sum().backward()
    _close("out", out.detach(), pinned.detach())
    got = [net[0].weight.grad.view(c1, -1), net[1].weight.grad, net[1].bias.grad, net[3].weight.grad.view(c2, c1), net[4].weight.grad,
           net[4].bias.grad]
    for nm, a, p_ in zip(("w1", "g1", "b1", "w2", "g2", "b2"), got, params):
        _close("d_" + nm, a, p_.grad)
    if dx:
        gf = torch.cat([xf.grad[:, cin - 3:cin], xf.grad[:, :cin - 3]], 1)      # back to [xyz | feats]
        _close("d_feats", gf[:, 3:], xd.grad[:, 3:])                            # the xyz columns carry no gradient downstream
    assert int(net[1].num_batches_tracked) == 1 and float(net[1].running_var.min()) > 0
