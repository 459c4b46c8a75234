"""CPU (fp64) checks of the two algebraic restatements the CUDA path relies on, independent of any kernel:

* SA2: the first layer over grouped rows x[r] = [feats(i) | xyz(i) - centre_j] equals a gather of per-point projections,
  y1[r] = Z[i] - Cc[j]; its weight / feature gradients follow from G[i] = sum of dY1 over the rows referencing point i and
  H[j] = sum of dY1 over the rows of centre j.
* SA1: dW2 = dY2^T h1 with dY2 = [winner] dsel - (a2 y2 + b2), y2 = W2 h1  equals  T1 - diag(a2) W2 (h1^T h1) - b2 (x) colsum(h1).

The reference formulation is plain autograd over the materialised grouped tensor (OPS/pointnet2_utils.py:319-328 followed by
the first Conv2d of OPS/pointnet2_modules.py:9-19)."""
import torch


def test_sa2_first_layer_through_linearity():
    g = torch.Generator().manual_seed(0)
    b, n, m, ns, c, n1 = 2, 40, 7, 5, 8, 6
    feats = torch.randn(b, n, c, generator=g, dtype=torch.float64, requires_grad=True)
    xyz = torch.randn(b, n, 3, generator=g, dtype=torch.float64)
    centres = torch.randn(b, m, 3, generator=g, dtype=torch.float64)
    idx = torch.randint(0, n, (b, m, ns), generator=g)
    w1 = torch.randn(n1, 3 + c, generator=g, dtype=torch.float64, requires_grad=True)     # reference column order [xyz | feats]
    flat = (idx + torch.arange(b).view(b, 1, 1) * n).reshape(-1)

    # reference: materialised grouped rows
    gx = xyz.reshape(b * n, 3)[flat] - centres.reshape(b * m, 3).repeat_interleave(ns, dim=0)
    x = torch.cat([gx, feats.reshape(b * n, c)[flat]], dim=1)
    y1_ref = x @ w1.t()
    d_y1 = torch.randn(y1_ref.shape, generator=g, dtype=torch.float64)
    dw_ref, df_ref = torch.autograd.grad(y1_ref, (w1, feats), d_y1)

    # restated: per-point projection, gather, per-point / per-centre sums
    w = w1.detach()
    z = torch.cat([feats.detach().reshape(b * n, c), xyz.reshape(b * n, 3)], dim=1) @ torch.cat([w[:, 3:], w[:, :3]], dim=1).t()
    cc = centres.reshape(b * m, 3) @ w[:, :3].t()
    y1 = z[flat] - cc.repeat_interleave(ns, dim=0)
    torch.testing.assert_close(y1, y1_ref.detach(), rtol=1e-12, atol=1e-12)
    G = torch.zeros(b * n, n1, dtype=torch.float64).index_add_(0, flat, d_y1)
    H = d_y1.view(b * m, ns, n1).sum(1)
    d_feats = (G @ w[:, 3:]).view(b, n, c)
    d_w1 = torch.cat([G.t() @ xyz.reshape(b * n, 3) - H.t() @ centres.reshape(b * m, 3), G.t() @ feats.detach().reshape(b * n, c)], dim=1)
    torch.testing.assert_close(d_feats, df_ref, rtol=1e-12, atol=1e-12)
    torch.testing.assert_close(d_w1, dw_ref, rtol=1e-12, atol=1e-12)


def test_sa1_dw2_gram_form():
    g = torch.Generator().manual_seed(1)
    groups, ns, n2, k = 9, 4, 10, 6
    rows = groups * ns
    h1 = torch.relu(torch.randn(rows, k, generator=g, dtype=torch.float64))
    w2 = torch.randn(n2, k, generator=g, dtype=torch.float64)
    y2 = h1 @ w2.t()
    garg = torch.randint(0, ns, (groups, n2), generator=g)
    dsel = torch.randn(groups, n2, generator=g, dtype=torch.float64)
    a2, b2 = torch.randn(n2, generator=g, dtype=torch.float64), torch.randn(n2, generator=g, dtype=torch.float64)
    slot = torch.arange(rows) % ns
    grp = torch.arange(rows) // ns
    d_y2 = (garg[grp] == slot.view(-1, 1)).double() * dsel[grp] - (a2 * y2 + b2)      # the operand form (include/sg4d.h section 3)
    want = d_y2.t() @ h1
    win = (torch.arange(groups).view(-1, 1) * ns + garg)                               # (groups, n2) winning row of every (g, c)
    t1 = torch.einsum("gc,gck->ck", dsel, h1[win])
    got = t1 - a2.view(-1, 1) * (w2 @ (h1.t() @ h1)) - b2.view(-1, 1) * h1.sum(0).view(1, -1)
    torch.testing.assert_close(got, want, rtol=1e-11, atol=1e-11)
