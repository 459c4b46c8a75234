"""Developer diagnostic (run on the GPU box): each backward kernel of csrc/mlp.cu against a PyTorch fp64
restatement of its formula on random operands.  Prints max abs errors; not part of the test-suite."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sg4d import _lib, mlp  # noqa: E402

dev = torch.device("cuda", 0)
torch.manual_seed(0)


def run(rows, n1, n2, group, kp):
    G = rows // group
    y1 = torch.randn(rows, n1, device=dev)
    y2 = torch.randn(rows, n2, device=dev)
    x = torch.randn(rows, kp, device=dev)
    w1 = torch.randn(n1, kp, device=dev) / kp ** 0.5
    w2 = torch.randn(n2, n1, device=dev) / n1 ** 0.5
    a2, b2 = 0.1 * torch.randn(n2, device=dev), 0.1 * torch.randn(n2, device=dev)
    dsel = torch.randn(G, n2, device=dev)
    garg = torch.randint(0, group, (G, n2), device=dev, dtype=torch.uint8)
    s1, t1 = torch.randn(n1, device=dev), 0.2 * torch.randn(n1, device=dev)
    i1, m1 = torch.rand(n1, device=dev) + 0.5, 0.1 * torch.randn(n1, device=dev)
    em1 = (-m1 * i1).contiguous()
    # references (fp64)
    dy2 = -(a2.double() * y2.double() + b2.double())
    dy2.view(G, group, n2).scatter_add_(1, garg.long().unsqueeze(1), dsel.double().unsqueeze(1))
    a1 = torch.relu(y1.double() * s1.double() + t1.double())
    dz1_ref = (dy2 @ w2.double()) * (a1 > 0)
    yh1 = y1.double() * i1.double() + em1.double()
    sums_ref = torch.stack([dz1_ref.sum(0), (dz1_ref * yh1).sum(0)])
    dw2_ref = dy2.t() @ a1

    dz1 = torch.empty(rows, n1, device=dev)
    part = torch.empty(_lib.load().sg4d_mlp_grid(rows) * 128 * 2, dtype=torch.float64, device=dev)
    _lib.call("sg4d_pool_bwd_da", x, rows, n2, n1, group, y2.data_ptr(), a2.data_ptr(), b2.data_ptr(), dsel.data_ptr(),
              garg.data_ptr(), mlp.pack_weight(w2.t()).data_ptr(), y1.data_ptr(), s1.data_ptr(), t1.data_ptr(), i1.data_ptr(),
              em1.data_ptr(), dz1.data_ptr(), part.data_ptr())
    sums = torch.empty(2, n1, device=dev)
    _lib.call("sg4d_partial_sums", x, n1, part.numel() // 2, part.data_ptr(), sums.data_ptr())
    torch.cuda.synchronize()
    print(f"rows={rows} n1={n1} n2={n2} group={group} kp={kp}")
    print("  pool_bwd_da  dz1 err %.3e (scale %.2f)  sums err %.3e (scale %.1f)" % (
        (dz1.double() - dz1_ref).abs().max().item(), dz1_ref.abs().max().item(),
        (sums.double() - sums_ref).abs().max().item(), sums_ref.abs().max().item()))
    dw2 = torch.empty(n2, n1, device=dev)
    _lib.call("sg4d_pool_bwd_dw", x, rows, n2, n1, group, y2.data_ptr(), a2.data_ptr(), b2.data_ptr(), dsel.data_ptr(),
              garg.data_ptr(), y1.data_ptr(), s1.data_ptr(), t1.data_ptr(), mlp._wgrad_partial(rows, n1, dev).data_ptr(),
              dw2.data_ptr())
    torch.cuda.synchronize()
    print("  pool_bwd_dw  dW2 err %.3e (scale %.1f)" % ((dw2.double() - dw2_ref).abs().max().item(), dw2_ref.abs().max().item()))
    d = dw2.double()
    print("     |dW2|max %.3e  nan %d  corr(ref) %.4f corr(ref^T) %.4f" % (
        d.abs().max().item(), int(torch.isnan(d).sum()),
        float((d * dw2_ref).sum() / (d.norm() * dw2_ref.norm() + 1e-30)),
        float((d * dw2_ref.t()).sum() / (d.norm() * dw2_ref.norm() + 1e-30)) if n1 == n2 else float("nan")))
    print("     got[0,:4]", d[0, :4].tolist(), " ref[0,:4]", dw2_ref[0, :4].tolist())
    # layer 1
    p1, q1, u1 = torch.randn(n1, device=dev), 0.1 * torch.randn(n1, device=dev), 0.1 * torch.randn(n1, device=dev)
    dz1r = torch.randn(rows, n1, device=dev)
    dy1 = p1.double() * dz1r.double() - (q1.double() * y1.double() + u1.double())
    dw1_ref = dy1.t() @ x.double()
    npad = 32 if kp <= 32 else (64 if kp <= 64 else (128 if kp <= 128 else 224))
    dw1 = torch.empty(n1, kp, device=dev)
    _lib.call("sg4d_inner_bwd_dw", x, rows, n1, kp, kp, y1.data_ptr(), dz1r.data_ptr(), p1.data_ptr(), q1.data_ptr(),
              u1.data_ptr(), x.data_ptr(), mlp._wgrad_partial(rows, npad, dev).data_ptr(), dw1.data_ptr())
    torch.cuda.synchronize()
    print("  inner_bwd_dw dW1 err %.3e (scale %.1f)" % ((dw1.double() - dw1_ref).abs().max().item(), dw1_ref.abs().max().item()))
    if kp >= 64:
        n = 64
        dx = torch.zeros(rows, kp, device=dev)
        _lib.call("sg4d_inner_bwd_dx", x, rows, n1, n, y1.data_ptr(), dz1r.data_ptr(), p1.data_ptr(), q1.data_ptr(),
                  u1.data_ptr(), mlp.pack_weight(w1[:, :n].t()).data_ptr(), dx.data_ptr(), kp, 0)
        torch.cuda.synchronize()
        dx_ref = dy1 @ w1[:, :n].double()
        print("  inner_bwd_dx dX  err %.3e (scale %.1f)" % ((dx[:, :n].double() - dx_ref).abs().max().item(), dx_ref.abs().max().item()))


for cfg in [(128, 128, 128, 32, 128), (1024, 128, 128, 32, 196), (2048, 64, 64, 16, 8), (4096, 64, 128, 32, 8), (128 * 300 + 64, 128, 64, 64, 132)]:
    run(*cfg)
