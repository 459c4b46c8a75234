"""Developer tool: run the MLP kernels once at benchmark shapes (for ncu captures on the GPU box)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sg4d import _lib, mlp
dev = torch.device("cuda", 0)
torch.manual_seed(0)
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 528 * 512 * 32
which = sys.argv[2] if len(sys.argv) > 2 else "fwd2"
n1, n2, group = 64, 128, 32
y1 = torch.randn(rows, n1, device=dev)
w2 = torch.randn(n2, n1, device=dev) / 8
s1, t1 = torch.randn(n1, device=dev), torch.randn(n1, device=dev)
g2 = torch.randn(n2, device=dev)
img = mlp.pack_weight(w2)
for _ in range(3):
    if which == "fwd2":
        y2, part, gsel, garg = mlp.linear_fwd(y1, n1, img, n2, scale=s1, shift=t1, group=group, gamma=g2)
    elif which == "fwd1":
        x = torch.randn(rows, 8, device=dev) if _ == 0 else x
        w1 = torch.randn(n1, 8, device=dev)
        y, part, _, _ = mlp.linear_fwd(x, 8, mlp.pack_weight(w1), n1)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
if which == "fwd2":
    mlp.linear_fwd(y1, n1, img, n2, scale=s1, shift=t1, group=group, gamma=g2)
e1.record()
torch.cuda.synchronize()
print(which, rows, "ms", e0.elapsed_time(e1))

if which in ("dw1", "da", "dw2"):
    n1, n2, group, kp = 128, 128, 64, 196
    G = rows // group
    y1 = torch.randn(rows, n1, device=dev); y2 = torch.randn(rows, n2, device=dev); x = torch.randn(rows, kp, device=dev)
    dz1 = torch.randn(rows, n1, device=dev)
    p1, q1, u1 = torch.randn(n1, device=dev), torch.randn(n1, device=dev), torch.randn(n1, device=dev)
    a2, b2 = torch.randn(n2, device=dev), torch.randn(n2, device=dev)
    dsel = torch.randn(G, n2, device=dev); garg = torch.randint(0, group, (G, n2), device=dev, dtype=torch.uint8)
    w2 = torch.randn(n2, n1, device=dev)
    img = mlp.pack_weight(w2.t())
    dw = torch.empty(n1, kp, device=dev); dw2 = torch.empty(n2, n1, device=dev)
    part = torch.empty(_lib.load().sg4d_mlp_partial_doubles(rows), dtype=torch.float64, device=dev)
    wp = mlp._wgrad_partial(rows, 224, dev)
    def go():
        if which == "dw1":
            _lib.call("sg4d_inner_bwd_dw", x, rows, n1, kp, kp, y1.data_ptr(), dz1.data_ptr(), p1.data_ptr(), q1.data_ptr(),
                      u1.data_ptr(), x.data_ptr(), wp.data_ptr(), dw.data_ptr())
        elif which == "dw2":
            _lib.call("sg4d_pool_bwd_dw", x, rows, n2, n1, group, y2.data_ptr(), a2.data_ptr(), b2.data_ptr(), dsel.data_ptr(),
                      garg.data_ptr(), y1.data_ptr(), p1.data_ptr(), q1.data_ptr(), wp.data_ptr(), dw2.data_ptr())
        else:
            _lib.call("sg4d_pool_bwd_da", x, rows, n2, n1, group, y2.data_ptr(), a2.data_ptr(), b2.data_ptr(), dsel.data_ptr(),
                      garg.data_ptr(), img.data_ptr(), y1.data_ptr(), p1.data_ptr(), q1.data_ptr(), p1.data_ptr(), u1.data_ptr(),
                      dz1.data_ptr(), part.data_ptr())
    for _ in range(3):
        go()
    torch.cuda.synchronize()
    e0.record(); go(); e1.record(); torch.cuda.synchronize()
    print(which, rows, "ms", e0.elapsed_time(e1))
