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
