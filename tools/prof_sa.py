"""Developer tool: one fused SA scale forward+backward at benchmark shapes (for ncu captures / quick timings).
   python tools/prof_sa.py sa1|sa2 [clouds] [reps]"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sg4d import _lib, mlp
from sg4d.pointnet2_ops.pointnet2_modules import build_shared_mlp
dev = torch.device("cuda", 0)
torch.manual_seed(0)
kind = sys.argv[1] if len(sys.argv) > 1 else "sa1"
b = int(sys.argv[2]) if len(sys.argv) > 2 else 528
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
if kind == "sa1":
    n, m, ns, c = 80000, 512, 32, 4
    pts = torch.rand(b, n, 7, device=dev)
    feats, foff = pts, 3
    net = build_shared_mlp([7, 64, 128]).to(dev)
else:
    n, m, ns, c = 512, 128, 64, 192
    pts = torch.rand(b, n, 3, device=dev)
    feats, foff = torch.randn(b, n, c, device=dev).requires_grad_(True), 0
    net = build_shared_mlp([195, 128, 128]).to(dev)
centers = pts[:, :m, :3].contiguous()
idx = torch.randint(0, n, (b, m, ns), device=dev, dtype=torch.int32).sort(dim=2).values.contiguous()
cnt = torch.full((b, m), ns, device=dev, dtype=torch.int32)
_lib.enable_timing(True)
for r in range(reps):
    out = mlp.fused_sa_scale(kind, pts, feats, foff, c, centers, idx, cnt, net)
    out.backward(torch.ones_like(out))
    t = _lib.drain_timing()
for (name, key), ms in sorted(t.items(), key=lambda kv: -sum(kv[1])):
    print(f"{name:30s} {sum(ms) / len(ms):8.3f} ms")
