"""Developer tool: time the GPU front-end (8 scenes -> 624 clouds of 80000 points) and the step on its crops vs synthetic crops."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sg4d import _lib, frontend, synthetic
dev = torch.device("cuda", 0)
S, P, N = 8, 200000, 80000
raw = [synthetic.make_raw_scene(i, n_obj=12, n_points=P) for i in range(S)]
pts = [r[0].to(dev) for r in raw]; msk = [r[1].to(dev) for r in raw]
for rep in range(3):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    _lib.enable_timing(True); _lib.drain_timing()
    e0.record()
    scenes = [frontend.prepare_scene(pts[k], msk[k], 12, N, N, pairs="unordered") for k in range(S)]
    e1.record(); torch.cuda.synchronize()
    t = _lib.drain_timing(); _lib.enable_timing(False)
print("front-end, 8 scenes:", round(e0.elapsed_time(e1), 2), "ms")
agg = {}
for (name, key), ms in t.items():
    agg[name] = agg.get(name, 0) + sum(ms)
for k, v in sorted(agg.items(), key=lambda kv: -kv[1]):
    print(f"  {k:28s} {v:7.3f} ms")
tot = scenes[0]["_debug"]["obj_totals"] if "_debug" in scenes[0] else None
