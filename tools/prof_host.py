"""Host-side profile of one-scene steps (the reference's batch_size=1 mode): where does the Python time go?"""
import cProfile
import json
import os
import pstats
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sg4d import parallel, synthetic  # noqa: E402
from sg4d.model import SGPNModelWrapper  # noqa: E402

dev = torch.device("cuda", 0)
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
cfg = json.load(open(os.path.join(ROOT, "4d-or_b200", "default_config.json")))
torch.manual_seed(0)
model = SGPNModelWrapper(cfg, 12, 15, torch.ones(12), torch.ones(15), [f"r{i}" for i in range(14)] + ["none"]).to(dev).train()
bucket = parallel.GradBucket(model)
batch = synthetic.to_device(synthetic.make_scene(5, n_obj=9, n_points_obj=4000, n_points_rel=8000, pairs="ordered"), dev)


def step():
    bucket.zero()
    loss = model.training_step(batch)
    loss.backward()
    return loss


for _ in range(5):
    step()
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for _ in range(20):
    step()
torch.cuda.synchronize()
pr.disable()
st = pstats.Stats(pr)
st.sort_stats("tottime").print_stats(28)
