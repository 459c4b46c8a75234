"""oracle/ref_gpu.py -- TEST / MEASUREMENT INFRASTRUCTURE, NOT PRODUCT CODE.

Runs the REFERENCE itself on the GPU: its own Python model (``SGPNModelWrapper`` and everything below it, imported
unmodified through oracle/ref_harness.py) over its own CUDA kernels compiled for sm_100a (oracle/_ref/pn2_ref_ext.so,
oracle/build_ref_ext.py).  bench.py times it next to sg4d on the same B200 and the same synthetic scenes
("Reference timed beside it"): this is the denominator of BASELINE.json's ">= 10x the reference
single-GPU scenes/s" target.  Third-party pieces that are not installed (pytorch_lightning, torch_geometric,
torch_scatter) are the harness's restatements, exactly as for the golden fixtures.
"""
import time

import torch

from . import build_ref_ext, ref_harness


def available():
    return ref_harness.available() and build_ref_ext.load_module() is not None


def build_model(device, seed=0):
    ext = build_ref_ext.load_module()
    if ext is None:
        raise RuntimeError("oracle/_ref/pn2_ref_ext.so is missing (built only where /root/reference is mounted)")
    ref_harness.install(ext)
    cfg = ref_harness.ref_config()
    cfg["MODEL"]["lambda_o"] = 0.1
    model = ref_harness.build_ref_model(cfg, seed=seed).to(device).train()
    model.reset_metrics()
    return model


def _step(model, batch, autocast):
    for p in model.parameters():
        p.grad = None
    with torch.autocast("cuda", dtype=torch.float16, enabled=autocast):
        loss = model.training_step(batch, 1)
    loss.backward()
    return loss


def time_scenes(model, scenes, reps=1, autocast=False, chunk=1):
    """scenes: list of device batch dicts (one scene each).  chunk = 1: one forward+backward per scene, like the
    reference's main.py (batch_size = 1); chunk > 1 is not offered by the reference's dataloader (its forward handles one
    scene) and is not timed here.  Returns (scenes/s, ms per scene)."""
    assert chunk == 1
    for sc in scenes[:1]:
        _step(model, sc, autocast)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    n = 0
    for _ in range(reps):
        for sc in scenes:
            _step(model, sc, autocast)
            n += 1
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    wall = (time.perf_counter() - t0) * 1e3 / n
    return 1e3 / ms, ms, wall
