"""oracle/weights.py -- TEST INFRASTRUCTURE.  Deterministic synthetic weights for a reference-layout
state_dict, generated from (key, shape) alone so that the container (where the reference Python runs
and the golden fixtures are made) and the GPU box (where only this repo exists) rebuild bit-identical
parameters without shipping 21 MB of tensors."""
import hashlib
import json
import os

import torch

_KEYS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "state_dict_keys.json")


def key_shapes(image=False):
    with open(_KEYS) as f:
        d = json.load(f)
    shapes = dict(d["no_gt"])
    if image:
        shapes.update(d["image_extra"])
    return shapes


def _gen(key, seed):
    h = hashlib.sha256(f"{seed}:{key}".encode()).digest()
    return torch.Generator().manual_seed(int.from_bytes(h[:7], "little"))


def synth_state_dict(shapes=None, seed=0, image=False):
    shapes = shapes or key_shapes(image)
    sd = {}
    for key, shape in shapes.items():
        g = _gen(key, seed)
        leaf = key.rsplit(".", 1)[1]
        if leaf == "num_batches_tracked":
            t = torch.zeros((), dtype=torch.int64)
        elif leaf == "running_mean":
            t = 0.1 * torch.randn(shape, generator=g)
        elif leaf == "running_var":
            t = 0.5 + torch.rand(shape, generator=g)
        elif len(shape) == 1 and leaf == "weight":      # BatchNorm scale: mixed signs on purpose
            t = (0.5 + torch.rand(shape, generator=g)) * torch.where(
                torch.rand(shape, generator=g) < 0.15, -1.0, 1.0)
        elif leaf == "bias":
            t = 0.1 * torch.randn(shape, generator=g)
        else:                                            # conv / linear weight
            fan_in = 1
            for s in shape[1:]:
                fan_in *= s
            t = torch.randn(shape, generator=g) * (1.5 / max(fan_in, 1)) ** 0.5
        sd[key] = t
    return sd
