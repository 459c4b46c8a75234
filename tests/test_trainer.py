"""Trainer / checkpoint logic mirroring the reference's main.py train mode (SGP/main.py:24-33,61-66): per-epoch
``epoch=<N>.ckpt`` files with Lightning's keys, resume from the newest one.  CPU only: a stub model stands in for the
network (the real model needs the GPU)."""
import os

import torch
import torch.nn as nn

from sg4d import trainer


class Stub(nn.Module):
    def __init__(self):
        super().__init__()
        torch.manual_seed(0)
        self.net = nn.Sequential(nn.Linear(4, 8), nn.ReLU(), nn.Linear(8, 3))
        self.dead = nn.Linear(2, 2)                       # never used: grad stays None, AdamW must skip it

    def training_step(self, batch, batch_idx=0):
        return nn.functional.cross_entropy(self.net(batch["x"]), batch["y"])

    validation_step = training_step

    def configure_optimizers(self):
        return torch.optim.AdamW(self.parameters(), lr=3e-3, weight_decay=1e-3)


def _batches(seed, n=5):
    g = torch.Generator().manual_seed(seed)
    return [{"x": torch.randn(6, 4, generator=g), "y": torch.randint(0, 3, (6,), generator=g)} for _ in range(n)]


def test_checkpoints_and_resume_reproduce_an_uninterrupted_run(tmp_path):
    train, val = _batches(1), _batches(2, 2)
    a = Stub()
    hist = trainer.fit(a, train, val, max_epochs=3, log_dir=str(tmp_path / "a"))
    assert [h["epoch"] for h in hist] == [0, 1, 2] and hist[-1]["global_step"] == 15 and "val_loss" in hist[0]
    assert sorted(os.listdir(tmp_path / "a" / "checkpoints")) == ["epoch=0.ckpt", "epoch=1.ckpt", "epoch=2.ckpt"]
    assert trainer.find_checkpoint_path(str(tmp_path / "a")).endswith("epoch=2.ckpt")
    ck = torch.load(tmp_path / "a" / "checkpoints" / "epoch=2.ckpt", weights_only=False)
    assert {"epoch", "global_step", "state_dict", "optimizer_states"} <= set(ck) and ck["epoch"] == 2

    b = Stub()
    trainer.fit(b, train, val, max_epochs=2, log_dir=str(tmp_path / "b"))
    c = Stub()                                               # a fresh process would start like this ...
    hist_c = trainer.fit(c, train, val, max_epochs=3, log_dir=str(tmp_path / "b"))   # ... and resume at epoch 2
    assert [h["epoch"] for h in hist_c] == [2]
    for (k, va), (_, vc) in zip(a.state_dict().items(), c.state_dict().items()):
        assert torch.equal(va, vc), k
    assert a.dead.weight.grad is None


def test_find_checkpoint_orders_by_epoch_number(tmp_path):
    d = tmp_path / "checkpoints"
    d.mkdir()
    for e in (9, 10, 2):
        (d / f"epoch={e}.ckpt").write_bytes(b"")
    assert trainer.find_checkpoint_path(str(tmp_path)).endswith("epoch=10.ckpt")     # numeric, not lexicographic
    assert trainer.find_checkpoint_path(str(tmp_path / "none")) is None


def test_bare_state_dict_files_load(tmp_path):
    a, b = Stub(), Stub()
    with torch.no_grad():
        a.net[0].weight.add_(1.0)
    p = tmp_path / "paper_weights.pth"
    torch.save(a.state_dict(), p)                            # the reference's paper weights are bare state_dicts
    assert trainer.load_checkpoint(str(p), b) == (-1, 0)
    assert torch.equal(a.net[0].weight, b.net[0].weight)
