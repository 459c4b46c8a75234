"""Minimal trainer with the checkpoint / resume behaviour of the reference's ``main.py`` train mode
(SGP/main.py:24-33,47-66), for environments without pytorch-lightning:

* ``model.configure_optimizers()`` (AdamW, lr = LR, weight_decay = W_DECAY; SGH/model/scene_graph_prediction_model.py:240-242),
* one ``training_step`` per batch, ``validation_step`` over the validation batches after every epoch,
* a checkpoint per epoch named ``epoch=<N>.ckpt`` under ``<log_dir>/checkpoints`` (``ModelCheckpoint(filename='{epoch}',
  save_top_k=-1, every_n_epochs=1)``), holding Lightning's keys ``epoch``, ``global_step``, ``state_dict``,
  ``optimizer_states`` -- a reference checkpoint's ``state_dict`` loads into the sg4d model and vice versa,
* resume from the newest ``epoch=<N>.ckpt`` (``find_checkpoint_path``).

* per-epoch relation metrics with the reference's per-take bookkeeping (``metrics.RelationMetrics``: precision / recall / F1
  per relation and take, macro and weighted averages; scene_graph_prediction_model.py:113-132,195-238),
* ``precision=16``: dynamic loss scaling with ``torch.cuda.amp.GradScaler``'s semantics (``LossScaler``: what Lightning's
  ``precision=16`` wraps around the optimizer step, SGP/main.py:64) for reduced-precision compute paths.

Multi-GPU: pass a ``parallel.GradBucket``; the gradients are all-reduced (mean) once per step.
"""
import glob
import os
import re

import torch


class LossScaler:
    """Dynamic loss scaling with the defaults and update rule of ``torch.cuda.amp.GradScaler`` (init_scale 2**16, growth 2x
    after 2000 consecutive finite steps, back-off 0.5 and a SKIPPED optimizer step when a gradient is inf / nan)."""

    def __init__(self, init_scale=2.0 ** 16, growth_factor=2.0, backoff_factor=0.5, growth_interval=2000, enabled=True):
        self.scale_value, self.growth_factor, self.backoff_factor = float(init_scale), growth_factor, backoff_factor
        self.growth_interval, self.enabled, self._good_steps = growth_interval, enabled, 0

    def scale(self, loss):
        return loss * self.scale_value if self.enabled else loss

    def step(self, optimizer, params=None):
        """Unscale the gradients in place, skip the step if any is non-finite, update the scale.  Returns True if stepped."""
        if not self.enabled:
            optimizer.step()
            return True
        params = [p for g in optimizer.param_groups for p in g["params"]] if params is None else list(params)
        grads = [p.grad for p in params if p.grad is not None]
        inv = 1.0 / self.scale_value
        finite = torch.ones((), dtype=torch.bool, device=grads[0].device) if grads else None
        for g in grads:
            g.mul_(inv)
            finite = finite & torch.isfinite(g).all()
        ok = bool(finite) if grads else True          # one host read per step, like GradScaler's found_inf
        if ok:
            optimizer.step()
            self._good_steps += 1
            if self._good_steps == self.growth_interval:
                self.scale_value *= self.growth_factor
                self._good_steps = 0
        else:
            self.scale_value *= self.backoff_factor
            self._good_steps = 0
        return ok

    def state_dict(self):
        return {"scale": self.scale_value, "growth_factor": self.growth_factor, "backoff_factor": self.backoff_factor,
                "growth_interval": self.growth_interval, "_growth_tracker": self._good_steps}

    def load_state_dict(self, sd):
        self.scale_value, self._good_steps = float(sd["scale"]), int(sd.get("_growth_tracker", 0))


def find_checkpoint_path(log_dir):
    """Newest ``epoch=<N>.ckpt`` in ``<log_dir>/checkpoints`` or None (SGP/main.py:24-33)."""
    best, best_epoch = None, -1
    for p in glob.glob(os.path.join(log_dir, "checkpoints", "*.ckpt")):
        m = re.search(r"=(\d+)\.ckpt$", os.path.basename(p))
        if m and int(m.group(1)) > best_epoch:
            best, best_epoch = p, int(m.group(1))
    return best


def _rank_world():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def save_checkpoint(path, model, optimizer, epoch, global_step):
    """Rank 0 writes (to a temporary file, then an atomic rename: a concurrent reader never sees a torn file);
    every rank waits for it."""
    rank, world = _rank_world()
    if rank == 0:
        os.makedirs(os.path.dirname(path), exist_ok=True)
        tmp = path + ".tmp"
        torch.save({"epoch": epoch, "global_step": global_step, "state_dict": model.state_dict(),
                    "optimizer_states": [optimizer.state_dict()]}, tmp)
        os.replace(tmp, path)
    if world > 1:
        import torch.distributed as dist
        dist.barrier()


def load_checkpoint(path, model, optimizer=None, map_location=None):
    """Loads a checkpoint written by this trainer or by Lightning (same keys); a bare ``state_dict`` file (the
    reference's paper weights, SGP/main.py:78) is accepted too.  Returns (epoch, global_step)."""
    ckpt = torch.load(path, map_location=map_location, weights_only=False)
    if "state_dict" not in ckpt:
        model.load_state_dict(ckpt)
        return -1, 0
    model.load_state_dict(ckpt["state_dict"])
    if optimizer is not None and ckpt.get("optimizer_states"):
        optimizer.load_state_dict(ckpt["optimizer_states"][0])
    return int(ckpt.get("epoch", -1)), int(ckpt.get("global_step", 0))


def fit(model, train_batches, val_batches=None, max_epochs=1, log_dir=None, bucket=None, on_epoch_end=None, precision=32,
        metrics=None):
    """Trains ``model`` (anything with ``training_step`` / ``validation_step`` / ``configure_optimizers``).

    ``train_batches`` / ``val_batches``: callables returning an iterable of batch dicts for one epoch (device
    tensors), or plain iterables.  ``precision=16``: dynamic loss scaling (``LossScaler``).  ``metrics``: a
    ``metrics.RelationMetrics``; the model's steps then record their relation predictions per take and every epoch record
    gets ``train_macro_f1`` / ``val_macro_f1`` (the reference's Epoch_Macro/*_F1).  Returns a list of per-epoch dicts.
    """
    optimizer = model.configure_optimizers()
    scaler = LossScaler(enabled=(precision == 16))
    if metrics is not None:
        model.metrics = metrics
    start_epoch, global_step = 0, 0
    if log_dir is not None:
        ckpt = find_checkpoint_path(log_dir)
        if ckpt is not None:
            last, global_step = load_checkpoint(ckpt, model, optimizer)
            start_epoch = last + 1
    history = []
    if bucket is not None:
        bucket.broadcast_state(model)          # every rank starts from rank 0's parameters and BatchNorm buffers

    def batches_of(src):
        return src() if callable(src) else src

    for epoch in range(start_epoch, max_epochs):
        model.train()
        tot, cnt = None, 0           # the loss is accumulated on the device and read back once per epoch
        for i, batch in enumerate(batches_of(train_batches)):
            if bucket is not None:
                bucket.zero()
            else:
                optimizer.zero_grad(set_to_none=True)
            loss = model.training_step(batch, i)
            scaler.scale(loss).backward()
            if bucket is not None:
                if bucket._hooks:
                    bucket.drop_unused()
                bucket.all_reduce_mean()
            scaler.step(optimizer)
            global_step += 1
            tot, cnt = loss.detach() if tot is None else tot + loss.detach(), cnt + 1
        rec = {"epoch": epoch, "train_loss": (float(tot) if tot is not None else 0.0) / max(cnt, 1), "global_step": global_step}
        if val_batches is not None:
            model.eval()
            with torch.no_grad():
                vt, vc = 0.0, 0
                for i, batch in enumerate(batches_of(val_batches)):
                    vt, vc = vt + float(model.validation_step(batch, i)), vc + 1
            rec["val_loss"] = vt / max(vc, 1)
        if metrics is not None:
            rec["train_macro_f1"] = metrics.evaluate("train")["macro_f1"]
            metrics.reset("train")
            if val_batches is not None:
                rec["val_macro_f1"] = metrics.evaluate("val")["macro_f1"]
                metrics.reset("val")
        if log_dir is not None:
            save_checkpoint(os.path.join(log_dir, "checkpoints", f"epoch={epoch}.ckpt"), model, optimizer, epoch, global_step)
        if on_epoch_end is not None:
            on_epoch_end(rec)
        history.append(rec)
    return history
