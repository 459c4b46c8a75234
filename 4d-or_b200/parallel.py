"""Scene-sharded data parallelism: one process per GPU, each rank runs the full model on its own
scenes, and the gradients of the live parameters are summed with ONE all-reduce per step
(NCCL over NVLink/NVSwitch on the GPU box, gloo in the CPU tests).

The reference has no multi-GPU path (``pl.Trainer(gpus=1)``, SGP/main.py:62); scenes are
independent units, so this is plain DDP-without-SyncBN semantics (BatchNorm statistics stay
per-rank).  The never-used ``backbone.fc_layer.*`` parameters (PN2/models/pointnet2_ssg_cls.py:87-96)
receive no gradient and are left out of the bucket: 3.89 M of the 5.23 M parameters travel.
"""
import os

import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """Initialise torch.distributed from torchrun's environment; returns (rank, world, local_rank)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
            dist.init_process_group(backend, device_id=torch.device("cuda", local_rank))
        else:
            dist.init_process_group(backend)
    return rank, world, local_rank


def shard_range(n_items, rank, world):
    """Contiguous block of scenes for ``rank`` (sizes differ by at most one)."""
    base, rem = divmod(n_items, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


class GradBucket:
    """Flat fp32 buffer over the gradients of the parameters that actually train."""

    def __init__(self, module, skip_substrings=("backbone.fc_layer.",)):
        self.params = [p for n, p in module.named_parameters()
                       if p.requires_grad and not any(s in n for s in skip_substrings)]
        self.numel = sum(p.numel() for p in self.params)
        p0 = self.params[0]
        self.flat = torch.zeros(self.numel, dtype=torch.float32, device=p0.device)
        # make every .grad a view into the flat buffer: no pack/unpack copies around the collective
        off = 0
        for p in self.params:
            n = p.numel()
            p.grad = self.flat[off:off + n].view_as(p)
            off += n

        # Parameters that never receive a gradient in this configuration (e.g. the last gconv's nn2 with
        # OBJ_PRED_FROM_GCN = false) must keep grad = None like on the single-GPU path -- a permanent zero view would
        # still be weight-decayed by AdamW.  Hooks mark who was touched by the first backward; drop_unused() unbinds the rest.
        self._touched = set()
        self._hooks = [p.register_post_accumulate_grad_hook(lambda q, s=self._touched: s.add(id(q))) for p in self.params]

    def drop_unused(self):
        """Call once after the first backward: parameters no gradient reached get grad = None (optimizers skip them)."""
        for h in self._hooks:
            h.remove()
        self._hooks = []
        for p in self.params:
            if id(p) not in self._touched:
                p.grad = None

    def broadcast_state(self, module, src=0):
        """Rank `src`'s parameters and buffers to every rank (call once before training)."""
        if dist.is_initialized() and dist.get_world_size() > 1:
            for t in list(module.parameters()) + list(module.buffers()):
                dist.broadcast(t.data, src=src)

    def zero(self):
        self.flat.zero_()

    def all_reduce_mean(self):
        """Sum over ranks then divide by world size (the loss of each rank is a mean over its scenes)."""
        if dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)
            self.flat.div_(dist.get_world_size())
        return self.flat


class DevicePrefetcher:
    """Double-buffered host -> device staging of batch dicts on a side stream, so that the upload of batch i+1
    overlaps the compute of batch i (the reference relies on Lightning's synchronous ``batch_to_device``).

        pf = DevicePrefetcher(device)
        pf.submit(host_batch)                 # pinned host tensors
        for ...:
            batch = pf.next()                 # device tensors; the current stream waits for their copy
            pf.submit(next_host_batch)        # starts copying while the step below runs
            step(batch)

    Point tensors ``*_points`` keep the (B, N, C) row layout underneath their (B, C, N) view, like
    ``synthetic.to_device``.  Two sets of device buffers are reused round-robin; a buffer is only overwritten
    after the compute stream has passed the ``next()`` that follows its last use.
    """

    def __init__(self, device, depth=2):
        self.device = torch.device(device)
        self.stream = torch.cuda.Stream(device=self.device)
        self.slots = [dict(buf={}, ready=None, free=None) for _ in range(depth)]
        self.head = self.tail = 0
        self.pending = 0

    @staticmethod
    def _base(k, v):
        return v.permute(0, 2, 1) if (v.dim() == 3 and k.endswith("_points")) else v

    def submit(self, host_batch):
        if self.pending == len(self.slots):
            raise RuntimeError("DevicePrefetcher: all buffers are in flight; call next() first")
        slot = self.slots[self.head]
        self.head = (self.head + 1) % len(self.slots)
        self.pending += 1
        out = {}
        with torch.cuda.stream(self.stream):
            if slot["free"] is not None:
                self.stream.wait_event(slot["free"])      # the compute stream is done with this buffer
            for k, v in host_batch.items():
                if not torch.is_tensor(v):
                    out[k] = v
                    continue
                src = self._base(k, v)
                dst = slot["buf"].get(k)
                if dst is None or dst.shape != src.shape or dst.dtype != src.dtype:
                    dst = torch.empty(src.shape, dtype=src.dtype, device=self.device)
                    slot["buf"][k] = dst
                dst.copy_(src, non_blocking=True)
                out[k] = self._base(k, dst)
            slot["ready"] = torch.cuda.Event()
            slot["ready"].record(self.stream)
        slot["out"] = out

    def next(self):
        if self.pending == 0:
            raise RuntimeError("DevicePrefetcher: nothing submitted")
        prev = self.slots[(self.tail - 1) % len(self.slots)]
        cur = torch.cuda.current_stream(self.device)
        if prev.get("out") is not None:
            # everything enqueued so far on the compute stream (the previous step) precedes the reuse of its buffer
            prev["free"] = torch.cuda.Event()
            prev["free"].record(cur)
        slot = self.slots[self.tail]
        self.tail = (self.tail + 1) % len(self.slots)
        self.pending -= 1
        cur.wait_event(slot["ready"])
        return slot["out"]
