"""CUDA-graphed training step: the ~260 C-ABI launches and ~500 small PyTorch launches of one forward + loss + backward are
captured ONCE and replayed as a single graph launch per step (the reference pays the per-launch host cost on every op:
at its real operating point -- batch_size = 1, 4000 / 8000 points, SGP/main.py:47-66 -- the step is host-bound).

    step = GraphedStep(model, example_batch, bucket)      # captures on the current device
    loss = step(batch)                                    # copies the batch into the static inputs, replays, returns the loss

Shapes are frozen by the example batch (scenes of the same object / edge / point counts); parameters, gradients
(``GradBucket``'s flat buffer), BatchNorm buffers and the RNG state of the dropout layers live at fixed addresses, so
optimizers and the NCCL all-reduce work on the same tensors as in eager mode.  Nothing in the step synchronises with the
host (``rows.EdgeCSR`` builds its CSR with sort + searchsorted), which is what makes it capturable.
"""
import torch


class GraphedStep:
    def __init__(self, model, example_batch, bucket=None, step_fn=None, warmup=3):
        self.model, self.bucket = model, bucket
        dev = next(model.parameters()).device
        self.static = {k: (v.clone() if torch.is_tensor(v) else v) for k, v in example_batch.items()}
        for k, v in example_batch.items():            # keep the (B, C, N) views over point-major storage
            if torch.is_tensor(v) and v.dim() == 3 and k.endswith("_points"):
                self.static[k] = v.permute(0, 2, 1).contiguous().permute(0, 2, 1)
        self._fn = step_fn or (lambda b: model.training_step(b))

        def run():
            if bucket is not None:
                bucket.zero()
            else:
                for p in model.parameters():
                    p.grad = None
            loss = self._fn(self.static)
            if loss.requires_grad:
                loss.backward()
            return loss.detach()

        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):                 # eager warm-up on a side stream (allocator / cuBLAS-free lazy state)
            for _ in range(warmup):
                run()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        if bucket is None:                            # gradients must exist at fixed addresses before the capture
            for p in model.parameters():
                if p.grad is not None:
                    p.grad.zero_()
        self.graph = torch.cuda.CUDAGraph()
        # captured on a HIGH-priority stream: kernel nodes keep the capture stream's priority, so work queued beside the replay
        # on ordinary streams (the next batch's upload / crop front-end) fills idle SMs instead of competing for them
        with torch.cuda.graph(self.graph, stream=torch.cuda.Stream(device=dev, priority=-1)):
            if bucket is not None:
                bucket.zero()
            else:
                for p in model.parameters():
                    if p.grad is not None:
                        p.grad.zero_()
            loss = self._fn(self.static)
            if loss.requires_grad:
                loss.backward()
            self.loss = loss.detach()

    def load(self, batch, non_blocking=True):
        """copy a batch (host or device tensors of the captured shapes) into the static inputs"""
        for k, v in batch.items():
            if torch.is_tensor(v):
                dst = self.static[k]
                if v.dim() == 3 and k.endswith("_points"):
                    dst.permute(0, 2, 1).copy_(v.permute(0, 2, 1), non_blocking=non_blocking)
                else:
                    dst.copy_(v, non_blocking=non_blocking)

    def __call__(self, batch=None):
        if batch is not None and batch is not self.static:
            self.load(batch)
        self.graph.replay()
        return self.loss
