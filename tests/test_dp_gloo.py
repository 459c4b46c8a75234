"""world_size-2 gloo test of the scene-sharded data-parallel plumbing (CPU, no kernels involved)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from sg4d import parallel
    r, w, _ = parallel.init_from_env("gloo")
    assert (r, w) == (rank, world)
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.ReLU(), torch.nn.Linear(16, 4))
    net.add_module("backbone_fc", torch.nn.Linear(4, 4))      # stands in for a parameter that never trains
    bucket = parallel.GradBucket(net, skip_substrings=("backbone_fc",))
    data = torch.arange(6 * 8, dtype=torch.float32).view(6, 8) / 10.0     # 6 "scenes"
    lo, hi = parallel.shard_range(6, rank, world)
    bucket.zero()
    net[:3](data[lo:hi]).pow(2).mean().backward()
    bucket.all_reduce_mean()
    out[rank] = bucket.flat.clone()
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gradient_allreduce_equals_full_batch():
    world, port = 2, _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    assert torch.equal(out[0], out[1])
    # single-process reference: mean over both shards' losses
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.ReLU(), torch.nn.Linear(16, 4))
    data = torch.arange(6 * 8, dtype=torch.float32).view(6, 8) / 10.0
    loss = 0.5 * (net(data[:3]).pow(2).mean() + net(data[3:]).pow(2).mean())
    loss.backward()
    want = torch.cat([p.grad.flatten() for p in net.parameters()])
    assert out[0].numel() == want.numel()          # the never-trained parameter is not in the bucket
    torch.testing.assert_close(out[0], want, rtol=1e-5, atol=1e-6)
