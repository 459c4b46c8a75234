"""Design study for the bucket-pruned FPS kernel (csrc/fps_bucket.cu): counts, on the benchmark's
synthetic clouds, how many spatial buckets a round really has to touch when a bucket is skipped
whenever its bounding-box distance to the new pick is >= its current maximum min-distance.

    python tools/fps_prune_sim.py [n_points] [bucket] [bits]
"""
import sys

import numpy as np
import torch

sys.path.insert(0, __file__.rsplit("/", 2)[0])
from sg4d import synthetic  # noqa: E402


def morton(q, bits):
    key = np.zeros(len(q), dtype=np.int64)
    for b in range(bits):
        for a in range(3):
            key |= ((q[:, a] >> b) & 1) << (3 * b + a)
    return key


def simulate(xyz, m, bucket, bits):
    n = len(xyz)
    lo, hi = xyz.min(0), xyz.max(0)
    q = np.clip(((xyz - lo) / np.maximum(hi - lo, 1e-20) * (1 << bits)).astype(np.int64), 0, (1 << bits) - 1)
    order = np.argsort(morton(q, bits), kind="stable")
    p = xyz[order]
    nb = (n + bucket - 1) // bucket
    pad = nb * bucket - n
    pp = np.concatenate([p, np.repeat(p[-1:], pad, 0)]) if pad else p
    pb = pp.reshape(nb, bucket, 3)
    blo, bhi = pb.min(1), pb.max(1)
    temp = np.full((nb, bucket), 1e10, dtype=np.float32)
    bmax = temp.max(1)
    c = xyz[0]
    active_hist, changed_hist = [], []
    for j in range(1, m):
        e = np.maximum(0, np.maximum(blo - c, c - bhi))
        bound = (e * e).sum(1)
        act = ~(bound >= bmax)
        d = ((pb[act] - c) ** 2).sum(2).astype(np.float32)
        new = np.minimum(temp[act], d)
        changed_hist.append(int((new < temp[act]).sum()))
        temp[act] = new
        bmax[act] = new.max(1)
        active_hist.append(int(act.sum()))
        b = int(bmax.argmax())
        c = pb[b, int(temp[b].argmax())]
    return nb, np.array(active_hist), np.array(changed_hist)


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 80000
    bucket = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    bits = int(sys.argv[3]) if len(sys.argv) > 3 else 5
    tot = []
    for seed in range(6):
        gen = torch.Generator().manual_seed(1234 + seed)
        cloud = synthetic.make_cloud(gen, n, 6).numpy()[:, :3].astype(np.float32)
        nb, act, chg = simulate(cloud, 512, bucket, bits)
        tot.append(act.sum() * bucket / n)
        print(f"seed {seed}: buckets {nb}; active/round mean {act.mean():.1f} median {np.median(act):.0f} "
              f"max {act.max()} last100 {act[-100:].mean():.1f}; point visits = {act.sum() * bucket / n:.1f} N "
              f"(brute force 511 N); really changed {chg.sum() / n:.1f} N")
    print("mean visits / N:", np.mean(tot))
