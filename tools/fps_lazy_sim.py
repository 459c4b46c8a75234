"""Design study: lazy bucket refresh for FPS (a bucket's stale max-min-distance is an upper bound; only buckets
whose bound can still win the round are brought up to date).  Counts point visits and bound checks."""
import sys
import numpy as np
import torch
sys.path.insert(0, __file__.rsplit("/", 2)[0])
from sg4d import synthetic  # noqa: E402
from tools.fps_prune_sim import morton  # noqa: E402


def simulate(xyz, m, bucket=64, bits=5):
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
    ub = temp.max(1)
    ver = np.zeros(nb, dtype=np.int64)      # centres applied so far
    centres = [xyz[0]]
    visits = checks = iters = 0
    for j in range(1, m):
        while True:
            fresh = ver == j
            if fresh.any():
                need = (~fresh) & (ub >= ub[fresh].max())
            else:                                   # nothing is up to date yet: start with the 8 largest bounds
                need = np.zeros(nb, dtype=bool)
                need[np.argsort(-ub)[:8]] = True
            iters += 1
            if not need.any():
                break
            for b in np.nonzero(need)[0]:
                cs = np.array(centres[ver[b]:j])
                e = np.maximum(0, np.maximum(blo[b] - cs, cs - bhi[b]))
                bound = (e * e).sum(1)
                checks += len(cs)
                use = cs[~(bound >= ub[b])]
                if len(use):
                    visits += 1
                    d = ((pb[b][None] - use[:, None]) ** 2).sum(2).astype(np.float32).min(0)
                    temp[b] = np.minimum(temp[b], d)
                    ub[b] = temp[b].max()
                ver[b] = j
        b = int(ub.argmax())
        centres.append(pb[b, int(temp[b].argmax())])
    return nb, visits, checks, iters


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 80000
    for seed in range(3):
        gen = torch.Generator().manual_seed(1234 + seed)
        cloud = synthetic.make_cloud(gen, n, 6).numpy()[:, :3].astype(np.float32)
        nb, visits, checks, iters = simulate(cloud, 512)
        print(f"seed {seed}: buckets {nb}; bucket visits {visits} = {visits * 64 / n:.1f} N point visits; "
              f"bound checks {checks} ({checks / 511 / nb:.2f} per bucket-round); argmax iterations {iters / 511:.2f} per round")
