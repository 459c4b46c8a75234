"""Times the SA1 index ops (spatial index build, FPS, ball query) old vs indexed on benchmark-shaped clouds.

    python tools/prof_index.py [clouds] [points] [stride]
"""
import sys

import torch

sys.path.insert(0, __file__.rsplit("/", 2)[0])
from sg4d import rows  # noqa: E402


def gpu_clouds(b, n, stride, dev, seed=0):
    """GPU re-statement of synthetic.make_cloud (Gaussian mixture, unit sphere) -- timing data only"""
    g = torch.Generator(device=dev).manual_seed(seed)
    centres = torch.rand(b, 4, 3, generator=g, device=dev) - 0.5
    sigma = 0.05 + 0.25 * torch.rand(b, 4, 1, generator=g, device=dev)
    comp = torch.randint(0, 4, (b, n), generator=g, device=dev)
    xyz = torch.gather(centres, 1, comp.unsqueeze(-1).expand(-1, -1, 3)) + \
        torch.gather(sigma, 1, comp.unsqueeze(-1)) * torch.randn(b, n, 3, generator=g, device=dev).clamp_(-3, 3)
    xyz -= xyz.mean(1, keepdim=True)
    xyz /= xyz.pow(2).sum(2).sqrt().amax(1, keepdim=True).unsqueeze(-1)
    pts = torch.rand(b, n, stride, generator=g, device=dev)
    pts[:, :, :3] = xyz
    return pts.contiguous()


def timeit(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, out


if __name__ == "__main__":
    b = int(sys.argv[1]) if len(sys.argv) > 1 else 528
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 80000
    stride = int(sys.argv[3]) if len(sys.argv) > 3 else 7
    dev = torch.device("cuda", 0)
    pts = gpu_clouds(b, n, stride, dev)
    m, radii, nss = 512, [0.1, 0.2], [16, 32]
    t_build, index = timeit(lambda: rows.SpatialIndex(pts))
    print(f"index build          {t_build:8.3f} ms")
    t_old, (idx_old, ctr) = timeit(lambda: rows.fps_rows(pts, m), reps=1)
    print(f"fps (on-chip)        {t_old:8.3f} ms")

    def fps_new():
        return rows.fps_rows(pts, m, rows.SpatialIndex(pts))
    t_new, (idx_new, ctr2) = timeit(fps_new)
    print(f"fps (build+indexed)  {t_new:8.3f} ms   -> indexed alone {t_new - t_build:8.3f} ms; equal: "
          f"{torch.equal(idx_old, idx_new) and torch.equal(ctr, ctr2)}")
    t_bq, (i0, c0) = timeit(lambda: rows.ball_query_rows(ctr, pts, radii, nss), reps=1)
    print(f"ball query (brute)   {t_bq:8.3f} ms")
    for prefix in (0, 1024, 2048, 4096, 8192, 16384):
        t, (i1, c1) = timeit(lambda: rows.ball_query_rows(ctr, pts, radii, nss, index, prefix=prefix))
        ok = all(torch.equal(a, b_) for a, b_ in zip(i0, i1)) and all(torch.equal(a, b_) for a, b_ in zip(c0, c1))
        unfinished = float(((c1[0] < nss[0]) | (c1[1] < nss[1])).float().mean())
        print(f"ball query (prefix {prefix:5d}) {t:8.3f} ms  equal: {ok}  centres short of nsample overall: {unfinished:.3f}")
