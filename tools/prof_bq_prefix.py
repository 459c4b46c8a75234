"""Developer tool: ball-query time (528 clouds x 80000 points, 512 centres, radii 0.1 / 0.2) vs the brute-force prefix length."""
import sys, torch
sys.path.insert(0, __file__.rsplit("/", 2)[0])
from sg4d import rows
from tools.prof_index import gpu_clouds, timeit
dev = torch.device("cuda", 0)
pts = gpu_clouds(528, 80000, 7, dev)
index = rows.SpatialIndex(pts)
_, new_xyz = rows.fps_rows(pts, 512, index)
ref = None
for prefix in (0, 512, 1024, 2048, 4096, 8192, 16384):
    ms, out = timeit(lambda: rows.ball_query_rows(new_xyz, pts, [0.1, 0.2], [16, 32], index, prefix=prefix))
    if ref is None:
        ref = out
    same = all(torch.equal(a, b) for a, b in zip(out[0], ref[0]))
    print(f"prefix {prefix:6d}: {ms:7.3f} ms  identical={same}")
