"""Measurement tool (not product code): times the REFERENCE's own CUDA kernels (oracle/_ref/pn2_ref_ext.so = its
_ext-src sources compiled unmodified for sm_100a) next to the sg4d kernels on the same B200, same clouds, and checks
that the index outputs agree bit for bit.

    python tools/prof_ref_kernels.py [clouds] [points]        (defaults: one scene's 66 edge clouds x 80000 points)
"""
import json
import sys

import torch

sys.path.insert(0, __file__.rsplit("/", 2)[0])
from oracle import build_ref_ext  # noqa: E402
from sg4d import rows  # noqa: E402
from tools.prof_index import gpu_clouds, timeit  # noqa: E402

if __name__ == "__main__":
    b = int(sys.argv[1]) if len(sys.argv) > 1 else 66
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 80000
    ref = build_ref_ext.load_module()
    if ref is None:
        raise SystemExit("oracle/_ref/pn2_ref_ext.so is missing (built in the container by __graft_entry__.build())")
    dev = torch.device("cuda", 0)
    pts = gpu_clouds(b, n, 7, dev)
    xyz = pts[:, :, :3].contiguous()
    m, radii, nss = 512, [0.1, 0.2], [16, 32]
    out = {"clouds": b, "points": n}

    t_ref, idx_ref = timeit(lambda: ref.furthest_point_sampling(xyz, m), reps=2)

    def ours_fps():
        index = rows.SpatialIndex(pts) if rows.wants_index(n) else None
        return rows.fps_rows(pts, m, index) + (index,)
    t_new, (idx_new, ctr, index) = timeit(ours_fps, reps=3)
    out["fps"] = {"reference_ms": t_ref, "sg4d_ms_incl_index_build": t_new, "bit_exact": bool(torch.equal(idx_ref, idx_new))}

    def ref_bq():
        return [ref.ball_query(ctr, xyz, r, ns) for r, ns in zip(radii, nss)]
    t_ref, bq_ref = timeit(ref_bq, reps=2)
    t_new, (bq_new, cnt) = timeit(lambda: rows.ball_query_rows(ctr, pts, radii, nss, index), reps=3)
    out["ball_query_2_radii"] = {"reference_ms": t_ref, "sg4d_ms": t_new,
                                 "bit_exact": all(torch.equal(a, c) for a, c in zip(bq_ref, bq_new))}

    feats_cm = pts[:, :, 3:].transpose(1, 2).contiguous()
    xyz_cm = xyz.transpose(1, 2).contiguous()

    def ref_group():
        res = []
        for i in bq_ref:
            gx = ref.group_points(xyz_cm, i)
            gx -= ctr.transpose(1, 2).unsqueeze(-1)
            res.append(torch.cat([gx, ref.group_points(feats_cm, i)], dim=1))
        return res
    t_ref, g_ref = timeit(ref_group, reps=2)
    t_new, g_new = timeit(lambda: [rows.group_rows(pts, pts, ctr, i, c, 4, 3, 8) for i, c in zip(bq_new, cnt)], reps=3)
    same = all(torch.equal(a.permute(0, 2, 3, 1), c[..., :7]) for a, c in zip(g_ref, g_new))
    out["query_and_group_sa1"] = {"reference_ms": t_ref, "sg4d_ms": t_new, "bit_exact": bool(same)}
    print(json.dumps(out, indent=1))
