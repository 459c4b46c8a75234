"""oracle/frontend_ref.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

numpy restatement of the reference's crop / sample step, SGH/dataset/data_preparation_utils.py:
  :104-125  per object: members of mask i + 1, bounding box -+ padding, down-sample, zero_mean
  :178-218  per edge: union of the two padded boxes, points STRICTLY inside, 4th feature mask1 * 1 + mask2 * 2, down-sample,
            zero_mean
  :12-18    zero_mean (torch.mean over the points, max of the row norms)
  :37-39    calculate_downsample_indices, the `np.random.choice(len, target, replace=True)` branch -- with the draws made
            explicit: index = candidates[min(floor(u * len), len - 1)] for given uniforms u (fp32 product), so that the GPU
            front-end can be checked bit for bit on the selected indices.
The other branch (len >= target: open3d voxel trace, then a draw WITHOUT replacement, :41-49) needs open3d, which is absent
offline: it is not restated, and the GPU front-end draws with replacement there as well.  Parity is therefore
"restatement only" for this step: the reference module cannot be imported here (open3d, helpers.configurations).
"""
import numpy as np
import torch


def zero_mean(point):
    """:12-18, on a (n, 3) float32 torch tensor."""
    mean = torch.mean(point, dim=0)
    point = point - mean.unsqueeze(0)
    dist = point.pow(2).sum(1).sqrt().max()
    return point / dist, mean, dist


def _draw(u, n_candidates):
    k = (u.astype(np.float32) * np.float32(n_candidates)).astype(np.int64)      # fp32 product, truncated
    return np.minimum(k, n_candidates - 1)


def prepare_scene(points, masks, n_obj, edges, u_obj, u_rel, padding=0.2):
    """points (P, S) float32, masks (P,) int, edges (2, E), u_* uniforms.  Returns dict of numpy / torch results."""
    points = np.asarray(points, dtype=np.float32)
    masks = np.asarray(masks)
    obj_points, obj_picked, boxes = [], [], []
    for i in range(n_obj):
        members = np.where(masks == i + 1)[0]                                    # :105
        ps = points[members]
        boxes.append((ps[:, :3].min(0) - np.float32(padding), ps[:, :3].max(0) + np.float32(padding)))   # :106-108
        choice = _draw(u_obj[i], len(members))                                   # :39 with explicit draws
        sel = members[choice]
        cloud = torch.from_numpy(points[sel].copy())
        cloud[:, :3], _, _ = zero_mean(cloud[:, :3])                             # :113
        obj_points.append(cloud)
        obj_picked.append(sel)
    rel_points, rel_picked, rel_counts = [], [], []
    for e in range(edges.shape[1]):
        a, b = int(edges[0, e]), int(edges[1, e])
        mask_ = ((masks == a + 1).astype(np.int32) * 1 + (masks == b + 1).astype(np.int32) * 2)[:, None]   # :188-190
        lo = np.minimum(boxes[a][0], boxes[b][0])                                # :193-196
        hi = np.maximum(boxes[a][1], boxes[b][1])
        inside = (points[:, 0] > lo[0]) & (points[:, 0] < hi[0]) & (points[:, 1] > lo[1]) & (points[:, 1] < hi[1]) & \
                 (points[:, 2] > lo[2]) & (points[:, 2] < hi[2])                 # :197-199
        cand = np.where(inside)[0]
        choice = _draw(u_rel[e], len(cand))
        sel = cand[choice]
        p4 = np.concatenate([points, mask_.astype(np.float32)], 1)               # :201
        cloud = torch.from_numpy(p4[sel].copy())
        cloud[:, :3], _, _ = zero_mean(cloud[:, :3])                             # :207
        rel_points.append(cloud)
        rel_picked.append(sel)
        rel_counts.append(len(cand))
    return {"obj_points": torch.stack(obj_points), "rel_points": torch.stack(rel_points),
            "obj_picked": np.stack(obj_picked), "rel_picked": np.stack(rel_picked), "edge_totals": np.asarray(rel_counts),
            "obj_box": np.stack([np.concatenate(bx) for bx in boxes])}
