"""Synthetic scenes with the shapes and statistics of the reference's batch dict (SURVEY.md 3.5,
8(d)); no dataset or checkpoint is reachable offline.

Per cloud: xyz ~ mixture of 4 Gaussians (sigma in [0.05, 0.3], centres in [-0.5, 0.5]^3), then
centred and scaled to the unit sphere exactly like the reference's ``zero_mean``
(SGH/dataset/data_preparation_utils.py:12-18) so that the radii 0.1/0.2/0.4 keep their meaning;
rgb ~ U[0,1); the edge clouds' mask channel in {0,1,2,3} with P = (.2,.4,.3,.1).  10 % of the clouds
get their last quarter overwritten by exact duplicates of earlier rows (the ``replace=True``
down-sampling case, data_preparation_utils.py:38-39: exact FPS distance ties) and 5 % get 10 % of
their rows zeroed (augmentation_utils.py:54: the FPS |p|^2 <= 1e-3 skip rule).

Edges: ``pairs='unordered'`` takes every i<j once (BASELINE.json: 66 edges for 12 objects);
``pairs='ordered'`` builds the reference's n(n-1) list (data_preparation_utils.py:127-133).
All tensors are created on the CPU from ``torch.Generator(seed = 1234 + scene_id)``.
"""
import torch


def _cloud_xyz(gen, n):
    k = 4
    centres = torch.rand(k, 3, generator=gen) - 0.5
    sigma = 0.05 + 0.25 * torch.rand(k, 1, generator=gen)
    comp = torch.randint(0, k, (n,), generator=gen)
    xyz = centres[comp] + sigma[comp] * torch.randn(n, 3, generator=gen).clamp_(-3.0, 3.0)
    xyz -= xyz.mean(dim=0, keepdim=True)                      # zero_mean: centroid 0 ...
    xyz /= xyz.pow(2).sum(1).sqrt().max()                     # ... and max radius 1
    return xyz


def make_cloud(gen, n, channels):
    """(n, channels) fp32 rows: xyz, rgb[, mask]."""
    rows = torch.empty(n, channels, dtype=torch.float32)
    rows[:, :3] = _cloud_xyz(gen, n)
    rows[:, 3:6] = torch.rand(n, 3, generator=gen)
    if channels > 6:
        rows[:, 6] = torch.multinomial(torch.tensor([.2, .4, .3, .1]), n, replacement=True, generator=gen).float()
    u = torch.rand(1, generator=gen).item()
    if u < 0.10 and n >= 8:       # duplicates of earlier rows
        q = n // 4
        src = torch.randint(0, n - q, (q,), generator=gen)
        rows[n - q:] = rows[src]
    elif u < 0.15 and n >= 10:    # zeroed rows
        z = torch.randperm(n, generator=gen)[: n // 10]
        rows[z] = 0.0
    return rows


def edge_list(n_obj, pairs="unordered"):
    if pairs == "unordered":
        e = [(i, j) for i in range(n_obj) for j in range(i + 1, n_obj)]
    elif pairs == "ordered":
        e = [(i, j) for i in range(n_obj) for j in range(n_obj) if i != j]
    else:
        raise ValueError(pairs)
    return torch.tensor(e, dtype=torch.int64).t().contiguous()


def make_scene(scene_id, n_obj=12, n_points_obj=80000, n_points_rel=80000, pairs="unordered",
               num_class=12, num_rel=15, image=False):
    """One scene as the reference's collated batch dict (CPU tensors)."""
    gen = torch.Generator().manual_seed(1234 + int(scene_id))
    edges = edge_list(n_obj, pairs)
    n_edge = edges.shape[1]
    obj = torch.stack([make_cloud(gen, n_points_obj, 6) for _ in range(n_obj)])
    rel = torch.stack([make_cloud(gen, n_points_rel, 7) for _ in range(n_edge)])
    one_hot = torch.zeros(n_edge, 12)
    a = torch.randint(0, 6, (n_edge,), generator=gen)
    b = torch.randint(0, 6, (n_edge,), generator=gen)
    one_hot[torch.arange(n_edge), a] = 1.0
    one_hot[torch.arange(n_edge), 6 + b] = 1.0
    batch = {
        "obj_points": obj.permute(0, 2, 1),     # (n_obj, 6, N) view of (n_obj, N, 6), like collate_fn
        "rel_points": rel.permute(0, 2, 1),
        "edge_indices": edges,
        "relation_objects_one_hot": one_hot,
        "gt_class": torch.randint(0, num_class, (n_obj,), generator=gen),
        "gt_rels": torch.randint(0, num_rel, (n_edge,), generator=gen),
        "scan_id": f"0_{int(scene_id)}",
    }
    if image:
        batch["full_image_features"] = 0.1 * torch.randn(6, 2048, generator=gen)
    return batch


def concat_scenes(scenes):
    """PyG-style concatenation of several scenes into one batch (edge indices offset per scene)."""
    out, off = {}, 0
    edges, scene_of_edge = [], []
    for s, sc in enumerate(scenes):
        edges.append(sc["edge_indices"] + off)
        scene_of_edge.append(torch.full((sc["edge_indices"].shape[1],), s, dtype=torch.int64))
        off += sc["obj_points"].shape[0]
    for key in ("obj_points", "rel_points"):
        rows = torch.cat([sc[key].permute(0, 2, 1) for sc in scenes], dim=0).contiguous()
        out[key] = rows.permute(0, 2, 1)
    out["edge_indices"] = torch.cat(edges, dim=1).contiguous()
    out["edge_scene"] = torch.cat(scene_of_edge)
    for key in ("relation_objects_one_hot", "gt_class", "gt_rels"):
        out[key] = torch.cat([sc[key] for sc in scenes], dim=0)
    if "full_image_features" in scenes[0]:
        out["full_image_features"] = torch.stack([sc["full_image_features"] for sc in scenes])
    out["scan_id"] = [sc["scan_id"] for sc in scenes]
    return out


def make_batch(first_scene, n_scenes, **kw):
    return concat_scenes([make_scene(first_scene + i, **kw) for i in range(n_scenes)])


def to_device(batch, device, non_blocking=False):
    out = {}
    for k, v in batch.items():
        if torch.is_tensor(v):
            if v.dim() == 3 and k.endswith("_points"):   # keep the (B, N, C) row layout underneath
                v = v.permute(0, 2, 1).to(device, non_blocking=non_blocking).permute(0, 2, 1)
            else:
                v = v.to(device, non_blocking=non_blocking)
        out[k] = v
    return out


def make_raw_scene(scene_id, n_obj=12, n_points=200000, background=0.3):
    """A raw scene for the crop / sample front-end (sg4d.frontend): points (P, 6) = xyz in metres + rgb, and per-point object
    masks (P,) int32 (0 = unlabelled, i + 1 = object i) -- the inputs of the reference's data_preparation
    (SGH/dataset/data_preparation_utils.py:52-76).  Objects are Gaussian blobs in a 4 m room, point counts drawn unevenly."""
    gen = torch.Generator().manual_seed(99000 + int(scene_id))
    share = torch.rand(n_obj, generator=gen) + 0.2
    n_bg = int(n_points * background)
    counts = (share / share.sum() * (n_points - n_bg)).long()
    counts[0] += n_points - n_bg - int(counts.sum())
    pts, masks = [], []
    for i in range(n_obj):
        c = 0.5 + 3.0 * torch.rand(3, generator=gen)
        s = 0.1 + 0.25 * torch.rand(3, generator=gen)
        pts.append(c + s * torch.randn(int(counts[i]), 3, generator=gen))
        masks.append(torch.full((int(counts[i]),), i + 1, dtype=torch.int32))
    pts.append(4.0 * torch.rand(n_bg, 3, generator=gen))
    masks.append(torch.zeros(n_bg, dtype=torch.int32))
    xyz, m = torch.cat(pts), torch.cat(masks)
    perm = torch.randperm(n_points, generator=gen)          # scanner order: objects are interleaved
    points = torch.cat([xyz, torch.rand(n_points, 3, generator=gen)], dim=1)[perm].contiguous()
    return points, m[perm].contiguous()
