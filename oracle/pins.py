"""oracle/pins.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Selection pinning for gradient parity.  A max-pool arg-max and a ReLU active set make the gradient of a set-abstraction scale
piecewise smooth: two correct fp32 evaluations of the same scale differ by O(1) in the entries whose selection sits within
rounding of a tie.  The checkers therefore evaluate the reference arithmetic with the selections THE PRODUCT MADE (recorded
through ``sg4d.mlp.CAPTURE``): ``pins_from_captures`` converts them into the oracle's channel-major layout for
``oracle.model_ref.forward(pins=...)``; ``grouped_fp64`` / ``h1_mask`` rebuild the grouped rows (OPS/pointnet2_utils.py:318-328)
and the first block's ReLU mask exactly as the kernels evaluate it.  Used by tests/ and by __graft_entry__.smoke().
"""
import torch


def grouped_fp64(pts, feats, foff, c, centers, idx):
    """reference grouped rows [xyz - centre | feats] (utils.py:319-328), fp64, differentiable w.r.t. feats"""
    b, n, _ = pts.shape
    m, ns = idx.shape[1], idx.shape[2]
    li = idx.long().view(b, m * ns)
    xyz = torch.gather(pts[:, :, :3].double(), 1, li.unsqueeze(-1).expand(-1, -1, 3)).view(b, m, ns, 3)
    xyz = (xyz.float() - centers.view(b, m, 1, 3)).double()           # the reference subtracts in fp32
    cols = [xyz]
    if c:
        f = feats[:, :, foff:foff + c]
        cols.append(torch.gather(f, 1, li.unsqueeze(-1).expand(-1, -1, c)).view(b, m, ns, c).double())
    return torch.cat(cols, dim=3).view(b * m * ns, 3 + c)


def h1_mask(cap, x32):
    """The first layer's ReLU mask exactly as the kernels evaluate it (fp32 fused multiply-adds in the kernel's order):
    the second pinned selection -- an activation within rounding of 0 may be clipped on one side only."""
    if cap["kind"] == "sa1":
        w1s, t1 = cap["w1s"], cap["stats1"][1]
        xa = torch.cat([x32, torch.zeros(x32.shape[0], 8 - x32.shape[1], device=x32.device)], 1)
        v = t1.expand(x32.shape[0], 64).clone()
        for j in range(8):
            v = torch.addcmul(v, xa[:, j:j + 1], w1s[j:j + 1])       # fma(x_j, w_j, v), j ascending (sa1_y1bn)
        return v > 0
    return torch.addcmul(cap["t1"], cap["y1"], cap["s1"]) > 0


def pins_from_captures(caps):
    """sg4d's selections (max-pool rows, the two ReLU active sets) of every set-abstraction scale, in the oracle's
    channel-major layout and keyed by the oracle's module prefixes.  Capture order = call order: the object encoder's five
    scales (SA1 x 2, SA2 x 2, SA3), then the edge encoder's."""
    pins, scales, inputs = {}, [], None
    for q in caps:
        if "mlp" in q:
            inputs = q
        elif "garg" in q:
            scales.append((inputs if q["kind"] in ("sa1", "sa2") else None, q))
            inputs = None
    assert len(scales) == 10, len(scales)
    names = [(enc, lvl, sc) for enc in ("obj_encoder", "rel_encoder") for lvl, sc in ((0, 0), (0, 1), (1, 0), (1, 1), (2, 0))]
    for (enc, lvl, sc), (inp, q) in zip(names, scales):
        g, c2 = q["garg"].shape
        if inp is not None:
            b, m, ns = inp["idx"].shape
            x32 = grouped_fp64(inp["pts"], inp["feats"] if inp["feats"] is not None else inp["pts"], inp["foff"], inp["c"],
                                      inp["centers"], inp["idx"]).float()
            h1 = h1_mask(q, x32)
        else:                                   # SA3 (GroupAll): one group of n points per cloud
            h1 = h1_mask(q, None)
            b, m, ns = g, 1, h1.shape[0] // g
        c2 = q["out"].shape[1]
        pins[f"{enc}.backbone.SA_modules.{lvl}.{sc}"] = (
            h1.view(b, m, ns, -1).permute(0, 3, 1, 2).float().cpu(),
            q["garg"][:, :c2].view(b, m, c2).permute(0, 2, 1).unsqueeze(-1).long().cpu(),
            (q["out"] > 0).view(b, m, c2).permute(0, 2, 1).float().cpu())
    return pins
