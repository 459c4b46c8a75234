"""Fused shared point-MLP of one set-abstraction scale on the tensor cores (``csrc/mlp.cu``).

``fused_shared_mlp(x, k0, group, mlp)`` evaluates the reference's
``[Conv2d(1x1) -> BatchNorm2d -> ReLU] x 2 -> max_pool2d over nsample`` (OPS/pointnet2_modules.py:9-19,
66-70) on the point-major grouped matrix ``x (R, K_padded)`` and returns the pooled ``(R / group, C_out)``
features, forward and backward:

  forward   2 x sg4d_linear_fwd (3xTF32 tcgen05 GEMM; BatchNorm+ReLU of the previous layer fused into the
            operand staging, statistics and max-pool fused into the epilogue) + 2 x sg4d_bn_finalize
  backward  sg4d_pool_bwd_da / sg4d_pool_bwd_dw / sg4d_inner_bwd_dw (/ sg4d_inner_bwd_dx): the max-pool, ReLU and
            BatchNorm backward formulas are evaluated inside the operand stagers from the saved pre-activations
            y1, y2 and a handful of per-channel constants computed here -- no dY tensor is ever materialised.

Saved for backward: x, y1, y2 (pre-activations), the pooled selection (value + row index) and per-channel vectors.

``fused_sa_scale`` goes from the ball-query indices to the pooled features without the grouped tensor (include/sg4d.h
section 4): ``_FusedSA1`` recomputes the K <= 7 first layer inside the operand stagers and forms its second-layer weight
gradient from the Gram matrix of the recomputed activations; ``_FusedSA2`` uses the first layer's linearity -- one small GEMM
per SOURCE POINT, the grouped activations are a gather ``Z[idx] - Cc[centre]``, and the backward pass reduces dY1 per source
point / per centre before three small GEMMs.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib


CAPTURE = None      # tests set this to a list: every fused scale appends its max-pool selection (garg) for pinned-selection checks


def _capture(**kw):
    if CAPTURE is not None:
        CAPTURE.append(kw)


def pack_weight(w2d):
    """(N, K) fp32 -> the pre-split, pre-swizzled shared-memory image sg4d_linear_fwd streams by bulk TMA."""
    w2d = w2d.contiguous()
    n, k = w2d.shape
    img = torch.empty(_lib.load().sg4d_weight_image_floats(n, k), dtype=torch.float32, device=w2d.device)
    _lib.call("sg4d_pack_weight", w2d, n, k, w2d.stride(0), w2d.data_ptr(), img.data_ptr())
    return img


def linear_fwd(a, k, wimg, n, scale=None, shift=None, group=0, gamma=None, store_y=True):
    """Y = act(a[:, :k]) @ W^T with fused statistics (and group max/min).  Returns (y, partial, gsel, garg)."""
    rows, lda = a.shape
    dev = a.device
    y = torch.empty(rows, n, dtype=torch.float32, device=dev) if store_y else None
    partial = torch.empty(_lib.load().sg4d_mlp_partial_doubles(rows), dtype=torch.float64, device=dev)
    gsel = garg = None
    if group:
        gsel = torch.empty(rows // group, n, dtype=torch.float32, device=dev)
        garg = torch.empty(rows // group, n, dtype=torch.uint8, device=dev)
    _lib.call("sg4d_linear_fwd", a, rows, k, lda, n, group, a.data_ptr(), _lib.ptr(scale), _lib.ptr(shift),
              wimg.data_ptr(), _lib.ptr(y), partial.data_ptr(), _lib.ptr(gamma), _lib.ptr(gsel), _lib.ptr(garg))
    return y, partial, gsel, garg


def bn_scale_shift(bn, partial, rows):
    """Batch statistics -> (scale, shift, mean, invstd); updates the module's running statistics exactly like
    nn.BatchNorm2d.forward in training mode.  In eval mode the running statistics are used instead."""
    n = bn.num_features
    dev = bn.weight.device
    if bn.training or not bn.track_running_stats:
        out = torch.empty(4, n, dtype=torch.float32, device=dev)
        track = bn.training and bn.track_running_stats
        momentum = 0.0 if bn.momentum is None else bn.momentum
        if track and bn.num_batches_tracked is not None:
            bn.num_batches_tracked.add_(1)
            if bn.momentum is None:
                momentum = 1.0 / float(bn.num_batches_tracked)
        _lib.call("sg4d_bn_finalize", partial, n, partial.numel() // 2, rows, partial.data_ptr(), bn.weight.data_ptr(),
                  bn.bias.data_ptr(), float(bn.eps), float(momentum), _lib.ptr(bn.running_mean if track else None),
                  _lib.ptr(bn.running_var if track else None), out[0].data_ptr(), out[1].data_ptr(), out[2].data_ptr(),
                  out[3].data_ptr())
        return out[0], out[1], out[2], out[3]
    invstd = torch.rsqrt(bn.running_var + bn.eps)
    scale = bn.weight.detach() * invstd
    return scale, bn.bias.detach() - bn.running_mean * scale, bn.running_mean, invstd


def supported(mlp, k_padded, group):
    """True when the scale can run on the fused tensor-core path (otherwise the generic path is used)."""
    if len(mlp) != 6:
        return False
    c1, b1, r1, c2, b2, r2 = mlp
    if not (isinstance(c1, nn.Conv2d) and isinstance(b1, nn.BatchNorm2d) and isinstance(c2, nn.Conv2d)
            and isinstance(b2, nn.BatchNorm2d) and isinstance(r1, nn.ReLU) and isinstance(r2, nn.ReLU)):
        return False
    if c1.bias is not None or c2.bias is not None or not b1.affine or not b2.affine:
        return False
    n1, n2 = c1.out_channels, c2.out_channels
    if n1 not in (64, 128) or n2 not in (64, 128) or k_padded > 224 or k_padded % 4:
        return False
    if group not in (1, 2, 4, 8, 16, 32, 64, 128) or (n2 == 64 and group > 64):
        return False
    return True


def bn_bwd_coeffs(scale, mean, invstd, d_beta, d_gamma, rows, batch):
    """(q, u, -mean * invstd) of the BatchNorm backward  dY = scale * dz - (q * y + u)  in one launch (sg4d_bn_bwd_coeffs)."""
    n = scale.numel()
    coef = torch.empty(3, n, dtype=torch.float32, device=scale.device)
    _lib.call("sg4d_bn_bwd_coeffs", scale, n, rows, 1 if batch else 0, d_beta.data_ptr(), d_gamma.data_ptr(), scale.data_ptr(),
              mean.data_ptr(), invstd.data_ptr(), coef.data_ptr())
    return coef[0], coef[1], coef[2]


def _wgrad_partial(rows, npad, dev):
    return torch.empty(_lib.load().sg4d_wgrad_partial_floats(rows, npad), dtype=torch.float32, device=dev)


class _FusedSharedMLP(torch.autograd.Function):
    """x (R, kp); w1 (n1, kp) already permuted / zero-padded to x's column layout; w2 (n2, n1)."""

    @staticmethod
    def forward(ctx, x, w1, g1, be1, w2, g2, be2, group, dx_cols, bn1, bn2):
        rows, kp = x.shape
        n1, n2 = w1.shape[0], w2.shape[0]
        y1, part1, _, _ = linear_fwd(x, kp, pack_weight(w1), n1)
        s1, t1, m1, i1 = bn_scale_shift(bn1, part1, rows)
        y2, part2, gsel, garg = linear_fwd(y1, n1, pack_weight(w2), n2, scale=s1, shift=t1, group=group, gamma=g2)
        s2, t2, m2, i2 = bn_scale_shift(bn2, part2, rows)
        out = torch.relu(torch.addcmul(t2, gsel, s2))
        _capture(kind="rows", garg=garg, gsel=gsel, y1=y1, y2=y2, s1=s1, t1=t1)
        ctx.save_for_backward(x, y1, y2, gsel, garg, out, w1, w2, s1, t1, m1, i1, s2, m2, i2)
        ctx.meta = (group, dx_cols, bn1.training or not bn1.track_running_stats,
                    bn2.training or not bn2.track_running_stats)
        return out

    @staticmethod
    def backward(ctx, d_out):
        x, y1, y2, gsel, garg, out, w1, w2, s1, t1, m1, i1, s2, m2, i2 = ctx.saved_tensors
        group, dx_cols, batch1, batch2 = ctx.meta
        rows, kp = x.shape
        n1, n2 = w1.shape[0], w2.shape[0]
        dev = x.device
        inv_r = 1.0 / rows

        # ---- layer 2: BatchNorm2 + ReLU + max-pool.  dY2 = dsel*[row is the pooled one] - (a2*y2 + b2)
        # one fused pass over the pooled tensors: dz = d_out*[out > 0], dsel = dz*s2, and the two per-channel sums
        groups = out.shape[0]
        if d_out.stride(1) != 1 or (d_out.stride(0) & 3) or (d_out.data_ptr() & 15):
            d_out = d_out.contiguous()
        lib = _lib.load()
        nparts = lib.sg4d_pool_bwd_prologue_parts() * n2
        part2 = torch.empty(nparts * 2, dtype=torch.float64, device=dev)
        dsel = torch.empty(groups, n2, dtype=torch.float32, device=dev)
        _lib.call("sg4d_pool_bwd_prologue", x, groups, n2, d_out.stride(0), d_out.data_ptr(), out.data_ptr(), gsel.data_ptr(),
                  s2.data_ptr(), m2.data_ptr(), i2.data_ptr(), dsel.data_ptr(), part2.data_ptr())
        sums2 = torch.empty(2, n2, dtype=torch.float32, device=dev)
        _lib.call("sg4d_partial_sums", x, n2, nparts, part2.data_ptr(), sums2.data_ptr())
        d_be2, d_g2 = sums2[0], sums2[1]
        a2, b2, _ = bn_bwd_coeffs(s2, m2, i2, d_be2, d_g2, rows, batch2)
        em1 = (-m1 * i1).contiguous()
        dz1 = torch.empty(rows, n1, dtype=torch.float32, device=dev)
        part = torch.empty(_lib.load().sg4d_mlp_partial_doubles(rows), dtype=torch.float64, device=dev)
        _lib.call("sg4d_pool_bwd_da", x, rows, n2, n1, group, y2.data_ptr(), a2.data_ptr(), b2.data_ptr(), dsel.data_ptr(),
                  garg.data_ptr(), pack_weight(w2.t()).data_ptr(), y1.data_ptr(), s1.data_ptr(), t1.data_ptr(),
                  i1.data_ptr(), em1.data_ptr(), dz1.data_ptr(), part.data_ptr())
        sums = torch.empty(2, n1, dtype=torch.float32, device=dev)
        _lib.call("sg4d_partial_sums", x, n1, part.numel() // 2, part.data_ptr(), sums.data_ptr())
        d_be1, d_g1 = sums[0], sums[1]
        d_w2 = torch.empty(n2, n1, dtype=torch.float32, device=dev)
        _lib.call("sg4d_pool_bwd_dw", x, rows, n2, n1, group, y2.data_ptr(), a2.data_ptr(), b2.data_ptr(), dsel.data_ptr(),
                  garg.data_ptr(), y1.data_ptr(), s1.data_ptr(), t1.data_ptr(), _wgrad_partial(rows, n1, dev).data_ptr(),
                  d_w2.data_ptr())

        # ---- layer 1: dY1 = p1*dz1 - (q1*y1 + u1)
        p1 = s1.contiguous()
        q1, u1, _ = bn_bwd_coeffs(s1, m1, i1, d_be1, d_g1, rows, batch1)
        npad = 32 if kp <= 32 else (64 if kp <= 64 else (128 if kp <= 128 else 224))
        d_w1 = torch.empty(n1, kp, dtype=torch.float32, device=dev)
        _lib.call("sg4d_inner_bwd_dw", x, rows, n1, kp, kp, y1.data_ptr(), dz1.data_ptr(), p1.data_ptr(), q1.data_ptr(),
                  u1.data_ptr(), x.data_ptr(), _wgrad_partial(rows, npad, dev).data_ptr(), d_w1.data_ptr())
        d_x = None
        if ctx.needs_input_grad[0]:
            # only the first dx_cols columns (the gathered features) carry a gradient downstream
            d_x = torch.empty(rows, kp, dtype=torch.float32, device=dev)
            d_x[:, dx_cols:].zero_()
            col = 0
            while col < dx_cols:
                n = 128 if dx_cols - col >= 128 else 64
                assert dx_cols - col >= n, "feature width must be a multiple of 64"
                wt = pack_weight(w1[:, col:col + n].t())
                _lib.call("sg4d_inner_bwd_dx", x, rows, n1, n, y1.data_ptr(), dz1.data_ptr(), p1.data_ptr(), q1.data_ptr(),
                          u1.data_ptr(), wt.data_ptr(), d_x.data_ptr(), kp, col)
                col += n
        return (d_x, d_w1, d_g1, d_be1, d_w2, d_g2, d_be2, None, None, None, None)


def _src_args(pts, feats, foff, c, centers, idx):
    """The grouped-row source arguments shared by section 4 of include/sg4d.h."""
    b, n, ps = pts.shape
    m, ns = idx.shape[1], idx.shape[2]
    f = pts if feats is None else feats
    return (b * m * ns, n, m, ns, ps, f.shape[2], int(foff), int(c), pts.data_ptr(), f.data_ptr(), centers.data_ptr(),
            idx.data_ptr())


def _bn_mode(bn):
    """(batch statistics?, update running statistics?, momentum) with nn.BatchNorm2d.forward's bookkeeping."""
    batch = bn.training or not bn.track_running_stats
    track = bn.training and bn.track_running_stats
    momentum = 0.0 if bn.momentum is None else bn.momentum
    if track and bn.num_batches_tracked is not None:
        bn.num_batches_tracked.add_(1)
        if bn.momentum is None:
            momentum = 1.0 / float(bn.num_batches_tracked)
    return batch, track, momentum


def _pool_bwd_consts(d_out, out, gsel, s2, m2, i2, rows, batch2, ref):
    """dsel and the BatchNorm2-backward constants a2, b2 (+ d_gamma2, d_beta2) of a pooled layer."""
    groups, n2 = out.shape
    dev = out.device
    if d_out.stride(1) != 1 or (d_out.stride(0) & 3) or (d_out.data_ptr() & 15):
        d_out = d_out.contiguous()
    nparts = _lib.load().sg4d_pool_bwd_prologue_parts() * n2
    part2 = torch.empty(nparts * 2, dtype=torch.float64, device=dev)
    dsel = torch.empty(groups, n2, dtype=torch.float32, device=dev)
    _lib.call("sg4d_pool_bwd_prologue", ref, groups, n2, d_out.stride(0), d_out.data_ptr(), out.data_ptr(), gsel.data_ptr(),
              s2.data_ptr(), m2.data_ptr(), i2.data_ptr(), dsel.data_ptr(), part2.data_ptr())
    sums2 = torch.empty(2, n2, dtype=torch.float32, device=dev)
    _lib.call("sg4d_partial_sums", ref, n2, nparts, part2.data_ptr(), sums2.data_ptr())
    d_be2, d_g2 = sums2[0], sums2[1]
    a2, b2, _ = bn_bwd_coeffs(s2, m2, i2, d_be2, d_g2, rows, batch2)
    return dsel, a2, b2, d_g2, d_be2


class _FusedSA1(torch.autograd.Function):
    """One scale of the first set-abstraction level, ball-query indices -> pooled features, without the grouped
    tensor and without y1 (csrc/mlp.cu, "Fused set-abstraction scales"): replaces QueryAndGroup + the shared MLP +
    max_pool2d (OPS/pointnet2_utils.py:300-337, OPS/pointnet2_modules.py:61-70) for K = 3 + c <= 7 inputs and no
    gradient into the points.  w1 (64, 3 + c) in the reference's column order [xyz | feats]."""

    @staticmethod
    def forward(ctx, pts, feats, centers, idx, w1, g1, be1, w2, g2, be2, foff, c, bn1, bn2):
        src = _src_args(pts, feats, foff, c, centers, idx)
        rows, dev, k = src[0], pts.device, 3 + c
        n2 = w2.shape[0]
        lib = _lib.load()
        w1 = w1.contiguous()
        nparts = lib.sg4d_sa_moments_parts()
        part = torch.empty(nparts * 36, dtype=torch.float64, device=dev)
        _lib.call("sg4d_sa_moments", pts, *src, part.data_ptr())
        moments = torch.empty(64, dtype=torch.float64, device=dev)
        stats1 = torch.empty(4, 64, dtype=torch.float32, device=dev)
        w1s = torch.empty(8, 64, dtype=torch.float32, device=dev)
        batch1, track1, mom1 = _bn_mode(bn1)
        _lib.call("sg4d_sa1_bn1", pts, k, nparts, part.data_ptr(), w1.data_ptr(), k, g1.data_ptr(), be1.data_ptr(), float(bn1.eps),
                  float(mom1), _lib.ptr(bn1.running_mean if (track1 or not batch1) else None),
                  _lib.ptr(bn1.running_var if (track1 or not batch1) else None), 0 if batch1 else 1, moments.data_ptr(),
                  stats1.data_ptr(), w1s.data_ptr())
        y2 = torch.empty(rows, n2, dtype=torch.float32, device=dev)
        part2 = torch.empty(lib.sg4d_mlp_partial_doubles(rows), dtype=torch.float64, device=dev)
        ns = idx.shape[2]
        gsel = torch.empty(rows // ns, n2, dtype=torch.float32, device=dev)
        garg = torch.empty(rows // ns, n2, dtype=torch.uint8, device=dev)
        _lib.call("sg4d_sa1_fwd", pts, *src, w1s.data_ptr(), stats1[1].data_ptr(), n2, pack_weight(w2).data_ptr(), y2.data_ptr(),
                  part2.data_ptr(), g2.data_ptr(), gsel.data_ptr(), garg.data_ptr())
        s2, t2, m2, i2 = bn_scale_shift(bn2, part2, rows)
        out = torch.relu(torch.addcmul(t2, gsel, s2))
        _capture(kind="sa1", garg=garg, gsel=gsel, y2=y2, stats1=stats1, moments=moments, w1s=w1s, out=out)
        ctx.save_for_backward(pts, feats, centers, idx, y2, gsel, garg, out, w1, w2, stats1, w1s, moments, s2, m2, i2)
        ctx.meta = (foff, c, batch1, bn2.training or not bn2.track_running_stats)
        return out

    @staticmethod
    def backward(ctx, d_out):
        pts, feats, centers, idx, y2, gsel, garg, out, w1, w2, stats1, w1s, moments, s2, m2, i2 = ctx.saved_tensors
        foff, c, batch1, batch2 = ctx.meta
        src = _src_args(pts, feats, foff, c, centers, idx)
        rows, dev, k = src[0], pts.device, 3 + c
        n2 = w2.shape[0]
        lib = _lib.load()
        dsel, a2, b2, d_g2, d_be2 = _pool_bwd_consts(d_out, out, gsel, s2, m2, i2, rows, batch2, pts)
        t1 = stats1[1]
        s1part = torch.empty(lib.sg4d_sa1_s1part_doubles(rows), dtype=torch.float64, device=dev)
        _lib.call("sg4d_sa1_bwd_da", pts, *src, w1s.data_ptr(), t1.data_ptr(), n2, y2.data_ptr(), a2.data_ptr(), b2.data_ptr(),
                  dsel.data_ptr(), garg.data_ptr(), pack_weight(w2.t()).data_ptr(), s1part.data_ptr())
        # dW2 = dY2^T h1 without the dY2 operand: T1 - diag(a2) W2 (h1^T h1) - b2 (x) colsum(h1)  (include/sg4d.h)
        d_w2 = torch.empty(n2, 64, dtype=torch.float32, device=dev)
        ws = torch.empty(lib.sg4d_sa1_bwd_dw2_gram_ws_floats(rows, n2), dtype=torch.float32, device=dev)
        _lib.call("sg4d_sa1_bwd_dw2_gram", pts, *src, w1s.data_ptr(), t1.data_ptr(), n2, w2.contiguous().data_ptr(), a2.data_ptr(),
                  b2.data_ptr(), dsel.data_ptr(), garg.data_ptr(), ws.data_ptr(), d_w2.data_ptr())
        d_w1 = torch.empty(64, k, dtype=torch.float32, device=dev)
        d_gb1 = torch.empty(2, 64, dtype=torch.float32, device=dev)
        _lib.call("sg4d_sa1_bwd_finalize", pts, k, rows, s1part.data_ptr(), moments.data_ptr(), w1.data_ptr(), k,
                  stats1.data_ptr(), 1 if batch1 else 0, d_w1.data_ptr(), k, d_gb1[0].data_ptr(), d_gb1[1].data_ptr())
        return (None, None, None, None, d_w1, d_gb1[0], d_gb1[1], d_w2, d_g2, d_be2, None, None, None, None)


class _FusedSA2(torch.autograd.Function):
    """One scale of a set-abstraction level whose input features carry a gradient.  The first layer is linear in the grouped
    row x = [feats(i) | xyz(i) - centre_j], so it is evaluated per SOURCE POINT (one small GEMM over the b*n points) and the
    grouped first-layer activations are a gather:  y1[r] = Z[i(r)] - Cc[j(r)].  The backward pass uses the same linearity:
    G[i] = sum of dY1 over the rows that reference point i, H[j] = sum of dY1 over the rows of centre j,
    dFeats = G W1f,  dW1 = G^T [feats | xyz] - H^T [0 | centre].  Neither the grouped tensor nor a GEMM with K = 3 + c over the
    b*m*nsample grouped rows exists in either direction.  w1 (n1, 3 + c), reference column order [xyz | feats]."""

    @staticmethod
    def forward(ctx, pts, feats, centers, idx, cnt, w1, g1, be1, w2, g2, be2, foff, c, bn1, bn2):
        from . import dense
        dev = pts.device
        b, n = pts.shape[0], pts.shape[1]
        m, ns = idx.shape[1], idx.shape[2]
        rows = b * m * ns
        n1, n2 = w1.shape[0], w2.shape[0]
        lib = _lib.load()
        w1g = F.pad(torch.cat([w1[:, 3:], w1[:, :3]], dim=1), (0, 1))        # column order [feats | xyz | 0]
        xin = torch.cat([feats[..., foff:foff + c].reshape(b * n, c), pts[..., :3].reshape(b * n, 3),
                         torch.zeros(b * n, 1, dtype=torch.float32, device=dev)], dim=1)
        cen4 = F.pad(centers.reshape(b * m, 3), (0, 1))
        z = dense._fwd(xin, c + 4, dense.pack(w1g), n1)[0]                    # per source point
        cc = dense._fwd(cen4, 4, dense.pack(w1g[:, c:].contiguous()), n1)[0]  # per centre
        y1 = torch.empty(rows, n1, dtype=torch.float32, device=dev)
        part1 = torch.empty(2 * lib.sg4d_gather_y1_parts(n1), dtype=torch.float64, device=dev)
        _lib.call("sg4d_gather_y1", pts, rows, n, m, ns, n1, z.data_ptr(), cc.data_ptr(), idx.data_ptr(), y1.data_ptr(),
                  part1.data_ptr())
        s1, t1, m1, i1 = bn_scale_shift(bn1, part1, rows)
        y2, part2, gsel, garg = linear_fwd(y1, n1, pack_weight(w2), n2, scale=s1, shift=t1, group=ns, gamma=g2)
        s2, t2, m2, i2 = bn_scale_shift(bn2, part2, rows)
        out = torch.relu(torch.addcmul(t2, gsel, s2))
        _capture(kind="sa2", garg=garg, gsel=gsel, y1=y1, y2=y2, s1=s1, t1=t1, out=out)
        ctx.save_for_backward(xin, cen4, idx, cnt, y1, y2, gsel, garg, out, w1g, w2, s1, t1, m1, i1, s2, m2, i2)
        ctx.meta = (b, n, c, bn1.training or not bn1.track_running_stats, bn2.training or not bn2.track_running_stats)
        ctx.foff, ctx.feat_width = foff, feats.shape[2]
        return out

    @staticmethod
    def backward(ctx, d_out):
        from . import dense
        xin, cen4, idx, cnt, y1, y2, gsel, garg, out, w1g, w2, s1, t1, m1, i1, s2, m2, i2 = ctx.saved_tensors
        b, n, c, batch1, batch2 = ctx.meta
        m, ns = idx.shape[1], idx.shape[2]
        rows, dev = b * m * ns, xin.device
        n1, n2 = w1g.shape[0], w2.shape[0]
        dsel, a2, b2, d_g2, d_be2 = _pool_bwd_consts(d_out, out, gsel, s2, m2, i2, rows, batch2, xin)
        em1 = (-m1 * i1).contiguous()
        dz1 = torch.empty(rows, n1, dtype=torch.float32, device=dev)
        part = torch.empty(_lib.load().sg4d_mlp_partial_doubles(rows), dtype=torch.float64, device=dev)
        _lib.call("sg4d_pool_bwd_da", xin, rows, n2, n1, ns, y2.data_ptr(), a2.data_ptr(), b2.data_ptr(), dsel.data_ptr(),
                  garg.data_ptr(), pack_weight(w2.t()).data_ptr(), y1.data_ptr(), s1.data_ptr(), t1.data_ptr(),
                  i1.data_ptr(), em1.data_ptr(), dz1.data_ptr(), part.data_ptr())
        sums = torch.empty(2, n1, dtype=torch.float32, device=dev)
        _lib.call("sg4d_partial_sums", xin, n1, part.numel() // 2, part.data_ptr(), sums.data_ptr())
        d_be1, d_g1 = sums[0], sums[1]
        d_w2 = torch.empty(n2, n1, dtype=torch.float32, device=dev)
        _lib.call("sg4d_pool_bwd_dw", xin, rows, n2, n1, ns, y2.data_ptr(), a2.data_ptr(), b2.data_ptr(), dsel.data_ptr(),
                  garg.data_ptr(), y1.data_ptr(), s1.data_ptr(), t1.data_ptr(), _wgrad_partial(rows, n1, dev).data_ptr(),
                  d_w2.data_ptr())
        p1 = s1.contiguous()
        q1, u1, _ = bn_bwd_coeffs(s1, m1, i1, d_be1, d_g1, rows, batch1)
        # dY1 = p1 .* dz1 - (q1 .* y1 + u1) is never stored: both sums below generate it on the fly
        g_sum = torch.empty(b * n, n1, dtype=torch.float32, device=dev)      # per source point (deterministic gather)
        _lib.call("sg4d_group_rows_grad_dy", xin, b, n, m, ns, n1, y1.data_ptr(), dz1.data_ptr(), p1.data_ptr(), q1.data_ptr(),
                  u1.data_ptr(), idx.data_ptr(), cnt.data_ptr(), g_sum.data_ptr())
        h_sum = torch.empty(b * m, n1, dtype=torch.float32, device=dev)      # per centre
        _lib.call("sg4d_group_sum_dy", xin, b * m, ns, n1, y1.data_ptr(), dz1.data_ptr(), p1.data_ptr(), q1.data_ptr(),
                  u1.data_ptr(), h_sum.data_ptr())
        d_w1g = dense._dw(b * n, n1, g_sum, 0, xin, c + 4)                    # (n1, c + 4): G^T [feats | xyz | 0]
        d_wc = dense._dw(b * m, n1, h_sum, 0, cen4, 4)                        # (n1, 4):     H^T [centre | 0]
        d_w1 = torch.cat([d_w1g[:, c:c + 3] - d_wc[:, :3], d_w1g[:, :c]], dim=1)   # reference order [xyz | feats]
        d_feats = None
        if ctx.needs_input_grad[1]:
            ncol = (c + 63) // 64 * 64
            wt = F.pad(w1g[:, :c].t(), (0, 0, 0, ncol - c))                 # (ncol, n1) = W1[:, feats]^T, zero rows of padding
            d = dense._dx(b * n, n1, g_sum, 0, ncol, dense.pack(wt))[:, :c].reshape(b, n, c)
            if ctx.foff == 0 and ctx.feat_width == c:
                d_feats = d
            else:                                                           # the scale reads a column window of a wider tensor
                d_feats = torch.zeros(b, n, ctx.feat_width, dtype=torch.float32, device=dev)
                d_feats[..., ctx.foff:ctx.foff + c] = d
        return (None, d_feats, None, None, None, d_w1, d_g1, d_be1, d_w2, d_g2, d_be2, None, None, None, None)


def _two_layer(mlp):
    if len(mlp) != 6:
        return None
    c1, b1, r1, c2, b2, r2 = mlp
    if not (isinstance(c1, nn.Conv2d) and isinstance(b1, nn.BatchNorm2d) and isinstance(c2, nn.Conv2d)
            and isinstance(b2, nn.BatchNorm2d) and isinstance(r1, nn.ReLU) and isinstance(r2, nn.ReLU)):
        return None
    if c1.bias is not None or c2.bias is not None or not b1.affine or not b2.affine:
        return None
    return c1, b1, c2, b2


def sa_scale_kind(mlp, c, ns, feats_need_grad, feat_stride, foff):
    """Which fused kernel family evaluates this scale: 'sa1', 'sa2' or None (-> materialised grouped rows)."""
    layers = _two_layer(mlp)
    if layers is None or ns not in (8, 16, 32, 64, 128):
        return None
    c1, b1, c2, b2 = layers
    n1, n2 = c1.out_channels, c2.out_channels
    if n2 not in (64, 128) or (n2 == 64 and ns > 64):
        return None
    if not feats_need_grad and c <= 4 and n1 == 64:
        return "sa1"
    if n1 in (64, 128) and c > 0 and c % 4 == 0:
        return "sa2"
    return None


def fused_sa_scale(kind, pts, feats, foff, c, centers, idx, cnt, mlp):
    """pts (B,n,S), feats (B,n,Sf) or None, centers (B,m,3), idx (B,m,ns) -> pooled (B*m, C_out)."""
    c1, b1, c2, b2 = _two_layer(mlp)
    _capture(kind=kind, pts=pts, feats=feats, foff=foff, c=c, centers=centers, idx=idx, cnt=cnt, mlp=mlp)
    w1 = c1.weight.view(c1.out_channels, 3 + c)
    w2 = c2.weight.view(c2.out_channels, c2.in_channels)
    if kind == "sa1":
        return _FusedSA1.apply(pts, feats, centers, idx, w1, b1.weight, b1.bias, w2, b2.weight, b2.bias, int(foff), int(c), b1, b2)
    if kind == "sa2":
        if feats.data_ptr() % 16:
            raise RuntimeError("fused SA scale: the feature tensor must be 16-byte aligned")
        return _FusedSA2.apply(pts, feats, centers, idx, cnt, w1, b1.weight, b1.bias, w2, b2.weight, b2.bias, int(foff), int(c),
                               b1, b2)
    raise ValueError(kind)


def fused_shared_mlp(x, k0, group, mlp, xyz_last=False):
    """x (R, K_padded) point-major grouped rows -> pooled (R / group, C_out).

    Column layout of x: [xyz(3) | feats(k0-3) | 0-pad] or, with xyz_last, [feats(k0-3) | xyz(3) | 0-pad]; the
    first conv's weight columns are permuted / padded to match (differentiable torch ops, so the parameter
    gradient comes back in the reference's channel order)."""
    c1, b1, _, c2, b2, _ = mlp
    kp = x.shape[1]
    w1 = c1.weight.view(c1.out_channels, k0)
    if xyz_last:
        w1 = torch.cat([w1[:, 3:], w1[:, :3]], dim=1)
    if kp != k0:
        w1 = F.pad(w1, (0, kp - k0))
    dx_cols = (k0 - 3) if (xyz_last and x.requires_grad) else 0
    if x.requires_grad and not xyz_last:
        raise RuntimeError("input gradients need the feature-first (xyz_last) column layout")
    return _FusedSharedMLP.apply(x, w1, b1.weight, b1.bias, c2.weight.view(c2.out_channels, c2.in_channels), b2.weight,
                                 b2.bias, int(group), int(dx_cols), b1, b2)
