"""Dense layers on the tensor-core engine (``csrc/mlp.cu``, include/sg4d.h section 5): what the reference runs as
``nn.Linear`` / ``Conv2d(1x1)`` + ``BatchNorm`` + ``ReLU`` through cuBLAS / cuDNN / ATen in

* the GroupAll level SA3          OPS/pointnet2_modules.py:130-146 (``shared_mlp``)
* the TripletGCN MLPs             SGH/model/gcns/network_TripletGCN.py:11-58 (``linear`` / ``linear_bn_relu``)
* the classifier heads            SGH/model/pointnets/network_PointNet.py:188-271 (``linear``)

Every product is a 3xTF32 tcgen05 GEMM (fp32-level accuracy); BatchNorm statistics come out of the GEMM epilogue, the
BatchNorm / ReLU backward formulas are evaluated inside the operand stagers of the backward GEMMs.  There is no
fallback: shapes the engine cannot take raise.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib
from . import mlp as fused


def _ceil(v, m):
    return (v + m - 1) // m * m


def _rows_matrix(x, name):
    if x.dim() != 2 or x.dtype != torch.float32:
        raise RuntimeError(f"{name} must be a 2-D float tensor")
    _lib.require_cuda(x)
    if x.stride(1) != 1 or x.stride(0) % 4 or x.data_ptr() % 16 or x.shape[1] % 4:
        k = x.shape[1]
        x = F.pad(x, (0, _ceil(k, 4) - k)).contiguous() if k % 4 else x.contiguous()
    return x


def pack(w2d):
    """(n, k) fp32, n % 64 == 0 -> panel-wise pre-split, pre-swizzled weight image."""
    w2d = w2d.contiguous()
    n, k = w2d.shape
    img = torch.empty(_lib.load().sg4d_dense_weight_floats(n, k), dtype=torch.float32, device=w2d.device)
    _lib.call("sg4d_dense_pack_weight", w2d, n, k, w2d.stride(0), w2d.data_ptr(), img.data_ptr())
    return img


def _fwd(x, k, wimg, n, scale=None, shift=None, bias=None, stats=False, group=0, gamma=None):
    rows = x.shape[0]
    dev = x.device
    y = torch.empty(rows, n, dtype=torch.float32, device=dev)
    partial = torch.empty(_lib.load().sg4d_dense_partial_doubles(rows, n), dtype=torch.float64, device=dev) if stats else None
    gsel = garg = None
    if group:
        gsel = torch.empty(rows // group, n, dtype=torch.float32, device=dev)
        garg = torch.empty(rows // group, n, dtype=torch.uint8, device=dev)
    _lib.call("sg4d_dense_fwd", x, rows, k, x.stride(0), n, x.data_ptr(), _lib.ptr(scale), _lib.ptr(shift), wimg.data_ptr(),
              _lib.ptr(bias), y.data_ptr(), n, _lib.ptr(partial), int(group), _lib.ptr(gamma), _lib.ptr(gsel), _lib.ptr(garg))
    return y, partial, gsel, garg


class _PaddedBN:
    """A BatchNorm module seen through zero-padded channel vectors (layer widths that are not multiples of 64 run with
    zero weights / gamma / beta in the padding channels, which therefore stay exactly zero end to end)."""

    def __init__(self, bn, n):
        self.bn, self.n, self.real = bn, n, bn.num_features
        self.training, self.track_running_stats, self.momentum, self.eps = bn.training, bn.track_running_stats, bn.momentum, bn.eps
        self.num_batches_tracked = bn.num_batches_tracked
        pad = n - self.real
        self.running_mean = F.pad(bn.running_mean, (0, pad)) if bn.running_mean is not None else None
        self.running_var = F.pad(bn.running_var, (0, pad), value=1.0) if bn.running_var is not None else None

    def write_back(self):
        if self.running_mean is not None and self.training and self.track_running_stats:
            self.bn.running_mean.copy_(self.running_mean[:self.real])
            self.bn.running_var.copy_(self.running_var[:self.real])


def _bn_stats(bn, partial, rows, n, gamma, beta):
    """(4, n) scale / shift / mean / invstd with nn.BatchNorm's training / eval bookkeeping."""
    dev = gamma.device
    batch, track, momentum = fused._bn_mode(bn)
    stats = torch.empty(4, n, dtype=torch.float32, device=dev)
    if batch:
        _lib.call("sg4d_dense_bn_finalize", partial, n, rows, partial.data_ptr(), gamma.data_ptr(), beta.data_ptr(),
                  float(bn.eps), float(momentum), _lib.ptr(bn.running_mean if track else None),
                  _lib.ptr(bn.running_var if track else None), stats.data_ptr())
    else:
        invstd = torch.rsqrt(bn.running_var + bn.eps)
        stats[0] = gamma.detach() * invstd
        stats[1] = beta.detach() - bn.running_mean * stats[0]
        stats[2] = bn.running_mean
        stats[3] = invstd
    if isinstance(bn, _PaddedBN):
        bn.write_back()
    return stats, batch


def _dx(rows, kk, a, mode, nout, w_t_img, a2=None, p1=None, q1=None, u1=None, e=None, es=None, et=None):
    """dX (rows, nout) = dY * W, optionally masked by [e*es + et > 0]."""
    dev = a.device
    dx = torch.empty(rows, nout, dtype=torch.float32, device=dev)
    partial = torch.empty(_lib.load().sg4d_dense_partial_doubles(rows, nout), dtype=torch.float64, device=dev) if e is not None else None
    _lib.call("sg4d_dense_bwd_dx", a, rows, kk, a.stride(0), nout, mode, a.data_ptr(), _lib.ptr(a2), _lib.ptr(p1), _lib.ptr(q1),
              _lib.ptr(u1), w_t_img.data_ptr(), _lib.ptr(e), e.stride(0) if e is not None else 0, _lib.ptr(es), _lib.ptr(et),
              dx.data_ptr(), nout, _lib.ptr(partial))
    return dx


def _dw(rows, m, a, mode, x, k, a2=None, p1=None, q1=None, u1=None, xs=None, xt=None):
    dev = a.device
    dw = torch.empty(m, k, dtype=torch.float32, device=dev)
    partial = torch.empty(_lib.load().sg4d_dense_wgrad_partial_floats(rows, m, k), dtype=torch.float32, device=dev)
    _lib.call("sg4d_dense_bwd_dw", a, rows, m, a.stride(0), k, mode, a.data_ptr(), _lib.ptr(a2), _lib.ptr(p1), _lib.ptr(q1),
              _lib.ptr(u1), x.data_ptr(), x.stride(0), _lib.ptr(xs), _lib.ptr(xt), partial.data_ptr(), dw.data_ptr(), k)
    return dw


def _colsum(a, n):
    rows = a.shape[0]
    part = torch.empty(_lib.load().sg4d_colsum_part_doubles(rows, n), dtype=torch.float64, device=a.device)
    out = torch.empty(2, n, dtype=torch.float32, device=a.device)
    _lib.call("sg4d_colsum", a, rows, n, a.data_ptr(), a.stride(0), part.data_ptr(), out.data_ptr())
    return out[0]


class _Linear(torch.autograd.Function):
    """y = act(x) W^T + b;  act = identity or ReLU (``pre_relu``).  x (rows, k) with k % 4 == 0."""

    @staticmethod
    def forward(ctx, x, w, b, pre_relu):
        rows, k = x.shape
        n = w.shape[0]
        npad = _ceil(n, 64)
        dev = x.device
        wp = F.pad(w, (0, k - w.shape[1], 0, npad - n))                    # zero rows / columns of padding
        bp = F.pad(b, (0, npad - n)) if b is not None else None
        ones = zeros = None
        if pre_relu:
            ones, zeros = torch.ones(k, device=dev), torch.zeros(k, device=dev)
        y, _, _, _ = _fwd(x, k, pack(wp), npad, scale=ones, shift=zeros, bias=bp)
        ctx.save_for_backward(x, wp, ones, zeros)
        ctx.meta = (n, w.shape[1], b is not None)
        return y[:, :n]

    @staticmethod
    def backward(ctx, dy):
        x, wp, ones, zeros = ctx.saved_tensors
        n, kw, has_bias = ctx.meta
        rows, k = x.shape
        npad = wp.shape[0]
        dyp = F.pad(dy, (0, npad - n)).contiguous() if npad != n or not dy.is_contiguous() else dy
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            kpad = _ceil(k, 64)
            wt = F.pad(wp.t(), (0, 0, 0, kpad - k))                         # (kpad, npad) = W^T with zero rows
            dx = _dx(rows, npad, dyp, 0, kpad, pack(wt), e=x if ones is not None else None, es=ones, et=zeros)[:, :k]
        if ctx.needs_input_grad[1]:
            dw = _dw(rows, npad, dyp, 0, x, k, xs=ones, xt=zeros)[:n, :kw]
        if has_bias and ctx.needs_input_grad[2]:
            db = _colsum(dyp, npad)[:n]
        return dx, dw, db, None


class _LinearBNReLU(torch.autograd.Function):
    """h = relu(bn(act(x) W^T)) with batch statistics from the GEMM epilogue (a bias in front of a BatchNorm cancels; its
    gradient is exactly zero).  n % 64 == 0."""

    @staticmethod
    def forward(ctx, x, w, gamma, beta, pre_relu, bn):
        rows, k = x.shape
        n = w.shape[0]
        dev = x.device
        wp = F.pad(w, (0, k - w.shape[1])) if w.shape[1] != k else w
        ones = zeros = None
        if pre_relu:
            ones, zeros = torch.ones(k, device=dev), torch.zeros(k, device=dev)
        y, partial, _, _ = _fwd(x, k, pack(wp), n, scale=ones, shift=zeros, stats=True)
        stats, batch = _bn_stats(bn, partial, rows, n, gamma, beta)
        h = torch.empty(rows, n, dtype=torch.float32, device=dev)
        _lib.call("sg4d_bn_relu_apply", x, rows, n, y.data_ptr(), n, stats[0].data_ptr(), stats[1].data_ptr(), h.data_ptr(), n)
        ctx.save_for_backward(x, wp, y, h, stats, ones, zeros)
        ctx.meta = (w.shape[1], batch)
        return h

    @staticmethod
    def backward(ctx, dh):
        x, wp, y, h, stats, ones, zeros = ctx.saved_tensors
        kw, batch = ctx.meta
        rows, k = x.shape
        n = wp.shape[0]
        dev = x.device
        dh = dh.contiguous()
        dzs = torch.empty(rows, n, dtype=torch.float32, device=dev)
        part = torch.empty(_lib.load().sg4d_colsum_part_doubles(rows, n), dtype=torch.float64, device=dev)
        sums = torch.empty(2, n, dtype=torch.float32, device=dev)
        _lib.call("sg4d_bn_relu_bwd", x, rows, n, dh.data_ptr(), n, h.data_ptr(), n, y.data_ptr(), n, stats.data_ptr(),
                  dzs.data_ptr(), n, part.data_ptr(), sums.data_ptr())
        d_beta, d_gamma = sums[0], sums[1]
        s, m, i = stats[0], stats[2], stats[3]
        q1, u1, _ = fused.bn_bwd_coeffs(s, m, i, d_beta, d_gamma, rows, batch)
        p1 = torch.ones_like(s)
        dx = dw = None
        if ctx.needs_input_grad[0]:
            kpad = _ceil(k, 64)
            wt = F.pad(wp.t(), (0, 0, 0, kpad - k))
            dx = _dx(rows, n, y, 3, kpad, pack(wt), a2=dzs, p1=p1, q1=q1, u1=u1, e=x if ones is not None else None, es=ones,
                     et=zeros)[:, :k]
        if ctx.needs_input_grad[1]:
            dw = _dw(rows, n, y, 3, x, k, a2=dzs, p1=p1, q1=q1, u1=u1, xs=ones, xt=zeros)[:, :kw]
        return dx, dw, d_gamma, d_beta, None, None


def linear(x, layer, pre_relu=False):
    """``nn.Linear`` forward/backward on the engine; ``pre_relu`` fuses a ReLU on the input into the operand stager."""
    x = _rows_matrix(x, "x")
    if pre_relu and x.shape[1] % 64:
        raise RuntimeError("linear: a fused input ReLU needs an input width that is a multiple of 64")
    return _Linear.apply(x, layer.weight, layer.bias, bool(pre_relu))


def linear_bn_relu(x, layer, bn, pre_relu=False):
    """``nn.Linear`` (or a bias-free 1x1 conv given as a (n, k) weight) -> ``BatchNorm`` -> ``ReLU``."""
    x = _rows_matrix(x, "x")
    w = layer.weight.view(layer.weight.shape[0], -1)
    if w.shape[0] % 64 or not bn.affine:
        raise RuntimeError("linear_bn_relu: the layer width must be a multiple of 64 and the BatchNorm affine")
    h = _LinearBNReLU.apply(x, w, bn.weight, bn.bias, bool(pre_relu), bn)
    if getattr(layer, "bias", None) is not None and layer.bias.requires_grad and torch.is_grad_enabled():
        # the bias cancels inside the BatchNorm: attach an exactly-zero gradient so that optimizers see it like the reference
        h = h + 0.0 * layer.bias.sum()
    return h


class _DensePooledMLP(torch.autograd.Function):
    """[conv1x1 -> BN -> ReLU] x 2 -> max over ``group`` consecutive rows, any widths that are multiples of 64 (128 beyond
    128) -- the GroupAll level SA3 (259 -> 256 -> 256, one group of 128 points per cloud).  Same scheme as
    ``mlp._FusedSharedMLP``: pre-activations y1, y2 are the only stored activations."""

    @staticmethod
    def forward(ctx, x, w1, g1, be1, w2, g2, be2, group, dx_cols, bn1, bn2):
        rows, kp = x.shape
        n1, n2 = w1.shape[0], w2.shape[0]
        y1, part1, _, _ = _fwd(x, kp, pack(w1), n1, stats=True)
        st1, batch1 = _bn_stats(bn1, part1, rows, n1, g1, be1)
        y2, part2, gsel, garg = _fwd(y1, n1, pack(w2), n2, scale=st1[0], shift=st1[1], stats=True, group=group, gamma=g2)
        st2, batch2 = _bn_stats(bn2, part2, rows, n2, g2, be2)
        out = torch.relu(torch.addcmul(st2[1], gsel, st2[0]))
        fused._capture(kind="dense", garg=garg, gsel=gsel, y1=y1, y2=y2, s1=st1[0], t1=st1[1], out=out)
        ctx.save_for_backward(x, y1, y2, gsel, garg, out, w1, w2, st1, st2)
        ctx.meta = (group, dx_cols, batch1, batch2)
        return out

    @staticmethod
    def backward(ctx, d_out):
        x, y1, y2, gsel, garg, out, w1, w2, st1, st2 = ctx.saved_tensors
        group, dx_cols, batch1, batch2 = ctx.meta
        rows, kp = x.shape
        n1, n2 = w1.shape[0], w2.shape[0]
        dev = x.device
        lib = _lib.load()
        s1, t1, m1, i1 = st1[0], st1[1], st1[2], st1[3]
        dsel, a2, b2, d_g2, d_be2 = fused._pool_bwd_consts(d_out, out, gsel, st2[0], st2[2], st2[3], rows, batch2, x)
        em1 = (-m1 * i1).contiguous()
        dz1 = torch.empty(rows, n1, dtype=torch.float32, device=dev)
        part = torch.empty(lib.sg4d_dense_partial_doubles(rows, n1), dtype=torch.float64, device=dev)
        _lib.call("sg4d_pool_bwd_da", x, rows, n2, n1, group, y2.data_ptr(), a2.data_ptr(), b2.data_ptr(), dsel.data_ptr(),
                  garg.data_ptr(), pack(w2.t().contiguous()).data_ptr(), y1.data_ptr(), s1.data_ptr(), t1.data_ptr(), i1.data_ptr(),
                  em1.data_ptr(), dz1.data_ptr(), part.data_ptr())
        # the statistics partials are laid out per 128-column panel, one (sum dz1, sum dz1 * xhat1) pair per epilogue thread
        npan = n1 // (128 if n1 % 128 == 0 else 64)
        per = part.numel() // npan
        sums = torch.empty(npan, 2, n1 // npan, dtype=torch.float32, device=dev)
        for pnl in range(npan):
            _lib.call("sg4d_partial_sums", x, n1 // npan, per // 2, part[pnl * per:].data_ptr(), sums[pnl].data_ptr())
        d_be1, d_g1 = sums[:, 0].reshape(n1), sums[:, 1].reshape(n1)
        d_w2 = torch.empty(n2, n1, dtype=torch.float32, device=dev)
        wpart = torch.empty(lib.sg4d_dense_wgrad_partial_floats(rows, n2, n1), dtype=torch.float32, device=dev)
        _lib.call("sg4d_dense_pool_bwd_dw", x, rows, n2, n1, group, y2.data_ptr(), a2.data_ptr(), b2.data_ptr(), dsel.data_ptr(),
                  garg.data_ptr(), y1.data_ptr(), s1.data_ptr(), t1.data_ptr(), wpart.data_ptr(), d_w2.data_ptr())
        p1 = s1.contiguous()
        q1, u1, _ = fused.bn_bwd_coeffs(s1, m1, i1, d_be1.contiguous(), d_g1.contiguous(), rows, batch1)
        d_w1 = _dw(rows, n1, y1, 3, x, kp, a2=dz1, p1=p1, q1=q1, u1=u1)
        d_x = None
        if ctx.needs_input_grad[0]:
            # only the first dx_cols columns (the features, feature-first layout) carry a gradient downstream
            d_x = torch.zeros(rows, kp, dtype=torch.float32, device=dev)
            ncol = _ceil(dx_cols, 64)
            wt = F.pad(w1[:, :dx_cols].t(), (0, 0, 0, ncol - dx_cols))
            d_x[:, :dx_cols] = _dx(rows, n1, y1, 3, ncol, pack(wt), a2=dz1, p1=p1, q1=q1, u1=u1)[:, :dx_cols]
        return (d_x, d_w1, d_g1, d_be1, d_w2, d_g2, d_be2, None, None, None, None)


def _width(n):
    return 64 if n <= 64 else _ceil(n, 128)


def pooled_shared_mlp(x, k0, group, mlp, xyz_last=False):
    """x (R, K_padded) rows -> pooled (R / group, C_out): a ``build_shared_mlp`` stack of two [conv1x1 -> BN -> ReLU]
    blocks of ANY widths up to 1280 followed by the max over ``group`` consecutive rows (the GroupAll level SA3, and every
    set-abstraction shape outside the 64 / 128-wide fused kernels).  Widths that are not 64 or a multiple of 128 run
    zero-padded.  Column layout of x as in ``mlp.fused_shared_mlp``."""
    layers = fused._two_layer(mlp)
    if layers is None:
        raise NotImplementedError("sg4d evaluates shared MLPs of the form [Conv2d(1x1, bias=False) -> BatchNorm2d -> ReLU] x 2; "
                                  "other stacks have no kernel path (and there is no PyTorch fallback)")
    if group not in (1, 2, 4, 8, 16, 32, 64, 128):
        raise NotImplementedError(f"max-pool over {group} rows: the fused pool needs a power of two <= 128")
    c1, b1, c2, b2 = layers
    n1, n2 = c1.out_channels, c2.out_channels
    p1, p2 = _width(n1), _width(n2)
    if p2 == 64 and group > 64:
        p2 = 128
    if p1 > 1280 or p2 > 1280:
        raise NotImplementedError("layer widths beyond 1280 channels")
    kp = x.shape[1]
    w1 = c1.weight.view(n1, k0)
    if xyz_last:
        w1 = torch.cat([w1[:, 3:], w1[:, :3]], dim=1)
    w1 = F.pad(w1, (0, kp - k0, 0, p1 - n1))
    w2 = F.pad(c2.weight.view(n2, n1), (0, p1 - n1, 0, p2 - n2))
    g1, be1 = F.pad(b1.weight, (0, p1 - n1)), F.pad(b1.bias, (0, p1 - n1))
    g2, be2 = F.pad(b2.weight, (0, p2 - n2)), F.pad(b2.bias, (0, p2 - n2))
    bn1 = _PaddedBN(b1, p1) if p1 != n1 else b1
    bn2 = _PaddedBN(b2, p2) if p2 != n2 else b2
    dx_cols = (k0 - 3) if (xyz_last and x.requires_grad) else 0
    if x.requires_grad and not xyz_last:
        raise RuntimeError("input gradients need the feature-first (xyz_last) column layout")
    out = _DensePooledMLP.apply(x, w1, g1, be1, w2, g2, be2, int(group), int(dx_cols), bn1, bn2)
    return out[:, :n2] if p2 != n2 else out
