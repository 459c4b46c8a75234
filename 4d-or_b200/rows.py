"""Point-major ("rows") operators of the model path: thin autograd-aware wrappers over section 2 of
``include/sg4d.h``.  A cloud is ``(n, stride)`` fp32 with xyz in columns 0..2; see the header for
why this layout replaces the reference's channel-major tensors on the internal path.
"""
import ctypes

import torch

from . import _lib

_ONCHIP_MAX_POINTS = 16 * 1024 * 12
INDEX_MIN_POINTS = 12288    # clouds larger than this go through the spatial index (csrc/spatial.cu)
BALL_QUERY_PREFIX = 2048    # points scanned by brute force before the index answers the remaining centres


def _rows_ok(t, name):
    if t.dtype != torch.float32 or not t.is_contiguous() or t.dim() != 3:
        raise RuntimeError(f"{name} must be a contiguous (B, n, stride) float tensor")
    _lib.require_cuda(t)


class SpatialIndex:
    """Device workspace holding the Morton-sorted copy + bucket boxes of a batch of clouds (include/sg4d.h)."""

    def __init__(self, pts):
        _rows_ok(pts, "pts")
        b, n, s = pts.shape
        lib = _lib.load()
        if not lib.sg4d_spatial_index_supported(n):
            raise RuntimeError(f"spatial index: unsupported cloud size n={n}")
        self.b, self.n = b, n
        self.ws = torch.empty(lib.sg4d_spatial_index_bytes(b, n) // 4, dtype=torch.float32, device=pts.device)
        _lib.call("sg4d_spatial_index_build", pts, b, n, s, pts.data_ptr(), self.ws.data_ptr())
        self.fps_consumed = False


def wants_index(n, min_points=None):
    """True when clouds of n points should go through the spatial index."""
    lim = INDEX_MIN_POINTS if min_points is None else min_points
    return n > lim and bool(_lib.load().sg4d_spatial_index_supported(n))


def fps_rows(pts, npoint, index=None):
    """FPS + gather of the picked xyz.  pts (B,n,S) -> idx (B,npoint) int32, new_xyz (B,npoint,3).
    Replaces ``furthest_point_sample`` + ``gather_operation`` + transposes
    (OPS/pointnet2_modules.py:50-59).  With ``index`` (a fresh SpatialIndex of pts) the bucket-pruned
    kernel runs; the result is bit-identical either way."""
    _rows_ok(pts, "pts")
    b, n, s = pts.shape
    idx = torch.empty(b, npoint, dtype=torch.int32, device=pts.device)
    new_xyz = torch.empty(b, npoint, 3, dtype=torch.float32, device=pts.device)
    if index is not None:
        if index.fps_consumed or (index.b, index.n) != (b, n):
            raise RuntimeError("fps_rows needs a fresh SpatialIndex of the same clouds")
        index.fps_consumed = True
        _lib.call("sg4d_fps_indexed", pts, b, n, npoint, s, pts.data_ptr(), index.ws.data_ptr(), idx.data_ptr(),
                  new_xyz.data_ptr())
        return idx, new_xyz
    tmp = torch.empty(b, n, dtype=torch.float32, device=pts.device) if n > _ONCHIP_MAX_POINTS else None
    _lib.call("sg4d_fps_rows", pts, b, n, npoint, s, pts.data_ptr(), _lib.ptr(tmp), idx.data_ptr(),
              new_xyz.data_ptr())
    return idx, new_xyz


def ball_query_rows(centers, pts, radii, nsamples, index=None, prefix=None):
    """All radii of an MSG level in one scan.  centers (B,m,3|S'), pts (B,n,S) ->
    ([idx_s (B,m,ns_s) int32], [cnt_s (B,m) int32]).  Replaces one ``ball_query`` per scale
    (OPS/pointnet2_utils.py:318).  With ``index``: brute-force prefix + spatial index (same result)."""
    _rows_ok(pts, "pts")
    _rows_ok(centers, "centers")
    b, n, s = pts.shape
    m, cs = centers.shape[1], centers.shape[2]
    k = len(radii)
    idx = [torch.empty(b, m, int(ns), dtype=torch.int32, device=pts.device) for ns in nsamples]
    cnt = [torch.empty(b, m, dtype=torch.int32, device=pts.device) for _ in nsamples]
    r_arr = (ctypes.c_float * k)(*[float(r) for r in radii])
    ns_arr = (ctypes.c_int * k)(*[int(v) for v in nsamples])
    idx_arr = (ctypes.c_void_p * k)(*[t.data_ptr() for t in idx])
    cnt_arr = (ctypes.c_void_p * k)(*[t.data_ptr() for t in cnt])
    if index is not None:
        if (index.b, index.n) != (b, n):
            raise RuntimeError("ball_query_rows: the SpatialIndex belongs to other clouds")
        _lib.call("sg4d_ball_query_rows_indexed", pts, b, n, m, s, cs, k, ctypes.cast(r_arr, ctypes.c_void_p),
                  ctypes.cast(ns_arr, ctypes.c_void_p), centers.data_ptr(), pts.data_ptr(), index.ws.data_ptr(),
                  int(BALL_QUERY_PREFIX if prefix is None else prefix), ctypes.cast(idx_arr, ctypes.c_void_p),
                  ctypes.cast(cnt_arr, ctypes.c_void_p))
        return idx, cnt
    _lib.call("sg4d_ball_query_rows", pts, b, n, m, s, cs, k, ctypes.cast(r_arr, ctypes.c_void_p),
              ctypes.cast(ns_arr, ctypes.c_void_p), centers.data_ptr(), pts.data_ptr(),
              ctypes.cast(idx_arr, ctypes.c_void_p), ctypes.cast(cnt_arr, ctypes.c_void_p))
    return idx, cnt


class _GroupRows(torch.autograd.Function):
    """out[b,j,k] = [xyz(pts[idx]) - centre | feats[idx, off:off+c] | 0-pad]  (B,m,ns,out_stride).
    Replaces 2x ``grouping_operation`` + subtract + ``cat`` (OPS/pointnet2_utils.py:319-328);
    backward replaces ``group_points_grad`` (:236-241) with a deterministic gather."""

    @staticmethod
    def forward(ctx, pts, feats, centers, idx, cnt, c, feat_offset, out_stride, xyz_last):
        b, n, ps = pts.shape
        m, ns = idx.shape[1], idx.shape[2]
        fs = feats.shape[2] if feats is not None else ps
        out = torch.empty(b, m, ns, out_stride, dtype=torch.float32, device=pts.device)
        _lib.call("sg4d_group_rows", pts, b, n, m, ns, c, ps, fs, feat_offset, out_stride, c if xyz_last else 0,
                  pts.data_ptr(), _lib.ptr(feats), centers.data_ptr(), idx.data_ptr(), out.data_ptr())
        ctx.save_for_backward(idx, cnt)
        ctx.meta = (b, n, m, ns, c, fs, feat_offset, out_stride, xyz_last)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        idx, cnt = ctx.saved_tensors
        b, n, m, ns, c, fs, feat_offset, out_stride, xyz_last = ctx.meta
        grad_feats = None
        if ctx.needs_input_grad[1]:
            if feat_offset != 0 or fs != c:
                raise RuntimeError("group_rows backward expects a dense (B,n,c) feature tensor")
            grad_out = grad_out.contiguous()
            grad_feats = torch.empty(b, n, c, dtype=torch.float32, device=grad_out.device)
            _lib.call("sg4d_group_rows_grad", grad_out, b, n, m, ns, c, out_stride, 0 if xyz_last else 3, 0,
                      grad_out.data_ptr(), idx.data_ptr(), cnt.data_ptr(), grad_feats.data_ptr())
        return None, grad_feats, None, None, None, None, None, None, None


def group_rows(pts, feats, centers, idx, cnt, c, feat_offset=0, out_stride=None, xyz_last=False):
    """xyz_last=False: columns [xyz | feats | 0-pad] (the reference's channel order);
    xyz_last=True:  columns [feats | xyz | 0-pad] (feature columns 16-byte aligned)."""
    _rows_ok(pts, "pts")
    if feats is not None:
        _rows_ok(feats, "feats")
    if out_stride is None:
        out_stride = 3 + c
    return _GroupRows.apply(pts, feats, centers, idx, cnt, int(c), int(feat_offset), int(out_stride), bool(xyz_last))


# ---------------------------------------------------------------------------------------- GNN

class EdgeCSR:
    """Destination- and source-sorted edge lists of one batch (built once, shared by all layers)."""

    def __init__(self, edge_index, n_nodes):
        if edge_index.dim() != 2 or edge_index.shape[0] != 2:
            raise RuntimeError("edge_index must have shape (2, E)")
        if edge_index.dtype != torch.long:      # the kernels read int64_t; anything else would be an out-of-bounds read
            edge_index = edge_index.long()
        src, dst = edge_index[0].contiguous(), edge_index[1].contiguous()
        self.src, self.dst, self.n_nodes, self.n_edges = src, dst, int(n_nodes), int(src.numel())
        self.by_dst = self._csr(dst)
        self.by_src = self._csr(src)

    def _csr(self, key):
        # sort + searchsorted: no device->host synchronisation (torch.bincount would need the maximum on the host)
        sorted_key, order = torch.sort(key, stable=True)
        bounds = torch.arange(self.n_nodes + 1, dtype=key.dtype, device=key.device)
        ptr = torch.searchsorted(sorted_key, bounds).to(torch.int32)
        return order.to(torch.int32).contiguous(), ptr


def _segment_sum(src, col0, d, order_ptr, n_nodes, col1=None):
    order, ptr = order_ptr
    out = torch.empty(n_nodes, d, dtype=torch.float32, device=src.device)
    _lib.call("sg4d_segment_sum", src, n_nodes, d, src.stride(0), col0, 0 if col1 is None else col1,
              src.data_ptr(), 0 if col1 is None else 1, order.data_ptr(), ptr.data_ptr(), out.data_ptr())
    return out


class _TripletGather(torch.autograd.Function):
    """[x[dst] | e | x[src]] (network_TripletGCN.py:45-46 + PyG __collect__)."""

    @staticmethod
    def forward(ctx, x, edge_feat, csr):
        x, edge_feat = x.contiguous(), edge_feat.contiguous()
        d, de = x.shape[1], edge_feat.shape[1]
        out = torch.empty(csr.n_edges, 2 * d + de, dtype=torch.float32, device=x.device)
        _lib.call("sg4d_triplet_gather", x, csr.n_edges, d, de, x.data_ptr(), edge_feat.data_ptr(),
                  csr.src.data_ptr(), csr.dst.data_ptr(), out.data_ptr())
        ctx.csr, ctx.dims = csr, (d, de)
        return out

    @staticmethod
    def backward(ctx, g):
        csr, (d, de) = ctx.csr, ctx.dims
        g = g.contiguous()
        gx = _segment_sum(g, 0, d, csr.by_dst, csr.n_nodes) + _segment_sum(g, d + de, d, csr.by_src, csr.n_nodes)
        return gx, g[:, d:d + de], None


class _MessageAggregate(torch.autograd.Function):
    """m[v] = sum_{e: dst[e]=v} (h[e,:dh] + h[e,dh+de:])  (network_TripletGCN.py:48-58)."""

    @staticmethod
    def forward(ctx, h, dh, de, csr):
        h = h.contiguous()
        ctx.csr, ctx.dims = csr, (dh, de, h.shape[1])
        return _segment_sum(h, 0, dh, csr.by_dst, csr.n_nodes, col1=dh + de)

    @staticmethod
    def backward(ctx, g):
        csr, (dh, de, w) = ctx.csr, ctx.dims
        ge = g.index_select(0, csr.dst)
        gh = torch.zeros(csr.n_edges, w, dtype=torch.float32, device=g.device)
        gh[:, :dh] = ge
        gh[:, dh + de:] = ge
        return gh, None, None, None


def triplet_gather(x, edge_feat, csr):
    _lib.require_cuda(x, edge_feat)
    return _TripletGather.apply(x, edge_feat, csr)


def message_aggregate(h, dim_hidden, dim_edge, csr):
    _lib.require_cuda(h)
    return _MessageAggregate.apply(h, int(dim_hidden), int(dim_edge), csr)
