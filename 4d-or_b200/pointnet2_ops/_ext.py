"""Python face of the native FFI: same nine names, argument order, dtype/contiguity checks, output
allocation and error type as the reference's pybind module ``pointnet2_ops._ext``
(``_ext-src/src/bindings.cpp:6-19`` and the host wrappers ``sampling.cpp``, ``ball_query.cpp``,
``group_points.cpp``), implemented by calls into ``libsg4d.so`` (``include/sg4d.h``, section 1).

Differences, all deliberate: outputs that the kernels fully overwrite are allocated with ``empty``
instead of ``zeros``; the FPS scratch ``temp`` is only materialised for clouds too large to stay on
chip; CPU tensors raise ``RuntimeError("CPU not supported")`` exactly like the reference.  All nine functions
are implemented (the three feature-propagation ops live in ``csrc/interpolate.cu``).
"""
import torch

from .. import _lib, rows

_ONCHIP_MAX_POINTS = 16 * 1024 * 12  # fps.cu: 16-CTA cluster x 1024 threads x 12 points


def _check(t, dtype, name):
    if not t.is_contiguous():
        raise RuntimeError(f"{name} must be a contiguous tensor")
    if t.dtype != dtype:
        kind = "float" if dtype == torch.float32 else "int"
        raise RuntimeError(f"{name} must be a {kind} tensor")


def furthest_point_sampling(points, nsamples):
    """(B,N,3) fp32 -> (B,nsamples) int32.  sampling.cpp:66-87."""
    _check(points, torch.float32, "points")
    _lib.require_cuda(points)
    b, n, _ = points.shape
    if rows.wants_index(n):   # large clouds: bucket-pruned FPS on a spatial index (same picks, bit for bit)
        return rows.fps_rows(points, nsamples, rows.SpatialIndex(points))[0]
    out = torch.empty(b, nsamples, dtype=torch.int32, device=points.device)
    tmp = torch.empty(b, n, dtype=torch.float32, device=points.device) if n > _ONCHIP_MAX_POINTS else None
    _lib.call("sg4d_furthest_point_sampling", points, b, n, nsamples, points.data_ptr(), _lib.ptr(tmp),
              out.data_ptr())
    return out


def gather_points(points, idx):
    """(B,C,N) fp32, (B,m) int32 -> (B,C,m).  sampling.cpp:15-38."""
    _check(points, torch.float32, "points")
    _check(idx, torch.int32, "idx")
    _lib.require_cuda(points, idx)
    b, c, n = points.shape
    m = idx.shape[1]
    out = torch.empty(b, c, m, dtype=torch.float32, device=points.device)
    _lib.call("sg4d_gather_points", points, b, c, n, m, points.data_ptr(), idx.data_ptr(), out.data_ptr())
    return out


def gather_points_grad(grad_out, idx, n):
    """(B,C,m), (B,m) -> (B,C,n).  sampling.cpp:40-64."""
    _check(grad_out, torch.float32, "grad_out")
    _check(idx, torch.int32, "idx")
    _lib.require_cuda(grad_out, idx)
    b, c, m = grad_out.shape
    out = torch.zeros(b, c, n, dtype=torch.float32, device=grad_out.device)
    _lib.call("sg4d_gather_points_grad", grad_out, b, c, n, m, grad_out.data_ptr(), idx.data_ptr(),
              out.data_ptr())
    return out


def ball_query(new_xyz, xyz, radius, nsample):
    """(B,m,3), (B,N,3) -> (B,m,nsample) int32.  ball_query.cpp:8-32."""
    _check(new_xyz, torch.float32, "new_xyz")
    _check(xyz, torch.float32, "xyz")
    _lib.require_cuda(new_xyz, xyz)
    b, n, _ = xyz.shape
    m = new_xyz.shape[1]
    if rows.wants_index(n) and nsample <= 64:
        return rows.ball_query_rows(new_xyz, xyz, [radius], [nsample], rows.SpatialIndex(xyz))[0][0]
    idx = torch.empty(b, m, nsample, dtype=torch.int32, device=xyz.device)
    _lib.call("sg4d_ball_query", xyz, b, n, m, float(radius), int(nsample), new_xyz.data_ptr(),
              xyz.data_ptr(), idx.data_ptr())
    return idx


def group_points(points, idx):
    """(B,C,N), (B,m,ns) -> (B,C,m,ns).  group_points.cpp:12-36."""
    _check(points, torch.float32, "points")
    _check(idx, torch.int32, "idx")
    _lib.require_cuda(points, idx)
    b, c, n = points.shape
    _, m, ns = idx.shape
    out = torch.empty(b, c, m, ns, dtype=torch.float32, device=points.device)
    _lib.call("sg4d_group_points", points, b, c, n, m, ns, points.data_ptr(), idx.data_ptr(), out.data_ptr())
    return out


def group_points_grad(grad_out, idx, n):
    """(B,C,m,ns), (B,m,ns) -> (B,C,n).  group_points.cpp:38-62."""
    _check(grad_out, torch.float32, "grad_out")
    _check(idx, torch.int32, "idx")
    _lib.require_cuda(grad_out, idx)
    b, c, m, ns = grad_out.shape
    out = torch.zeros(b, c, n, dtype=torch.float32, device=grad_out.device)
    _lib.call("sg4d_group_points_grad", grad_out, b, c, n, m, ns, grad_out.data_ptr(), idx.data_ptr(),
              out.data_ptr())
    return out


def three_nn(unknowns, knows):
    """(B,n,3), (B,m,3) -> [dist2 (B,n,3) fp32, idx (B,n,3) int32].  interpolate.cpp:14-40."""
    _check(unknowns, torch.float32, "unknowns")
    _check(knows, torch.float32, "knows")
    _lib.require_cuda(unknowns, knows)
    b, n, _ = unknowns.shape
    m = knows.shape[1]
    dist2 = torch.empty(b, n, 3, dtype=torch.float32, device=unknowns.device)
    idx = torch.empty(b, n, 3, dtype=torch.int32, device=unknowns.device)
    _lib.call("sg4d_three_nn", unknowns, b, n, m, unknowns.data_ptr(), knows.data_ptr(), dist2.data_ptr(), idx.data_ptr())
    return [dist2, idx]


def three_interpolate(points, idx, weight):
    """(B,c,m), (B,n,3) int32, (B,n,3) -> (B,c,n).  interpolate.cpp:42-70."""
    _check(points, torch.float32, "points")
    _check(idx, torch.int32, "idx")
    _check(weight, torch.float32, "weight")
    _lib.require_cuda(points, idx, weight)
    b, c, m = points.shape
    n = idx.shape[1]
    out = torch.empty(b, c, n, dtype=torch.float32, device=points.device)
    _lib.call("sg4d_three_interpolate", points, b, c, m, n, points.data_ptr(), idx.data_ptr(), weight.data_ptr(),
              out.data_ptr())
    return out


def three_interpolate_grad(grad_out, idx, weight, m):
    """(B,c,n), (B,n,3), (B,n,3) -> (B,c,m).  interpolate.cpp:71-99."""
    _check(grad_out, torch.float32, "grad_out")
    _check(idx, torch.int32, "idx")
    _check(weight, torch.float32, "weight")
    _lib.require_cuda(grad_out, idx, weight)
    b, c, n = grad_out.shape
    out = torch.zeros(b, c, m, dtype=torch.float32, device=grad_out.device)
    _lib.call("sg4d_three_interpolate_grad", grad_out, b, c, n, m, grad_out.data_ptr(), idx.data_ptr(),
              weight.data_ptr(), out.data_ptr())
    return out
