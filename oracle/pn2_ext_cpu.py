"""oracle/pn2_ext_cpu.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

ctypes front-end to ``oracle/_build/libpn2_oracle.so`` (the plain-C restatement of the reference
CUDA kernels, ``oracle/pn2_oracle.c``) exposing the same nine function names, argument order and
allocation rules as the reference's pybind module ``pointnet2_ops._ext``
(``_ext-src/src/bindings.cpp:6-19``), but on CPU tensors.  Because the names match it can be
registered as ``sys.modules['pointnet2_ops._ext']`` underneath the reference's *unmodified*
``pointnet2_utils.py`` / ``pointnet2_modules.py`` (see ``oracle/ref_harness.py``).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs import this.
"""
import ctypes
import os
import subprocess

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libpn2_oracle.so")


def build(force=False):
    """Compile the C restatement (gcc, a second or two)."""
    src = os.path.join(_HERE, "pn2_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "_build/libpn2_oracle.so"],
                              stdout=subprocess.DEVNULL)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_SO)
        i, f, p = ctypes.c_int, ctypes.c_float, ctypes.c_void_p
        L.oracle_opt_n_threads.argtypes = [i]
        L.oracle_opt_n_threads.restype = i
        L.oracle_furthest_point_sampling.argtypes = [i, i, i, p, p, p]
        L.oracle_gather_points.argtypes = [i, i, i, i, p, p, p]
        L.oracle_gather_points_grad.argtypes = [i, i, i, i, p, p, p]
        L.oracle_ball_query.argtypes = [i, i, i, f, i, p, p, p]
        L.oracle_group_points.argtypes = [i, i, i, i, i, p, p, p]
        L.oracle_group_points_grad.argtypes = [i, i, i, i, i, p, p, p]
        L.oracle_three_nn.argtypes = [i, i, i, p, p, p, p]
        L.oracle_three_interpolate.argtypes = [i, i, i, i, p, p, p, p]
        L.oracle_three_interpolate_grad.argtypes = [i, i, i, i, p, p, p, p]
        for fn in ("oracle_furthest_point_sampling", "oracle_gather_points",
                   "oracle_gather_points_grad", "oracle_ball_query", "oracle_group_points",
                   "oracle_group_points_grad", "oracle_three_nn", "oracle_three_interpolate",
                   "oracle_three_interpolate_grad"):
            getattr(L, fn).restype = None
        _lib = L
    return _lib


def _chk(t, dtype, name):
    # mirrors CHECK_CONTIGUOUS / CHECK_IS_FLOAT / CHECK_IS_INT (_ext-src/include/utils.h:5-25)
    if not t.is_contiguous():
        raise RuntimeError(f"{name} must be a contiguous tensor")
    if t.dtype != dtype:
        raise RuntimeError(f"{name} must be a {'float' if dtype == torch.float32 else 'int'} tensor")
    if t.is_cuda:
        raise RuntimeError("oracle runs on CPU tensors only")


def opt_n_threads(n):
    return lib().oracle_opt_n_threads(int(n))


def furthest_point_sampling(points, nsamples):
    _chk(points, torch.float32, "points")
    b, n, _ = points.shape
    out = torch.zeros(b, nsamples, dtype=torch.int32)
    tmp = torch.full((b, n), 1e10, dtype=torch.float32)
    lib().oracle_furthest_point_sampling(b, n, nsamples, points.data_ptr(), tmp.data_ptr(),
                                         out.data_ptr())
    return out


def gather_points(points, idx):
    _chk(points, torch.float32, "points")
    _chk(idx, torch.int32, "idx")
    b, c, n = points.shape
    m = idx.shape[1]
    out = torch.zeros(b, c, m, dtype=torch.float32)
    lib().oracle_gather_points(b, c, n, m, points.data_ptr(), idx.data_ptr(), out.data_ptr())
    return out


def gather_points_grad(grad_out, idx, n):
    _chk(grad_out, torch.float32, "grad_out")
    _chk(idx, torch.int32, "idx")
    b, c, m = grad_out.shape
    out = torch.zeros(b, c, n, dtype=torch.float32)
    lib().oracle_gather_points_grad(b, c, n, m, grad_out.data_ptr(), idx.data_ptr(), out.data_ptr())
    return out


def ball_query(new_xyz, xyz, radius, nsample):
    _chk(new_xyz, torch.float32, "new_xyz")
    _chk(xyz, torch.float32, "xyz")
    b, n, _ = xyz.shape
    m = new_xyz.shape[1]
    idx = torch.zeros(b, m, nsample, dtype=torch.int32)
    lib().oracle_ball_query(b, n, m, float(radius), nsample, new_xyz.data_ptr(), xyz.data_ptr(),
                            idx.data_ptr())
    return idx


def group_points(points, idx):
    _chk(points, torch.float32, "points")
    _chk(idx, torch.int32, "idx")
    b, c, n = points.shape
    _, npoints, nsample = idx.shape
    out = torch.zeros(b, c, npoints, nsample, dtype=torch.float32)
    lib().oracle_group_points(b, c, n, npoints, nsample, points.data_ptr(), idx.data_ptr(),
                              out.data_ptr())
    return out


def group_points_grad(grad_out, idx, n):
    _chk(grad_out, torch.float32, "grad_out")
    _chk(idx, torch.int32, "idx")
    b, c, npoints, nsample = grad_out.shape
    out = torch.zeros(b, c, n, dtype=torch.float32)
    lib().oracle_group_points_grad(b, c, n, npoints, nsample, grad_out.data_ptr(), idx.data_ptr(),
                                   out.data_ptr())
    return out


def three_nn(unknowns, knows):
    _chk(unknowns, torch.float32, "unknowns")
    _chk(knows, torch.float32, "knows")
    b, n, _ = unknowns.shape
    m = knows.shape[1]
    dist2 = torch.zeros(b, n, 3, dtype=torch.float32)
    idx = torch.zeros(b, n, 3, dtype=torch.int32)
    lib().oracle_three_nn(b, n, m, unknowns.data_ptr(), knows.data_ptr(), dist2.data_ptr(), idx.data_ptr())
    return [dist2, idx]


def three_interpolate(points, idx, weight):
    _chk(points, torch.float32, "points")
    _chk(idx, torch.int32, "idx")
    _chk(weight, torch.float32, "weight")
    b, c, m = points.shape
    n = idx.shape[1]
    out = torch.zeros(b, c, n, dtype=torch.float32)
    lib().oracle_three_interpolate(b, c, m, n, points.data_ptr(), idx.data_ptr(), weight.data_ptr(), out.data_ptr())
    return out


def three_interpolate_grad(grad_out, idx, weight, m):
    _chk(grad_out, torch.float32, "grad_out")
    _chk(idx, torch.int32, "idx")
    _chk(weight, torch.float32, "weight")
    b, c, n = grad_out.shape
    out = torch.zeros(b, c, m, dtype=torch.float32)
    lib().oracle_three_interpolate_grad(b, c, n, m, grad_out.data_ptr(), idx.data_ptr(), weight.data_ptr(),
                                        out.data_ptr())
    return out
