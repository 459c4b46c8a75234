"""ctypes binding of ``libsg4d.so`` (``include/sg4d.h``).

The library is the product: there is NO fallback.  If the shared object is missing or a call returns
a non-zero status a ``RuntimeError`` is raised (the reference's launchers print and ``exit(-1)``,
``_ext-src/include/cuda_utils.h:30-39``; an exception is the Python-visible equivalent of its
``AT_ASSERT`` host checks, ``_ext-src/include/utils.h:5-25``).

Tensors cross the boundary as raw device pointers (``Tensor.data_ptr()``) plus the current CUDA
stream handle; no torch types appear in the C signatures.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.environ.get("SG4D_LIBRARY") or os.path.join(_HERE, "libsg4d.so")     # SG4D_LIBRARY: a developer (debug) build

_i, _i64, _f, _p = ctypes.c_int, ctypes.c_int64, ctypes.c_float, ctypes.c_void_p

# name -> argtypes; every function returns an int status.  Kept in the order of include/sg4d.h.
SIGNATURES = {
    "sg4d_furthest_point_sampling": [_i, _i, _i, _p, _p, _p, _p],
    "sg4d_gather_points": [_i, _i, _i, _i, _p, _p, _p, _p],
    "sg4d_gather_points_grad": [_i, _i, _i, _i, _p, _p, _p, _p],
    "sg4d_ball_query": [_i, _i, _i, _f, _i, _p, _p, _p, _p],
    "sg4d_group_points": [_i, _i, _i, _i, _i, _p, _p, _p, _p],
    "sg4d_group_points_grad": [_i, _i, _i, _i, _i, _p, _p, _p, _p],
    "sg4d_three_nn": [_i, _i, _i, _p, _p, _p, _p, _p],
    "sg4d_three_interpolate": [_i, _i, _i, _i, _p, _p, _p, _p, _p],
    "sg4d_three_interpolate_grad": [_i, _i, _i, _i, _p, _p, _p, _p, _p],
    "sg4d_fps_rows": [_i, _i, _i, _i, _p, _p, _p, _p, _p],
    "sg4d_ball_query_rows": [_i, _i, _i, _i, _i, _i, _p, _p, _p, _p, _p, _p, _p],
    "sg4d_group_rows": [_i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _p, _p, _p, _p, _p, _p],
    "sg4d_group_rows_grad": [_i, _i, _i, _i, _i, _i, _i, _i, _p, _p, _p, _p, _p],
    "sg4d_group_rows_grad_dy": [_i, _i, _i, _i, _i, _p, _p, _p, _p, _p, _p, _p, _p, _p],
    "sg4d_gather_y1": [_i64, _i, _i, _i, _i, _p, _p, _p, _p, _p, _p],
    "sg4d_group_sum_dy": [_i64, _i, _i, _p, _p, _p, _p, _p, _p, _p],
    "sg4d_spatial_index_build": [_i, _i, _i, _p, _p, _p],
    "sg4d_fps_indexed": [_i, _i, _i, _i, _p, _p, _p, _p, _p],
    "sg4d_ball_query_rows_indexed": [_i, _i, _i, _i, _i, _i, _p, _p, _p, _p, _p, _i, _p, _p, _p],
    "sg4d_triplet_gather": [_i64, _i, _i, _p, _p, _p, _p, _p, _p],
    "sg4d_segment_sum": [_i, _i, _i64, _i, _i, _p, _i, _p, _p, _p, _p],
    "sg4d_pack_weight": [_i, _i, _i, _p, _p, _p],
    "sg4d_linear_fwd": [_i64, _i, _i, _i, _i, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p],
    "sg4d_bn_finalize": [_i, _i, _i64, _p, _p, _p, _f, _f, _p, _p, _p, _p, _p, _p, _p],
    "sg4d_pool_bwd_da": [_i64, _i, _i, _i, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p],
    "sg4d_pool_bwd_dw": [_i64, _i, _i, _i, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p],
    "sg4d_inner_bwd_dx": [_i64, _i, _i, _p, _p, _p, _p, _p, _p, _p, _i, _i, _p],
    "sg4d_inner_bwd_dw": [_i64, _i, _i, _i, _p, _p, _p, _p, _p, _p, _p, _p, _p],
    "sg4d_pool_bwd_prologue": [_i64, _i, _i, _p, _p, _p, _p, _p, _p, _p, _p, _p],
    "sg4d_partial_sums": [_i, _i, _p, _p, _p],
    "sg4d_bn_bwd_coeffs": [_i, _i64, _i, _p, _p, _p, _p, _p, _p, _p],
    "sg4d_sa_moments": [_i64] + [_i] * 7 + [_p] * 5 + [_p],
    "sg4d_sa1_bn1": [_i, _i, _p, _p, _i, _p, _p, _f, _f, _p, _p, _i, _p, _p, _p, _p],
    "sg4d_sa1_fwd": [_i64] + [_i] * 7 + [_p] * 6 + [_i] + [_p] * 6 + [_p],
    "sg4d_sa1_bwd_da": [_i64] + [_i] * 7 + [_p] * 6 + [_i] + [_p] * 7 + [_p],
    "sg4d_sa1_bwd_dw2": [_i64] + [_i] * 7 + [_p] * 6 + [_i] + [_p] * 7 + [_p],
    "sg4d_sa1_bwd_dw2_gram": [_i64] + [_i] * 7 + [_p] * 6 + [_i] + [_p] * 7 + [_p],
    "sg4d_sa1_bwd_finalize": [_i, _i64, _p, _p, _p, _i, _p, _i, _p, _i, _p, _p, _p],
    "sg4d_linear_fwd_grouped": [_i64] + [_i] * 7 + [_p] * 4 + [_i] + [_p] * 3 + [_p],
    "sg4d_dense_pack_weight": [_i, _i, _i, _p, _p, _p],
    "sg4d_dense_fwd": [_i64, _i, _i, _i, _p, _p, _p, _p, _p, _p, _i, _p, _i, _p, _p, _p, _p],
    "sg4d_dense_bn_finalize": [_i, _i64, _p, _p, _p, _f, _f, _p, _p, _p, _p],
    "sg4d_bn_relu_apply": [_i64, _i, _p, _i, _p, _p, _p, _i, _p],
    "sg4d_bn_relu_bwd": [_i64, _i, _p, _i, _p, _i, _p, _i, _p, _p, _i, _p, _p, _p],
    "sg4d_colsum": [_i64, _i, _p, _i, _p, _p, _p],
    "sg4d_dense_bwd_dx": [_i64, _i, _i, _i, _i, _p, _p, _p, _p, _p, _p, _p, _i, _p, _p, _p, _i, _p, _p],
    "sg4d_dense_bwd_dw": [_i64, _i, _i, _i, _i, _p, _p, _p, _p, _p, _p, _i, _p, _p, _p, _p, _i, _p],
    "sg4d_dense_pool_bwd_dw": [_i64, _i, _i, _i, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p],
    "sg4d_frontend_objects": [_i, _i, _i, _p, _p, _f, _p, _p, _p, _p, _p],
    "sg4d_frontend_edges": [_i, _i, _i, _p, _p, _p, _p, _p, _p, _p, _p, _p],
    "sg4d_frontend_sample": [_i, _i, _i, _i, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p],
    "sg4d_inner_bwd_dw_grouped": [_i64] + [_i] * 7 + [_p] * 4 + [_i] + [_p] * 7 + [_i, _p],
}
OTHER_SYMBOLS = ["sg4d_abi_version", "sg4d_error_string", "sg4d_check_device", "sg4d_mlp_grid",
                 "sg4d_weight_image_floats", "sg4d_wgrad_partial_floats", "sg4d_mlp_partial_doubles", "sg4d_pool_bwd_prologue_parts", "sg4d_spatial_index_bytes",
                 "sg4d_spatial_index_supported", "sg4d_sa_moments_parts", "sg4d_sa1_s1part_doubles",
                 "sg4d_dense_weight_floats", "sg4d_dense_partial_doubles", "sg4d_colsum_part_doubles",
                 "sg4d_dense_wgrad_partial_floats", "sg4d_frontend_workspace_bytes", "sg4d_set_compute_precision",
                 "sg4d_get_compute_precision", "sg4d_gather_y1_parts",
                 "sg4d_sa1_bwd_dw2_gram_ws_floats"]

_lib = None


def load():
    """Load libsg4d.so (once).  Raises if it has not been built -- there is no CPU/PyTorch fallback."""
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise RuntimeError(
                f"{SO_PATH} is missing: build it with `python 4d-or_b200/build.py` "
                "(or `python -c 'import __graft_entry__ as g; g.build()'`). sg4d has no fallback path.")
        lib = ctypes.CDLL(SO_PATH)
        for name, args in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.argtypes, fn.restype = args, _i
        lib.sg4d_abi_version.restype = _i
        lib.sg4d_error_string.argtypes, lib.sg4d_error_string.restype = [_i], ctypes.c_char_p
        lib.sg4d_check_device.restype = _i
        lib.sg4d_mlp_grid.argtypes, lib.sg4d_mlp_grid.restype = [_i64], _i
        lib.sg4d_mlp_partial_doubles.argtypes, lib.sg4d_mlp_partial_doubles.restype = [_i64], _i64
        lib.sg4d_weight_image_floats.argtypes, lib.sg4d_weight_image_floats.restype = [_i, _i], _i64
        lib.sg4d_wgrad_partial_floats.argtypes, lib.sg4d_wgrad_partial_floats.restype = [_i64, _i], _i64
        lib.sg4d_pool_bwd_prologue_parts.argtypes, lib.sg4d_pool_bwd_prologue_parts.restype = [], _i
        lib.sg4d_spatial_index_bytes.argtypes, lib.sg4d_spatial_index_bytes.restype = [_i, _i], _i64
        lib.sg4d_spatial_index_supported.argtypes, lib.sg4d_spatial_index_supported.restype = [_i], _i
        lib.sg4d_sa_moments_parts.argtypes, lib.sg4d_sa_moments_parts.restype = [], _i
        lib.sg4d_sa1_s1part_doubles.argtypes, lib.sg4d_sa1_s1part_doubles.restype = [_i64], _i64
        lib.sg4d_dense_weight_floats.argtypes, lib.sg4d_dense_weight_floats.restype = [_i, _i], _i64
        lib.sg4d_dense_partial_doubles.argtypes, lib.sg4d_dense_partial_doubles.restype = [_i64, _i], _i64
        lib.sg4d_colsum_part_doubles.argtypes, lib.sg4d_colsum_part_doubles.restype = [_i64, _i], _i64
        lib.sg4d_dense_wgrad_partial_floats.argtypes, lib.sg4d_dense_wgrad_partial_floats.restype = [_i64, _i, _i], _i64
        lib.sg4d_frontend_workspace_bytes.argtypes, lib.sg4d_frontend_workspace_bytes.restype = [_i, _i, _i], _i64
        lib.sg4d_set_compute_precision.argtypes, lib.sg4d_set_compute_precision.restype = [_i], _i
        lib.sg4d_get_compute_precision.argtypes, lib.sg4d_get_compute_precision.restype = [], _i
        lib.sg4d_gather_y1_parts.argtypes, lib.sg4d_gather_y1_parts.restype = [_i], _i
        lib.sg4d_sa1_bwd_dw2_gram_ws_floats.argtypes, lib.sg4d_sa1_bwd_dw2_gram_ws_floats.restype = [_i64, _i], _i64
        if lib.sg4d_abi_version() != 1:
            raise RuntimeError("libsg4d.so ABI version mismatch; rebuild it")
        _lib = lib
    return _lib


def _check(status, name):
    if status != 0:
        msg = load().sg4d_error_string(status).decode()
        raise RuntimeError(f"{name} failed with status {status}: {msg}")


def stream_ptr(t):
    return torch.cuda.current_stream(t.device).cuda_stream


LAUNCH_COUNT = 0      # C-ABI calls issued (each launches at least one kernel); read by bench.py
_TIMING = None        # when a list: (name, first int args, start event, end event) per call


def enable_timing(on=True):
    """Per-call CUDA-event timing on the launching stream (bench.py's roofline numbers)."""
    global _TIMING
    _TIMING = [] if on else None


def drain_timing():
    """-> {(name, key): [ms, ...]} for the calls recorded since enable_timing(); synchronises."""
    global _TIMING
    rec, _TIMING = _TIMING or [], ([] if _TIMING is not None else None)
    torch.cuda.synchronize()
    out = {}
    for name, key, e0, e1 in rec:
        out.setdefault((name, key), []).append(e0.elapsed_time(e1))
    return out


_FN = {}


def call(name, ref, *args):
    """Invoke ``name`` on the current stream of ``ref``'s device."""
    global LAUNCH_COUNT
    fn = _FN.get(name)
    if fn is None:
        fn = _FN[name] = getattr(load(), name)
    index = ref.device.index
    if index is not None and index != torch._C._cuda_getDevice():
        with torch.cuda.device(ref.device):      # rare: a tensor of another device than the current one
            return call(name, ref, *args)
    LAUNCH_COUNT += 1
    # raw handle of torch's current stream on that device (what torch.cuda.current_stream().cuda_stream returns,
    # without building a Stream object on every call)
    stream = torch._C._cuda_getCurrentRawStream(torch._C._cuda_getDevice() if index is None else index)
    if _TIMING is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _check(fn(*args, stream), name)
        e1.record()
        _TIMING.append((name, tuple(a for a in args if isinstance(a, int) and 0 < a < (1 << 31)), e0, e1))
    else:
        _check(fn(*args, stream), name)


def ptr(t):
    return 0 if t is None else t.data_ptr()


class precision:
    """``with sg4d.precision("bf16"): ...`` -- bf16 operands / fp32 accumulation in every tensor-core layer (BASELINE.json
    configs[3]; the reference's counterpart is Lightning's ``precision=16`` autocast, SGP/main.py:62-64).  Process-wide."""
    MODES = {"fp32": 0, "f32": 0, 32: 0, "bf16": 1, 16: 1}

    def __init__(self, mode):
        self.mode = self.MODES[mode]

    def __enter__(self):
        self.prev = load().sg4d_get_compute_precision()
        _check(load().sg4d_set_compute_precision(self.mode), "sg4d_set_compute_precision")
        return self

    def __exit__(self, *a):
        load().sg4d_set_compute_precision(self.prev)


def set_precision(mode):
    _check(load().sg4d_set_compute_precision(precision.MODES[mode]), "sg4d_set_compute_precision")


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            # same wording as the reference's host wrappers (sampling.cpp:83, ball_query.cpp:28, ...)
            raise RuntimeError("CPU not supported")
