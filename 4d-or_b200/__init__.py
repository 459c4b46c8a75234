"""sg4d -- B200-native (sm_100a) implementation of 4D-OR's scene-graph prediction hot path.

The directory is named ``4d-or_b200`` (not an importable identifier); import it as ``sg4d`` through the
``sg4d.py`` shim at the repository root.  Layout:

  csrc/            hand-written CUDA kernels + the C ABI (``include/sg4d.h``) -> ``libsg4d.so``
  _lib.py          ctypes binding of the C ABI (no fallback: missing library = error)
  pointnet2_ops/   mirror of the reference's operator API (``_ext``, ``pointnet2_utils``, ``pointnet2_modules``)
  rows.py          point-major fused operators used by the model path
  mlp.py           fused set-abstraction scales on the tensor cores (ball-query indices -> pooled features)
  dense.py         dense layers on the same engine (SA3, TripletGCN MLPs, heads)
  frontend.py      GPU crop / sample front-end (scene -> object / edge clouds)
  metrics.py       per-take relation metrics (macro / weighted precision, recall, F1)
  model/           mirror of the reference's model API (``SGPNModelWrapper`` and its sub-modules)
  parallel.py      scene-sharded data parallelism (one NCCL gradient all-reduce per step), host->device prefetcher
  trainer.py       fit loop with the reference's per-epoch checkpoints / resume (no Lightning needed)
  synthetic.py     synthetic scenes of the benchmark shapes
"""
from . import _lib  # noqa: F401
from ._lib import precision, set_precision  # noqa: F401

__all__ = ["_lib", "pointnet2_ops", "rows", "mlp", "dense", "frontend", "metrics", "model", "parallel", "synthetic", "trainer",
           "precision", "set_precision"]


def library_path():
    return _lib.SO_PATH
