"""Drop-in mirror of the reference's ``pointnet2_ops`` package (operator API): same module names
(``pointnet2_utils``, ``pointnet2_modules``, ``_ext``), same callables, same shapes and dtypes
(OPS/pointnet2_utils.py, OPS/pointnet2_modules.py), backed by ``libsg4d.so``.
"""
from . import _ext, pointnet2_modules, pointnet2_utils  # noqa: F401

__version__ = "3.0.0+sg4d"
