"""Import shim: exposes the package directory ``4d-or_b200/`` (whose name is not a Python identifier)
as the importable package ``sg4d``.  ``import sg4d`` / ``import sg4d.model`` / ``from sg4d.pointnet2_ops
import pointnet2_utils`` all resolve into that directory."""
import os as _os

__path__ = [_os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "4d-or_b200")]
__package__ = "sg4d"
if __spec__ is not None:
    __spec__.submodule_search_locations = list(__path__)
with open(_os.path.join(__path__[0], "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(__path__[0], "__init__.py"), "exec"))
