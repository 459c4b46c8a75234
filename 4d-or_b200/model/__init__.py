"""Host-side mirror of the reference's scene-graph model (module and attribute names, constructor
signatures and ``state_dict`` keys preserved) on top of the sg4d kernels."""
from .scene_graph_prediction_model import SGPNModelWrapper  # noqa: F401
from .inference import dump_scan_relations, infer_scans  # noqa: F401
