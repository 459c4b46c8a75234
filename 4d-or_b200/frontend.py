"""GPU crop / sample front-end (``csrc/frontend.cu``, include/sg4d.h section 6).

``prepare_scene`` turns one scene -- points (P, 6) = xyz + rgb and per-point object masks -- into the batch dict the model
consumes (``obj_points`` (n_obj, 6, N), ``rel_points`` (E, 7, N_rel) as (B, C, N) views of point-major tensors,
``edge_indices``), i.e. what the reference's ``data_preparation`` (SGH/dataset/data_preparation_utils.py:52-240) builds on the
CPU with numpy / open3d before the hot path.  Everything runs on the device the scene lives on; the random draws are
uniforms generated on that device (or passed in, which is how the parity test pins them).
"""
import torch

from . import _lib, synthetic


def prepare_scene(points, masks, n_obj, num_points, num_points_union, pairs="ordered", padding=0.2, u_obj=None, u_rel=None,
                  generator=None, return_debug=False, out_obj=None, out_rel=None):
    """points (P, S) float32 cuda, masks (P,) int32 cuda with values 0..n_obj.  Returns the batch dict (+ debug tensors).
    out_obj (n_obj, num_points, S) / out_rel (E, num_points_union, S + 1): optional contiguous destinations, e.g. this
    scene's slices of a whole batch's tensors (no concatenation copy afterwards)."""
    _lib.require_cuda(points, masks)
    if points.dtype != torch.float32 or points.dim() != 2 or not points.is_contiguous() or masks.dtype != torch.int32:
        raise RuntimeError("points must be a contiguous (P, S) float tensor and masks an int32 tensor")
    dev = points.device
    P, S = points.shape
    edges = synthetic.edge_list(n_obj, pairs).to(dev)      # fully connected (:128-133); 'unordered' = BASELINE's 66 edges
    E = edges.shape[1]
    lib = _lib.load()
    ws = torch.empty(lib.sg4d_frontend_workspace_bytes(P, n_obj, E) // 4 + 1, dtype=torch.int32, device=dev)
    obj_list = torch.empty(n_obj, P, dtype=torch.int32, device=dev)
    totals = torch.empty(n_obj + 1, dtype=torch.int32, device=dev)
    obj_box = torch.empty(n_obj, 6, dtype=torch.float32, device=dev)
    _lib.call("sg4d_frontend_objects", points, P, S, n_obj, points.data_ptr(), masks.data_ptr(), float(padding), ws.data_ptr(),
              obj_list.data_ptr(), totals.data_ptr(), obj_box.data_ptr())
    edge_list = torch.empty(E, P, dtype=torch.int32, device=dev)
    edge_totals = torch.empty(E, 2, dtype=torch.int32, device=dev)
    edge_box = torch.empty(E, 6, dtype=torch.float32, device=dev)
    _lib.call("sg4d_frontend_edges", points, P, S, E, points.data_ptr(), masks.data_ptr(), edges.data_ptr(), obj_box.data_ptr(),
              ws.data_ptr(), edge_list.data_ptr(), edge_totals.data_ptr(), edge_box.data_ptr())
    if u_obj is None:
        u_obj = torch.rand(n_obj, num_points, device=dev, generator=generator)
    if u_rel is None:
        u_rel = torch.rand(E, num_points_union, device=dev, generator=generator)

    def sample(clouds, n, lst, tot, e, u, fout, out):
        if out is None:
            out = torch.empty(clouds, n, fout, dtype=torch.float32, device=dev)
        elif out.shape != (clouds, n, fout) or out.dtype != torch.float32 or not out.is_contiguous() or out.device != dev:
            raise RuntimeError(f"destination must be a contiguous float32 ({clouds}, {n}, {fout}) tensor on {dev}")
        picked = torch.empty(clouds, n, dtype=torch.int32, device=dev) if return_debug else None
        mean = torch.empty(clouds, 3, dtype=torch.float32, device=dev)
        dist = torch.empty(clouds, dtype=torch.float32, device=dev)
        scratch = torch.empty(clouds * (((n + 255) // 256) * 6 + 1 + n) + 2, dtype=torch.int32, device=dev)
        _lib.call("sg4d_frontend_sample", points, P, S, clouds, n, points.data_ptr(), masks.data_ptr(), lst.data_ptr(), tot.data_ptr(),
                  _lib.ptr(e), u.data_ptr(), out.data_ptr(), _lib.ptr(picked), mean.data_ptr(), dist.data_ptr(), scratch.data_ptr())
        return out, picked, mean, dist

    obj, obj_pick, _, _ = sample(n_obj, num_points, obj_list, totals, None, u_obj.contiguous(), S, out_obj)
    rel, rel_pick, rel_mean, rel_dist = sample(E, num_points_union, edge_list, edge_totals, edges, u_rel.contiguous(), S + 1, out_rel)
    batch = {"obj_points": obj.permute(0, 2, 1), "rel_points": rel.permute(0, 2, 1), "edge_indices": edges}
    if return_debug:
        batch["_debug"] = {"obj_picked": obj_pick, "rel_picked": rel_pick, "obj_box": obj_box, "edge_box": edge_box,
                           "obj_totals": totals, "edge_totals": edge_totals[:, 1], "rel_mean": rel_mean, "rel_dist": rel_dist}
    return batch
