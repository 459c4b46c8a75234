"""Inference epilogue of the reference's ``main.py`` (mode 'infer', SGP/main.py:92-115) around
``SGPNModelWrapper.predict_step`` (SGH/model/scene_graph_prediction_model.py:157-177): per scan, the predicted
(subject, predicate, object) triples without the 'none' class, and the ``scan_relations_<name>_<split>.json`` file
every downstream consumer reads (role prediction, phase recognition).

The reference compares ``rel == none_id`` on device tensors inside a Python loop (one device->host sync per edge);
here the arg-max is read back once per scan.
"""
import json
import os

import torch


@torch.no_grad()
def infer_scans(model, batches):
    """``{scan_id: [(subject, predicate, object), ...]}`` for an iterable of single-scan batch dicts (device tensors;
    ``objs_json`` maps 1-based instance ids to names, ``scan_id`` is the key -- the reference's collate contract)."""
    was_training = model.training
    model.eval()
    try:
        results = {}
        for i, batch in enumerate(batches):
            scan_id, relations = model.predict_step(batch, i)
            results[scan_id] = relations
        return results
    finally:
        model.train(was_training)


def dump_scan_relations(results, name, split="test", directory="."):
    """Writes ``scan_relations_<name>_<split>.json`` (SGP/main.py:111-115): tuples become JSON lists."""
    path = os.path.join(directory, f"scan_relations_{name}_{split}.json")
    with open(path, "w") as f:
        json.dump({k: [list(t) for t in v] for k, v in results.items()}, f)
    return path
