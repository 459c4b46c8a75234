"""Per-take and epoch relation metrics with the reference's bookkeeping
(SGH/model/scene_graph_prediction_model.py:113-132 ``reset_metrics`` / ``update_metrics``, :195-238 ``evaluate_predictions``).

The reference calls ``sklearn.metrics.classification_report(gts, preds, labels=range(n_rel), output_dict=True)`` per take and
over all takes and logs precision / recall / F1 per relation plus the ``macro avg`` and ``weighted avg`` rows (which it names
Epoch_Macro / Epoch_Micro).  ``classification_report`` below computes exactly those numbers (zero_division = 0, like sklearn's
default with a warning) without the sklearn dependency; tests/test_host_logic.py checks it against sklearn where installed.
Predictions are accumulated ON THE DEVICE and read back once per epoch (the reference syncs every step).
"""
from collections import defaultdict

import torch


def classification_report(gts, preds, n_labels, names=None):
    """-> {name: {'precision', 'recall', 'f1-score', 'support'}, 'macro avg': ..., 'weighted avg': ..., 'accuracy': float}"""
    gts = torch.as_tensor(gts, dtype=torch.long).flatten().cpu()
    preds = torch.as_tensor(preds, dtype=torch.long).flatten().cpu()
    names = list(names) if names is not None else [str(i) for i in range(n_labels)]
    conf = torch.zeros(n_labels, n_labels, dtype=torch.long)
    ok = (gts >= 0) & (gts < n_labels) & (preds >= 0) & (preds < n_labels)
    conf.index_put_((gts[ok], preds[ok]), torch.ones(int(ok.sum()), dtype=torch.long), accumulate=True)
    tp = conf.diag().double()
    support = conf.sum(1).double() + torch.bincount(gts[~ok & (gts >= 0) & (gts < n_labels)], minlength=n_labels).double()
    pred_cnt = conf.sum(0).double()
    prec = torch.where(pred_cnt > 0, tp / pred_cnt.clamp(min=1), torch.zeros_like(tp))
    rec = torch.where(support > 0, tp / support.clamp(min=1), torch.zeros_like(tp))
    f1 = torch.where(prec + rec > 0, 2 * prec * rec / (prec + rec).clamp(min=1e-300), torch.zeros_like(tp))
    out = {nm: {"precision": float(prec[i]), "recall": float(rec[i]), "f1-score": float(f1[i]), "support": int(support[i])}
           for i, nm in enumerate(names)}
    tot = float(support.sum())
    w = support / tot if tot > 0 else torch.zeros_like(support)
    out["macro avg"] = {"precision": float(prec.mean()), "recall": float(rec.mean()), "f1-score": float(f1.mean()), "support": int(tot)}
    out["weighted avg"] = {"precision": float((prec * w).sum()), "recall": float((rec * w).sum()), "f1-score": float((f1 * w).sum()),
                           "support": int(tot)}
    out["accuracy"] = float((gts == preds).double().mean()) if len(gts) else 0.0
    return out


class RelationMetrics:
    """``self.{train,val}_take_rel_{preds,gts}`` of the reference model, kept as device tensors until the epoch ends."""

    def __init__(self, relation_names):
        self.names = list(relation_names)
        self.reset()

    def reset(self, split=None):
        for s in (("train", "val") if split is None else (split,)):
            setattr(self, f"{s}_preds", defaultdict(list))
            setattr(self, f"{s}_gts", defaultdict(list))

    def update(self, batch, rel_pred, split="train"):
        """rel_pred (E, n_rel) log-probabilities; batch['take_idx'] (scalar per scene, or (E,) per edge for concatenated scenes)"""
        pred = rel_pred.detach().argmax(1)
        gts = batch["gt_rels"].detach()
        take = batch.get("take_idx", 0)
        preds_d, gts_d = getattr(self, f"{split}_preds"), getattr(self, f"{split}_gts")
        if torch.is_tensor(take) and take.numel() > 1:
            for t in take.unique().tolist():
                sel = take == t
                preds_d[int(t)].append(pred[sel])
                gts_d[int(t)].append(gts[sel])
        else:
            t = int(take) if not torch.is_tensor(take) else int(take.item())
            preds_d[t].append(pred)
            gts_d[t].append(gts)

    def evaluate(self, split):
        """-> {'takes': {take: report}, 'all': report, 'macro_f1': float} (one device->host read per take)"""
        preds_d, gts_d = getattr(self, f"{split}_preds"), getattr(self, f"{split}_gts")
        takes, all_p, all_g = {}, [], []
        for t in sorted(preds_d.keys()):
            p, g = torch.cat(preds_d[t]).cpu(), torch.cat(gts_d[t]).cpu()
            takes[t] = classification_report(g, p, len(self.names), self.names)
            all_p.append(p)
            all_g.append(g)
        rep = classification_report(torch.cat(all_g), torch.cat(all_p), len(self.names), self.names) if all_p else None
        return {"takes": takes, "all": rep, "macro_f1": rep["macro avg"]["f1-score"] if rep else 0.0}
