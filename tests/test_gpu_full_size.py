"""Parity at BENCHMARK size (BASELINE.json configs[2] shape: 12 object + 66 edge clouds of 80 000 points per scene).

(1) one full scene through the sg4d model (train-mode BatchNorm) vs the CPU oracle (oracle/model_ref.py = the
    reference's op sequence in stock PyTorch over the C restatement of its kernels) on the same inputs and weights:
    the 256-d encoder features of all 78 clouds at 1e-4, logits / loss at the conditioning-aware bound;
(2) every set-abstraction scale of that scene (2 encoders x 4 scales, captured with its real FPS / ball-query
    indices: up to 1.08 M grouped rows = 57 tiles per CTA through the persistent kernels) re-run forward + backward
    and compared IN FULL -- every pooled feature, every parameter gradient, the feature gradient -- with the fp64
    evaluation with pinned selections (tests/sa_ref.py) at 1e-4, no outliers."""
import json
import os

import pytest
import torch

import sa_ref
from oracle import model_ref, weights

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CFG = json.load(open(os.path.join(ROOT, "tests", "golden", "no_gt.json")))


def test_full_size_scene_against_oracle_and_pinned_fp64(cuda):
    from sg4d import mlp, synthetic
    from sg4d.model import SGPNModelWrapper
    sd = weights.synth_state_dict(seed=0)
    cfg = json.loads(json.dumps(CFG))
    cfg["MODEL"]["lambda_o"] = 0.1
    m = SGPNModelWrapper(cfg, 12, 15, torch.ones(12), torch.ones(15), [f"r{i}" for i in range(14)] + ["none"])
    m.load_state_dict(sd)
    m.to(cuda).train()
    m.obj_predictor.dropout.eval()
    m.rel_predictor.dropout.eval()
    batch = synthetic.make_scene(3, n_obj=12, n_points_obj=80000, n_points_rel=80000)
    db = synthetic.to_device(batch, cuda)
    mlp.CAPTURE = []
    try:
        outs = m(db, return_meta_data=True)
        caps = [q for q in mlp.CAPTURE if "mlp" in q]
    finally:
        mlp.CAPTURE = None
    loss = m.loss(outs[0], outs[1], db)
    loss.backward()
    assert len(caps) == 8 and sorted(q["kind"] for q in caps) == ["sa1"] * 4 + ["sa2"] * 4
    assert max(q["idx"].numel() for q in caps) == 66 * 512 * 32

    # ---- (1) the CPU oracle on the same scene
    s = model_ref.clone_state(sd)
    with torch.no_grad():
        want = model_ref.forward(s, batch, training=True, dropout=False)
    for name, a, b in (("obj_feature", outs[2], want[2]), ("rel_feature", outs[3], want[3])):
        err = float((a.detach().cpu() - b).abs().max())
        assert err <= 1e-4, (name, err)
    # logits pass through BatchNorm1d over 12 / 66 rows (GCN): the measured conditioning bound of tests/test_gpu_model.py
    for name, a, b in (("obj_cls", outs[0], want[0]), ("rel_cls", outs[1], want[1])):
        assert float((a.detach().cpu() - b).abs().max()) <= 5e-4, name
    want_loss = model_ref.loss_fn(want[0], want[1], batch, torch.ones(12), torch.ones(15), 0.1)
    assert abs(float(loss) - float(want_loss)) <= 5e-4
    for p_ in m.parameters():
        p_.grad = None

    # ---- (2) every scale at full size, in full, against the pinned fp64 evaluation
    report = {}
    for i, q in enumerate(caps):
        net = q["mlp"]
        net.zero_grad(set_to_none=True)
        sa_ref.scale_parity(q["kind"], q["pts"], q["feats"] if q["feats"] is not q["pts"] else None, q["foff"], q["c"],
                            q["centers"], q["idx"], q["cnt"], net, 100 + i)
        report[f"{q['kind']}[{i}] rows={q['idx'].numel()}"] = list(sa_ref.FAILS)
        print(f"{q['kind']}[{i}] rows={q['idx'].numel()}: " + ", ".join(
            f"{n} {e:.1e}/{w:.1e}" + (f" (torch fp32 {e32:.1e})" if e32 is not None else "") for n, e, w, e32 in sa_ref.LOG))
        sa_ref.FAILS.clear()
        sa_ref.LOG.clear()
        torch.cuda.empty_cache()
    assert not any(report.values()), report
